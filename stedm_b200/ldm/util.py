"""Plugin seam of the reference (ldm/util.py:78-93): classes are looked up by dotted ``target:`` strings.

The B200 drop-in keeps that mechanism and adds one thing: ``target`` strings that name the REFERENCE's classes
for the sampling path (``ldm.…``, ``networks.…``, ``modules.…``) are redirected to the same-named classes of
this package, so the reference's hydra configs (conf/diffusion/**) work unchanged.
"""
import importlib

_REDIRECT_ROOTS = ("ldm.", "networks.", "modules.")
PACKAGE = "stedm_b200"


def resolve_target(target: str) -> str:
    if target.startswith(PACKAGE + "."):
        return target
    if target.startswith(_REDIRECT_ROOTS):
        return f"{PACKAGE}.{target}"
    return target


def get_obj_from_str(string, reload=False):
    module, cls = resolve_target(string).rsplit(".", 1)
    mod = importlib.import_module(module)
    if reload:
        mod = importlib.reload(mod)
    return getattr(mod, cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**dict(config.get("params", dict()) or {}))


def exists(x):
    return x is not None


def default(val, d):
    if val is not None:
        return val
    return d() if callable(d) else d
