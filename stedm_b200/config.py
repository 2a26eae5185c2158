"""Hydra-free loader for the conf/ tree (hydra and omegaconf are not available offline): composes the
``defaults:`` lists of conf/config_predict.yaml the way hydra does for this tree and applies ``a.b=c`` overrides."""
import copy
import os

import yaml

CONF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conf")


class AttrDict(dict):
    """dict with attribute access (what the reference reads from its DictConfig nodes)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def _wrap(node):
    if isinstance(node, dict):
        return AttrDict({k: _wrap(v) for k, v in node.items()})
    if isinstance(node, list):
        return [_wrap(v) for v in node]
    return node


def _load_group(rel_dir, name, choices, prefix):
    with open(os.path.join(CONF_DIR, rel_dir, name + ".yaml")) as fh:
        node = yaml.safe_load(fh) or {}
    defaults = node.pop("defaults", [])
    out = {}
    for d in defaults:
        if d == "_self_":
            out.update(node)
        elif isinstance(d, dict):
            for group, choice in d.items():
                if group.startswith("override "):
                    continue
                key = f"{prefix}{group}"
                choice = choices.get(key, choice)
                sub_dir = os.path.join(rel_dir, group) if rel_dir else group
                out[group] = _load_group(sub_dir, choice, choices, key + ".")
    if "_self_" not in defaults:
        merged = dict(node)
        merged.update(out)
        out = merged
    return out


def load_config(overrides=(), config_name="config_predict"):
    """``overrides``: hydra-style strings, e.g. ["style_agg=mean", "ddim_steps=50", "diffusion.image_size=64"]."""
    choices, sets = {}, []
    groups = {"data", "location", "diffusion", "style_sampling", "style_agg", "diffusion.unet_config",
              "diffusion.first_stage_config", "diffusion.cond_stage_config"}
    for ov in overrides:
        k, v = ov.lstrip("+").split("=", 1)
        if k in groups:
            choices[k] = v
        else:
            sets.append((k, yaml.safe_load(v)))
    cfg = _load_group("", config_name, choices, "")
    for k, v in sets:
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return _wrap(copy.deepcopy(cfg))
