"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def build_config(latent, n_style=1, steps=50, precision="bf16", agg="mean"):
    from stedm_b200.config import load_config
    sampling = "mp" if n_style > 1 else "augmented"
    ov = [f"style_agg={agg}", f"style_sampling={sampling}", f"ddim_steps={steps}", f"diffusion.image_size={latent}",
          f"data.patch_size={latent * 4}", f"precision={precision}"]
    cfg = load_config(ov)
    if n_style > 1:
        cfg.style_sampling.num_patches = n_style
    return cfg


_MODELS = {}


def build_model(latent, n_style=1, precision="bf16", device="cuda", steps=50):
    """LDM_Diffusion (product code) with fixture weights, cached per (latent, n_style)."""
    from stedm_b200.modules.ldm_diffusion import LDM_Diffusion
    from stedm_b200.utils.fixture import apply_fixture_weights
    key = (latent, n_style, device)
    if key not in _MODELS:
        cfg = build_config(latent, n_style, steps, precision)
        m = LDM_Diffusion(cfg, precision=precision, load_first_stage_ckpt=False)
        apply_fixture_weights(m._model, seed=0)
        m = m.to(device).eval()
        _MODELS.clear()
        _MODELS[key] = m
    m = _MODELS[key]
    m._cfg.ddim_steps = steps
    m._model.set_precision(precision)
    return m


def oracle_state_dict(model):
    """fp32 CPU state dict of a product model under the oracle's key names."""
    sd = {}
    for k, v in model.state_dict().items():
        if k.startswith("_agg_block."):
            continue
        k = k.replace("agg_block._embedder.", "agg_block.embedder.")
        sd[k] = v.detach().cpu().float() if v.is_floating_point() else v.detach().cpu()
    return sd


def rel_err(a, b):
    """max|a-b| / max|b| — the 'relative' figure the bf16-mode bar is stated in."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def max_abs(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max())


def psnr(a, b, peak=2.0):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    mse = float(((a - b) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)
