"""The CPU oracle against golden vectors produced by the reference's own code (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import stedm_oracle as O
from stedm_b200.utils.fixture import fixture_tensor
from tests.util import load_golden, max_abs

import json
import os


@pytest.fixture(scope="module")
def fixture_sd(golden_dir):
    """Fixture weights for every reference parameter name (shapes from the committed key list)."""
    keys = json.load(open(os.path.join(golden_dir, "reference_state_dict_keys.json")))
    sd = {}
    for k, shape in keys.items():
        if k.startswith(("model_ema.", "_agg_block.")) or not (k.endswith(".weight") or k.endswith(".bias")):
            continue
        if "relative_position" in k:
            continue
        kk = k.replace("agg_block._embedder.", "agg_block.embedder.")
        sd[kk] = fixture_tensor(k, shape, seed=0)
    return sd


def test_cfg_rescale_dims_quirk():
    """std over dims (1,2) = channels and height, NOT width: shape (B,1,1,W) (ddim.py:182-183)."""
    g = torch.Generator().manual_seed(0)
    e_c, e_u = torch.randn(2, 3, 8, 5, generator=g), torch.randn(2, 3, 8, 5, generator=g)
    out = O.cfg_combine(e_c, e_u, 1.5)
    e_w = e_u + 1.5 * (e_c - e_u)
    col = 3
    r = e_c[:, :, :, col].reshape(2, -1).std(dim=1) / e_w[:, :, :, col].reshape(2, -1).std(dim=1)
    want = e_w[:, :, :, col] * r[:, None, None] * 0.7 + 0.3 * e_c[:, :, :, col]
    assert max_abs(out[:, :, :, col], want) < 1e-6


def test_oracle_eps_and_step_small(fixture_sd):
    g = load_golden("small_b2_l32")
    _, _, x_T = O.synthetic_batch(2, 128, 2, 0)
    cond = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["c_crossattn"])]}
    unc = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["uc_crossattn"])]}
    t = torch.full((2,), 981, dtype=torch.long)
    with torch.no_grad():
        e_c = O.apply_model(fixture_sd, x_T, t, cond)
        e_u = O.apply_model(fixture_sd, x_T, t, unc)
    assert max_abs(e_c, g["eps_c_981"]) < 1e-5 and max_abs(e_u, g["eps_u_981"]) < 1e-5
    tab = O.ddim_tables(50)
    e = O.cfg_combine(e_c, e_u, 1.5)
    x1, p0 = O.ddim_update(x_T, e, tab["a_t"][49], tab["a_prev"][49], tab["sigma"][49], tab["sqrt_one_minus_a"][49])
    assert max_abs(x1, g["x_after_1"]) < 1e-4 and max_abs(p0, g["pred_x0_step0"]) < 1e-2  # pred_x0 ~ 84x larger


def test_oracle_decode_small(fixture_sd):
    g = load_golden("small_b2_l32")
    z = torch.from_numpy(g["z_final"])
    with torch.no_grad():
        zq, idx = O.vq_quantize(z, fixture_sd["first_stage_model.quantize.embedding.weight"])
        assert (idx.numpy() == g["vq_idx"]).all()
        dec = O.decode_first_stage(fixture_sd, z)
    assert max_abs(dec, g["dec_quant"]) < 1e-4
    assert (O.to_uint8(dec) == g["img_u8"]).mean() > 0.999


def test_oracle_conditioning_small(fixture_sd):
    g = load_golden("small_b2_l32")
    seg, style, _ = O.synthetic_batch(2, 128, 2, 0)
    c = O.get_conditioning(fixture_sd, seg, style)
    assert max_abs(c["c_concat"][0], g["c_concat"]) < 1e-6
    assert max_abs(c["c_crossattn"][0], g["c_crossattn"]) < 1e-4


def test_oracle_plms_matches_reference_sampler_golden():
    """oracle.plms_sample == the reference's PLMSSampler run on the stand-in model of oracle/make_golden.py
    (tests/golden/plms.npz): pins timestep order, multistep coefficients, plain CFG and the x_prev update."""
    import torch
    from oracle import stedm_oracle as O
    eps = lambda x, t, c: torch.tanh(0.3 * x + c) * (1.0 + t.float().view(-1, 1, 1, 1) / 1000.0)
    g = torch.Generator().manual_seed(21)
    x_T = torch.randn(2, 3, 8, 8, generator=g)
    c, uc = torch.randn(2, 3, 8, 8, generator=g) * 0.5, torch.zeros(2, 3, 8, 8)
    gold = load_golden("plms")
    got1 = O.plms_sample(lambda x, t: eps(x, t, c), x_T, S=20)
    got3 = O.plms_sample(lambda x, t: eps(x, t, c), x_T, S=20, cfg_scale=3.0, uncond_eps_fn=lambda x, t: eps(x, t, uc))
    assert max_abs(got1, gold["plms_s20_cfg1"]) < 1e-6 and max_abs(got3, gold["plms_s20_cfg3"]) < 1e-6


def test_oracle_dpm_solver_matches_reference_golden(fixture_sd):
    """oracle.dpm_solver_sample (multistep order 2, data prediction, plain CFG) == the reference's DPMSolverSampler run on
    its own S_ZSS_DM with fixture weights (tests/golden/dpm_solver.npz), S = 12: ten second-order steps, first-order
    init and final step."""
    import torch
    from oracle import stedm_oracle as O
    g = load_golden("small_b2_l32")
    cond = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["c_crossattn"])]}
    unc = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["uc_crossattn"])]}
    _, _, x_T = O.synthetic_batch(2, 128, 2, 0)
    ac, _ = O.alphas_cumprod_linear()
    with torch.no_grad():
        got = O.dpm_solver_sample(lambda x, t: O.apply_model(fixture_sd, x, t, cond), x_T, ac, 12, cfg_scale=1.5,
                                  uncond_eps_fn=lambda x, t: O.apply_model(fixture_sd, x, t, unc))
    want = load_golden("dpm_solver")["dpm_s12"]
    assert max_abs(got, want) < 1e-5 * float(abs(want).max())


def test_oracle_ddpm_ancestral_matches_reference_golden(fixture_sd):
    """a17: oracle.ddpm_ancestral_sample vs the reference's LatentDiffusion.sample(timesteps=5) with the stored noise
    (tests/golden/ancestral.npz), plain and with quantize_denoised."""
    g = load_golden("small_b2_l32")
    gold = load_golden("ancestral")
    _, _, x_T = O.synthetic_batch(2, 128, 2, 0)
    cond = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["c_crossattn"])]}
    noises = torch.from_numpy(gold["noises"])
    T = int(gold["timesteps"])
    with torch.no_grad():
        z = O.ddpm_ancestral_sample(fixture_sd, cond, x_T, T, noises)
        zq = O.ddpm_ancestral_sample(fixture_sd, cond, x_T * 40.0, T, noises, quantize_denoised=True)
    assert max_abs(z, gold["ancestral_t5"]) < 1e-4
    assert max_abs(zq, gold["ancestral_t5_quant"]) < 1e-4 * float(np.abs(gold["ancestral_t5_quant"]).max())
