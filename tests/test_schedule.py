"""Oracle and product schedule math against the reference's own functions (tests/golden/sched.npz was written by
oracle/make_golden.py from ldm/modules/diffusionmodules/util.py) and SURVEY.md §A.4's analytic known answers."""
import numpy as np
import torch

from oracle import stedm_oracle as O
from tests.util import load_golden


def test_known_answers():
    ac32, ac64 = O.alphas_cumprod_linear()
    assert abs(ac64[1] - 0.996994152597893) < 1e-15
    assert abs(ac64[981] - 0.00014004594585901366) < 1e-18
    assert ac32[981] == np.float32(1.4004594e-4) and ac32[0] == np.float32(0.9985)
    ts = O.ddim_timesteps(50)
    assert ts[0] == 1 and ts[1] == 21 and ts[-1] == 981 and len(ts) == 50
    assert len(O.ddim_timesteps(128)) == 143            # c = 1000 // 128 = 7
    tab = O.ddim_tables(50)
    assert abs(1.0 / np.sqrt(tab["a_t"][49]) - 84.501564) < 1e-3
    assert tab["a_prev"][0] == np.float32(0.9985) and abs(tab["sqrt_one_minus_a"][0] - 0.054825746) < 1e-8
    e = O.timestep_embedding(torch.tensor([981]), 128)[0]
    assert np.allclose(e[:3].numpy(), [0.6800, 0.2858, 0.8710], atol=1e-4)
    assert np.allclose(e[64:67].numpy(), [0.7333, 0.9583, 0.4913], atol=1e-4)


def test_oracle_matches_reference_tables():
    g = load_golden("sched")
    ac32, _ = O.alphas_cumprod_linear()
    assert (ac32 == g["alphas_cumprod"]).all()
    for S in (50, 128, 20):
        tab = O.ddim_tables(S)
        assert (tab["timesteps"] == g[f"ts_{S}"]).all()
        assert (tab["a_t"] == g[f"a_{S}"]).all()
        assert (tab["a_prev"] == g[f"a_prev_{S}"].astype(np.float32)).all()
        assert (tab["sqrt_one_minus_a"] == g[f"sqrt1m_{S}"]).all()
    assert np.allclose(O.ddim_tables(50, eta=0.5)["sigma"], g["sigma_50_eta05"].astype(np.float32), rtol=1e-6)
    assert (O.timestep_embedding(torch.tensor([981, 481, 1]), 128).numpy() == g["temb_128"]).all()


def test_product_schedule_matches_reference():
    """The product's host-side schedule code (stedm_b200.ldm...util) against the same golden tables."""
    from stedm_b200.ldm.modules.diffusionmodules.util import (make_beta_schedule, make_ddim_sampling_parameters,
                                                              make_ddim_timesteps)
    g = load_golden("sched")
    betas = make_beta_schedule("linear", 1000, linear_start=0.0015, linear_end=0.0205)
    ac = torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)
    assert (ac.numpy() == g["alphas_cumprod"]).all()
    for S in (50, 128, 20):
        ts = make_ddim_timesteps("uniform", S, 1000, verbose=False)
        assert (ts == g[f"ts_{S}"]).all()
        sig, a, ap = make_ddim_sampling_parameters(ac, ts, 0.0, verbose=False)
        assert (a == g[f"a_{S}"]).all() and (ap == g[f"a_prev_{S}"]).all() and (sig == 0).all()
        assert (np.sqrt(1.0 - a) == g[f"sqrt1m_{S}"]).all()
    sig, _, _ = make_ddim_sampling_parameters(ac, g["ts_50"], 0.5, verbose=False)
    # eta > 0 is outside the north-star config; the reference mixes fp32 tensors into this numpy expression, so
    # agreement is to fp32 rounding, which is what torch.full(..., sigmas[index]) keeps anyway
    assert np.allclose(sig, g["sigma_50_eta05"], rtol=1e-6)
