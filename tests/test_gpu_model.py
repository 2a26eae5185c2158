"""End-to-end parity of the native sampling path on a B200 against golden vectors produced by the reference's own
code (tests/golden/*.npz, oracle/make_golden.py) and against the CPU oracle run live on the same seeded inputs.

Bars (north_star): per-step eps within 1e-4 max-abs in fp32 mode and 2e-2 relative (max|d|/max|ref|) in bf16
mode; decoded images at PSNR >= 40 dB (peak-to-peak 2) against the reference.
"""
import numpy as np
import pytest
import torch

from oracle import stedm_oracle as O
from tests.util import build_model, load_golden, max_abs, oracle_state_dict, psnr, rel_err

pytestmark = pytest.mark.gpu

FP32_EPS_BAR = 1e-4
BF16_EPS_BAR = 2e-2
PSNR_BAR = 40.0


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    """Library code outside the kernels (torchvision Swin) must not use TF32 while parity is measured."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _no_split_k():
    """Context: force the single-pass K loop (the latency mode's split-K changes the fp32 summation order with the
    number of output tiles, i.e. with the batch; bit-equality across batch sizes is a property of the throughput path)."""
    import contextlib
    from stedm_b200 import ops

    @contextlib.contextmanager
    def ctx():
        saved = ops.SPLIT_K[0]
        ops.SPLIT_K[0] = False
        try:
            yield
        finally:
            ops.SPLIT_K[0] = saved
    return ctx()


def _small_inputs():
    g = load_golden("small_b2_l32")
    seg, style, x_T = O.synthetic_batch(2, 128, 2, 0)
    return g, seg, style, x_T


def _cond(g, key, dev="cuda"):
    return {"c_concat": [torch.from_numpy(g["c_concat"]).to(dev)], "c_crossattn": [torch.from_numpy(g[key]).to(dev)]}


def test_conditioning_matches_reference():
    g, seg, style, _ = _small_inputs()
    m = build_model(32, n_style=2, precision="fp32")
    batch = {"image": torch.zeros(2, 128, 128, 3).cuda(), "segmentation": seg.cuda(), "style_imgs": style.cuda()}
    z, c = m._model.get_input(batch, "image")
    assert tuple(z.shape) == (2, 3, 32, 32)
    assert max_abs(c["c_concat"][0], g["c_concat"]) < 1e-6
    assert max_abs(c["c_crossattn"][0], g["c_crossattn"]) < FP32_EPS_BAR   # native fp32 style encoder vs the reference's
    unc = dict(batch, style_imgs=torch.zeros_like(batch["style_imgs"]) - 2)
    _, cu = m._model.get_input(unc, "image")
    assert max_abs(cu["c_crossattn"][0], g["uc_crossattn"]) < FP32_EPS_BAR


@pytest.mark.parametrize("t", [981, 481, 1])
def test_eps_fp32_mode(t):
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision="fp32")
    tt = torch.full((2,), t, dtype=torch.long, device="cuda")
    e_c = m._model.apply_model(x_T.cuda(), tt, _cond(g, "c_crossattn"))
    e_u = m._model.apply_model(x_T.cuda(), tt, _cond(g, "uc_crossattn"))
    assert max_abs(e_c, g[f"eps_c_{t}"]) < FP32_EPS_BAR, max_abs(e_c, g[f"eps_c_{t}"])
    assert max_abs(e_u, g[f"eps_u_{t}"]) < FP32_EPS_BAR


@pytest.mark.parametrize("t", [981, 481, 1])
def test_eps_bf16_mode(t):
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision="bf16")
    tt = torch.full((2,), t, dtype=torch.long, device="cuda")
    e_c = m._model.apply_model(x_T.cuda(), tt, _cond(g, "c_crossattn"))
    e_u = m._model.apply_model(x_T.cuda(), tt, _cond(g, "uc_crossattn"))
    r_c, r_u = rel_err(e_c, g[f"eps_c_{t}"]), rel_err(e_u, g[f"eps_u_{t}"])
    print(f"bf16 eps rel err t={t}: cond {r_c:.3e} uncond {r_u:.3e}")
    assert r_c < BF16_EPS_BAR and r_u < BF16_EPS_BAR, (r_c, r_u)


def test_unet_forward_api_equals_apply_model():
    """UNetModel.forward(cat([x, c_concat]), t, context) == apply_model(x, t, cond) (ddpm.py:1414-1417)."""
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision="fp32")
    tt = torch.full((2,), 481, dtype=torch.long, device="cuda")
    cond = _cond(g, "c_crossattn")
    a = m._model.apply_model(x_T.cuda(), tt, cond)
    b = m._model.model.diffusion_model(torch.cat([x_T.cuda(), cond["c_concat"][0]], 1), tt, context=cond["c_crossattn"][0])
    assert torch.equal(a, b)


def _sample(m, g, x_T, steps_limit=None, **kw):
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    model = m._model
    sampler = DDIMSampler(model, **kw)
    sampler.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
    cond, unc = _cond(g, "c_crossattn"), _cond(g, "uc_crossattn")
    img = x_T.cuda()
    total = sampler.ddim_timesteps.shape[0]
    xs = {}
    for i, step in enumerate(np.flip(sampler.ddim_timesteps)):
        if steps_limit is not None and i >= steps_limit:
            break
        ts = torch.full((img.shape[0],), int(step), device="cuda", dtype=torch.long)
        img, p0 = sampler.p_sample_ddim(img, cond, ts, index=total - i - 1, unconditional_guidance_scale=1.5,
                                        unconditional_conditioning=unc)
        xs[i + 1] = img
        if i == 0:
            xs["p0"] = p0
    return xs


def test_ddim_first_steps_fp32():
    g, _, _, x_T = _small_inputs()
    xs = _sample(build_model(32, n_style=2, precision="fp32"), g, x_T, steps_limit=3)
    for k in (1, 2, 3):
        assert rel_err(xs[k], g[f"x_after_{k}"]) < 1e-4, (k, rel_err(xs[k], g[f"x_after_{k}"]))
    assert rel_err(xs["p0"], g["pred_x0_step0"]) < 1e-4


def test_ddim_first_steps_bf16():
    g, _, _, x_T = _small_inputs()
    xs = _sample(build_model(32, n_style=2, precision="bf16"), g, x_T, steps_limit=3)
    for k in (1, 2, 3):
        r = rel_err(xs[k], g[f"x_after_{k}"])
        print(f"bf16 x after {k} steps rel err {r:.3e}")
        assert r < BF16_EPS_BAR


@pytest.mark.parametrize("precision,bar", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_full_ddim50_sample_log(precision, bar):
    """The whole DDIM-50 CFG loop through LatentDiffusion.sample_log vs the reference's final latent."""
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision=precision)
    z, inter = m._model.sample_log(_cond(g, "c_crossattn"), batch_size=2, ddim=True, ddim_steps=50, eta=0.0,
                                   log_every_t=1000, x_T=x_T.cuda(), unconditional_conditioning=_cond(g, "uc_crossattn"),
                                   unconditional_guidance_scale=1.5)
    assert len(inter["x_inter"]) == 3 and len(inter["pred_x0"]) == 3        # x_T, index == total-1, and index 0 (0 % log_every_t == 0)
    r = rel_err(z, g["z_final"])
    print(f"{precision} DDIM-50 final latent rel err {r:.3e}")
    assert r < bar, r


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_first_stage(precision):
    g, _, _, _ = _small_inputs()
    m = build_model(32, n_style=2, precision=precision)
    z = torch.from_numpy(g["z_final"]).cuda()
    dec_n = m._model.decode_first_stage(z, force_not_quantize=True)
    dec_q = m._model.decode_first_stage(z)
    p_n, p_q = psnr(dec_n.clamp(-1, 1), np.clip(g["dec_noquant"], -1, 1)), psnr(dec_q.clamp(-1, 1), np.clip(g["dec_quant"], -1, 1))
    print(f"{precision} decode PSNR: noquant {p_n:.1f} dB, quant {p_q:.1f} dB; max|d| {max_abs(dec_q, g['dec_quant']):.3e}")
    assert p_n >= PSNR_BAR and p_q >= PSNR_BAR
    if precision == "fp32":
        assert max_abs(dec_n, g["dec_noquant"]) < 1e-3 and max_abs(dec_q, g["dec_quant"]) < 1e-3
        from stedm_b200 import ops
        u8 = ops.image_to_uint8(dec_q.contiguous()).cpu().numpy()
        assert (u8 == g["img_u8"]).mean() > 0.995


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_predict_path_end_to_end(precision):
    """prepare_batch -> get_input x2 -> sample_log -> decode -> uint8 (modules/ldm_diffusion.py:76-96)."""
    g, seg, style, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision=precision)
    B, P = 2, 128
    seg_oh = seg.permute(0, 3, 1, 2).contiguous()
    batch = (torch.zeros(B, 3, P, P).cuda(), seg_oh.cuda(), None, style.permute(0, 1, 4, 2, 3).contiguous().cuda(),
             torch.arange(B))
    u8 = m.generate(m.prepare_batch(batch), x_T=x_T.cuda()).cpu().numpy()
    assert u8.shape == (B, P, P, 3) and u8.dtype == np.uint8
    ref = g["img_u8"].astype(np.float64)
    mse = ((u8.astype(np.float64) - ref) ** 2).mean()
    p = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"{precision} end-to-end uint8 PSNR vs reference: {p:.1f} dB")
    assert p >= PSNR_BAR      # bf16 measured 42.6 dB (round 1)


def test_batch_shard_invariance():
    """A rank-sharded run is bit-identical to the single-GPU run on the same per-sample inputs."""
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision="bf16")
    tt = torch.full((2,), 481, dtype=torch.long, device="cuda")
    with _no_split_k():      # (the planner splits K differently for 1 and 2 samples: compare like with like)
        full = m._model.apply_model(x_T.cuda(), tt, _cond(g, "c_crossattn"))
        for r in range(2):
            cond = {"c_concat": [torch.from_numpy(g["c_concat"][r:r + 1]).cuda()],
                    "c_crossattn": [torch.from_numpy(g["c_crossattn"][r:r + 1]).cuda()]}
            part = m._model.apply_model(x_T[r:r + 1].cuda(), tt[:1], cond)
            assert torch.equal(part[0], full[r])
    # default planning (split-K where few tiles leave SMs idle): the same sums in another fp32 order
    auto = m._model.apply_model(x_T.cuda(), tt, _cond(g, "c_crossattn"))
    r = rel_err(auto, full)
    print(f"split-K (default at this batch) vs single-pass eps rel err {r:.3e}")
    assert r < BF16_EPS_BAR and rel_err(auto, g["eps_c_481"]) < BF16_EPS_BAR     # bf16 rounding noise, like any reordering


def test_bench_geometry_b64_l64_rows_equal_b4_golden_run():
    """Parity at the BENCHMARKED geometry (BASELINE configs[1]: batch 64, latent 64, bf16; trunk at 64 samples, the
    rest at 128, where the tile / cluster / halo plan differs from the B = 4 golden case): one guided DDIM step and one
    apply_model at B = 64 whose samples 0-3 are the inputs of tests/golden/c1_b4_l64.npz.  Rows 0-3 must be BIT-EQUAL to
    the B = 4 run (shard invariance at the bench geometry) and within the bf16 bar of the reference's golden eps / x."""
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    g = load_golden("c1_b4_l64")
    _, _, x4 = O.synthetic_batch(4, 256, 1, 1)
    B = 64
    gen = torch.Generator().manual_seed(640)
    x = torch.cat([x4, torch.randn(B - 4, 3, 64, 64, generator=gen)], 0).cuda()
    cc = torch.cat([torch.from_numpy(g["c_concat"]), torch.randn(B - 4, 3, 64, 64, generator=gen)], 0).cuda()
    ca = torch.cat([torch.from_numpy(g["c_crossattn"]), torch.randn(B - 4, 512, generator=gen)], 0).cuda()
    cu = torch.cat([torch.from_numpy(g["uc_crossattn"]), torch.randn(B - 4, 512, generator=gen)], 0).cuda()
    m = build_model(64, n_style=1, precision="bf16")
    model = m._model
    mk = lambda a, b, n: {"c_concat": [a[:n].contiguous()], "c_crossattn": [b[:n].contiguous()]}
    with _no_split_k():
        for t in (981, 481):
            tt = torch.full((B,), t, dtype=torch.long, device="cuda")
            e64 = model.apply_model(x, tt, mk(cc, ca, B))
            e4 = model.apply_model(x[:4].contiguous(), tt[:4], mk(cc, ca, 4))
            assert torch.equal(e64[:4], e4), f"eps rows 0-3 at B=64 differ from the B=4 run (t={t})"
            r = rel_err(e64[:4], g[f"eps_c_{t}"])
            print(f"B=64 L=64 bf16 eps rows 0-3 vs reference golden, t={t}: rel err {r:.3e}")
            assert r < BF16_EPS_BAR, r
        s = DDIMSampler(model)
        s.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
        ts = torch.full((B,), 981, dtype=torch.long, device="cuda")
        x64, p64 = s.p_sample_ddim(x, mk(cc, ca, B), ts, index=49, unconditional_guidance_scale=1.5,
                                   unconditional_conditioning=mk(cc, cu, B))
        x4o, p4o = s.p_sample_ddim(x[:4].contiguous(), mk(cc, ca, 4), ts[:4], index=49, unconditional_guidance_scale=1.5,
                                   unconditional_conditioning=mk(cc, cu, 4))
    assert torch.equal(x64[:4], x4o) and torch.equal(p64[:4], p4o), "guided step rows 0-3 at B=64 differ from the B=4 run"
    r = rel_err(x64[:4], g["x_after_1"])
    print(f"B=64 L=64 bf16 guided step rows 0-3 vs reference golden x_after_1: rel err {r:.3e}")
    assert r < BF16_EPS_BAR, r


@pytest.mark.parametrize("L,B", [(64, 4), (32, 2)])
def test_groupnorm_in_operand_path_matches_separate_apply(L, B):
    """A guided DDIM step with GroupNorm + SiLU applied inside the consumer convolutions (ops.GN_FUSION, every site the
    kernel takes: 16x16 / 32x32 maps with >= 256 output channels) against the step with the separate apply kernel.
    With the 1x1 skip convolutions as separate launches both variants run the same K order and the step is BIT-EQUAL;
    with them folded into the 3x3 convolution's K loop the unfused launch walks K tap-major for the layers with many skip
    slabs, the fused one block-major — the same sum in another fp32 order (bound: 2e-3 of max |x|)."""
    from stedm_b200 import engine, ops
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    m = build_model(L, n_style=1, precision="bf16")
    unet = m._model.model.diffusion_model
    gen = torch.Generator().manual_seed(L + B)
    x = torch.randn(B, 3, L, L, generator=gen).cuda()
    cc = torch.randn(B, 3, L, L, generator=gen).cuda()
    c = {"c_concat": [cc], "c_crossattn": [torch.randn(B, 512, generator=gen).cuda()]}
    u = {"c_concat": [cc], "c_crossattn": [torch.randn(B, 512, generator=gen).cuda()]}
    ts = torch.full((B,), 481, dtype=torch.long, device="cuda")

    def run(flag):
        saved = ops.GN_FUSION[0]
        ops.GN_FUSION[0] = flag
        try:
            n0 = ops.LAUNCHES[0]
            s = DDIMSampler(m._model, use_cuda_graph=False)
            s.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
            out = s.p_sample_ddim(x, c, ts, index=24, unconditional_guidance_scale=1.5, unconditional_conditioning=u)
            return out, ops.LAUNCHES[0] - n0
        finally:
            ops.GN_FUSION[0] = saved

    saved_skip, saved_k = engine.FUSE_SKIP[0], ops.SPLIT_K[0]
    ops.SPLIT_K[0] = False        # fused launches never split K: compare against single-pass unfused launches
    try:
        engine.FUSE_SKIP[0] = False
        unet.invalidate_packed()
        (xf, pf), n_f = run(True)
        (xu, pu), n_u = run(False)
        print(f"L={L} B={B}: {n_u} launches with separate GroupNorm apply, {n_f} with the fused operand path")
        assert torch.equal(xf, xu) and torch.equal(pf, pu)
        assert n_f < n_u
    finally:
        engine.FUSE_SKIP[0] = saved_skip
        unet.invalidate_packed()
    (xf, _), _ = run(True)
    (xu, _), _ = run(False)
    ops.SPLIT_K[0] = saved_k
    r = rel_err(xf, xu)
    print(f"L={L} B={B}: fused vs separate GroupNorm with fused skip convolutions, rel err {r:.3e}")
    assert r < 2e-3


def test_cuda_graph_sampler_matches_eager():
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision="bf16")
    kw = dict(batch_size=2, ddim=True, ddim_steps=20, eta=0.0, log_every_t=1000, x_T=x_T.cuda(),
              unconditional_conditioning=_cond(g, "uc_crossattn"), unconditional_guidance_scale=1.5)
    m._model.use_cuda_graph = False
    z0, _ = m._model.sample_log(_cond(g, "c_crossattn"), **kw)
    m._model.use_cuda_graph = True
    m._model.model.diffusion_model.__dict__.pop("_graph_cache", None)
    try:
        z1, i1 = m._model.sample_log(_cond(g, "c_crossattn"), **kw)     # first call per signature: eager warm-up
        z2, i2 = m._model.sample_log(_cond(g, "c_crossattn"), **kw)     # captures the whole 20-step loop, replays it
        kw3 = dict(kw, x_T=(x_T * 0.5).cuda())
        z3, _ = m._model.sample_log(_cond(g, "c_crossattn"), **kw3)     # replay with another x_T
        cache = m._model.model.diffusion_model.__dict__["_graph_cache"]
        assert any(k[0] == "loop" for k in cache if isinstance(k, tuple) and k and k[0] == "loop")
    finally:
        m._model.use_cuda_graph = False
    z3e, _ = m._model.sample_log(_cond(g, "c_crossattn"), **kw3)
    assert torch.equal(z0, z1) and torch.equal(z0, z2) and torch.equal(z3, z3e)
    assert len(i2["x_inter"]) == len(i1["x_inter"]) and all(torch.equal(a, b) for a, b in zip(i1["pred_x0"], i2["pred_x0"]))
    # logged intermediates come out of the captured loop too (ddim.py:155-157: index % log_every_t == 0 or the first step)
    kw5 = dict(kw, log_every_t=5)
    m._model.use_cuda_graph = False
    _, ie = m._model.sample_log(_cond(g, "c_crossattn"), **kw5)
    m._model.use_cuda_graph = True
    try:
        m._model.sample_log(_cond(g, "c_crossattn"), **kw5)
        _, ig = m._model.sample_log(_cond(g, "c_crossattn"), **kw5)
    finally:
        m._model.use_cuda_graph = False
    assert len(ig["x_inter"]) == len(ie["x_inter"]) == 6 and len(ig["pred_x0"]) == 6
    assert all(torch.equal(a, b) for a, b in zip(ie["x_inter"], ig["x_inter"]))
    assert all(torch.equal(a, b) for a, b in zip(ie["pred_x0"], ig["pred_x0"]))
    # per-step path (a caller-driven loop, or any option that needs host-side work per step) still uses the per-pass graph
    m._model.use_cuda_graph = True
    try:
        z4, _ = m._model.sample_log(_cond(g, "c_crossattn"), **dict(kw, img_callback=lambda p, i: None))
    finally:
        m._model.use_cuda_graph = False
    assert torch.equal(z0, z4)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_shared_trunk_is_bit_identical(precision):
    """Guided step with the encoder trunk evaluated once for (cond, uncond) == the batched 2B pass, bit for bit; with
    the decoder's shared skip channels convolved once (engine.SPLIT_CONCAT) only the fp32 summation order differs."""
    from stedm_b200 import engine
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision=precision)
    cond, unc = _cond(g, "c_crossattn"), _cond(g, "uc_crossattn")
    ts = torch.full((2,), 481, device="cuda", dtype=torch.long)

    def step(share):
        s = DDIMSampler(m._model, share_trunk=share)
        s.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
        return s.p_sample_ddim(x_T.cuda(), cond, ts, index=24, unconditional_guidance_scale=1.5,
                               unconditional_conditioning=unc)

    saved, saved_style = engine.SPLIT_CONCAT[0], engine.SHARE_STYLE_CONV[0]
    from stedm_b200 import ops as _ops
    saved_k, _ops.SPLIT_K[0] = _ops.SPLIT_K[0], False      # the trunk runs at B, the two-pass reference at 2B
    try:
        engine.SPLIT_CONCAT[0] = engine.SHARE_STYLE_CONV[0] = False
        two_pass, shared = step(False), step(True)
        assert torch.equal(two_pass[0], shared[0]) and torch.equal(two_pass[1], shared[1])
        engine.SPLIT_CONCAT[0] = True
        split = step(True)
        # ResBlockStyle's first convolution once per distinct input (conv + bias in fp32, the style embedding added by
        # rows_add_emb): (acc + bias) + emb instead of acc + (bias + emb) — one fp32 rounding apart
        engine.SPLIT_CONCAT[0], engine.SHARE_STYLE_CONV[0] = False, True
        style = step(True)
    finally:
        engine.SPLIT_CONCAT[0], engine.SHARE_STYLE_CONV[0] = saved, saved_style
        _ops.SPLIT_K[0] = saved_k
    r = rel_err(split[0], two_pass[0])
    print(f"{precision} split-concat guided step vs unsplit rel err {r:.3e}")
    assert r < (1e-5 if precision == "fp32" else 5e-3)
    r2 = rel_err(style[0], two_pass[0])
    print(f"{precision} shared style-block convolution vs two passes rel err {r2:.3e}")
    assert r2 < (1e-5 if precision == "fp32" else 5e-3)


def test_shared_style_block_convolution_at_latent64():
    """Guided step at latent 64 (16 x 16 middle maps: whole 128-pixel tiles per sample) with ResBlockStyle's first
    convolution evaluated once per distinct input (engine.SHARE_STYLE_CONV: conv + bias in fp32, rows_add_emb adds the
    cond / uncond style embeddings) against the step that convolves both halves: one fp32 rounding apart before the bf16
    store, and both inside the bf16 bar of the reference's golden x_after_1 (tests/golden/c1_b4_l64.npz)."""
    from stedm_b200 import engine
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    g = load_golden("c1_b4_l64")
    _, _, x4 = O.synthetic_batch(4, 256, 1, 1)
    m = build_model(64, n_style=1, precision="bf16")
    cond = {"c_concat": [torch.from_numpy(g["c_concat"]).cuda()], "c_crossattn": [torch.from_numpy(g["c_crossattn"]).cuda()]}
    unc = {"c_concat": [torch.from_numpy(g["c_concat"]).cuda()], "c_crossattn": [torch.from_numpy(g["uc_crossattn"]).cuda()]}
    ts = torch.full((4,), 981, dtype=torch.long, device="cuda")

    def step(flag):
        saved = engine.SHARE_STYLE_CONV[0]
        engine.SHARE_STYLE_CONV[0] = flag
        try:
            s = DDIMSampler(m._model, use_cuda_graph=False)
            s.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
            return s.p_sample_ddim(x4.cuda(), cond, ts, index=49, unconditional_guidance_scale=1.5, unconditional_conditioning=unc)
        finally:
            engine.SHARE_STYLE_CONV[0] = saved

    (xa, _), (xb, _) = step(True), step(False)
    r = rel_err(xa, xb)
    print(f"shared style-block convolution vs both halves convolved, guided step at latent 64: rel err {r:.3e}")
    assert r < 5e-3
    assert rel_err(xa, g["x_after_1"]) < BF16_EPS_BAR and rel_err(xb, g["x_after_1"]) < BF16_EPS_BAR


def test_split_groupnorm_of_shared_skips_is_bit_identical():
    """engine.SPLIT_GN: decoder ResBlocks normalise the skip-only channels of [h | skip] once per distinct skip sample
    (ops.gn_apply_split) and feed the convolution two sources — same values, same K order: the guided eps must not move
    by a bit (latent 64: all nine concat sites, with and without the split convolution)."""
    from stedm_b200 import engine
    g = load_golden("c1_b4_l64")
    _, _, x4 = O.synthetic_batch(4, 256, 1, 1)
    m = build_model(64, n_style=1, precision="bf16")
    unet = m._model.model.diffusion_model
    cc = torch.from_numpy(g["c_concat"]).cuda()
    ctx = torch.cat([torch.from_numpy(g["uc_crossattn"]), torch.from_numpy(g["c_crossattn"])]).cuda()
    ts = torch.full((4,), 481, dtype=torch.long, device="cuda")
    outs = {}
    for flag in (True, False):
        saved = engine.SPLIT_GN[0]
        engine.SPLIT_GN[0] = flag
        try:
            outs[flag] = unet.forward_split(x4.cuda(), cc, ts, ctx).clone()
        finally:
            engine.SPLIT_GN[0] = saved
    assert outs[True].shape[0] == 8 and torch.equal(outs[True], outs[False])


def test_latent128_eps_and_decode_vs_oracle():
    """BASELINE configs[3] geometry (512^2 image, latent 128: full-row 128-pixel tiles, 16 384-token decoder
    attention) against the CPU oracle run live on the same fixture weights: batch 1, and the same sample as row 0 of a
    batch of 8 (bit-equal to the batch-1 run: shard invariance at this geometry too)."""
    m = build_model(128, n_style=1, precision="bf16")
    model = m._model
    sd = oracle_state_dict(model)
    seg, style, x_T = O.synthetic_batch(1, 512, 1, seed=5)
    g = torch.Generator().manual_seed(17)
    cond = {"c_concat": [torch.randn(1, 3, 128, 128, generator=g)], "c_crossattn": [torch.randn(1, 512, generator=g)]}
    t = torch.full((1,), 481, dtype=torch.long)
    with torch.no_grad():
        want = O.apply_model(sd, x_T, t, cond)
    with _no_split_k():
        got = model.apply_model(x_T.cuda(), t.cuda(), {k: [v[0].cuda()] for k, v in cond.items()})
        r = rel_err(got, want)
        print(f"latent-128 bf16 eps rel err {r:.3e}")
        assert r < BF16_EPS_BAR
        B = 8
        x8 = torch.cat([x_T, torch.randn(B - 1, 3, 128, 128, generator=g)], 0).cuda()
        c8 = {"c_concat": [torch.cat([cond["c_concat"][0], torch.randn(B - 1, 3, 128, 128, generator=g)], 0).cuda()],
              "c_crossattn": [torch.cat([cond["c_crossattn"][0], torch.randn(B - 1, 512, generator=g)], 0).cuda()]}
        got8 = model.apply_model(x8, torch.full((B,), 481, dtype=torch.long, device="cuda"), c8)
    assert torch.equal(got8[:1], got), "row 0 of the B=8 latent-128 pass differs from the B=1 run"
    z = x_T * 60
    with torch.no_grad():
        want_img = O.decode_first_stage(sd, z, force_not_quantize=True)
    img = model.decode_first_stage(z.cuda(), force_not_quantize=True)
    p = psnr(img.clamp(-1, 1), want_img.clamp(-1, 1))
    print(f"latent-128 bf16 decode PSNR {p:.1f} dB")
    assert tuple(img.shape) == (1, 3, 512, 512) and p >= PSNR_BAR
    z2 = torch.cat([z, torch.randn(1, 3, 128, 128, generator=g) * 60], 0).cuda()
    img2 = model.decode_first_stage(z2, force_not_quantize=True)
    assert torch.equal(img2[:1], img), "row 0 of the B=2 latent-128 decode differs from the B=1 run"


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_multi_style_aggregation_her2_shape(precision):
    """BASELINE configs[2]: N = 10 style patches per sample aggregated by Agg_Mean (conf/style_sampling/mp.yaml); the
    style encoder runs natively (stedm_b200.style_engine) and is compared with the oracle's torchvision swin_v2_t."""
    m = build_model(32, n_style=10, precision=precision)
    model = m._model
    sd = oracle_state_dict(model)
    seg, style, _ = O.synthetic_batch(2, 128, 10, seed=3)
    batch = {"image": torch.zeros(2, 128, 128, 3).cuda(), "segmentation": seg.cuda(), "style_imgs": style.cuda()}
    _, c = model.get_input(batch, "image")
    with torch.no_grad():
        want = O.get_conditioning(sd, seg, style)
    assert tuple(c["c_crossattn"][0].shape) == (2, 512)
    err = max_abs(c["c_crossattn"][0], want["c_crossattn"][0])
    if precision == "fp32":
        assert err < FP32_EPS_BAR, err
    else:
        rel = err / float(want["c_crossattn"][0].abs().max())
        print(f"bf16 style feature (N=10) rel err {rel:.3e}")
        assert rel < BF16_EPS_BAR, rel
    assert max_abs(c["c_concat"][0], want["c_concat"][0]) < 1e-6


def test_ddpm_ancestral_sampler_matches_reference_golden(monkeypatch):
    """Kept API: LatentDiffusion.sample -> p_sample_loop -> p_sample -> p_mean_variance
    (ddpm.py:1050-1110, 1168-1235), 5 ancestral steps with the per-step noise the reference run drew
    (tests/golden/ancestral.npz, oracle/make_golden.py --only ancestral), plain and with quantize_denoised."""
    import stedm_b200.ldm.models.diffusion.ddpm as native_ddpm
    g, _, _, x_T = _small_inputs()
    gold = load_golden("ancestral")
    T = int(gold["timesteps"])
    m = build_model(32, n_style=2, precision="fp32")
    model = m._model
    cond = _cond(g, "c_crossattn")
    for name, scale, quant in (("ancestral_t5", 1.0, False), ("ancestral_t5_quant", 40.0, True)):
        it = iter(torch.from_numpy(gold["noises"]).cuda())
        monkeypatch.setattr(native_ddpm, "noise_like", lambda shape, device, repeat=False: next(it).clone())
        z = model.sample(cond, batch_size=2, x_T=x_T.cuda() * scale, timesteps=T, verbose=False, quantize_denoised=quant)
        want = torch.from_numpy(gold[name])
        if not quant:
            assert rel_err(z, want) < 1e-4, rel_err(z, want)
        else:
            # a pixel whose two nearest codes are ~equidistant may snap differently: rare, and large when it happens
            bad = ((z.cpu() - want).abs() > 1e-3 * float(want.abs().max())).float().mean()
            print(f"ancestral quantize_denoised: fraction of differing values {float(bad):.2e}")
            assert float(bad) < 0.01
    # bf16 mode runs the same loop inside the bf16 eps bar
    monkeypatch.undo()
    it = iter(torch.from_numpy(gold["noises"]).cuda())
    monkeypatch.setattr(native_ddpm, "noise_like", lambda shape, device, repeat=False: next(it).clone())
    mb = build_model(32, n_style=2, precision="bf16")
    zb = mb._model.sample(cond, batch_size=2, x_T=x_T.cuda(), timesteps=T, verbose=False)
    r = rel_err(zb, gold["ancestral_t5"])
    print(f"bf16 ancestral 5-step rel err {r:.3e}")
    assert r < BF16_EPS_BAR
    # p_sample at t = 0 adds no noise (ddpm.py:1103)
    monkeypatch.undo()
    t0 = torch.zeros(2, dtype=torch.long, device="cuda")
    a = mb._model.p_sample(x_T.cuda(), cond, t0)
    b = mb._model.p_sample(x_T.cuda(), cond, t0)
    assert torch.equal(a, b)


@pytest.mark.parametrize("B,L", [(3, 32), (1, 64)])
def test_ragged_batches_guided_step_vs_oracle(B, L):
    """Odd batch at latent 32 (the shared-trunk broadcast is not tile aligned -> two-pass path, partial last tile)
    and batch 1 at latent 64: one guided DDIM step against the CPU oracle."""
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    m = build_model(L, n_style=1, precision="bf16")
    model = m._model
    sd = oracle_state_dict(model)
    g = torch.Generator().manual_seed(100 + B)
    x_T = torch.randn(B, 3, L, L, generator=g)
    cc = torch.randn(B, 3, L, L, generator=g)
    cond = {"c_concat": [cc], "c_crossattn": [torch.randn(B, 512, generator=g)]}
    unc = {"c_concat": [cc], "c_crossattn": [torch.randn(B, 512, generator=g)]}
    with torch.no_grad():
        want, _ = O.ddim_sample(sd, cond, unc, x_T, S=50, cfg_scale=1.5, max_steps=1)
    s = DDIMSampler(model)
    s.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
    ts = torch.full((B,), 981, device="cuda", dtype=torch.long)
    cu = lambda c: {k: [v[0].cuda()] for k, v in c.items()}
    got, _ = s.p_sample_ddim(x_T.cuda(), cu(cond), ts, index=49, unconditional_guidance_scale=1.5,
                             unconditional_conditioning=cu(unc))
    r = rel_err(got, want)
    print(f"B={B} L={L} guided step rel err {r:.3e}")
    assert r < BF16_EPS_BAR


def test_predict_entry_point_writes_pngs(tmp_path):
    """`python -m stedm_b200.predict` (the predict_diff.py counterpart): synthetic dataset, hydra-style overrides,
    PNGs keyed by dataset index."""
    import subprocess
    import sys
    from tests.util import ROOT
    out = str(tmp_path / "pred")
    cmd = [sys.executable, "-m", "stedm_b200.predict", "style_agg=mean", "style_sampling=augmented", "ddim_steps=4",
           "diffusion.image_size=32", "data.patch_size=128", "+synthetic=3", f"+predict_dir={out}"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    from PIL import Image
    for i in range(3):
        img = Image.open(f"{out}/img_{i:05d}.png")
        assert img.size == (128, 128) and img.mode == "RGB"
        assert Image.open(f"{out}/seg_{i:05d}.png").size == (128, 128)


@pytest.mark.parametrize("precision,z_bar,psnr_bar", [("fp32", 1e-3, 40.0), ("bf16", 5e-2, 40.0)])
def test_config0_flowers_b4_256_full_path_vs_reference_golden(precision, z_bar, psnr_bar):
    """BASELINE configs[0] (flowers proof of concept: batch 4, 256^2, DDIM-50, cfg 1.5) end to end against the
    golden produced by the reference's own code on CPU (tests/golden/c1_b4_l64.npz)."""
    g = load_golden("c1_b4_l64")
    seg, style, x_T = O.synthetic_batch(4, 256, 1, 1)
    m = build_model(64, n_style=1, precision=precision)
    cond = {"c_concat": [torch.from_numpy(g["c_concat"]).cuda()], "c_crossattn": [torch.from_numpy(g["c_crossattn"]).cuda()]}
    unc = {"c_concat": [torch.from_numpy(g["c_concat"]).cuda()], "c_crossattn": [torch.from_numpy(g["uc_crossattn"]).cuda()]}
    tt = torch.full((4,), 481, dtype=torch.long, device="cuda")
    e = m._model.apply_model(x_T.cuda(), tt, cond)
    r_e = rel_err(e, g["eps_c_481"])
    assert (max_abs(e, g["eps_c_481"]) < FP32_EPS_BAR) if precision == "fp32" else (r_e < BF16_EPS_BAR)
    z, _ = m._model.sample_log(cond, batch_size=4, ddim=True, ddim_steps=50, eta=0.0, log_every_t=1000, x_T=x_T.cuda(),
                               unconditional_conditioning=unc, unconditional_guidance_scale=1.5)
    r_z = rel_err(z, g["z_final"])
    img = m._model.decode_first_stage(torch.from_numpy(g["z_final"]).cuda())
    p_dec = psnr(img.clamp(-1, 1), np.clip(g["dec_quant"].astype(np.float32), -1, 1))
    img_e2e = m._model.decode_first_stage(z)
    p_e2e = psnr(img_e2e.clamp(-1, 1), np.clip(g["dec_quant"].astype(np.float32), -1, 1))
    print(f"config0 {precision}: eps rel {r_e:.2e}  DDIM-50 latent rel {r_z:.2e}  decode PSNR {p_dec:.1f} dB  "
          f"end-to-end PSNR {p_e2e:.1f} dB")
    assert r_z < z_bar and p_dec >= PSNR_BAR and p_e2e >= psnr_bar


def test_ddim_eta_mask_original_steps_and_quantize_branches():
    """The optional branches of ddim.py around the same U-Net call: eta > 0 (sigma*noise, :206-209),
    mask / x0 in-painting (:143-146), use_original_steps (:188-191) and quantize_denoised (:200-201), each against the
    oracle's restatement of the step tail fed with the engine's own eps."""
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision="fp32")
    model = m._model
    cond, unc = _cond(g, "c_crossattn"), _cond(g, "uc_crossattn")
    x = x_T.cuda()
    s = DDIMSampler(model)
    s.make_schedule(ddim_num_steps=50, ddim_eta=0.7, verbose=False)
    ts = torch.full((2,), int(s.ddim_timesteps[30]), device="cuda", dtype=torch.long)
    e = O.cfg_combine(model.apply_model(x, ts, cond).cpu(), model.apply_model(x, ts, unc).cpu(), 1.5)
    # eta > 0: same RNG stream as the reference (one randn per step)
    torch.manual_seed(5)
    xp, p0 = s.p_sample_ddim(x, cond, ts, index=30, unconditional_guidance_scale=1.5, unconditional_conditioning=unc)
    torch.manual_seed(5)
    noise = torch.randn(x.shape, device="cuda").cpu()
    sig = float(s.ddim_sigmas[30])
    assert sig > 0
    wx, wp = O.ddim_update(x_T, e, s.ddim_alphas[30], s.ddim_alphas_prev[30], sig, s.ddim_sqrt_one_minus_alphas[30], noise)
    assert max_abs(xp, wx) < 2e-4 and max_abs(p0, wp) < 2e-4
    # original 1000-step tables: index = t
    t_idx = 481
    ts2 = torch.full((2,), t_idx, device="cuda", dtype=torch.long)
    e2 = O.cfg_combine(model.apply_model(x, ts2, cond).cpu(), model.apply_model(x, ts2, unc).cpu(), 1.5)
    torch.manual_seed(6)
    xp, p0 = s.p_sample_ddim(x, cond, ts2, index=t_idx, use_original_steps=True, unconditional_guidance_scale=1.5,
                             unconditional_conditioning=unc)
    torch.manual_seed(6)
    noise = torch.randn(x.shape, device="cuda").cpu()
    wx, wp = O.ddim_update(x_T, e2, model.alphas_cumprod[t_idx].cpu(), model.alphas_cumprod_prev[t_idx].cpu(),
                           s.ddim_sigmas_for_original_num_steps[t_idx].cpu(),
                           model.sqrt_one_minus_alphas_cumprod[t_idx].cpu(), noise)
    assert max_abs(xp, wx) < 2e-4 and max_abs(p0, wp) < 2e-4
    # quantize_denoised: pred_x0 snapped to the codebook, x_prev rebuilt from it
    s0 = DDIMSampler(model)
    s0.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
    xq, pq = s0.p_sample_ddim(x * 40, cond, ts, index=30, quantize_denoised=True, unconditional_guidance_scale=1.5,
                              unconditional_conditioning=unc)
    eq = O.cfg_combine(model.apply_model(x * 40, ts, cond).cpu(), model.apply_model(x * 40, ts, unc).cpu(), 1.5)
    _, wp = O.ddim_update(x_T * 40, eq, s0.ddim_alphas[30], s0.ddim_alphas_prev[30], 0.0, s0.ddim_sqrt_one_minus_alphas[30])
    cb = model.first_stage_model.quantize.embedding.weight.detach().cpu().float()
    wq, _ = O.vq_quantize(wp, cb)
    a_prev = torch.tensor(float(s0.ddim_alphas_prev[30]))
    wxq = a_prev.sqrt() * wq + (1.0 - a_prev).sqrt() * eq
    assert float((pq.cpu() != wq).float().mean()) < 0.01          # a code flips only when two codes are ~equidistant
    assert max_abs(xq.cpu() - a_prev.sqrt() * pq.cpu(), wxq - a_prev.sqrt() * wq) < 2e-3
    # in-painting: masked region follows q_sample(x0, t), the sampler runs and returns the right shapes
    mask = torch.zeros(2, 1, 32, 32, device="cuda")
    mask[..., :16] = 1
    z, inter = s0.sample(4, 2, (3, 32, 32), conditioning=cond, verbose=False, x_T=x, mask=mask, x0=torch.zeros_like(x),
                         eta=0.0, unconditional_guidance_scale=1.5, unconditional_conditioning=unc)
    assert tuple(z.shape) == (2, 3, 32, 32) and bool(torch.isfinite(z).all())
    z2, _ = s0.ddim_sampling(cond, (2, 3, 32, 32), x_T=x, ddim_use_original_steps=True, timesteps=3,
                             unconditional_guidance_scale=1.5, unconditional_conditioning=unc)
    assert bool(torch.isfinite(z2).all())


def test_plms_sampler_matches_oracle():
    """PLMSSampler (plms.py) on the native U-Net: first five steps (improved Euler, then Adams-Bashforth orders 2-4)
    against oracle.plms_sample — itself bit-equal to the reference's sampler class — with the oracle U-Net as eps."""
    from stedm_b200.ldm.models.diffusion.plms import PLMSSampler
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision="fp32")
    model = m._model
    sd = oracle_state_dict(model)
    cond, unc = _cond(g, "c_crossattn"), _cond(g, "uc_crossattn")
    s = PLMSSampler(model)
    s.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
    tr = np.flip(s.ddim_timesteps)
    img, old = x_T.cuda(), []
    n_steps = 5
    for i in range(n_steps):
        ts = torch.full((2,), int(tr[i]), device="cuda", dtype=torch.long)
        tn = torch.full((2,), int(tr[i + 1]), device="cuda", dtype=torch.long)
        img, p0, e_t = s.p_sample_plms(img, cond, ts, index=50 - i - 1, unconditional_guidance_scale=1.5,
                                       unconditional_conditioning=unc, old_eps=old, t_next=tn)
        old.append(e_t)
        if len(old) >= 4:
            old.pop(0)
    oc, ou = _cond(g, "c_crossattn", "cpu"), _cond(g, "uc_crossattn", "cpu")
    with torch.no_grad():
        want = O.plms_sample(lambda x, t: O.apply_model(sd, x, t, oc), x_T, S=50, cfg_scale=1.5,
                             uncond_eps_fn=lambda x, t: O.apply_model(sd, x, t, ou), max_steps=n_steps)
    assert max_abs(img, want) < 5e-4, max_abs(img, want)
    z, inter = s.sample(50, 2, (3, 32, 32), conditioning=cond, verbose=False, x_T=x_T.cuda(), timesteps=None,
                        unconditional_guidance_scale=1.5, unconditional_conditioning=unc, log_every_t=10)
    assert tuple(z.shape) == (2, 3, 32, 32) and bool(torch.isfinite(z).all()) and len(inter["x_inter"]) == 7
    with pytest.raises(ValueError):
        s.make_schedule(ddim_num_steps=50, ddim_eta=0.5, verbose=False)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_with_spatial_transformer_matches_reference(precision):
    """use_spatial_transformer=True, context_dim=1024: the reference's UNetModel with a SpatialTransformer in
    middle_block[2] (golden from the reference itself, oracle/make_golden.py --only st) vs the native engine."""
    from tests.test_engine_glue import build_st_unet, st_inputs
    g = load_golden("spatial_transformer")
    x, ctx, _, _, _ = st_inputs()
    unet = build_st_unet().model.diffusion_model.cuda()
    unet.set_precision(precision)
    for t in (981, 1):
        eps = unet(x.cuda(), torch.full((2,), t, device="cuda", dtype=torch.long), ctx.cuda())
        want = g[f"unet_eps_{t}"]
        if precision == "fp32":
            assert max_abs(eps, want) < FP32_EPS_BAR, max_abs(eps, want)
        else:
            r = rel_err(eps, want)
            print(f"SpatialTransformer U-Net bf16 eps rel err t={t}: {r:.3e}")
            assert r < BF16_EPS_BAR, r


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_spatial_transformer_cross_attention_over_style_tokens(precision):
    """Stand-alone SpatialTransformer.forward(x, context) with (B, N, 512) style tokens, N = 10 and 1: cross-attention on
    the tcgen05 flash kernel (separate key/value length) vs the reference class's output."""
    from tests.test_engine_glue import build_st_module, st_inputs
    g = load_golden("spatial_transformer")
    _, _, xs, c10, c1 = st_inputs()
    st = build_st_module().st.cuda()
    st.set_precision(precision)
    for c, name in ((c10, "st_ctx10"), (c1, "st_ctx1")):
        out = st(xs.cuda(), c.cuda())
        if precision == "fp32":
            assert max_abs(out, g[name]) < 2e-4, max_abs(out, g[name])
        else:
            r = rel_err(out, g[name])
            print(f"SpatialTransformer cross-attention bf16 rel err {name}: {r:.3e}")
            assert r < BF16_EPS_BAR, r


def test_odd_latent_size_falls_back_to_cuda_cores_and_matches_oracle():
    """A 96 x 96 image (latent 24; maps 24, 12, 6 in the U-Net, 24, 48, 96 in the decoder, 24/12/6/3 in the style encoder)
    cannot be tiled by the tcgen05 kernel: in bf16 mode those layers run on the CUDA-core implicit-GEMM kernel (no
    fused skip / statistics), everything else (GroupNorm, attention, step kernel, VQ) is unchanged.  The path must stay
    functional for every size the reference accepts and keep the bf16 bars."""
    L, B = 24, 1
    m = build_model(L, n_style=1, precision="bf16")
    model = m._model
    sd = oracle_state_dict(model)
    seg, style, x_T = O.synthetic_batch(B, 4 * L, 1, seed=9)
    batch = {"image": torch.zeros(B, 4 * L, 4 * L, 3).cuda(), "segmentation": seg.cuda(), "style_imgs": style.cuda()}
    _, c = model.get_input(batch, "image")
    with torch.no_grad():
        want_c = O.get_conditioning(sd, seg, style)
    assert rel_err(c["c_crossattn"][0], want_c["c_crossattn"][0]) < BF16_EPS_BAR
    t = torch.full((B,), 481, dtype=torch.long)
    eps = model.apply_model(x_T.cuda(), t.cuda(), c)
    with torch.no_grad():
        want = O.apply_model(sd, x_T, t, want_c)
    r = rel_err(eps, want)
    print(f"latent-24 (CUDA-core fallback) bf16 eps rel err {r:.3e}")
    assert r < BF16_EPS_BAR, r
    img = model.decode_first_stage(x_T.cuda() * 40)
    with torch.no_grad():
        want_img = O.decode_first_stage(sd, x_T * 40)
    p = psnr(img.clamp(-1, 1), want_img.clamp(-1, 1))
    print(f"latent-24 bf16 decode PSNR {p:.1f} dB")
    assert tuple(img.shape) == (B, 3, 96, 96) and p >= PSNR_BAR


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_dpm_solver_sampler_matches_reference(precision):
    """DPMSolverSampler (dpm_solver/sampler.py) on the native U-Net at fractional model times vs the reference's own
    sampler run on the reference model (tests/golden/dpm_solver.npz, S = 12 and 16)."""
    from stedm_b200.ldm.models.diffusion.dpm_solver import DPMSolverSampler
    g, _, _, x_T = _small_inputs()
    m = build_model(32, n_style=2, precision=precision)
    cond, unc = _cond(g, "c_crossattn"), _cond(g, "uc_crossattn")
    gold = load_golden("dpm_solver")
    for S in (12, 16):
        z, _ = DPMSolverSampler(m._model).sample(S, 2, (3, 32, 32), conditioning=cond, verbose=False, x_T=x_T.cuda(),
                                                 unconditional_guidance_scale=1.5, unconditional_conditioning=unc)
        r = rel_err(z, gold[f"dpm_s{S}"])
        print(f"DPM-Solver S={S} {precision} final latent rel err {r:.3e}")
        assert r < (1e-4 if precision == "fp32" else BF16_EPS_BAR), r
