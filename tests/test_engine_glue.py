"""Host-side engine logic on CPU: the product's UNetRunner / DecoderRunner driven through a torch stand-in for the
kernel layer (tests/fake_ops.py) must reproduce the reference's golden eps and decoded image.  This pins weight
repacking, channel padding, concat order, embedding-table offsets and block order without needing a GPU; the
kernels themselves are checked by the -m gpu tests."""
import pytest
import torch

from tests import fake_ops
from tests.util import build_config, load_golden, max_abs


@pytest.fixture(scope="module")
def cpu_model():
    from stedm_b200.modules.ldm_diffusion import LDM_Diffusion
    from stedm_b200.utils.fixture import apply_fixture_weights
    m = LDM_Diffusion(build_config(32, n_style=2), load_first_stage_ckpt=False)
    apply_fixture_weights(m._model, seed=0)
    return m.eval()


@pytest.fixture()
def patched(monkeypatch):
    from stedm_b200 import engine
    monkeypatch.setattr(engine, "ops", fake_ops)
    return engine


@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_unet_runner_glue(cpu_model, patched, precision, bar):
    from oracle import stedm_oracle as O
    g = load_golden("small_b2_l32")
    _, _, x_T = O.synthetic_batch(2, 128, 2, 0)
    runner = patched.UNetRunner(cpu_model._model.model.diffusion_model, precision)
    t = torch.full((2,), 481, dtype=torch.long)
    with torch.no_grad():
        eps = runner(x_T, torch.from_numpy(g["c_concat"]), t, torch.from_numpy(g["c_crossattn"]))
        # batched guidance: (cond | uncond) in one pass must equal the two separate passes
        x2, t2 = torch.cat([x_T, x_T]), torch.cat([t, t])
        cc = torch.from_numpy(g["c_concat"])
        ctx = torch.cat([torch.from_numpy(g["c_crossattn"]), torch.from_numpy(g["uc_crossattn"])])
        eps2 = runner(x2, torch.cat([cc, cc]), t2, ctx)
        # shared encoder trunk: B inputs, 2B contexts -> the same 2B eps
        eps3 = runner(x_T, cc, t, ctx)
        # one timestep for the whole batch: embeddings computed for one row and broadcast
        assert max_abs(runner(x_T, cc, t, ctx, uniform_t=True), eps3) < (5e-5 if precision == "fp32" else 5e-2)
    # (the CPU stand-in's conv picks batch-dependent algorithms, so bf16 emulation is not bit-stable here; the
    #  bit-exactness of the real kernels is asserted on the GPU in tests/test_gpu_model.py)
    assert max_abs(eps3, eps2) < (1e-5 if precision == "fp32" else 5e-2)
    scale = float(torch.from_numpy(g["eps_c_481"]).abs().max())
    assert max_abs(eps, g["eps_c_481"]) < bar * (scale if precision == "bf16" else 1.0)
    assert max_abs(eps2[:2], g["eps_c_481"]) < bar * (scale if precision == "bf16" else 1.0)
    assert max_abs(eps2[2:], g["eps_u_481"]) < bar * (scale if precision == "bf16" else 1.0)


def test_decoder_runner_glue(cpu_model, patched):
    g = load_golden("small_b2_l32")
    runner = patched.DecoderRunner(cpu_model._model.first_stage_model, "fp32")
    z = torch.from_numpy(g["z_final"])
    with torch.no_grad():
        assert max_abs(runner(z), g["dec_quant"]) < 1e-3
        assert max_abs(runner(z, force_not_quantize=True), g["dec_noquant"]) < 1e-3


# ---------------------------------------------------------------------------------------------- SpatialTransformer
ST_UNET_KW = dict(image_size=32, in_channels=6, out_channels=3, model_channels=128, attention_resolutions=[32, 16, 8],
                  num_res_blocks=2, channel_mult=[1, 4, 8], num_heads=8, use_spatial_transformer=True, context_dim=1024)


def st_inputs():
    g = torch.Generator().manual_seed(31)                 # the draw order of oracle/make_golden.py gen_spatial_transformer
    x = torch.randn(2, 6, 32, 32, generator=g)
    ctx = torch.randn(2, 512, generator=g) * 0.5
    xs = torch.randn(2, 256, 16, 16, generator=g)
    c10 = torch.randn(2, 10, 512, generator=g)
    c1 = torch.randn(2, 1, 512, generator=g)
    return x, ctx, xs, c10, c1


def build_st_unet():
    from stedm_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    from stedm_b200.utils.fixture import apply_fixture_weights
    holder = torch.nn.Module()
    holder.model = torch.nn.Module()
    holder.model.diffusion_model = UNetModel(**ST_UNET_KW).eval()
    apply_fixture_weights(holder, seed=0)
    return holder


def build_st_module():
    from stedm_b200.ldm.modules.attention import SpatialTransformer
    from stedm_b200.utils.fixture import apply_fixture_weights
    h = torch.nn.Module()
    h.st = SpatialTransformer(256, 4, 64, depth=2, context_dim=512).eval()
    apply_fixture_weights(h, seed=0)
    return h


@pytest.fixture()
def patched_all(monkeypatch):
    from stedm_b200 import engine, style_engine
    monkeypatch.setattr(engine, "ops", fake_ops)
    monkeypatch.setattr(style_engine, "ops", fake_ops)
    return engine


def test_spatial_transformer_oracle_matches_reference_golden():
    """oracle.unet_forward / spatial_transformer on fixture weights == the reference's own classes
    (tests/golden/spatial_transformer.npz from oracle/make_golden.py --only st)."""
    from oracle import stedm_oracle as O
    g = load_golden("spatial_transformer")
    x, ctx, xs, c10, c1 = st_inputs()
    sd = {k: v.detach().float() for k, v in build_st_unet().state_dict().items()}
    with torch.no_grad():
        eps = O.unet_forward(sd, x, torch.full((2,), 981, dtype=torch.long), ctx)
    assert max_abs(eps, g["unet_eps_981"]) < 1e-5
    sd2 = {k: v.detach().float() for k, v in build_st_module().state_dict().items()}
    with torch.no_grad():
        assert max_abs(O.spatial_transformer(xs, sd2, "st.", 4, c10), g["st_ctx10"]) < 1e-5


@pytest.mark.parametrize("precision,bar", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_spatial_transformer_glue(patched_all, precision, bar):
    """Host logic of PackedSpatialTransformer (fused qkv weights, head split, cross-attention k/v packing, GEGLU halves,
    residual order) and its wiring into the U-Net's middle block, through the torch stand-in for the kernels."""
    g = load_golden("spatial_transformer")
    x, ctx, xs, c10, c1 = st_inputs()
    unet = build_st_unet().model.diffusion_model
    runner = patched_all.UNetRunner(unet, precision)
    with torch.no_grad():
        eps = runner(x[:, :3].contiguous(), x[:, 3:].contiguous(), torch.full((2,), 981, dtype=torch.long), ctx)
    scale = float(abs(g["unet_eps_981"]).max()) if precision == "bf16" else 1.0
    assert max_abs(eps, g["unet_eps_981"]) < bar * scale, max_abs(eps, g["unet_eps_981"])
    st = build_st_module().st
    r = patched_all.PackedSpatialTransformer(st, patched_all.Precision(precision))
    pool = patched_all.StatsPool(1, 2, "cpu")
    for c, name in ((c10, "st_ctx10"), (c1, "st_ctx1")):
        with torch.no_grad():
            h = xs.permute(0, 2, 3, 1).contiguous().to(r.prec.act)
            out = r(h, pool, c).float().permute(0, 3, 1, 2)
        scale = float(abs(g[name]).max()) if precision == "bf16" else 1.0
        assert max_abs(out, g[name]) < bar * scale, (name, max_abs(out, g[name]))


# ------------------------------------------------------------------------------ split of the decoder's concat conv
@pytest.mark.parametrize("c0,c1,want", [(1024, 1024, 1024),   # 64 ch / group: no group straddles -> split at c0
                                        (1024, 512, 1088),    # 48 ch / group: group 1008..1055 straddles -> next slab
                                        (512, 512, 512),
                                        (512, 128, 0),        # 20 ch / group: only 64 shared channels -> not worth it
                                        (128, 128, 0)])       # 128 shared channels at 64^2: HBM cost > FLOPs saved
def test_split_point_of_concat_conv(patched, c0, c1, want):
    """PackedResBlock._split_point: the channels of the normalised [h | skip] concat above the split must belong to
    GroupNorm groups that lie wholly inside the skip half (their values are then shared by cond / uncond)."""
    blk = patched.PackedResBlock.__new__(patched.PackedResBlock)
    blk.c1 = type("C", (), {"tc": True, "tc_ok": staticmethod(lambda x0, x1=None: True)})()
    x0, x1 = torch.empty(4, 16, 16, c0), torch.empty(2, 16, 16, c1)
    sp = blk._split_point(x0, x1)
    assert sp == want
    # the GroupNorm of the skip-only channels is shared whatever their number (engine.SPLIT_GN)
    cpg_, sg = (c0 + c1) // 32, blk._shared_from(x0, x1)
    assert sg == -(-(-(-c0 // cpg_) * cpg_) // 64) * 64 and (not sp or sg == sp)
    if sp:
        cpg = (c0 + c1) // 32
        assert sp % 64 == 0 and sp >= c0 and (sp // cpg) * cpg >= c0 and all((g * cpg >= c0) for g in range(-(-sp // cpg), 32))
    assert blk._split_point(x0, torch.empty(4, 16, 16, c1)) == 0      # unguided pass: nothing is shared
