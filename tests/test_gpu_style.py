"""Native style encoder on a B200: every kernel of stedm_b200/csrc/style_encoder.cu and the relaxed plain-GEMM mode of
stedm_conv_tc against a torch restatement (tests/fake_ops.py mirrors torchvision's shifted_window_attention), then
the whole StyleEncoderRunner and the aggregation blocks against torchvision's swin_v2_t — the library the reference
itself calls (networks/s_zss_dm.py:19-20, networks/agg_blocks.py).

Bars: fp32 mode max-abs 1e-4 on the 512-d style feature (the north_star fp32 bar for eps, applied here to the
conditioning vector); bf16 mode 2e-2 relative (max|d| / max|ref|)."""
import math

import pytest
import torch
import torch.nn.functional as F
import torchvision

from tests import fake_ops
from tests.util import max_abs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from stedm_b200 import ops as _ops
    return _ops


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def bf(x):
    return x.to(torch.bfloat16).float()


def _swin(seed=0, logit_hi=5.0):
    """swin_v2_t with every parameter off its default.  logit_hi: upper end of the logit-scale parameter range;
    5.0 reaches past the clamp at log(100) (fp32 tests), the bf16 tests stay at <= e^3 = 20 (the init is log 10):
    a cosine logit multiplied by 100 turns the 2^-9 operand rounding into 0.1 absolute on the softmax input."""
    torch.manual_seed(seed)
    m = torchvision.models.get_model("swin_v2_t")
    m.head = torch.nn.Linear(768, 512)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            elif "norm" in n and n.endswith("weight"):
                p.uniform_(0.5, 1.5)
            elif n.endswith("logit_scale"):
                p.uniform_(1.0, logit_hi)
    return m.eval()


# ---------------------------------------------------------------------------------------------- plain GEMM mode
@pytest.mark.parametrize("B,H,W,K,N,act", [
    (2, 32, 32, 96, 288, 0),     # partial K slab (96 = 64 + 32) and partial last channel tile (288 = 4 x 64 + 32)
    (2, 32, 32, 96, 96, 0),      # one 128-wide tile, 96 stored
    (3, 16, 16, 96, 384, 1),     # GELU epilogue
    (2, 16, 16, 384, 96, 0),
    (4, 8, 8, 192, 576, 0),
    (5, 4, 4, 768, 3072, 1),     # eight samples per tile, partial M tile, GELU
    (2, 4, 4, 3072, 768, 0),     # deep K
    (1, 64, 64, 96, 288, 0),
])
def test_conv_tc_plain_gemm_relaxed_shapes(ops, B, H, W, K, N, act):
    g = torch.Generator().manual_seed(K * 7 + N)
    x = bf(torch.randn(B, H, W, K, generator=g))
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    want = x @ w.t() + b
    if act:
        want = F.gelu(want)
    got = ops.conv(x.to(torch.bfloat16).cuda(), w.to(torch.bfloat16).cuda().contiguous(), b.cuda(), N, 1,
                   out_dtype=torch.float32, tensor_core=True, act=act)
    assert tuple(got.shape) == (B, H, W, N)
    assert max_abs(got, want) < 2e-3, max_abs(got, want)
    got16 = ops.conv(x.to(torch.bfloat16).cuda(), w.to(torch.bfloat16).cuda().contiguous(), b.cuda(), N, 1,
                     out_dtype=torch.bfloat16, tensor_core=True, act=act)
    assert max_abs(got16.float(), want) < 3e-2


def test_conv_simt_gelu(ops):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 8, 8, 96, generator=g)
    w = torch.randn(384, 96, generator=g) / 10
    b = torch.randn(384, generator=g)
    want = F.gelu(x @ w.t() + b)
    got = ops.conv(x.cuda(), w.t().contiguous().cuda(), b.cuda(), 384, 1, tensor_core=False, act=1)
    assert max_abs(got, want) < 1e-5


# ---------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("P", [32, 128, 256])
def test_patch_embed_ln(ops, P):
    g = torch.Generator().manual_seed(P)
    img = torch.rand(3, P, P, 3, generator=g) * 2 - 1
    w = torch.randn(48, 96, generator=g) / 7
    b, ga, be = torch.randn(96, generator=g), torch.rand(96, generator=g) + 0.5, torch.randn(96, generator=g)
    want, _ = fake_ops.patch_embed_ln(img, w, b, ga, be, 1e-5)
    f32, b16 = ops.patch_embed_ln(img.cuda(), w.cuda(), b.cuda(), ga.cuda(), be.cuda(), 1e-5, want_bf16=True)
    assert max_abs(f32, want) < 2e-5
    assert max_abs(b16.float(), want) < 3e-2


@pytest.mark.parametrize("C", [96, 192, 384, 768, 256, 1024, 8])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm(ops, C, dt):
    g = torch.Generator().manual_seed(C)
    rows = 1000 + 3   # not a multiple of the rows a block owns
    x = (torch.randn(rows, C, generator=g) * 3 + 1).to(dt)
    r = torch.randn(rows, C, generator=g)
    ga, be = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    want, _ = fake_ops.layernorm(x, r, ga, be, 1e-5)
    f32, b16 = ops.layernorm(x.cuda(), r.cuda(), ga.cuda(), be.cuda(), 1e-5, want_bf16=True)
    assert max_abs(f32, want) < 2e-5
    assert max_abs(b16.float(), want) < 5e-2
    want2, _ = fake_ops.layernorm(x, None, ga, be, 1e-5)
    f32, none = ops.layernorm(x.cuda(), None, ga.cuda(), be.cuda(), 1e-5)
    assert none is None and max_abs(f32, want2) < 2e-5


@pytest.mark.parametrize("H,W,heads,shift", [
    (16, 16, 3, 0), (16, 16, 3, 4),      # whole windows, shifted: four regions in the last window row / column
    (32, 16, 6, 4),                       # non-square map
    (8, 8, 12, 4),                        # one window: the shift is dropped (window covers the map)
    (4, 4, 24, 4),                        # map smaller than the window: padded keys carry the qkv bias
    (12, 20, 3, 4),                       # padded AND shifted (e.g. 224-pixel inputs)
])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_window_attention(ops, H, W, heads, shift, dt):
    g = torch.Generator().manual_seed(H * 100 + W + heads)
    C = heads * 32
    qkv = torch.randn(2, H, W, 3 * C, generator=g).to(dt)
    ls = torch.rand(heads, generator=g) * 20 + 1
    rb = torch.rand(heads, 64, 64, generator=g) * 16
    qb = torch.randn(3 * C, generator=g) * 0.3
    qb[C:2 * C] = 0
    want = fake_ops.window_attention(qkv.float(), ls, rb, qb, heads, shift)
    got = ops.window_attention(qkv.cuda(), ls.cuda(), rb.cuda(), qb.cuda(), heads, shift)
    assert got.dtype == dt and tuple(got.shape) == (2, H, W, C)
    # bf16: the tensor-core kernel rounds the normalised, logit-scaled q (|q| up to 21 here), k and the softmax
    # numerators to bf16 before its two MMAs; outputs are O(1-3)
    assert max_abs(got.float(), want) < (2e-5 if dt == torch.float32 else 4e-2), max_abs(got.float(), want)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_patch_merge_gather(ops, dt):
    x = torch.randn(3, 8, 12, 96).to(dt)
    assert torch.equal(ops.patch_merge_gather(x.cuda()).cpu(), fake_ops.patch_merge_gather(x))


@pytest.mark.parametrize("T,C", [(64, 768), (16, 768), (5, 96), (256, 1024)])
def test_ln_meanpool(ops, T, C):
    g = torch.Generator().manual_seed(T + C)
    x = torch.randn(3, T, C, generator=g) * 2 + 0.5
    ga, be = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    assert max_abs(ops.ln_meanpool(x.cuda(), ga.cuda(), be.cuda(), 1e-5), fake_ops.ln_meanpool(x, ga, be, 1e-5)) < 2e-5


@pytest.mark.parametrize("mode", ["mean", "max"])
def test_set_reduce(ops, mode):
    x = torch.randn(3, 10, 512)
    assert max_abs(ops.set_reduce(x.cuda(), mode), fake_ops.set_reduce(x, mode)) < 1e-6


def test_linear_relu_flags_and_wide_k(ops):
    g = torch.Generator().manual_seed(9)
    x, w, b = torch.randn(3, 5120, generator=g), torch.randn(512, 5120, generator=g) / 70, torch.randn(512, generator=g)
    want = F.relu(F.linear(F.relu(x), w, b))
    assert max_abs(ops.linear(x.cuda(), w.cuda(), b.cuda(), act_in="relu", relu_out=True), want) < 1e-4


# ---------------------------------------------------------------------------------------------- whole encoder
@pytest.mark.parametrize("P,B", [(128, 3), (256, 2), (64, 2)])
def test_style_encoder_fp32_matches_torchvision(P, B):
    from stedm_b200.style_engine import StyleEncoderRunner
    m = _swin(P).cuda()
    torch.manual_seed(P + 1)
    imgs = (torch.rand(B, P, P, 3) * 2 - 1).cuda()
    with torch.no_grad():
        want = m(imgs.permute(0, 3, 1, 2).contiguous())
        got = StyleEncoderRunner(m, "fp32")(imgs)
    assert tuple(got.shape) == (B, 512)
    assert max_abs(got, want) < 1e-4, max_abs(got, want)


@pytest.mark.parametrize("P,B", [(128, 3), (256, 4), (512, 1)])
def test_style_encoder_bf16_matches_torchvision(P, B):
    from stedm_b200.style_engine import StyleEncoderRunner
    m = _swin(P + 7, logit_hi=3.0).cuda()
    torch.manual_seed(P + 2)
    imgs = (torch.rand(B, P, P, 3) * 2 - 1).cuda()
    with torch.no_grad():
        want = m(imgs.permute(0, 3, 1, 2).contiguous())
        got = StyleEncoderRunner(m, "bf16")(imgs)
    rel = max_abs(got, want) / float(want.abs().max())
    print(f"style encoder bf16 rel err P={P}: {rel:.3e}")
    assert rel < 2e-2, rel


def test_style_encoder_is_batch_invariant_and_chunks():
    """A sample's feature is bit-identical whatever batch (or chunk of a large batch) it is computed in."""
    from stedm_b200.style_engine import StyleEncoderRunner
    m = _swin(11, logit_hi=3.0).cuda()
    torch.manual_seed(12)
    imgs = (torch.rand(6, 128, 128, 3) * 2 - 1).cuda()
    r = StyleEncoderRunner(m, "bf16")
    with torch.no_grad():
        full = r(imgs)
        assert torch.equal(full[2:4], r(imgs[2:4]))
        r.MAX_CHUNK_TOKENS = 2 * 32 * 32            # force two-image chunks
        assert torch.equal(full, r(imgs))


@pytest.mark.parametrize("agg", ["mean", "max", "linear"])
def test_agg_blocks_match_reference_formulas(agg):
    """Agg_Mean / Agg_Max / Agg_Linear (agg_blocks.py) on the native encoder vs the same reductions over torchvision."""
    from types import SimpleNamespace
    from stedm_b200.networks import agg_blocks
    m = _swin(21)
    cfg = SimpleNamespace(name="mp", num_patches=3)
    cls = {"mean": agg_blocks.Agg_Mean, "max": agg_blocks.Agg_Max, "linear": agg_blocks.Agg_Linear}[agg]
    torch.manual_seed(22)
    blk = cls(cfg, m).cuda().eval()
    blk.set_precision("fp32")
    style = (torch.rand(2, 3, 128, 128, 3) * 2 - 1).cuda()
    with torch.no_grad():
        f = m(style.reshape(6, 128, 128, 3).permute(0, 3, 1, 2).contiguous()).view(2, 3, 512)
        want = {"mean": lambda: f.mean(1), "max": lambda: f.max(1)[0],
                "linear": lambda: blk._linear_block(f.reshape(2, -1))}[agg]()
        got = blk(style)
    assert max_abs(got, want) < 1e-4, max_abs(got, want)


# ---------------------------------------------------------------------------------------------- style_agg=svit
def test_spt_patchify_assemble_token_mean(ops):
    g = torch.Generator().manual_seed(3)
    style = torch.rand(2, 3, 32, 32, 3, generator=g)
    assert torch.equal(ops.spt_patchify(style.cuda(), 8).cpu(), fake_ops.spt_patchify(style, 8))
    patches = torch.randn(2, 16, 64, generator=g)
    cls, pos = torch.randn(64, generator=g), torch.randn(18, 64, generator=g)
    for dt in (torch.float32, torch.bfloat16):
        got = ops.svit_assemble(patches.to(dt).cuda(), cls.cuda(), pos.cuda(), 128)
        assert max_abs(got, fake_ops.svit_assemble(patches.to(dt), cls, pos, 128)) < 1e-6
    x = torch.randn(3, 128, 256, generator=g)
    assert max_abs(ops.token_mean(x.cuda(), 66), fake_ops.token_mean(x, 66)) < 1e-6


@pytest.mark.parametrize("T,heads", [(66, 12), (258, 2), (1026, 3)])
def test_attention_tc_diagonal_mask_and_padded_rows(ops, T, heads):
    """LSA on the tcgen05 flash kernel: masked diagonal, learned temperature as the scale, token buffer padded to
    whole 128-row tiles (pad rows hold garbage on input and stay zero on output)."""
    g = torch.Generator().manual_seed(T)
    t_pad = (T + 127) // 128 * 128
    inner = heads * 64
    qkv = torch.randn(2, t_pad, 3 * inner, generator=g).to(torch.bfloat16)
    qkv[:, T:] = 1000.0                                    # must not leak into real rows
    want = fake_ops.attention_simt(qkv, qkv, qkv, heads, 64, T, 0, inner, 2 * inner, 3 * inner, 64, 0.21,
                                   torch.float32, mask_diag=True, batch_tokens=t_pad)
    q = qkv.cuda()
    got = ops.attention_tc(q, q, q, heads, 64, T, (t_pad * 3 * inner, 64, 3 * inner), 0.21, q_off=0, k_off=inner,
                           v_off=2 * inner, mask_diag=True, batch_tokens=t_pad)
    assert tuple(got.shape) == (2, t_pad, inner) and float(got[:, T:].abs().max()) == 0.0
    assert max_abs(got.float(), want) < 2e-2
    s = ops.attention_simt(q.float(), q.float(), q.float(), heads, 64, T, 0, inner, 2 * inner, 3 * inner, 64, 0.21,
                           torch.float32, mask_diag=True, batch_tokens=t_pad)
    assert max_abs(s, want) < 1e-4


@pytest.mark.parametrize("name", ["svit_p64_n2_mean", "svit_p128_n1_cls"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_svit_matches_reference_golden(name, precision):
    """Native sViT vs the reference's own networks/vit_set.py output on fixture weights (tests/golden/svit.npz)."""
    from oracle import stedm_oracle as O
    from tests.test_style_engine_glue import SVIT_CASES, build_svit
    from tests.util import load_golden
    P, ns, pool, seed = SVIT_CASES[name]
    blk = build_svit(P, ns, pool).agg_block.cuda()
    blk.set_precision(precision)
    _, style, _ = O.synthetic_batch(2, P, ns, seed)
    got = blk(style.cuda())
    want = load_golden("svit")[name]
    err = max_abs(got, want)
    if precision == "fp32":
        assert err < 1e-4, err
    else:
        rel = err / float(abs(want).max())
        print(f"sViT bf16 rel err {name}: {rel:.3e}")
        assert rel < 2e-2, rel


def test_svit_full_size_vs_oracle_and_model_wiring():
    """style_agg=svit at the path's real shape (256^2, 1026 tokens, N = 10 'mp' patches) through S_ZSS_DM.get_input."""
    from oracle import stedm_oracle as O
    from stedm_b200.modules.ldm_diffusion import LDM_Diffusion
    from stedm_b200.utils.fixture import apply_fixture_weights
    from tests.util import build_config, oracle_state_dict
    cfg = build_config(64, n_style=10, agg="svit")
    m = LDM_Diffusion(cfg, precision="bf16", load_first_stage_ckpt=False)
    apply_fixture_weights(m._model, seed=0)
    m = m.cuda().eval()
    seg, style, _ = O.synthetic_batch(2, 256, 10, seed=4)
    batch = {"image": torch.zeros(2, 256, 256, 3).cuda(), "segmentation": seg.cuda(), "style_imgs": style.cuda()}
    _, c = m._model.get_input(batch, "image")
    sd = {k: v for k, v in oracle_state_dict(m._model).items() if k.startswith("agg_block.")}
    with torch.no_grad():
        want = O.svit_aggregate(sd, style, heads=12, patch=8, pool="mean")
    got = c["c_crossattn"][0]
    rel = max_abs(got, want) / float(want.abs().max())
    print(f"sViT 256^2 N=10 bf16 rel err {rel:.3e}")
    assert tuple(got.shape) == (2, 512) and rel < 2e-2, rel
    m._model.set_precision("fp32")
    _, c = m._model.get_input(batch, "image")
    assert max_abs(c["c_crossattn"][0], want) < 1e-4
