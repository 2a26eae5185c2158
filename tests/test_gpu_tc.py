"""The tcgen05/TMEM/TMA implicit-GEMM convolution (stedm_conv_tc) against torch fp32 convolution of the SAME
bf16-rounded operands (so only the accumulation order differs), over every tile geometry the path uses:
W >= 128, W | 128 with several rows per tile, several samples per tile, partial last tile, two-source concat,
broadcast second source, every BN instantiation, fused bias + embedding + residual epilogue, bf16/fp32 output."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.util import max_abs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from stedm_b200 import ops as _ops
    return _ops


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def tc_w(w):
    """OIHW -> bf16 [Cout][kh*kw*Cin] (tap-major, channel-minor)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()


def bf(x):
    return x.to(torch.bfloat16).float()


CASES = [
    # B, H, W, Cin, Cout, k
    (2, 16, 16, 128, 256, 3),     # 16x8 pixel tiles, BN=256
    (3, 8, 8, 64, 128, 3),        # two samples per tile, partial last tile (M = 192), BN=128
    (1, 64, 64, 64, 128, 3),      # 64x2 tiles
    (1, 2, 128, 64, 64, 3),       # full-row tiles, BN=64
    (1, 4, 256, 64, 16, 3),       # half-row tiles, BN=16
    (2, 32, 32, 192, 512, 1),     # 1x1, K not a power of two
    (1, 16, 16, 1024, 1024, 3),   # deep K (144 slabs), 4 n-tiles
    (5, 4, 4, 64, 48, 3),         # eight samples per tile, cout = 3 x 16
    # halo mode (tiles of >= 2 whole rows of one sample: one activation box per (channel block, horizontal tap))
    (2, 32, 32, 128, 512, 3),     # 32x4 tiles + 2 halo rows, BN=256 CTA pairs
    (1, 24, 16, 64, 128, 3),      # 16x8 tiles, three tiles: the pair's odd tile is all out of bounds
    (3, 16, 16, 64, 16, 3),       # BN=16, two halo boxes in the ring
    (1, 16, 32, 192, 64, 3),      # non-square map, three channel blocks, BN=64
    (1, 8, 16, 64, 64, 3),        # H < tile rows + halo: falls back to one slab per tap
    (2, 64, 64, 128, 128, 3),     # 64x2 tiles + 2 halo rows (32 KB boxes, two per ring), BN=128 CTA pairs
    (1, 16, 128, 64, 128, 3),     # 128-wide map: full-row tiles, one slab per tap
    (1, 8, 256, 64, 64, 3),       # 256-wide map, BN=64
    (1, 128, 128, 64, 16, 3),     # image-head shape: BN=16 on a 128-wide map
]


@pytest.fixture()
def tile2d():
    """The opt-in 2-D pixel tiling (16 x 8 pixel tiles + halo boxes on maps at least 32 wide): STEDM_TC_TILE2D is read per call."""
    import os
    old = os.environ.get("STEDM_TC_TILE2D")
    os.environ["STEDM_TC_TILE2D"] = "1"
    yield
    if old is None:
        os.environ.pop("STEDM_TC_TILE2D", None)
    else:
        os.environ["STEDM_TC_TILE2D"] = old


@pytest.mark.parametrize("B,H,W,cin,cout,k", [(2, 32, 32, 128, 512, 3), (2, 64, 64, 128, 128, 3), (1, 16, 128, 64, 128, 3),
                                              (1, 8, 256, 64, 64, 3), (1, 128, 128, 64, 16, 3), (3, 64, 32, 192, 256, 3)])
def test_conv_tc_2d_tiles_match_fp32(ops, tile2d, B, H, W, cin, cout, k):
    test_conv_tc_matches_fp32(ops, B, H, W, cin, cout, k)


@pytest.mark.parametrize("B,H,W,cin,cout,k", CASES)
def test_conv_tc_matches_fp32(ops, B, H, W, cin, cout, k):
    g = torch.Generator().manual_seed(B * 1000 + H + cin)
    x = bf(torch.randn(B, cin, H, W, generator=g))
    w = bf(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    b = torch.randn(cout, generator=g)
    want = F.conv2d(x, w, b, padding=k // 2)
    got = ops.conv(nhwc(x).to(torch.bfloat16).cuda(), tc_w(w).cuda(), b.cuda(), cout, k, out_dtype=torch.float32,
                   tensor_core=True)
    torch.cuda.synchronize()
    err = max_abs(nchw(got.cpu()), want)
    assert err < 2e-3, err
    got_bf = ops.conv(nhwc(x).to(torch.bfloat16).cuda(), tc_w(w).cuda(), b.cuda(), cout, k,
                      out_dtype=torch.bfloat16, tensor_core=True)
    assert max_abs(nchw(got_bf.float().cpu()), want) < 4e-2


@pytest.mark.parametrize("k", [1, 3])
def test_conv_tc_concat_emb_residual(ops, k):
    g = torch.Generator().manual_seed(77 + k)
    B, c0, c1, co, hw = 4, 128, 64, 256, 16
    x0, x1 = bf(torch.randn(B, c0, hw, hw, generator=g)), bf(torch.randn(2, c1, hw, hw, generator=g))
    w = bf(torch.randn(co, c0 + c1, k, k, generator=g) / math.sqrt((c0 + c1) * k * k))
    b, emb_all = torch.randn(co, generator=g), torch.randn(B, 700, generator=g)
    res = bf(torch.randn(B, co, hw, hw, generator=g))
    want = (F.conv2d(torch.cat([x0, torch.cat([x1, x1], 0)], 1), w, b, padding=k // 2)
            + emb_all[:, 256:256 + co, None, None] + res)
    emb = emb_all.cuda()[:, 256:256 + co]
    for res_dt in (torch.bfloat16, torch.float32):
        got = ops.conv(nhwc(x0).to(torch.bfloat16).cuda(), tc_w(w).cuda(), b.cuda(), co, k,
                       x1=nhwc(x1).to(torch.bfloat16).cuda(), emb=emb, residual=nhwc(res).to(res_dt).cuda(),
                       out_dtype=torch.float32, tensor_core=True)
        assert max_abs(nchw(got.cpu()), want) < 2e-3


@pytest.mark.parametrize("hw", [32, 64])
def test_conv_tc_2d_tiles_concat_emb_residual_stats(ops, tile2d, hw):
    """The 2-D pixel tiles (maps at least 32 wide) with everything the epilogue does: two-source concat with a broadcast
    second source, embedding rows, bf16 residual (requested a chunk ahead), bf16 NHWC output through the staging
    transpose, GroupNorm tile statistics — against torch, and bit-equal to the full-row tiling (STEDM_TC_TILE2D=0 is a
    load-time switch, so the reference here is the same convolution run as one 3x3 tap-major GEMM over im2col'ed rows...
    which has another K order; compare with tolerance instead and pin the statistics)."""
    g = torch.Generator().manual_seed(hw)
    B, c0, c1, co = 4, 128, 64, 256
    x0, x1 = bf(torch.randn(B, c0, hw, hw, generator=g)), bf(torch.randn(2, c1, hw, hw, generator=g))
    w = bf(torch.randn(co, c0 + c1, 3, 3, generator=g) / math.sqrt((c0 + c1) * 9))
    b, emb = torch.randn(co, generator=g), torch.randn(B, co, generator=g)
    res = bf(torch.randn(B, co, hw, hw, generator=g))
    want = F.conv2d(torch.cat([x0, torch.cat([x1, x1], 0)], 1), w, b, padding=1) + emb[:, :, None, None] + res
    m_tiles = B * hw * hw // 128
    stats = torch.zeros((m_tiles, co, 2), device="cuda")
    got = ops.conv(nhwc(x0).to(torch.bfloat16).cuda(), tc_w(w).cuda(), b.cuda(), co, 3, x1=nhwc(x1).to(torch.bfloat16).cuda(),
                   emb=emb.cuda(), residual=nhwc(res).to(torch.bfloat16).cuda(), out_dtype=torch.bfloat16, tensor_core=True,
                   stats_out=stats)
    assert max_abs(nchw(got.float().cpu()), want) < 2 ** -8 * float(want.abs().max()) + 3e-3     # bf16 output rounding
    # per-sample channel sums from the tile statistics (whatever pixels a tile holds, a sample owns hw*hw/128 of them):
    # statistics of the stored bf16 tensor
    gs = nchw(got.float().cpu()).double()
    per_sample = stats.cpu().double().reshape(B, hw * hw // 128, co, 2).sum(1)
    assert max_abs(per_sample[..., 0], gs.sum((2, 3))) < 2e-2
    rel = ((per_sample[..., 1] - (gs ** 2).sum((2, 3))).abs() / (gs ** 2).sum((2, 3))).max()
    assert float(rel) < 1e-5


def test_conv_tc_stride2_via_im2col(ops):
    """Downsample (3x3, stride 2, pad 1) = im2col_3x3_s2 + 1x1 tensor-core GEMM."""
    g = torch.Generator().manual_seed(5)
    B, c, hw = 2, 128, 32
    x = bf(torch.randn(B, c, hw, hw, generator=g))
    w = bf(torch.randn(c, c, 3, 3, generator=g) / math.sqrt(9 * c))
    b = torch.randn(c, generator=g)
    want = F.conv2d(x, w, b, stride=2, padding=1)
    cols = ops.im2col_3x3_s2(nhwc(x).to(torch.bfloat16).cuda())
    got = ops.conv(cols, tc_w(w).cuda(), b.cuda(), c, 1, out_dtype=torch.float32, tensor_core=True)
    assert max_abs(nchw(got.cpu()), want) < 2e-3


@pytest.mark.parametrize("B,cin,cout,H,W", [(2, 128, 128, 32, 32), (3, 64, 512, 64, 64), (1, 192, 64, 16, 32),
                                            (2, 512, 512, 32, 32), (1, 64, 128, 256, 256)])
def test_conv_tc_stride2_gather_in_the_tma_unit(ops, B, cin, cout, H, W):
    """Downsample (openaimodel.py:164-166: 3x3, stride 2, pad 1) inside conv_tc: tensor-map boxes traversed with element
    stride 2 sample the tile's input positions, out-of-bounds coordinate -1 is the padding; equal to the im2col + GEMM
    form bit for bit (same K order) and to torch."""
    g = torch.Generator().manual_seed(B * 7 + cin)
    x = bf(torch.randn(B, cin, H, W, generator=g))
    w = bf(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    b = torch.randn(cout, generator=g)
    emb = torch.randn(B, cout, generator=g)
    want = F.conv2d(x, w, b, stride=2, padding=1) + emb[:, :, None, None]
    xd = nhwc(x).to(torch.bfloat16).cuda()
    m_tiles = B * (H // 2) * (W // 2) // 128
    stats = torch.zeros((max(1, m_tiles), cout, 2), device="cuda") if ((H // 2) * (W // 2)) % 128 == 0 and cout % 64 == 0 else None
    got = ops.conv(xd, tc_w(w).cuda(), b.cuda(), cout, 3, stride=2, emb=emb.cuda(), out_dtype=torch.float32,
                   tensor_core=True, stats_out=stats, split_k=False)
    assert tuple(got.shape) == (B, H // 2, W // 2, cout)
    assert max_abs(nchw(got.cpu()), want) < 3e-3, max_abs(nchw(got.cpu()), want)
    ref = ops.conv(ops.im2col_3x3_s2(xd), tc_w(w).cuda(), b.cuda(), cout, 1, emb=emb.cuda(), out_dtype=torch.float32,
                   tensor_core=True, split_k=False)
    assert torch.equal(got, ref)
    auto = ops.conv(xd, tc_w(w).cuda(), b.cuda(), cout, 3, stride=2, emb=emb.cuda(), out_dtype=torch.float32, tensor_core=True)
    assert max_abs(auto, got) < 1e-4 * max(1.0, float(want.abs().max()))
    if stats is not None:
        t = nhwc(want).reshape(-1, 128, cout)
        assert max_abs(stats[..., 0].cpu(), t.sum(1)) < 0.05


def test_conv_tc_rejects_unsupported(ops):
    x = torch.zeros(1, 12, 12, 64, device="cuda", dtype=torch.bfloat16)   # width 12 does not tile 128 pixels
    w = torch.zeros(64, 9 * 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.conv(x, w, None, 64, 3, tensor_core=True)


def test_conv_tc_nchw_head(ops):
    """The eps / image heads (128 -> 3): weight rows zero-padded to one 16-wide UMMA tile, the epilogue stores the
    3 real channels as NCHW fp32."""
    g = torch.Generator().manual_seed(9)
    B, c, hw = 3, 128, 32
    x = bf(torch.randn(B, c, hw, hw, generator=g))
    w = bf(torch.randn(3, c, 3, 3, generator=g) / math.sqrt(9 * c))
    b = torch.randn(3, generator=g)
    want = F.conv2d(x, w, b, padding=1)
    wp = torch.zeros(16, c, 3, 3)
    wp[:3] = w
    bp = torch.zeros(16)
    bp[:3] = b
    got = ops.conv(nhwc(x).to(torch.bfloat16).cuda(), tc_w(wp).cuda(), bp.cuda(), 16, 3, out_dtype=torch.float32,
                   tensor_core=True, out_nchw=True, cout_store=3)
    assert tuple(got.shape) == (B, 3, hw, hw)
    assert max_abs(got.cpu(), want) < 2e-3


@pytest.mark.parametrize("heads,ch,T,B", [(8, 128, 256, 2), (4, 64, 64, 3), (8, 128, 1024, 1), (2, 128, 100, 2),
                                          (1, 64, 300, 1)])
def test_attention_tc_legacy_layout(ops, heads, ch, T, B):
    """Fused tcgen05 flash attention vs QKVAttentionLegacy math (openaimodel.py:378-394) on bf16-rounded q, k, v."""
    g = torch.Generator().manual_seed(heads * 100 + T)
    Cc = heads * ch
    qkv = bf(torch.randn(B, 3 * Cc, T, generator=g))
    q, k, v = qkv.reshape(B * heads, 3 * ch, T).split(ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
    want = torch.einsum("bts,bcs->bct", w, v).reshape(B, Cc, T).permute(0, 2, 1)          # [B, T, C]
    buf = qkv.permute(0, 2, 1).contiguous().to(torch.bfloat16).cuda()                      # [B, T, 3C] head-major
    got = ops.attention_tc(buf, buf, buf, heads, ch, T, (T * 3 * Cc, 3 * ch, 3 * Cc), 1 / math.sqrt(ch),
                           q_off=0, k_off=ch, v_off=2 * ch)
    torch.cuda.synchronize()
    err = max_abs(got.float().cpu(), want)
    assert err < 2e-2 * max(1.0, float(want.abs().max())), err


@pytest.mark.parametrize("B,T,Tkv", [(2, 1024, 0), (1, 4096, 0), (3, 200, 0), (2, 256, 77), (2, 100, 0), (1, 128, 300)])
def test_attention_tc_wide_single_head_d512(ops, B, T, Tkv):
    """The VAE decoder's AttnBlock (model.py:178-202: one head, d = 512, scale 512^-1/2) on the wide flash kernel — Q
    resident, K / V streamed in 64-channel slabs, output channels split over two CTAs, no T x T tensor: against fp32
    softmax(q k^T / sqrt(d)) v of the same bf16-rounded q | k | v buffer; ragged T (keys masked, rows not stored) and a
    separate key / value length."""
    g = torch.Generator().manual_seed(T + B)
    C = 512
    qkv = bf(torch.randn(B, T, 3 * C, generator=g))
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    buf = qkv.to(torch.bfloat16).cuda()
    scale = C ** -0.5
    if Tkv:
        kv = bf(torch.randn(B, Tkv, 2 * C, generator=g))
        k, v = kv[..., :C], kv[..., C:]
        kvd = kv.to(torch.bfloat16).cuda()
        got = ops.attention_tc(buf, kvd, kvd, 1, C, T, (T * 3 * C, C, 3 * C), scale, q_off=0, k_off=0, v_off=C,
                               tokens_kv=Tkv, kv_strides=(Tkv * 2 * C, C, 2 * C))
    else:
        got = ops.attention_tc(buf, buf, buf, 1, C, T, (T * 3 * C, C, 3 * C), scale, q_off=0, k_off=C, v_off=2 * C)
    torch.cuda.synchronize()
    # peaky logits too: scale q up for half of the samples' rows so that the running maximum keeps moving
    w = torch.softmax(torch.einsum("btc,bsc->bts", q, k) * scale, dim=-1)
    want = torch.einsum("bts,bsc->btc", w, v)
    assert tuple(got.shape) == (B, T, C)
    err = max_abs(got.float().cpu(), want)
    assert err < 2e-2 * max(1.0, float(want.abs().max())), err
    assert ops.attention_tc_supported(512, T, heads=1) and not ops.attention_tc_supported(512, T, heads=2)


def test_attention_tc_wide_moving_maximum(ops):
    """Online softmax with a maximum that keeps growing along the keys (the O rescale path in TMEM) and rows whose
    maximum never moves after the first tile (the warp-uniform skip)."""
    g = torch.Generator().manual_seed(9)
    B, T, C = 1, 512, 512
    q = bf(torch.randn(B, T, C, generator=g))
    k = bf(torch.randn(B, T, C, generator=g) * torch.linspace(0.2, 3.0, T)[None, :, None])   # later keys: larger logits
    k[:, :, :] = torch.where(torch.arange(T)[None, :, None] < 128, k, k)                   # (first tile unchanged)
    q[:, :64] *= 0.0                                                                        # flat rows: maximum 0 everywhere
    v = bf(torch.randn(B, T, C, generator=g))
    buf = torch.cat([q, k, v], -1).to(torch.bfloat16).cuda()
    got = ops.attention_tc(buf, buf, buf, 1, C, T, (T * 3 * C, C, 3 * C), C ** -0.5, q_off=0, k_off=C, v_off=2 * C)
    w = torch.softmax(torch.einsum("btc,bsc->bts", q, k) * C ** -0.5, dim=-1)
    want = torch.einsum("bts,bsc->btc", w, v)
    err = max_abs(got.float().cpu(), want)
    assert err < 2e-2 * max(1.0, float(want.abs().max())), err


def test_softmax_rows_scaled_bf16(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(300, 1000, generator=g) * 3
    want = torch.softmax(x * 0.25, dim=-1)
    out = torch.empty(300, 1000, dtype=torch.bfloat16, device="cuda")
    ops.softmax_rows(x.cuda(), 0.25, out=out)
    assert max_abs(out.float().cpu(), want) < 2e-3
    y = x.cuda()
    ops.softmax_rows(y, 0.25)
    assert max_abs(y.cpu(), want) < 1e-6


def test_conv_tc_folded_upsample(ops):
    """nearest-x2 upsample + 3x3 conv as four 2x2 sub-pixel phase convs on the low-resolution input
    (openaimodel.py:123-132 / model.py:53-57) against F.interpolate + conv2d."""
    from stedm_b200.engine import PackedConv, Precision
    g = torch.Generator().manual_seed(21)
    B, c, co, hw = 2, 128, 256, 16
    x = bf(torch.randn(B, c, hw, hw, generator=g))
    w = torch.randn(co, c, 3, 3, generator=g) / math.sqrt(9 * c)
    b = torch.randn(co, generator=g)
    want = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    pc = PackedConv(w.cuda(), b.cuda(), Precision("bf16"), fold_upsample=True)
    got = pc(nhwc(x).to(torch.bfloat16).cuda(), upsample=True, out_dtype=torch.float32)
    assert tuple(got.shape) == (B, 2 * hw, 2 * hw, co)
    # the phase weights are sums of fp32 taps rounded once to bf16 (not sums of bf16-rounded taps): ~1e-2 abs
    assert max_abs(nchw(got.cpu()), want) < 3e-2


@pytest.mark.parametrize("cout,hw,B", [(256, 16, 3), (128, 32, 2), (64, 16, 2)])
def test_conv_tc_fused_groupnorm_statistics(ops, cout, hw, B):
    """The epilogue's per-(tile, channel) sums, folded per sample, equal the GroupNorm statistics of the conv
    output; the normalised result matches F.group_norm of that output."""
    g = torch.Generator().manual_seed(cout + hw)
    c = 128
    x = bf(torch.randn(B, c, hw, hw, generator=g))
    w = bf(torch.randn(cout, c, 3, 3, generator=g) / math.sqrt(9 * c))
    b = torch.randn(cout, generator=g)
    res = bf(torch.randn(B, cout, hw, hw, generator=g))
    m_tiles = B * hw * hw // 128
    tiles = torch.full((m_tiles, cout, 2), float("nan"), device="cuda")
    y = ops.conv(nhwc(x).to(torch.bfloat16).cuda(), tc_w(w).cuda(), b.cuda(), cout, 3, residual=nhwc(res).to(torch.bfloat16).cuda(),
                 out_dtype=torch.bfloat16, tensor_core=True, stats_out=tiles)
    want_y = F.conv2d(x, w, b, padding=1) + res
    assert max_abs(nchw(y.float().cpu()), want_y) < 2 ** -8 * float(want_y.abs().max()) + 3e-3
    # the statistics describe the STORED (bf16-rounded) tensor — the values the GroupNorm will normalise
    ys = nchw(y.float().cpu()).double()
    sums = tiles.cpu().double().reshape(B, hw * hw // 128, cout, 2).sum(1)              # per sample, per channel
    assert max_abs(sums[..., 0], ys.sum((2, 3))) < 5e-3
    assert float(((sums[..., 1] - (ys ** 2).sum((2, 3))).abs() / (ys ** 2).sum((2, 3))).max()) < 1e-5
    folded = ops.gn_fold_tiles((tiles, cout, 1, m_tiles, hw * hw // 128, B), None, B)
    gamma, beta = torch.randn(cout, generator=g), torch.randn(cout, generator=g)
    got = ops.gn_apply(y, None, folded, gamma.cuda(), beta.cuda(), 1e-5, True, torch.float32, n_chunks=1)
    want = F.silu(F.group_norm(nchw(y.float().cpu()), 32, gamma, beta, 1e-5))
    assert max_abs(nchw(got.cpu()), want) < 2e-2


@pytest.mark.parametrize("B,hw,cin,cout,k", [(1, 16, 1024, 1024, 3), (2, 16, 512, 256, 3), (1, 32, 256, 128, 1),
                                             (1, 8, 2048, 1024, 3)])
def test_conv_tc_split_k_small_batch(ops, B, hw, cin, cout, k):
    """Few output tiles + deep K: the launch splits K across the idle SMs (workspace + finish kernel); the result
    must equal the single-pass one up to fp32 summation order, with every epilogue term applied exactly once."""
    g = torch.Generator().manual_seed(B + cin)
    x = bf(torch.randn(B, cin, hw, hw, generator=g))
    w = bf(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    b = torch.randn(cout, generator=g)
    res = bf(torch.randn(B, cout, hw, hw, generator=g))
    emb = torch.randn(B, cout, generator=g)
    want = F.conv2d(x, w, b, padding=k // 2) + res + emb[:, :, None, None]
    xd, wd = nhwc(x).to(torch.bfloat16).cuda(), tc_w(w).cuda()
    assert ops.SPLIT_K[0], "split-K is the default for launches with few output tiles"
    got = ops.conv(xd, wd, b.cuda(), cout, k, emb=emb.cuda(), residual=nhwc(res).to(torch.bfloat16).cuda(),
                   out_dtype=torch.float32, tensor_core=True)
    assert max_abs(nchw(got.cpu()), want) < 3e-3
    single = ops.conv(xd, wd, b.cuda(), cout, k, emb=emb.cuda(), residual=nhwc(res).to(torch.bfloat16).cuda(),
                      out_dtype=torch.float32, tensor_core=True, split_k=False)
    assert max_abs(got, single) < 1e-4 * max(1.0, float(want.abs().max()))   # the same sum in another fp32 order
    again = ops.conv(xd, wd, b.cuda(), cout, k, emb=emb.cuda(), residual=nhwc(res).to(torch.bfloat16).cuda(),
                     out_dtype=torch.float32, tensor_core=True)
    assert torch.equal(got, again)                                            # fixed-order reduce: deterministic
    # the finish pass publishes the GroupNorm tile statistics of a split launch, like the single-pass epilogue
    if (hw * hw) % 128 == 0:
        m_tiles = B * hw * hw // 128
        tiles, tiles1 = torch.zeros((m_tiles, cout, 2), device="cuda"), torch.zeros((m_tiles, cout, 2), device="cuda")
        got2 = ops.conv(xd, wd, b.cuda(), cout, k, out_dtype=torch.float32, tensor_core=True, stats_out=tiles)
        ops.conv(xd, wd, b.cuda(), cout, k, out_dtype=torch.float32, tensor_core=True, stats_out=tiles1, split_k=False)
        assert got2._stats_written
        want2 = F.conv2d(x, w, b, padding=k // 2)
        assert max_abs(nchw(got2.cpu()), want2) < 3e-3
        t = nhwc(want2).reshape(m_tiles, 128, cout)
        assert max_abs(tiles[..., 0].cpu(), t.sum(1)) < 0.05 and max_abs(tiles, tiles1) < 0.05


@pytest.mark.parametrize("B,hw,cmid,cs0,cs1,co,sb", [
    (4, 16, 128, 128, 64, 256, 2),     # two-source skip, second source broadcast (b % 2), BN=256 pair
    (2, 32, 128, 256, 0, 128, 0),      # one skip source, BN=128
    (3, 8, 64, 64, 64, 64, 0),         # several samples per tile, partial last tile, BN=64
])
def test_conv_tc_fused_skip_connection(ops, B, hw, cmid, cs0, cs1, co, sb):
    """ResBlock tail in one launch: conv3x3(a) + bias + skip_1x1([x0 | x1]) + skip bias, the skip GEMM's weights appended
    along K (openaimodel.py:246-256, 288); with emb and fused GroupNorm statistics of the sum."""
    g = torch.Generator().manual_seed(B * 10 + hw)
    a = bf(torch.randn(B, cmid, hw, hw, generator=g))
    s0 = bf(torch.randn(B, cs0, hw, hw, generator=g))
    s1 = bf(torch.randn(sb or B, cs1, hw, hw, generator=g)) if cs1 else None
    w3 = bf(torch.randn(co, cmid, 3, 3, generator=g) / math.sqrt(9 * cmid))
    w1 = bf(torch.randn(co, cs0 + cs1, 1, 1, generator=g) / math.sqrt(cs0 + cs1))
    b3, b1 = torch.randn(co, generator=g), torch.randn(co, generator=g)
    skip_in = s0 if s1 is None else torch.cat([s0, s1.repeat(B // s1.shape[0], 1, 1, 1)], 1)
    want = F.conv2d(a, w3, b3, padding=1) + F.conv2d(skip_in, w1, b1)
    wcat = torch.cat([tc_w(w3), tc_w(w1)], 1).contiguous()
    dev = lambda t: None if t is None else nhwc(t).to(torch.bfloat16).cuda()
    stats = torch.empty((B * hw * hw // 128, co, 2), device="cuda") if (hw * hw) % 128 == 0 and co % 64 == 0 else None
    got = ops.conv(dev(a), wcat.cuda(), (b3 + b1).cuda(), co, 3, out_dtype=torch.float32, tensor_core=True,
                   skip_x0=dev(s0), skip_x1=dev(s1), stats_out=stats)
    assert max_abs(nchw(got.cpu()), want) < 3e-3, max_abs(nchw(got.cpu()), want)
    if stats is not None:
        tiles = nhwc(want).reshape(-1, 128, co)
        assert max_abs(stats[..., 0].cpu(), tiles.sum(1)) < 0.05 and max_abs(stats[..., 1].cpu(), (tiles ** 2).sum(1)) < 0.5


def test_conv_tc_channel_slice_input_and_broadcast_residual(ops):
    """conv([h | s]) as conv_h(a[..., :c0]) + conv_s(a[:Bs, ..., c0:]): channel-slice views of one NHWC tensor as inputs
    (x0_pix_stride) and the fp32 partial of the shared half added as a residual broadcast over b % Bs (res_batch)."""
    g = torch.Generator().manual_seed(123)
    B, Bs, c0, c1, co, hw = 4, 2, 128, 64, 256, 16
    h = bf(torch.randn(B, c0, hw, hw, generator=g))
    s = bf(torch.randn(Bs, c1, hw, hw, generator=g))
    w = bf(torch.randn(co, c0 + c1, 3, 3, generator=g) / math.sqrt(9 * (c0 + c1)))
    b = torch.randn(co, generator=g)
    a_full = torch.cat([h, s.repeat(B // Bs, 1, 1, 1)], 1)
    want = F.conv2d(a_full, w, b, padding=1)
    a = nhwc(a_full).to(torch.bfloat16).cuda()                       # [B, H, W, c0 + c1]
    part = ops.conv(a[:Bs, :, :, c0:], tc_w(w[:, c0:]).cuda(), None, co, 3, out_dtype=torch.float32, tensor_core=True)
    assert tuple(part.shape) == (Bs, hw, hw, co)
    got = ops.conv(a[..., :c0], tc_w(w[:, :c0]).cuda(), b.cuda(), co, 3, residual=part, out_dtype=torch.float32,
                   tensor_core=True)
    assert max_abs(nchw(got.cpu()), want) < 3e-3, max_abs(nchw(got.cpu()), want)
    whole = ops.conv(a, tc_w(w).cuda(), b.cuda(), co, 3, out_dtype=torch.float32, tensor_core=True)
    assert max_abs(got, whole) < 1e-4                                 # only the fp32 summation order differs


# ---------------------------------------------------------------------------------------------------------------------
# GroupNorm + SiLU inside the convolution's operand path (stedm_conv_desc.gn_coef)
# ---------------------------------------------------------------------------------------------------------------------
def _tile_sums(x_nhwc_bf16):
    """What a producer's epilogue writes: fp32 (sum, sum of squares) per (128-pixel tile, channel) of its bf16 output."""
    b, h, w, c = x_nhwc_bf16.shape
    t = x_nhwc_bf16.float().reshape(b * h * w // 128, 128, c)
    tiles = torch.stack([t.sum(1), (t * t).sum(1)], -1).contiguous()
    return (tiles.cuda(), c, 1, tiles.shape[0], h * w // 128, b)


GN_CASES = [
    # B, H, W, c0, c1, x1 batch, cout, silu
    (4, 16, 16, 1024, 0, 0, 1024, True),     # the U-Net's deep layers: 16x8 tiles + 2 halo rows, 16 channel blocks
    (2, 32, 32, 512, 0, 0, 512, True),       # 32x4 tiles, 24 KB boxes
    (4, 16, 16, 512, 256, 2, 512, True),     # concat with a broadcast second source; 24-channel groups straddle blocks
    (3, 16, 16, 256, 0, 0, 256, False),      # no SiLU, odd sample count
    (1, 24, 16, 128, 0, 0, 256, True),       # three tiles: the pair's second tile is all out of bounds
]


@pytest.mark.parametrize("B,H,W,c0,c1,xb,cout,silu", GN_CASES)
def test_conv_tc_groupnorm_in_operand_path_is_bit_equal_to_apply_then_conv(ops, B, H, W, c0, c1, xb, cout, silu):
    """conv(x, gn_coef=...) == conv(gn_apply(x)) bit for bit (same formula, same bf16 rounding, same K order), and both
    match torch's fp32 GroupNorm + SiLU + conv of the same bf16 operands."""
    g = torch.Generator().manual_seed(B * 1000 + c0 + c1)
    C = c0 + c1
    x0 = (torch.randn(B, H, W, c0, generator=g) * 2 + 0.5).to(torch.bfloat16)
    x1 = (torch.randn(xb, H, W, c1, generator=g) * 0.7 - 1).to(torch.bfloat16) if c1 else None
    w = bf(torch.randn(cout, C, 3, 3, generator=g) / math.sqrt(9 * C))
    bias, emb = torch.randn(cout, generator=g), torch.randn(B, cout, generator=g)
    res = torch.randn(B, H, W, cout, generator=g).to(torch.bfloat16)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    s0, s1 = _tile_sums(x0), (_tile_sums(x1) if c1 else None)
    assert ops.conv_gn_fusable(B, H, W, C, cout)
    coef = ops.gn_fold_tiles(s0, s1, B, coef_for=(gamma.cuda(), beta.cuda(), 1e-5, H * W))
    folded = ops.gn_fold_tiles(s0, s1, B)
    x0d, x1d = x0.cuda(), (x1.cuda() if c1 else None)
    a = ops.gn_apply(x0d, x1d, folded, gamma.cuda(), beta.cuda(), 1e-5, silu, torch.bfloat16, n_chunks=1)
    kw = dict(emb=emb.cuda(), residual=res.cuda(), out_dtype=torch.bfloat16, tensor_core=True)
    m_tiles = B * H * W // 128
    t_ref, t_fus = torch.zeros((m_tiles, cout, 2), device="cuda"), torch.zeros((m_tiles, cout, 2), device="cuda")
    ref = ops.conv(a, tc_w(w).cuda(), bias.cuda(), cout, 3, stats_out=t_ref, split_k=False, **kw)
    fus = ops.conv(x0d, tc_w(w).cuda(), bias.cuda(), cout, 3, x1=x1d, stats_out=t_fus, gn_coef=coef, gn_silu=silu, **kw)
    torch.cuda.synchronize()
    assert torch.equal(ref, fus), float((ref.float() - fus.float()).abs().max())
    assert torch.equal(t_ref, t_fus)
    # and against torch: GroupNorm in fp32 on the bf16 inputs, operand rounded to bf16, fp32 convolution
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float().repeat(B // xb, 1, 1, 1)], -1)
    n = F.group_norm(xin.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    n = bf(F.silu(n) if silu else n)
    want = F.conv2d(n, w, bias, padding=1) + emb[:, :, None, None] + res.float().permute(0, 3, 1, 2)
    err = max_abs(nchw(fus.float().cpu()), want)
    assert err < 0.06 * max(1.0, float(want.abs().max()) / 8), err


def test_conv_tc_groupnorm_in_operand_path_with_fused_skip_and_channel_slices(ops):
    """The ResBlock tail conv3x3(SiLU(GN(h))) + skip1x1([x0 | x1]) with the normalisation in the kernel (the raw skip slabs
    ride in the same ring, three per set), and the split [h | skip] form: x1 as a channel SLICE of the skip tensor with a
    coefficient-table offset (gn_c_off) for the shared part."""
    g = torch.Generator().manual_seed(77)
    B, Bs, hw, cm, co = 4, 2, 16, 512, 512
    h = (torch.randn(B, hw, hw, cm, generator=g) * 1.5).to(torch.bfloat16)
    k0 = torch.randn(B, hw, hw, 192, generator=g).to(torch.bfloat16)
    k1 = torch.randn(Bs, hw, hw, 64, generator=g).to(torch.bfloat16)
    w3 = bf(torch.randn(co, cm, 3, 3, generator=g) / math.sqrt(9 * cm))
    w1 = bf(torch.randn(co, 256, 1, 1, generator=g) / 16)
    bias = torch.randn(co, generator=g)
    gamma, beta = torch.randn(cm, generator=g), torch.randn(cm, generator=g)
    sh = _tile_sums(h)
    coef = ops.gn_fold_tiles(sh, None, B, coef_for=(gamma.cuda(), beta.cuda(), 1e-5, hw * hw))
    a = ops.gn_apply(h.cuda(), None, ops.gn_fold_tiles(sh, None, B), gamma.cuda(), beta.cuda(), 1e-5, True,
                     torch.bfloat16, n_chunks=1)
    wcat = torch.cat([tc_w(w3), tc_w(w1)], 1).contiguous().cuda()
    assert ops.conv_gn_fusable(B, hw, hw, cm, co, 256)
    kw = dict(out_dtype=torch.float32, tensor_core=True, skip_x0=k0.cuda(), skip_x1=k1.cuda())
    ref = ops.conv(a, wcat, bias.cuda(), co, 3, split_k=False, **kw)
    fus = ops.conv(h.cuda(), wcat, bias.cuda(), co, 3, gn_coef=coef, **kw)
    assert torch.equal(ref, fus), float((ref - fus).abs().max())
    # split concat: GN over [h (512) | s (512, Bs samples)], shared channels [sp, 1024) convolved once per distinct sample
    s = (torch.randn(Bs, hw, hw, 512, generator=g) + 0.3).to(torch.bfloat16)
    C, sp = 1024, 576           # 32-channel groups: none straddles, but split one K slab into the skip half
    wc = bf(torch.randn(co, C, 3, 3, generator=g) / math.sqrt(9 * C))
    g2, b2 = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ss = _tile_sums(s)
    coef2 = ops.gn_fold_tiles(sh, ss, B, coef_for=(g2.cuda(), b2.cuda(), 1e-5, hw * hw))
    a2 = ops.gn_apply(h.cuda(), s.cuda(), ops.gn_fold_tiles(sh, ss, B), g2.cuda(), b2.cuda(), 1e-5, True, torch.bfloat16,
                      n_chunks=1)
    wv = wc.permute(0, 2, 3, 1)                                                   # [co, 3, 3, C]
    w_lo = wv[..., :sp].reshape(co, -1).to(torch.bfloat16).contiguous().cuda()
    w_hi = wv[..., sp:].reshape(co, -1).to(torch.bfloat16).contiguous().cuda()
    part_ref = ops.conv(a2[:Bs, :, :, sp:], w_hi, None, co, 3, out_dtype=torch.float32, tensor_core=True, split_k=False)
    ref2 = ops.conv(a2[..., :sp], w_lo, bias.cuda(), co, 3, residual=part_ref, out_dtype=torch.bfloat16, tensor_core=True,
                    split_k=False)
    sd = s.cuda()
    part = ops.conv(sd[:, :, :, sp - cm:], w_hi, None, co, 3, out_dtype=torch.float32, tensor_core=True, gn_coef=coef2,
                    gn_c_off=sp)
    assert torch.equal(part, part_ref), float((part - part_ref).abs().max())
    fus2 = ops.conv(h.cuda(), w_lo, bias.cuda(), co, 3, x1=sd[:, :, :, :sp - cm], residual=part, out_dtype=torch.bfloat16,
                    tensor_core=True, gn_coef=coef2)
    assert torch.equal(ref2, fus2), float((ref2.float() - fus2.float()).abs().max())


def test_conv_tc_groupnorm_in_operand_path_rejects_unsupported_shapes(ops):
    """64-wide maps (32 KB boxes), 128-wide channel tiles and maps with several samples per tile are not fusable: the
    planner says so and a forced call fails with a message instead of computing something else."""
    assert not ops.conv_gn_fusable(2, 64, 64, 128, 256)     # six 32 KB boxes do not fit beside the weight ring
    assert not ops.conv_gn_fusable(2, 16, 16, 256, 128)     # BN = 128
    assert not ops.conv_gn_fusable(4, 8, 8, 256, 256)       # two samples per pixel tile
    x = torch.zeros(2, 16, 16, 256, dtype=torch.bfloat16, device="cuda")
    coef = torch.zeros(2, 256, 2, device="cuda")
    w = torch.zeros(128, 9 * 256, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="gn_coef"):
        ops.conv(x, w, None, 128, 3, tensor_core=True, gn_coef=coef)


# ---------------------------------------------------------------------------------------------------------------------
# out-of-bounds canaries (compute-sanitizer is not available on the GPU pool): outputs are views into larger buffers
# filled with a sentinel; the kernels of this round must leave everything outside their outputs untouched
# ---------------------------------------------------------------------------------------------------------------------
def _guarded(shape, dtype, pad=4096):
    n = 1
    for s_ in shape:
        n *= s_
    buf = torch.full((n + 2 * pad,), 12345.0 if dtype.is_floating_point else 77, dtype=dtype, device="cuda")
    return buf, buf[pad:pad + n].view(*shape), pad


def _untouched(buf, pad, dtype):
    ref = 12345.0 if dtype.is_floating_point else 77
    ref = torch.tensor(ref, dtype=dtype).item()
    return bool((buf[:pad] == ref).all()) and bool((buf[-pad:] == ref).all())


def test_round2_kernels_do_not_write_outside_their_outputs(ops):
    g = torch.Generator().manual_seed(31)
    # (1) split-K convolution with statistics: partial workspace, finish pass, odd sample count
    B, hw, cin, cout = 3, 16, 512, 256
    x = torch.randn(B, hw, hw, cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(cout, 9 * cin, generator=g) / 70).to(torch.bfloat16).cuda()
    obuf, out, pad = _guarded((B, hw, hw, cout), torch.bfloat16)
    sbuf, stats, spad = _guarded((B * hw * hw // 128, cout, 2), torch.float32)
    ops.conv(x, w, None, cout, 3, out_dtype=torch.bfloat16, tensor_core=True, out=out, stats_out=stats)
    torch.cuda.synchronize()
    assert _untouched(obuf, pad, torch.bfloat16) and _untouched(sbuf, spad, torch.float32)
    assert bool(torch.isfinite(out.float()).all()) and bool(torch.isfinite(stats).all()) and float(stats[..., 1].min()) >= 0
    # (2) GroupNorm in the operand path: output + statistics + coefficient table
    t = x.float().reshape(B * hw * hw // 128, 128, cin)
    src = (torch.stack([t.sum(1), (t * t).sum(1)], -1).contiguous(), cin, 1, B * hw * hw // 128, hw * hw // 128, B)
    coef = ops.gn_fold_tiles(src, None, B, coef_for=(torch.ones(cin, device="cuda"), torch.zeros(cin, device="cuda"), 1e-5, hw * hw))
    assert tuple(coef.shape) == (B, cin, 2) and bool(torch.isfinite(coef).all())
    obuf, out, pad = _guarded((B, hw, hw, cout), torch.bfloat16)
    sbuf, stats, spad = _guarded((B * hw * hw // 128, cout, 2), torch.float32)
    ops.conv(x, w, None, cout, 3, out_dtype=torch.bfloat16, tensor_core=True, out=out, stats_out=stats, gn_coef=coef)
    torch.cuda.synchronize()
    assert _untouched(obuf, pad, torch.bfloat16) and _untouched(sbuf, spad, torch.float32)
    # (3) stride-2 convolution
    obuf, out, pad = _guarded((B, hw // 2, hw // 2, cout), torch.float32)
    ops.conv(x, w, None, cout, 3, stride=2, out_dtype=torch.float32, tensor_core=True, out=out)
    torch.cuda.synchronize()
    assert _untouched(obuf, pad, torch.float32) and bool(torch.isfinite(out).all())


def test_attention_tc_wide_ragged_tokens_write_only_their_rows(ops):
    """T = 300 (three query tiles: the CTA pair of the last one has an all-out-of-range partner), output rows >= T of the
    padded buffer and everything around it stay untouched."""
    B, T, C = 2, 300, 512
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(B, T, 3 * C, generator=g) * 0.7).to(torch.bfloat16).cuda()
    import ctypes as Cc
    from stedm_b200 import _lib
    obuf, out, pad = _guarded((B, T, C), torch.bfloat16)
    lib = _lib.load()
    es = 2
    rc = lib.stedm_attention_tc(qkv.data_ptr(), qkv.data_ptr() + C * es, qkv.data_ptr() + 2 * C * es, out.data_ptr(), B, 1, T, C,
                                T * 3 * C, C, 3 * C, C ** -0.5, T * C, 0, 0, 0, 0, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert _untouched(obuf, pad, torch.bfloat16)
    q, k, v = (t_.float().cpu() for t_ in qkv.split(C, dim=-1))
    want = torch.einsum("bts,bsc->btc", torch.softmax(torch.einsum("btc,bsc->bts", q, k) * C ** -0.5, -1), v)
    assert max_abs(out.float().cpu(), want) < 2e-2 * max(1.0, float(want.abs().max()))
