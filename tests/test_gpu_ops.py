"""Per-kernel parity on a real B200: every C-ABI entry point except the tensor-core convolution, against the CPU
oracle / plain torch fp32 restatements of the reference ops on the same seeded inputs."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stedm_oracle as O
from tests.util import max_abs, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from stedm_b200 import _lib, ops as _ops
    assert _lib.load().stedm_device_supported() == 1, "not an sm_100 device"
    return _ops


def _gen(seed=0):
    return torch.Generator().manual_seed(seed)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


# ---------------------------------------------------------------------------------------------- K11
@pytest.mark.parametrize("shape", [(3, 3, 64, 64), (2, 3, 32, 20), (1, 3, 128, 128), (2, 4, 7, 33)])
def test_cfg_ddim_step_guided(ops, shape):
    g = _gen(1)
    e_c, e_u, x = (torch.randn(shape, generator=g) for _ in range(3))
    tab = O.ddim_tables(50)
    for index in (49, 25, 0):
        a_t, a_p, s1m = tab["a_t"][index], tab["a_prev"][index], tab["sqrt_one_minus_a"][index]
        e = O.cfg_combine(e_c, e_u, 1.5)
        want_x, want_p0 = O.ddim_update(x, e, a_t, a_p, 0.0, s1m)
        got_x, got_p0 = ops.cfg_ddim_step(e_c.cuda(), e_u.cuda(), x.cuda(), a_t, a_p, 0.0, s1m, cfg_scale=1.5)
        assert rel_err(got_x, want_x) < 2e-6 and rel_err(got_p0, want_p0) < 2e-6, index


def test_cfg_ddim_step_unguided_and_noise(ops):
    g = _gen(2)
    e_c, x, noise = (torch.randn(2, 3, 64, 64, generator=g) for _ in range(3))
    tab = O.ddim_tables(50, eta=0.5)
    i = 30
    want_x, want_p0 = O.ddim_update(x, e_c, tab["a_t"][i], tab["a_prev"][i], tab["sigma"][i],
                                    tab["sqrt_one_minus_a"][i], noise)
    got_x, got_p0 = ops.cfg_ddim_step(e_c.cuda(), None, x.cuda(), tab["a_t"][i], tab["a_prev"][i], tab["sigma"][i],
                                      tab["sqrt_one_minus_a"][i], noise=noise.cuda())
    assert rel_err(got_x, want_x) < 2e-6 and rel_err(got_p0, want_p0) < 2e-6


# ---------------------------------------------------------------------------------------------- K7
@pytest.mark.parametrize("c0,c1,hw,dt", [(128, 0, 64, torch.float32), (512, 128, 32, torch.float32),
                                         (1024, 512, 16, torch.bfloat16), (2048, 0, 8, torch.bfloat16),
                                         (256, 0, 33, torch.float32)])
def test_group_norm_silu_concat(ops, c0, c1, hw, dt):
    g = _gen(3)
    B = 3
    x0 = torch.randn(B, c0, hw, hw, generator=g) * 2 + 0.5
    x1 = torch.randn(B, c1, hw, hw, generator=g) - 1 if c1 else None
    gamma, beta = torch.randn(c0 + c1, generator=g), torch.randn(c0 + c1, generator=g)
    x0d = nhwc(x0).to(dt).cuda()
    x1d = nhwc(x1).to(dt).cuda() if c1 else None
    full = torch.cat([nchw(x0d.float().cpu())] + ([nchw(x1d.float().cpu())] if c1 else []), 1)
    for silu, eps in ((True, 1e-5), (False, 1e-6)):
        want = F.group_norm(full, 32, gamma, beta, eps)
        want = F.silu(want) if silu else want
        got = ops.group_norm(x0d, x1d, gamma.cuda(), beta.cuda(), eps, silu, torch.float32)
        assert max_abs(nchw(got.cpu()), want) < 2e-5 * max(1.0, float(want.abs().max()))
    got_bf = ops.group_norm(x0d, x1d, gamma.cuda(), beta.cuda(), 1e-5, True, torch.bfloat16)
    want = F.silu(F.group_norm(full, 32, gamma, beta, 1e-5))
    assert rel_err(nchw(got_bf.float().cpu()), want) < 1e-2


def test_group_norm_broadcast_skip(ops):
    """x1 (encoder skip) has half the batch of x0 and is broadcast as b % B1 (batched guidance)."""
    g = _gen(4)
    x0, x1 = torch.randn(4, 64, 16, 16, generator=g), torch.randn(2, 64, 16, 16, generator=g)
    gamma, beta = torch.randn(128, generator=g), torch.randn(128, generator=g)
    want = F.silu(F.group_norm(torch.cat([x0, torch.cat([x1, x1], 0)], 1), 32, gamma, beta, 1e-5))
    got = ops.group_norm(nhwc(x0).cuda(), nhwc(x1).cuda(), gamma.cuda(), beta.cuda(), 1e-5, True, torch.float32)
    assert max_abs(nchw(got.cpu()), want) < 5e-5


@pytest.mark.parametrize("c0,c1,hw,split_c", [(128, 128, 16, 128), (512, 128, 8, 576), (1024, 512, 8, 1088),
                                               (64, 192, 12, 64), (512, 256, 8, 576)])
def test_group_norm_split_for_shared_skip(ops, c0, c1, hw, split_c):
    """gn_apply_split: the skip-only channels of the normalised concat once per distinct skip sample, the rest per
    sample — bit for bit the slices of the one-tensor apply (ragged channel counts: 17 / 46 / 80 items per pixel)."""
    g = _gen(40)
    B, bs = 4, 2
    x0 = nhwc(torch.randn(B, c0, hw, hw, generator=g) * 2 + 0.5).to(torch.bfloat16).cuda()
    x1 = nhwc(torch.randn(bs, c1, hw, hw, generator=g) - 1).to(torch.bfloat16).cuda()
    gamma, beta = torch.randn(c0 + c1, generator=g).cuda(), torch.randn(c0 + c1, generator=g).cuda()
    stats = ops.gn_stats(x0, x1, None)
    full = ops.gn_apply(x0, x1, stats, gamma, beta, 1e-5, True, torch.bfloat16)
    lo, hi = ops.gn_apply_split(x0, x1, stats, gamma, beta, 1e-5, True, split_c)
    assert tuple(lo.shape) == (B, hw, hw, split_c) and tuple(hi.shape) == (bs, hw, hw, c0 + c1 - split_c)
    assert torch.equal(lo, full[..., :split_c])
    assert torch.equal(hi, full[:bs, :, :, split_c:]) and torch.equal(hi, full[bs:, :, :, split_c:])
    with pytest.raises(RuntimeError, match="groups without x0 channels"):
        ops.gn_apply_split(x0, x1, stats, gamma, beta, 1e-5, True, c0 - 8)     # channel c0 - 8 shares a group with x0


# ---------------------------------------------------------------------------------------------- SIMT conv
def _simt_w(w):
    co = w.shape[0]
    return w.permute(0, 2, 3, 1).reshape(co, -1).t().contiguous()


@pytest.mark.parametrize("cin,cout,k,hw,stride,up", [(8, 128, 3, 32, 1, False), (128, 128, 3, 16, 2, False),
                                                     (64, 32, 3, 8, 1, True), (128, 3, 3, 32, 1, False),
                                                     (96, 64, 1, 16, 1, False), (4, 64, 1, 16, 1, False)])
def test_conv_simt(ops, cin, cout, k, hw, stride, up):
    g = _gen(5)
    B = 2
    x = torch.randn(B, cin, hw, hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if up else x
    want = F.conv2d(xin, w, b, stride=stride, padding=k // 2)
    got = ops.conv(nhwc(x).cuda(), _simt_w(w).cuda(), b.cuda(), cout, k, stride=stride, upsample=up,
                   tensor_core=False)
    assert max_abs(nchw(got.cpu()), want) < 2e-5
    got2 = ops.conv(nhwc(x).cuda(), _simt_w(w).cuda(), b.cuda(), cout, k, stride=stride, upsample=up,
                    tensor_core=False, out_nchw=True, out_dtype=torch.float32)
    assert max_abs(got2.cpu(), want) < 2e-5


def test_conv_simt_concat_emb_residual(ops):
    g = _gen(6)
    B, c0, c1, co, hw = 4, 64, 32, 96, 16
    x0, x1 = torch.randn(B, c0, hw, hw, generator=g), torch.randn(2, c1, hw, hw, generator=g)
    w = torch.randn(co, c0 + c1, 3, 3, generator=g) / 30
    b, emb_all, res = torch.randn(co, generator=g), torch.randn(B, 300, generator=g), torch.randn(B, co, hw, hw, generator=g)
    want = F.conv2d(torch.cat([x0, torch.cat([x1, x1], 0)], 1), w, b, padding=1) + emb_all[:, 100:100 + co, None, None] + res
    emb = emb_all.cuda()[:, 100:100 + co]
    got = ops.conv(nhwc(x0).cuda(), _simt_w(w).cuda(), b.cuda(), co, 3, x1=nhwc(x1).cuda(), emb=emb,
                   residual=nhwc(res).cuda(), tensor_core=False)
    assert max_abs(nchw(got.cpu()), want) < 3e-5


# ---------------------------------------------------------------------------------------------- attention (CUDA cores)
@pytest.mark.parametrize("heads,ch,T", [(8, 128, 64), (4, 64, 256), (1, 512, 100)])
def test_attention_simt_legacy_layout(ops, heads, ch, T):
    g = _gen(7)
    B, Cc = 2, heads * ch
    qkv = torch.randn(B, 3 * Cc, T, generator=g)
    q, k, v = qkv.reshape(B * heads, 3 * ch, T).split(ch, dim=1)          # openaimodel.py:387
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
    want = torch.einsum("bts,bcs->bct", w, v).reshape(B, Cc, T)
    buf = qkv.permute(0, 2, 1).contiguous().cuda()                         # [B, T, 3C] token-major
    got = ops.attention_simt(buf, buf, buf, heads, ch, T, 0, ch, 2 * ch, 3 * Cc, 3 * ch, 1 / math.sqrt(ch),
                             torch.float32)
    assert max_abs(got.cpu().permute(0, 2, 1), want) < 2e-5


# ---------------------------------------------------------------------------------------------- K8
def test_timestep_embedding_and_linear(ops):
    t = torch.tensor([981, 481, 1, 0, 999], dtype=torch.long)
    want = O.timestep_embedding(t, 128)
    got = ops.timestep_embedding(t.cuda(), 128)
    assert max_abs(got, want) < 2e-4          # |t*f| up to ~1e3: sin/cos argument rounding
    g = _gen(8)
    for B, K, N, silu in ((5, 128, 512, False), (11, 512, 1000, True), (1, 512, 11392, True)):
        x, w, b = torch.randn(B, K, generator=g), torch.randn(N, K, generator=g) / math.sqrt(K), torch.randn(N, generator=g)
        want = F.linear(F.silu(x) if silu else x, w, b)
        got = ops.linear(x.cuda(), w.cuda(), b.cuda(), silu_in=silu)
        assert max_abs(got, want) < 2e-5


# ---------------------------------------------------------------------------------------------- K12 / K13 / tail / movement
def test_vq_nearest(ops):
    g = _gen(9)
    cb = torch.randn(8192, 3, generator=g) * 64
    z = torch.randn(2, 3, 32, 32, generator=g) * 90
    want_zq, want_idx = O.vq_quantize(z, cb)
    zq, idx = ops.vq_nearest(z.cuda(), cb.cuda(), return_indices=True)
    agree = (idx.cpu().long() == want_idx).float().mean()
    assert agree > 0.999, agree                      # near-ties may resolve differently under fp32 reassociation
    same = (idx.cpu().long() == want_idx).reshape(2, 1, 32, 32).expand(-1, 3, -1, -1)
    assert max_abs(zq.cpu()[same], want_zq[same]) == 0.0
    # exact ties: duplicated codebook rows must resolve to the first index like torch.argmin
    cb2 = torch.cat([cb[:100], cb[:100]], 0)
    _, idx2 = ops.vq_nearest(z.cuda(), cb2.cuda(), return_indices=True)
    assert int(idx2.max()) < 100
    # ragged pixel count (not a multiple of the 4 pixels a warp scans together), 4-channel codebook, NaN pixel -> code 0
    z3 = torch.randn(1, 4, 7, 5, generator=g) * 3
    cb3 = torch.randn(300, 4, generator=g) * 2
    want3, idx3w = O.vq_quantize(z3, cb3)
    zq3, idx3 = ops.vq_nearest(z3.cuda(), cb3.cuda(), return_indices=True)
    assert (idx3.cpu().long() == idx3w).float().mean() > 0.97 and tuple(zq3.shape) == (1, 4, 7, 5)
    zn = z.clone()
    zn[0, :, 3, 4] = float("nan")
    _, idxn = ops.vq_nearest(zn.cuda(), cb.cuda(), return_indices=True)
    assert int(idxn[3 * 32 + 4]) == 0 and int(idxn.max()) < 8192 and int(idxn.min()) >= 0


def test_spatial_rescale(ops):
    seg, _, _ = O.synthetic_batch(3, 128, 1, 5)
    w = torch.randn(3, 2, 1, 1, generator=_gen(10))
    sd = {"cond_stage_model.channel_mapper.weight": w}
    x = seg.permute(0, 3, 1, 2).contiguous()
    want = O.spatial_rescaler(sd, x)
    got = ops.spatial_rescale(x.cuda(), w.reshape(3, 2).contiguous().cuda(), 2)
    assert max_abs(got, want) < 1e-6
    xr = torch.rand(2, 2, 64, 64, generator=_gen(11))
    assert max_abs(ops.spatial_rescale(xr.cuda(), w.reshape(3, 2).contiguous().cuda(), 2), O.spatial_rescaler(sd, xr)) < 1e-6


def test_image_to_uint8(ops):
    img = torch.randn(2, 3, 64, 64, generator=_gen(12)) * 0.8
    img[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.999999])
    want = O.to_uint8(img)
    got = ops.image_to_uint8(img.cuda()).cpu().numpy()
    assert (got == want).all()


def test_movement_kernels(ops):
    g = _gen(13)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(2, 8, 6, 64, generator=g).to(dt)                     # NHWC
        up = ops.upsample_nearest2x(x.cuda()).cpu()
        want = nhwc(F.interpolate(nchw(x.float()), scale_factor=2, mode="nearest")).to(dt)
        assert torch.equal(up, want)
        cols = ops.im2col_3x3_s2(x.cuda()).cpu().float()
        unf = F.unfold(nchw(x.float()), 3, padding=1, stride=2)              # [B, C*9, L] channel-major
        unf = unf.reshape(2, 64, 9, 4, 3).permute(0, 3, 4, 2, 1).reshape(2, 4, 3, 9 * 64)
        assert torch.equal(cols, unf)
    a, b = torch.randn(2, 3, 16, 16, generator=g), torch.randn(2, 3, 16, 16, generator=g)
    p = ops.pack_nchw_to_nhwc(a.cuda(), b.cuda(), 64, torch.bfloat16).cpu().float()
    assert torch.equal(p[..., :6], nhwc(torch.cat([a, b], 1)).to(torch.bfloat16).float()) and float(p[..., 6:].abs().max()) == 0
    y = torch.randn(2, 16, 16, 40, generator=g)
    assert torch.equal(ops.nhwc_to_nchw_f32(y.cuda()).cpu(), nchw(y))


def test_no_cpu_fallback(ops):
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(1, 8), torch.zeros(4, 8), None)


@pytest.mark.parametrize("cols", [256, 1026, 4096, 5000])
@pytest.mark.parametrize("mask", [False, True])
def test_softmax_rows_warp_and_block_kernels(cols, mask):
    """Row softmax (openaimodel.py:392, model.py:189-190) on both kernels (warp per row; block per row for >= 2048
    columns), fp32 in place and bf16 out, with sViT's diagonal mask (vit_set.py:52-54)."""
    from stedm_b200 import ops
    g = torch.Generator().manual_seed(cols)
    rows = 2 * cols if mask else 77
    x = torch.randn(rows, cols, generator=g) * 3
    ref_in = x * 0.37
    if mask:
        ref_in = ref_in.view(2, cols, cols).masked_fill(torch.eye(cols, dtype=torch.bool), float("-inf")).view(rows, cols)
    want = torch.softmax(ref_in, -1)
    period = cols if mask else 0
    got16 = ops.softmax_rows(x.clone().cuda(), 0.37, out=torch.empty(rows, cols, device="cuda", dtype=torch.bfloat16),
                             mask_diag_period=period)
    got32 = ops.softmax_rows(x.clone().cuda(), 0.37, mask_diag_period=period)
    assert float((got32.cpu() - want).abs().max()) < 1e-6
    assert float((got16.float().cpu() - want).abs().max()) < 4e-3


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_geglu(dt):
    from stedm_b200 import ops
    from tests import fake_ops
    x = (torch.randn(37, 5, 2 * 264) * 2).to(dt)
    got = ops.geglu(x.cuda())
    assert tuple(got.shape) == (37, 5, 264)
    want = fake_ops.geglu(x.float())
    err = ((got.float().cpu() - want).abs() / want.abs().clamp_min(1.0)).max()      # bf16: half an ulp of the result
    assert float(err) < (1e-5 if dt == torch.float32 else 5e-3)


@pytest.mark.parametrize("T,N,heads,d", [(256, 10, 8, 128), (64, 1, 4, 64), (300, 77, 2, 64), (128, 200, 2, 128)])
def test_attention_tc_cross_attention(T, N, heads, d):
    """CrossAttention (attention.py:169-193) on the flash kernel: tokens_kv != tokens, k/v in their own buffer."""
    from stedm_b200 import ops
    from tests import fake_ops
    g = torch.Generator().manual_seed(T + N)
    inner = heads * d
    q = torch.randn(2, T, inner, generator=g).to(torch.bfloat16)
    kv = torch.randn(2, N, 2 * inner, generator=g).to(torch.bfloat16)
    scale = d ** -0.5
    want = fake_ops.attention_simt(q, kv, kv, heads, d, T, 0, 0, inner, inner, d, scale, torch.float32,
                                   tokens_kv=N, kv_token_stride=2 * inner)
    got = ops.attention_tc(q.cuda(), kv.cuda(), kv.cuda(), heads, d, T, (T * inner, d, inner), scale, q_off=0, k_off=0,
                           v_off=inner, tokens_kv=N, kv_strides=(N * 2 * inner, d, 2 * inner))
    assert float((got.float().cpu() - want).abs().max()) < 2e-2
    s = ops.attention_simt(q.float().cuda(), kv.float().cuda(), kv.float().cuda(), heads, d, T, 0, 0, inner, inner, d,
                           scale, torch.float32, tokens_kv=N, kv_token_stride=2 * inner)
    assert float((s.cpu() - want).abs().max()) < 1e-4


def test_rows_add_emb_broadcast_and_statistics(ops):
    """out[b] = src[b % B] + emb[b] with the per-tile GroupNorm statistics of the (bf16-rounded) result: the expansion of
    ResBlockStyle's shared first convolution to the cond / uncond halves of a guided batch."""
    g = _gen(41)
    B, G, H, W, C = 3, 2, 16, 16, 256
    src = torch.randn(B, H, W, C, generator=g) * 2
    emb = torch.randn(G * B, C, generator=g)
    out = ops.rows_add_emb(src.cuda(), emb.cuda(), G)
    want = (src.repeat(G, 1, 1, 1) + emb[:, None, None, :]).to(torch.bfloat16)
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == (G * B, H, W, C)
    assert torch.equal(out.cpu(), want)
    tiles, c, reps, m_tiles, tps, nb = out._gn_tiles
    assert (c, reps, m_tiles, tps, nb) == (C, 1, G * B * H * W // 128, H * W // 128, G * B)
    t = want.float().reshape(m_tiles, 128, C)
    assert max_abs(tiles[..., 0].cpu(), t.sum(1)) < 2e-3 and max_abs(tiles[..., 1].cpu(), (t * t).sum(1)) < 2e-2
    out32 = ops.rows_add_emb(src.cuda(), emb.cuda(), G, out_dtype=torch.float32, want_stats=False)
    assert torch.equal(out32.cpu(), src.repeat(G, 1, 1, 1) + emb[:, None, None, :]) and out32._gn_tiles is None
    # ragged: 320 channels (a partial second channel block), 8 x 8 maps (tiles straddle samples, no statistics),
    # 3 x 64 = 192 rows (a partial last tile)
    src = torch.randn(1, 8, 8, 320, generator=g)
    emb = torch.randn(3, 320, generator=g)
    out = ops.rows_add_emb(src.cuda(), emb.cuda(), 3)
    assert out._gn_tiles is None
    assert torch.equal(out.cpu(), (src.repeat(3, 1, 1, 1) + emb[:, None, None, :]).to(torch.bfloat16))
