"""Torch-tensor front end of the C ABI: argument checking, output allocation, current-stream plumbing.

PyTorch is used only for device memory and streams.  Every function launches hand-written sm_100a kernels from
libstedm_b200.so and raises if the library is missing or the tensors are not CUDA tensors — there is no fallback.
Activations are NHWC; ``F32``/``BF16`` tags follow include/stedm_b200.h.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import BF16, F32, ConvDesc

_DT = {torch.float32: F32, torch.bfloat16: BF16}
_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16}

LAUNCHES = [0]  # number of native kernel launches issued through this module (bench.py's gpu_launches)
SPLITK_MAX_PIXELS = 148 * 128 * 4  # above this the output tiles alone fill the persistent grid: never split-K
# Split-K for launches whose output tiles would leave more than half of the SMs idle (small per-GPU batches — the
# reference ships 2-8 images per GPU, conf/location/cluster.yaml:3-5): the K loop is split across the idle SMs into fp32
# partial tiles that a second pass sums IN SPLIT ORDER (deterministic), applying the epilogue and the fused GroupNorm
# tile statistics once.  On by default: the library's planner decides per launch (stedm_conv_tc_workspace_bytes > 0).
# It changes the fp32 summation order as a function of how many tiles a launch has, i.e. of the batch size: results stay
# within fp32 reassociation of the single-pass path, but a sample is bit-identical across batch sizes / rank shards only
# among launches planned the same way (always true for per-rank batches >= 16 at latent 64; tests pin the single-pass
# path with SPLIT_K[0] = False).
SPLIT_K = [os.environ.get("STEDM_SPLIT_K", "1") != "0"]


def enable_split_k(flag=True):
    """Split-K for launches with few output tiles (latency mode for small batches); see SPLIT_K."""
    SPLIT_K[0] = bool(flag)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError("stedm_b200 ops take CUDA tensors only (no CPU fallback)")
            if not t.is_contiguous():
                raise RuntimeError("stedm_b200 ops take contiguous tensors")


def _call(name, *args):
    lib = _lib.load()
    _lib.check(getattr(lib, name)(*args), name)
    LAUNCHES[0] += 1


def tag(dtype):
    return _DT[dtype]


def torch_dtype(tag_):
    return _TORCH_DT[tag_]


# ------------------------------------------------------------------------------------------------ K11
def cfg_ddim_step(e_c, e_u, x, a_t, a_prev, sigma_t, sqrt_one_minus_at, cfg_scale=1.0, phi=0.7, noise=None,
                  out_x_prev=None, out_pred_x0=None):
    """ddim.py:177-209 in one kernel.  NCHW fp32 tensors; ``e_u=None`` -> unguided."""
    _cuda(e_c, e_u, x, noise)
    assert e_c.dtype == torch.float32 and x.dtype == torch.float32 and e_c.shape == x.shape and x.dim() == 4
    b, c, h, w = x.shape
    x_prev = torch.empty_like(x) if out_x_prev is None else out_x_prev
    pred_x0 = torch.empty_like(x) if out_pred_x0 is None else out_pred_x0
    _call("stedm_cfg_ddim_step", _ptr(e_c), _ptr(e_u), _ptr(x), _ptr(noise), _ptr(x_prev), _ptr(pred_x0), b, c, h, w,
          0 if e_u is None else 1, float(cfg_scale), float(phi), float(a_t), float(a_prev), float(sigma_t),
          float(sqrt_one_minus_at), _stream())
    return x_prev, pred_x0


# ------------------------------------------------------------------------------------------------ K7
_GN_CHUNKS = {}


def gn_num_chunks(hw, channels):
    """Number of per-sample partial-statistics chunks (a function of the per-sample shape only)."""
    key = (hw, channels)
    if key not in _GN_CHUNKS:
        n = _lib.load().stedm_gn_num_chunks(hw, channels)
        if n <= 0:
            raise RuntimeError(f"stedm_gn_num_chunks({hw}, {channels}) failed")
        _GN_CHUNKS[key] = n
    return _GN_CHUNKS[key]


def gn_stats(x0, x1, stats=None):
    """Per-chunk (sum, sumsq) per (sample, group) of the concat [x0 | x1]: double [B, n_chunks, 32, 2]."""
    _cuda(x0, x1, stats)
    b, h, w, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[-1]
    x1b = 0 if x1 is None or x1.shape[0] == b else x1.shape[0]
    n = gn_num_chunks(h * w, c0 + c1)
    if stats is None:
        stats = torch.empty((b, n, 32, 2), device=x0.device, dtype=torch.float64)
    assert stats.dtype == torch.float64 and stats.numel() >= b * n * 64
    _call("stedm_gn_stats", _ptr(x0), _ptr(x1), _DT[x0.dtype], b, x1b, h * w, c0, c1, _ptr(stats), _stream())
    return stats


def gn_fold_tiles(src0, src1, batch, out=None, coef_for=None):
    """Fold the per-tile statistics written by the producing convolutions (conv(..., stats_out=...)) into the
    per-sample GroupNorm partials [batch, 1, 32, 2] (double).  src = (tiles fp32 [reps*rep_stride, c, 2], c, reps,
    rep_stride, tiles_per_sample, batch_of_source); src1 = None for a single-source input.

    ``coef_for`` = (gamma, beta, eps, hw): instead of the partials, return the per-(sample, channel) (scale, shift)
    table fp32 [batch, c0+c1, 2] a consumer convolution applies in its own operand path (conv(..., gn_coef=...))."""
    t1, c1, r1, rs1, tps1, b1 = src1 if src1 is not None else (None, 0, 1, 0, 1, 1)
    t0, c0, r0, rs0, tps0, b0 = src0
    coef, gamma, beta, eps, hw = None, None, None, 0.0, 0
    if coef_for is not None:
        gamma, beta, eps, hw = coef_for
        coef = torch.empty((batch, c0 + c1, 2), device=t0.device, dtype=torch.float32)
    elif out is None:
        out = torch.empty((batch, 1, 32, 2), device=t0.device, dtype=torch.float64)
    _cuda(t0, t1, out, gamma, beta)
    _call("stedm_gn_fold_tiles", _ptr(t0), c0, r0, rs0, tps0, b0, _ptr(t1), c1, r1, rs1, tps1, b1, batch, _ptr(out),
          _ptr(gamma), _ptr(beta), float(eps), int(hw), _ptr(coef), _stream())
    return coef if coef_for is not None else out


def gn_apply(x0, x1, stats, gamma, beta, eps, silu, out_dtype, n_chunks=0):
    _cuda(x0, x1, stats, gamma, beta)
    b, h, w, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[-1]
    x1b = 0 if x1 is None or x1.shape[0] == b else x1.shape[0]
    out = torch.empty((b, h, w, c0 + c1), device=x0.device, dtype=out_dtype)
    _call("stedm_gn_apply", _ptr(x0), _ptr(x1), _DT[x0.dtype], b, x1b, h * w, c0, c1, _ptr(stats), n_chunks, _ptr(gamma),
          _ptr(beta), float(eps), 1 if silu else 0, _ptr(out), _DT[out_dtype], _stream())
    return out


def gn_apply_split(x0, x1, stats, gamma, beta, eps, silu, split_c, n_chunks=0):
    """gn_apply for a concat whose second source x1 holds fewer (distinct) samples than x0: returns (lo, hi) with
    lo = channels [0, split_c) for every sample and hi = channels [split_c, C) once per DISTINCT x1 sample — the channels
    from split_c on lie in groups of x1 channels only, so they do not depend on x0.  [lo | hi broadcast] == gn_apply(...)."""
    _cuda(x0, x1, stats, gamma, beta)
    b, h, w, c0 = x0.shape
    bs, c1 = x1.shape[0], x1.shape[-1]
    assert x0.dtype == torch.bfloat16 and x1.dtype == torch.bfloat16 and bs < b
    lo = torch.empty((b, h, w, split_c), device=x0.device, dtype=torch.bfloat16)
    hi = torch.empty((bs, h, w, c0 + c1 - split_c), device=x0.device, dtype=torch.bfloat16)
    _call("stedm_gn_apply_split", _ptr(x0), _ptr(x1), b, bs, h * w, c0, c1, _ptr(stats), n_chunks, _ptr(gamma), _ptr(beta),
          float(eps), 1 if silu else 0, split_c, _ptr(lo), _ptr(hi), _stream())
    return lo, hi


def group_norm(x0, x1, gamma, beta, eps, silu, out_dtype, stats=None):
    """GroupNorm(32) (+SiLU) of the channel concat [x0 | x1]; returns the normalised NHWC tensor."""
    stats = gn_stats(x0, x1, stats)
    return gn_apply(x0, x1, stats, gamma, beta, eps, silu, out_dtype)


# ------------------------------------------------------------------------------------------------ conv
def conv(x0, weight, bias, cout, ksize, *, x1=None, emb=None, residual=None, out_dtype=None, stride=1,
         upsample=False, out_nchw=False, tensor_core=True, out=None, cout_store=0, up_phase=None, stats_out=None,
         act=0, skip_x0=None, skip_x1=None, gn_coef=None, gn_c_off=0, gn_silu=True, split_k=True):
    """Implicit-GEMM convolution of NHWC ``[x0 | x1]``.  ``tensor_core`` selects stedm_conv_tc (bf16 weights
    [cout][k*k*cin]) or stedm_conv_simt (fp32 weights [k*k*cin][cout]).

    ``gn_coef`` (fp32 [samples, C_norm, 2] from gn_fold_tiles(coef_for=...)): x0 / x1 are the RAW inputs of a GroupNorm
    (+ SiLU when ``gn_silu``) whose channel ``gn_c_off`` is this convolution's input channel 0; the normalisation runs
    inside the kernel's operand path (only shapes conv_gn_fusable() accepts)."""
    def _slice_stride(x):
        """A channel slice a[..., lo:hi] of a wider NHWC tensor (same pixels, stride(2) channels apart) -> pixel stride."""
        if x is None or not x.is_cuda or x.is_contiguous():
            return 0
        bb, hh, ww, cc = x.shape
        ps = x.stride(2)
        assert x.stride(3) == 1 and ps >= cc and x.stride(1) == ww * ps and x.stride(0) == hh * ww * ps, x.stride()
        assert tensor_core, "channel-slice inputs are a tensor-core path feature"
        return ps
    x0_pix_stride, x1_pix_stride = _slice_stride(x0), _slice_stride(x1)
    # strided sources were checked above; everything else must be contiguous CUDA memory
    _cuda(None if x0_pix_stride else x0, None if x1_pix_stride else x1, weight, bias, residual, skip_x0, skip_x1, gn_coef)
    if not x0.is_cuda or (x1 is not None and not x1.is_cuda):
        raise RuntimeError("stedm_b200 ops take CUDA tensors only (no CPU fallback)")
    if emb is not None:  # a column slice of the stacked embedding table: rows strided, columns dense
        assert emb.is_cuda and emb.dtype == torch.float32 and emb.stride(1) == 1 and emb.shape[1] == cout
        assert emb.shape[0] in (1, x0.shape[0])          # one row per sample, or one row broadcast to all samples
    b, h, w, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[-1]
    out_dtype = out_dtype or x0.dtype
    uh, uw = (2 * h, 2 * w) if (upsample or up_phase is not None) else (h, w)
    oh, ow = uh // stride, uw // stride
    if out is None:
        shape = (b, cout_store or cout, oh, ow) if out_nchw else (b, oh, ow, cout)
        out = torch.empty(shape, device=x0.device, dtype=out_dtype)
    d = ConvDesc()
    d.x0, d.x1, d.weight, d.bias = _ptr(x0), _ptr(x1), _ptr(weight), _ptr(bias)
    d.emb, d.residual, d.out = _ptr(emb), _ptr(residual), _ptr(out)
    d.stats_out = _ptr(stats_out)
    d.c0, d.c1, d.in_dtype = c0, c1, _DT[x0.dtype]
    d.batch, d.in_h, d.in_w = b, h, w
    d.x1_batch = 0 if x1 is None or x1.shape[0] == b else x1.shape[0]
    d.ksize, d.stride, d.upsample = ksize, stride, 1 if upsample else 0
    d.emb_stride = 0 if (emb is None or emb.shape[0] == 1) else emb.stride(0)
    d.res_dtype = F32 if residual is None else _DT[residual.dtype]
    d.out_dtype, d.out_nchw, d.cout = _DT[out.dtype], 1 if out_nchw else 0, cout
    d.cout_store = cout_store if out_nchw else 0
    d.tap_mode, d.phase = (0, 0) if up_phase is None else (1, up_phase)
    d.act = act
    d.x0_pix_stride, d.x1_pix_stride = x0_pix_stride, x1_pix_stride
    if gn_coef is not None:
        assert tensor_core and gn_coef.dtype == torch.float32 and gn_coef.dim() == 3 and gn_coef.shape[2] == 2
        assert gn_coef.shape[0] >= b and gn_coef.shape[1] >= gn_c_off + c0 + c1, (gn_coef.shape, b, gn_c_off, c0, c1)
        d.gn_coef, d.gn_cstride, d.gn_c_off, d.gn_silu = _ptr(gn_coef), gn_coef.shape[1], gn_c_off, 1 if gn_silu else 0
    if residual is not None and residual.shape[0] != b:   # broadcast residual (b % res_batch), tensor-core path
        assert tensor_core and b % residual.shape[0] == 0 and residual.shape[1:] == (oh, ow, cout)
        d.res_batch = residual.shape[0]
    skip_c = 0
    if skip_x0 is not None:     # fused 1x1 skip convolution over [skip_x0 | skip_x1] (weights appended along K)
        assert tensor_core and skip_x0.shape[1:3] == (h, w) and skip_x0.shape[0] == b and skip_x0.dtype == x0.dtype
        d.skip_x0, d.skip_x1 = _ptr(skip_x0), _ptr(skip_x1)
        d.skip_c0, d.skip_c1 = skip_x0.shape[-1], 0 if skip_x1 is None else skip_x1.shape[-1]
        d.skip_x1_batch = 0 if skip_x1 is None or skip_x1.shape[0] == b else skip_x1.shape[0]
        skip_c = d.skip_c0 + d.skip_c1
    out._stats_written = stats_out is not None
    if tensor_core and split_k and SPLIT_K[0] and gn_coef is None and b * oh * ow <= SPLITK_MAX_PIXELS:
        # few output tiles and a deep K loop: split-K over the idle SMs through a caller-owned workspace (the finish pass
        # applies the epilogue and publishes the GroupNorm tile statistics)
        need = _splitk_workspace_bytes(d)
        if need > 0:
            ws = torch.empty((need,), device=x0.device, dtype=torch.uint8)
            d.workspace, d.workspace_bytes = _ptr(ws), need
    if tensor_core:
        ntaps = 4 if up_phase is not None else ksize * ksize
        assert weight.dtype == torch.bfloat16 and weight.numel() == cout * (ntaps * (c0 + c1) + skip_c), \
            (weight.shape, cout, ksize, c0, c1, skip_c)
        _call("stedm_conv_tc", C.byref(d), _stream())
    else:
        assert weight.dtype == torch.float32 and weight.numel() == cout * ksize * ksize * (c0 + c1), \
            (weight.shape, cout, ksize, c0, c1)
        _call("stedm_conv_simt", C.byref(d), _stream())
    return out


def rows_add_emb(src, emb, groups, out_dtype=torch.bfloat16, want_stats=True):
    """[B, H, W, C] fp32 ``src`` -> [groups * B, H, W, C]: sample b of the output = src[b % B] + emb[b] (per channel), with
    the per-tile GroupNorm statistics of the result (``out._gn_tiles`` as conv(..., stats_out=...) leaves them)."""
    _cuda(src, emb)
    b, h, w, c = src.shape
    assert src.dtype == torch.float32 and emb.dtype == torch.float32 and emb.shape == (groups * b, c) and emb.stride(1) == 1
    out = torch.empty((groups * b, h, w, c), device=src.device, dtype=out_dtype)
    tiles = None
    if want_stats and (h * w) % 128 == 0:
        tiles = torch.empty((groups * b * h * w // 128, c, 2), device=src.device, dtype=torch.float32)
    _call("stedm_rows_add_emb", _ptr(src), b * h * w, _ptr(emb), emb.stride(0), _ptr(out), _DT[out_dtype], groups * b * h * w,
          h * w, c, _ptr(tiles), _stream())
    out._gn_tiles = None if tiles is None else (tiles, c, 1, tiles.shape[0], h * w // 128, groups * b)
    return out


_SPLITK_WS = {}


def _splitk_workspace_bytes(d):
    """stedm_conv_tc_workspace_bytes, cached per launch geometry (the planner is a pure function of these fields)."""
    key = (d.batch, d.in_h, d.in_w, d.c0, d.c1, d.cout, d.ksize, d.stride, d.tap_mode, d.skip_c0 + d.skip_c1 if d.skip_x0 else 0)
    need = _SPLITK_WS.get(key)
    if need is None:
        need = _SPLITK_WS[key] = int(_lib.load().stedm_conv_tc_workspace_bytes(C.byref(d)))
    return need


_GN_FUSABLE = {}
# GroupNorm + SiLU inside the consumer convolution's operand path (stedm_conv_desc.gn_coef).  Bit-identical to the separate
# apply kernel and 13 % fewer launches, but measured 1.4 % SLOWER end to end on the same box (65.5 vs 66.45 images/s,
# tools/gpu_ab_bench.sh, round 2): every activation element is transformed once per output-channel tile and halo row
# (x5 at 1024 channels) inside a power-capped tensor-core kernel, while the tensors of the 16x16 / 32x32 sites it covers
# are L2-resident for the stand-alone kernel anyway.  Opt-in: STEDM_GN_FUSION=1.
GN_FUSION = [os.environ.get("STEDM_GN_FUSION", "0") == "1"]


def conv_gn_fusable(batch, h, w, cin, cout, skip_c=0):
    """Does stedm_conv_tc take a 3x3 convolution of this shape with GroupNorm applied in its operand path?  Asked of
    the library's own planner (stedm_conv_tc_plan with gn_coef set) once per shape.  A capability query: whether the
    engine USES the fused path is GN_FUSION's business."""
    key = (batch, h, w, cin, cout, skip_c)
    ok = _GN_FUSABLE.get(key)
    if ok is None:
        d = ConvDesc()
        d.c0, d.c1, d.in_dtype, d.batch, d.in_h, d.in_w = cin, 0, BF16, batch, h, w
        d.ksize, d.stride, d.cout, d.out_dtype = 3, 1, cout, BF16
        d.skip_c0, d.skip_x0 = skip_c, (1 if skip_c else 0)
        d.gn_coef, d.gn_cstride, d.gn_silu = 1, cin, 1
        out8 = (C.c_int32 * 8)()
        ok = _GN_FUSABLE[key] = _lib.load().stedm_conv_tc_plan(C.byref(d), out8) == 0
    return ok


def gemm_simt(a, b, c, m, n, k, lda, ldb, ldc, b_is_nk, nb, nh, a_strides, b_strides, c_strides, alpha=1.0,
              a_off=0, b_off=0, c_off=0):
    """Batched CUDA-core GEMM on raw views (element offsets/strides); see include/stedm_b200.h."""
    _cuda(a, b, c)
    _call("stedm_gemm_simt", _ptr(a) + a_off * a.element_size(), _ptr(b) + b_off * b.element_size(),
          _ptr(c) + c_off * c.element_size(), _DT[a.dtype], _DT[b.dtype], _DT[c.dtype], m, n, k, lda, ldb, ldc,
          1 if b_is_nk else 0, nb, nh, a_strides[0], a_strides[1], b_strides[0], b_strides[1], c_strides[0],
          c_strides[1], float(alpha), _stream())


def softmax_rows(x, scale=1.0, out=None, mask_diag_period=0):
    """softmax(scale * x) over the last dim of fp32 ``x`` (used as scratch); result in ``out`` (fp32/bf16) or in place.
    ``mask_diag_period`` = T: rows are queries of [.., T, T] score matrices, the diagonal is excluded (sViT LSA)."""
    _cuda(x, out)
    assert x.dtype == torch.float32 and (out is None or out.shape == x.shape)
    cols = x.shape[-1]
    _call("stedm_softmax_rows", _ptr(x), _ptr(out), F32 if out is None else _DT[out.dtype], x.numel() // cols, cols,
          float(scale), int(mask_diag_period), _stream())
    return x if out is None else out


def attention_simt(q_src, k_src, v_src, heads, head_dim, tokens, q_off, k_off, v_off, token_stride, head_stride,
                   scale, out_dtype, mask_diag=False, batch_tokens=None, tokens_kv=None, kv_token_stride=None):
    """fp32-accumulate attention on CUDA cores with a materialised score matrix (parity mode).
    q/k/v live in [B, T, token_stride] buffers at channel offset ``*_off + head*head_stride``; for the U-Net's
    legacy head-major qkv layout head_stride = 3*head_dim, for separate q/k/v tensors head_stride = head_dim."""
    b = q_src.shape[0]
    hs = head_stride
    bt = batch_tokens or tokens        # rows per sample in the q/k/v/out buffers (> tokens when padded)
    # cross-attention: tokens_kv keys / values per sample in their own buffer (row stride kv_token_stride)
    tk = tokens_kv or tokens
    kts = kv_token_stride or token_stride
    kv_sb = (tk if tokens_kv else bt) * kts
    s = torch.empty((b, heads, tokens, tk), device=q_src.device, dtype=torch.float32)
    gemm_simt(q_src, k_src, s, tokens, tk, head_dim, token_stride, kts, tk, True, b, heads,
              (bt * token_stride, hs), (kv_sb, hs), (heads * tokens * tk, tokens * tk),
              alpha=scale, a_off=q_off, b_off=k_off)
    softmax_rows(s, mask_diag_period=tokens if mask_diag else 0)
    out = torch.zeros((b, bt, heads * head_dim), device=q_src.device, dtype=out_dtype)
    gemm_simt(s, v_src, out, tokens, head_dim, tk, tk, kts, heads * head_dim, False, b, heads,
              (heads * tokens * tk, tokens * tk), (kv_sb, hs),
              (bt * heads * head_dim, head_dim), b_off=v_off)
    return out


def attention_tc_supported(head_dim, tokens, heads=None):
    """Shapes the fused tcgen05 attention kernels take (stedm_attention_tc): 64 / 128 channels per head, or ONE 512-wide
    head (the VAE decoder's AttnBlock: the wide kernel splits the output channels over two CTAs)."""
    return head_dim in (64, 128) or (head_dim == 512 and heads == 1)


def attention_tc(q, k, v, heads, head_dim, tokens, strides, scale, q_off=0, k_off=0, v_off=0, mask_diag=False,
                 batch_tokens=None, tokens_kv=0, kv_strides=(0, 0, 0)):
    """Fused tcgen05 flash attention; q/k/v are bf16 views described by element offsets and (b, h, t) strides.
    ``batch_tokens`` > tokens: the output keeps that many (zero) rows per sample; ``mask_diag``: sViT's LSA mask."""
    _cuda(q, k, v)
    b = q.shape[0]
    bt = batch_tokens or tokens
    alloc = torch.empty if bt == tokens else torch.zeros
    out = alloc((b, bt, heads * head_dim), device=q.device, dtype=torch.bfloat16)
    es = 2
    _call("stedm_attention_tc", _ptr(q) + q_off * es, _ptr(k) + k_off * es, _ptr(v) + v_off * es, _ptr(out), b, heads,
          tokens, head_dim, strides[0], strides[1], strides[2], float(scale), bt * heads * head_dim,
          1 if mask_diag else 0, tokens_kv, kv_strides[0], kv_strides[1], kv_strides[2], _stream())
    return out


# ------------------------------------------------------------------------------------------------ movement
def upsample_nearest2x(x):
    _cuda(x)
    b, h, w, c = x.shape
    out = torch.empty((b, 2 * h, 2 * w, c), device=x.device, dtype=x.dtype)
    _call("stedm_upsample_nearest2x", _ptr(x), _ptr(out), _DT[x.dtype], b, h, w, c, _stream())
    return out


def im2col_3x3_s2(x):
    _cuda(x)
    b, h, w, c = x.shape
    out = torch.empty((b, h // 2, w // 2, 9 * c), device=x.device, dtype=x.dtype)
    _call("stedm_im2col_3x3_s2", _ptr(x), _ptr(out), _DT[x.dtype], b, h, w, c, _stream())
    return out


def pack_nchw_to_nhwc(x0, x1, c_pad, out_dtype):
    """[x0 | x1] NCHW fp32 -> NHWC ``out_dtype`` zero-padded to c_pad channels (ddpm.py:1414-1415 concat)."""
    _cuda(x0, x1)
    assert x0.dtype == torch.float32 and (x1 is None or x1.dtype == torch.float32)
    b, c0, h, w = x0.shape
    c1 = 0 if x1 is None else x1.shape[1]
    out = torch.empty((b, h, w, c_pad), device=x0.device, dtype=out_dtype)
    _call("stedm_pack_nchw_to_nhwc", _ptr(x0), c0, _ptr(x1), c1, _ptr(out), _DT[out_dtype], b, h * w, c_pad, _stream())
    return out


def nhwc_to_nchw_f32(x):
    _cuda(x)
    b, h, w, c = x.shape
    out = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
    _call("stedm_nhwc_to_nchw_f32", _ptr(x), _DT[x.dtype], _ptr(out), b, h * w, c, _stream())
    return out


# ------------------------------------------------------------------------------------------------ K8
def timestep_embedding(t, dim):
    """util.py:151-171 for int64 timesteps, or float32 ones (DPM-Solver's fractional model times)."""
    _cuda(t)
    assert t.dtype in (torch.int64, torch.float32)
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=torch.float32)
    name = "stedm_timestep_embedding" if t.dtype == torch.int64 else "stedm_timestep_embedding_f32"
    _call(name, _ptr(t), _ptr(out), t.shape[0], dim, _stream())
    return out


def linear(x, weight, bias, silu_in=False, act_in=None, relu_out=False):
    """g(bias + f(x) @ weight.T): f = SiLU (``silu_in``) / ReLU (``act_in='relu'``), g = ReLU (``relu_out``)."""
    _cuda(x, weight, bias)
    act = (1 if silu_in else 0) | (2 if act_in == "relu" else 0) | (4 if relu_out else 0)
    assert x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 2
    b, k = x.shape
    n = weight.shape[0]
    assert weight.shape[1] == k
    out = torch.empty((b, n), device=x.device, dtype=torch.float32)
    _call("stedm_linear", _ptr(x), _ptr(weight), _ptr(bias), _ptr(out), b, k, n, act, _stream())
    return out


# ------------------------------------------------------------------------------------------------ style encoder
ACT_NONE, ACT_GELU = 0, 1


def patch_embed_ln(img, w, bias, gamma, beta, eps, want_f32=True, want_bf16=False):
    """features.0 of swin_v2_t on NHWC fp32 images [B, P, P, 3]; w fp32 [48][E].  Returns (fp32 | None, bf16 | None)."""
    _cuda(img, w, bias, gamma, beta)
    assert img.dtype == torch.float32 and img.dim() == 4 and img.shape[3] == 3 and img.shape[1] == img.shape[2]
    b, p = img.shape[0], img.shape[1]
    e = w.shape[1]
    assert w.shape[0] == 48 and w.dtype == torch.float32
    shape = (b, p // 4, p // 4, e)
    of = torch.empty(shape, device=img.device, dtype=torch.float32) if want_f32 else None
    ob = torch.empty(shape, device=img.device, dtype=torch.bfloat16) if want_bf16 else None
    _call("stedm_patch_embed_ln", _ptr(img), _ptr(w), _ptr(bias), _ptr(gamma), _ptr(beta), float(eps), _ptr(of),
          _ptr(ob), b, p, 4, e, _stream())
    return of, ob


def layernorm(x, residual, gamma, beta, eps, want_f32=True, want_bf16=False):
    """[residual +] LayerNorm(x) over the last dim; x fp32/bf16, residual fp32.  Returns (fp32 | None, bf16 | None)."""
    _cuda(x, residual, gamma, beta)
    c = x.shape[-1]
    rows = x.numel() // c
    assert residual is None or (residual.dtype == torch.float32 and residual.shape == x.shape)
    of = torch.empty(x.shape, device=x.device, dtype=torch.float32) if want_f32 else None
    ob = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _call("stedm_layernorm", _ptr(x), _DT[x.dtype], _ptr(residual), _ptr(gamma), _ptr(beta), float(eps), _ptr(of),
          _ptr(ob), rows, c, _stream())
    return of, ob


def window_attention(qkv, logit_scale, rel_bias, qkv_bias, heads, shift, window=8):
    """ShiftedWindowAttentionV2 core on qkv [B, H, W, 3*heads*32] -> [B, H, W, heads*32] (same dtype)."""
    _cuda(qkv, logit_scale, rel_bias, qkv_bias)
    b, h, w, c3 = qkv.shape
    hd = c3 // (3 * heads)
    assert c3 == 3 * heads * hd and rel_bias.shape == (heads, window * window, window * window)
    assert logit_scale.numel() == heads and logit_scale.dtype == torch.float32 and rel_bias.dtype == torch.float32
    out = torch.empty((b, h, w, heads * hd), device=qkv.device, dtype=qkv.dtype)
    _call("stedm_window_attention", _ptr(qkv), _DT[qkv.dtype], _ptr(logit_scale), _ptr(rel_bias), _ptr(qkv_bias),
          _ptr(out), b, h, w, heads, hd, window, shift, _stream())
    return out


def patch_merge_gather(x):
    _cuda(x)
    b, h, w, c = x.shape
    out = torch.empty((b, h // 2, w // 2, 4 * c), device=x.device, dtype=x.dtype)
    _call("stedm_patch_merge_gather", _ptr(x), _ptr(out), _DT[x.dtype], b, h, w, c, _stream())
    return out


def ln_meanpool(x, gamma, beta, eps):
    """mean over tokens of LayerNorm(x): fp32 [B, T, C] -> [B, C]."""
    _cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.dim() == 3
    b, t, c = x.shape
    out = torch.empty((b, c), device=x.device, dtype=torch.float32)
    _call("stedm_ln_meanpool", _ptr(x), _ptr(gamma), _ptr(beta), float(eps), _ptr(out), b, t, c, _stream())
    return out


def set_reduce(x, mode):
    """mean ('mean') or max ('max') over dim 1 of fp32 [B, N, F]."""
    _cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 3
    b, n, f = x.shape
    out = torch.empty((b, f), device=x.device, dtype=torch.float32)
    _call("stedm_set_reduce", _ptr(x), _ptr(out), b, n, f, {"mean": 0, "max": 1}[mode], _stream())
    return out


def geglu(x):
    """[..., 2F] = [x | gate] -> [..., F] = x * gelu(gate) (attention.py:37-44)."""
    _cuda(x)
    f = x.shape[-1] // 2
    out = torch.empty(x.shape[:-1] + (f,), device=x.device, dtype=x.dtype)
    _call("stedm_geglu", _ptr(x), _ptr(out), _DT[x.dtype], x.numel() // (2 * f), f, _stream())
    return out


def spt_patchify(style_imgs, patch):
    """SPT patch tokens of 'b n h w c' fp32 style images -> fp32 [b, (P/patch)^2, patch*patch*3*n]."""
    _cuda(style_imgs)
    assert style_imgs.dtype == torch.float32 and style_imgs.dim() == 5 and style_imgs.shape[-1] == 3
    b, ns, p, p2, _ = style_imgs.shape
    assert p == p2 and p % patch == 0
    out = torch.empty((b, (p // patch) ** 2, patch * patch * 3 * ns), device=style_imgs.device, dtype=torch.float32)
    _call("stedm_spt_patchify", _ptr(style_imgs), _ptr(out), b, ns, p, patch, _stream())
    return out


def svit_assemble(patches, cls, pos, t_pad):
    """[cls | t_emb=0 | patches] + pos_embedding -> fp32 [b, t_pad, dim] (rows past the sequence are zero)."""
    _cuda(patches, cls, pos)
    b, n, dim = patches.shape
    assert cls.numel() == dim and pos.shape[-1] == dim and pos.numel() >= (n + 2) * dim and pos.dtype == torch.float32
    out = torch.empty((b, t_pad, dim), device=patches.device, dtype=torch.float32)
    _call("stedm_svit_assemble", _ptr(patches), _DT[patches.dtype], _ptr(cls), _ptr(pos), _ptr(out), b, n, t_pad, dim,
          _stream())
    return out


def token_mean(x, tokens):
    """mean over the first ``tokens`` rows of fp32 [b, t_pad, c]."""
    _cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 3
    b, t_pad, c = x.shape
    out = torch.empty((b, c), device=x.device, dtype=torch.float32)
    _call("stedm_token_mean", _ptr(x), _ptr(out), b, tokens, t_pad, c, _stream())
    return out


# ------------------------------------------------------------------------------------------------ K12 / K13 / tail
def vq_nearest(z, codebook, return_indices=False):
    _cuda(z, codebook)
    assert z.dtype == torch.float32 and codebook.dtype == torch.float32
    b, c, h, w = z.shape
    zq = torch.empty_like(z)
    idx = torch.empty((b * h * w,), device=z.device, dtype=torch.int32) if return_indices else None
    _call("stedm_vq_nearest", _ptr(z), _ptr(codebook), _ptr(zq), _ptr(idx), b, c, h * w, codebook.shape[0], _stream())
    return (zq, idx) if return_indices else zq


def spatial_rescale(seg, weight, n_stages):
    _cuda(seg, weight)
    assert seg.dtype == torch.float32 and weight.dtype == torch.float32
    b, cin, p, p2 = seg.shape
    assert p == p2
    cout = weight.shape[0]
    l = p >> n_stages
    out = torch.empty((b, cout, l, l), device=seg.device, dtype=torch.float32)
    _call("stedm_spatial_rescale", _ptr(seg), _ptr(weight), _ptr(out), b, cin, cout, p, n_stages, _stream())
    return out


def image_to_uint8(img):
    _cuda(img)
    assert img.dtype == torch.float32
    b, c, h, w = img.shape
    out = torch.empty((b, h, w, c), device=img.device, dtype=torch.uint8)
    _call("stedm_image_to_uint8", _ptr(img), _ptr(out), b, c, h * w, _stream())
    return out
