"""Native execution engine: walks a parameter-holding module tree once, repacks the reference-named weights into
kernel layouts, and runs the eps U-Net / VQ decoder as a sequence of C-ABI kernel launches (stedm_b200.ops).

Two precisions (north_star):
  * ``bf16`` — throughput mode: NHWC bf16 activations, tcgen05/TMEM/TMA implicit-GEMM convolutions
    (stedm_conv_tc), fp32 accumulation, fp32 GroupNorm statistics / softmax / embeddings, fp32 eps output.
  * ``fp32`` — parity mode (1e-4 max-abs eps bar): NHWC fp32 activations, every contraction on the FFMA
    implicit-GEMM kernel (stedm_conv_simt).
Weights stay visible as nn.Parameters under the reference's names; the packed copies here are derived, private,
and rebuilt whenever the owner invalidates them (load_state_dict / .to()).
"""
import math
import os

import torch

from . import ops

PAD_TC = 64   # channel padding granularity of the tensor-core path (one 128-byte K slab of bf16)
FUSE_SKIP = [True]   # fold ResBlock 1x1 skip convolutions into the second 3x3 conv (A/B switch for measurements)
SPLIT_CONCAT = [True]  # decoder conv1([h | skip]): the skip channels shared by cond / uncond convolved once per distinct sample
SHARE_STYLE_CONV = [True]  # ResBlockStyle under guidance: its first convolution once per distinct input, the style embeddings added after
SPLIT_MIN_SHARED = 384  # ... when at least this many input channels are shared (PackedResBlock._split_point); 128 measured: -1.3 % images/s
SPLIT_GN = [os.environ.get("STEDM_SPLIT_GN", "1") != "0"]  # ... and their GroupNorm + SiLU written once per distinct sample (ops.gn_apply_split), whatever their number
PAD_SIMT = 4


class Precision:
    def __init__(self, name):
        if name not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {name!r}")
        self.name = name
        self.tc = name == "bf16"
        self.act = torch.bfloat16 if self.tc else torch.float32
        self.cpad = PAD_TC if self.tc else PAD_SIMT


def _round_up(v, m):
    return (v + m - 1) // m * m


def tc_geometry_ok(b, h, w, x1_batch=0):
    """Can stedm_conv_tc tile this NHWC map?  128 consecutive pixels of the flattened (b, y, x) index must form one TMA
    box {w_t, h_t, b_t}: true for the path's power-of-two maps (the shipped 512^2 / BASELINE's 256^2), false e.g. for a
    96-wide latent.  Maps that do not tile run on the CUDA-core implicit-GEMM kernel instead (same results, far
    slower): the layer stays functional for any size the reference accepts."""
    if w >= 128:
        if w % 128:
            return False
        tb = 1
    else:
        if 128 % w:
            return False
        rows = 128 // w
        if h >= rows:
            if h % rows:
                return False
            tb = 1
        else:
            if rows % h:
                return False
            tb = rows // h
    if x1_batch and x1_batch != b and (x1_batch * h * w) % 128 and tb != 1:
        return False
    return True


class PackedConv:
    """One convolution (nn.Conv2d 3x3/1x1 or nn.Conv1d k=1) repacked for the selected kernel.

    OIHW fp32 (reference layout) ->
      tensor-core: bf16 [Cout_pad][kh*kw*Cin_pad]   (K-major: tap-major, channel-minor)
      SIMT:        fp32 [kh*kw*Cin_pad][Cout]
    ``cin_split`` = (c0, c1) when the input is a two-source concat (each part padded separately)."""

    def __init__(self, weight, bias, prec, *, stride=1, force_simt=False, cin_split=None, cout_pad=None,
                 fold_upsample=False, pad=None):
        self._src = (weight, bias, stride, cin_split)     # for the lazily built CUDA-core twin (odd map sizes)
        self._twin = None
        w = weight.detach()
        if w.dim() == 3:  # Conv1d k=1
            w = w[:, :, :, None]
        cout, cin, kh, kw = w.shape
        assert kh == kw and kh in (1, 3)
        self.ksize, self.stride = kh, stride
        self.tc = prec.tc and not force_simt
        pad = pad or (PAD_TC if self.tc else PAD_SIMT)
        parts = cin_split or (cin,)
        assert sum(parts) == cin
        w = w.permute(0, 2, 3, 1).float()                      # [Cout, kh, kw, Cin]
        chunks, o = [], 0
        for c in parts:
            part = w[..., o:o + c]
            cp = _round_up(c, pad)
            if cp != c:
                part = torch.nn.functional.pad(part, (0, cp - c))
            chunks.append(part)
            o += c
        w = torch.cat(chunks, dim=-1) if len(chunks) > 1 else chunks[0]
        self.cin_pad = w.shape[-1]
        self.cout_real = cout
        self.cout = cout
        b = None if bias is None else bias.detach().float()
        if self.tc:
            cp = cout_pad or _round_up(cout, 16)
            if cp != cout:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, cp - cout))
                if b is not None:
                    b = torch.nn.functional.pad(b, (0, cp - cout))
                self.cout = cp
            self.weight = w.reshape(self.cout, -1).to(torch.bfloat16).contiguous()
            self._split_cache = {}
            self.phase_weights = None
            if fold_upsample:
                # nearest-x2 upsample followed by a 3x3 conv == four 2x2 convs on the low-resolution input, one per
                # output sub-pixel phase (py, px): rows/cols of the 3x3 kernel that read the same source pixel are
                # summed (in fp32, before the bf16 rounding): py=0 -> {-1: w0, 0: w1+w2}, py=1 -> {0: w0+w1, +1: w2}
                assert kh == 3 and len(parts) == 1
                groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
                self.phase_weights = []
                for py in (0, 1):
                    for px in (0, 1):
                        taps = []
                        for ra in groups[py]:
                            for cb in groups[px]:
                                taps.append(sum(w[:, r, c, :] for r in ra for c in cb))      # [Cout, Cin]
                        wp = torch.stack(taps, dim=1)                                          # [Cout, 4, Cin]
                        self.phase_weights.append(wp.reshape(self.cout, -1).to(torch.bfloat16).contiguous())
        else:
            if cout_pad and cout_pad != cout:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, cout_pad - cout))
                if b is not None:
                    b = torch.nn.functional.pad(b, (0, cout_pad - cout))
                self.cout = cout_pad
            self.weight = w.reshape(self.cout, -1).t().contiguous()   # [K][Cout]
        self.bias = None if b is None else b.contiguous()
        self.prec = prec

    def _tile_stats(self, x, want, reps=1):
        """Buffer for the GroupNorm statistics the epilogue folds into this convolution (None if not wanted or the
        128-pixel tiles would straddle samples).  x = the tensor whose pixels index the GEMM rows."""
        b, h, w_, _ = x.shape
        if not want or (h * w_) % 128 != 0 or self.cout % 64 != 0:
            return None, None
        m_tiles = b * h * w_ // 128
        tiles = torch.empty((reps * m_tiles, self.cout, 2), device=x.device, dtype=torch.float32)
        return tiles, (tiles, self.cout, reps, m_tiles, h * w_ // 128, b)

    def split_weights(self, sp):
        """The same convolution as two K ranges, conv([x_lo | x_hi]) = conv_lo(x[..., :sp]) + conv_hi(x[..., sp:]), for
        a split channel ``sp`` (a multiple of the K slab): used when the channels above ``sp`` are shared by the cond /
        uncond halves of a guided batch (PackedResBlock.__call__)."""
        if sp not in self._split_cache:
            assert self.tc and sp % PAD_TC == 0 and 0 < sp < self.cin_pad and not getattr(self, "has_skip", False)
            w = self.weight.view(self.cout, self.ksize * self.ksize, self.cin_pad)
            self._split_cache[sp] = (w[..., :sp].reshape(self.cout, -1).contiguous(),
                                     w[..., sp:].reshape(self.cout, -1).contiguous())
        return self._split_cache[sp]

    def fuse_skip(self, skip):
        """Fold a 1x1 skip convolution on the block input (ResBlock.skip_connection, openaimodel.py:246-256;
        ResnetBlock.nin_shortcut, model.py:104-119) into this convolution: its [Cout][Cin_skip] weights are appended
        along K and its bias added, so out = conv(a) + skip(x) accumulates in one TMEM tile — no separate launch, no
        skip tensor written and re-read as a residual.  Tensor-core path only."""
        assert self.tc and skip.tc and skip.ksize == 1 and skip.cout == self.cout and self.stride == 1
        self.weight = torch.cat([self.weight, skip.weight], dim=1).contiguous()
        if skip.bias is not None:
            self.bias = skip.bias.clone() if self.bias is None else (self.bias + skip.bias).contiguous()
        self.has_skip = True

    def tc_ok(self, x0, x1=None):
        """Does the tensor-core kernel take this call's map (the GEMM rows are the OUTPUT pixels for stride 2)?"""
        b, h, w_, _ = x0.shape
        if self.stride == 2:
            h, w_ = h // 2, w_ // 2
        return tc_geometry_ok(b, h, w_, 0 if x1 is None else x1.shape[0])

    def _simt_twin(self):
        """The same convolution packed for stedm_conv_simt (fp32 weights, the tensor-core path's channel padding)."""
        if self._twin is None:
            weight, bias, stride, cin_split = self._src
            self._twin = PackedConv(weight, bias, self.prec, stride=stride, force_simt=True, cin_split=cin_split,
                                    pad=PAD_TC)
        return self._twin

    def gn_fusable(self, x0, x1=None, skip=None):
        """Can this call run with the GroupNorm (+ SiLU) of its input applied inside the kernel (ops.conv gn_coef)?"""
        if not (self.tc and self.ksize == 3 and self.stride == 1 and self.tc_ok(x0, x1)):
            return False
        b, h, w_, c0 = x0.shape
        cin = c0 + (0 if x1 is None else x1.shape[-1])
        skip_c = 0 if skip is None else skip[0].shape[-1] + (0 if skip[1] is None else skip[1].shape[-1])
        return cin % PAD_TC == 0 and ops.conv_gn_fusable(b, h, w_, cin, self.cout, skip_c)

    def __call__(self, x0, x1=None, emb=None, residual=None, out_dtype=None, upsample=False, out_nchw=False,
                 want_stats=False, skip=None, gn=None):
        """``gn`` = (coef table, silu): x0 / x1 are the raw inputs of a GroupNorm applied in the kernel's operand path
        (only after gn_fusable() said yes)."""
        out_dtype = out_dtype or self.prec.act
        gkw = {} if gn is None else dict(gn_coef=gn[0], gn_silu=gn[1])
        if self.tc and not self.tc_ok(x0, x1):
            assert skip is None and gn is None, "callers check tc_ok() before asking for a fused skip / GroupNorm"
            out = self._simt_twin()(x0, x1, emb=emb, residual=residual, out_dtype=out_dtype, upsample=upsample,
                                    out_nchw=out_nchw)
            out._gn_tiles = None
            return out
        if skip is not None:
            assert self.tc and getattr(self, "has_skip", False) and not upsample and self.stride == 1
            tiles, meta = self._tile_stats(x0, want_stats)
            out = ops.conv(x0, self.weight, self.bias, self.cout, self.ksize, x1=x1, emb=emb, residual=residual,
                           out_dtype=out_dtype, tensor_core=True, stats_out=tiles, skip_x0=skip[0], skip_x1=skip[1], **gkw)
            out._gn_tiles = meta if getattr(out, "_stats_written", False) else None
            return out
        if self.tc:
            assert gn is None or (self.stride == 1 and not upsample and not out_nchw)
            if self.stride == 2:
                # Downsample (openaimodel.py:164-166): the stride-2 gather happens in the TMA unit (boxes traversed with
                # element stride 2 over the full-resolution input) — no im2col tensor is written
                assert x1 is None and self.ksize == 3
                b, h, w_, _ = x0.shape
                tiles, meta = self._tile_stats(x0.new_empty((b, h // 2, w_ // 2, 0)), want_stats)
                out = ops.conv(x0, self.weight, self.bias, self.cout, 3, emb=emb, residual=residual,
                               out_dtype=out_dtype, tensor_core=True, stats_out=tiles, stride=2)
                out._gn_tiles = meta if getattr(out, "_stats_written", False) else None
                return out
            if upsample and self.phase_weights is not None:
                assert x1 is None and emb is None and residual is None and not out_nchw
                b, h, w_, _ = x0.shape
                out = torch.empty((b, 2 * h, 2 * w_, self.cout), device=x0.device, dtype=out_dtype)
                tiles, meta = self._tile_stats(x0, want_stats, reps=4)
                for ph, wp in enumerate(self.phase_weights):
                    ops.conv(x0, wp, self.bias, self.cout, 3, out_dtype=out_dtype, tensor_core=True, out=out,
                             up_phase=ph, stats_out=tiles)
                out._gn_tiles = meta if getattr(out, "_stats_written", False) else None
                return out
            if upsample:
                assert x1 is None
                x0 = ops.upsample_nearest2x(x0)
            tiles, meta = self._tile_stats(x0, want_stats and not out_nchw)
            out = ops.conv(x0, self.weight, self.bias, self.cout, self.ksize, x1=x1, emb=emb, residual=residual,
                           out_dtype=out_dtype, tensor_core=True, out_nchw=out_nchw,
                           cout_store=self.cout_real if out_nchw else 0, stats_out=tiles, **gkw)
            out._gn_tiles = meta if getattr(out, "_stats_written", False) else None
            return out
        assert gn is None
        return ops.conv(x0, self.weight, self.bias, self.cout, self.ksize, x1=x1, emb=emb, residual=residual,
                        out_dtype=out_dtype, stride=self.stride, upsample=upsample, out_nchw=out_nchw,
                        tensor_core=False)


class PackedNorm:
    def __init__(self, gn, eps):
        self.gamma = gn.weight.detach().float().contiguous()
        self.beta = gn.bias.detach().float().contiguous()
        self.eps = eps

    def coefs(self, x0, x1):
        """Per-(sample, channel) (scale, shift) table of this GroupNorm over [x0 | x1] for a consumer convolution that
        normalises in its own operand path; None unless the statistics of every source came out of its producer's
        epilogue (the same condition under which __call__ skips the statistics pass)."""
        t0 = getattr(x0, "_gn_tiles", None)
        t1 = getattr(x1, "_gn_tiles", None) if x1 is not None else None
        if t0 is None or (x1 is not None and t1 is None):
            return None
        return ops.gn_fold_tiles(t0, t1, x0.shape[0], coef_for=(self.gamma, self.beta, self.eps, x0.shape[1] * x0.shape[2]))

    def split(self, x0, x1, sp, silu, stats):
        """__call__ for a concat whose skip half x1 is shared by the halves of a guided batch: (a[..., :sp] for every
        sample, a[..., sp:] once per distinct skip sample), see ops.gn_apply_split."""
        t0 = getattr(x0, "_gn_tiles", None)
        t1 = getattr(x1, "_gn_tiles", None)
        if t0 is not None and t1 is not None:
            folded = ops.gn_fold_tiles(t0, t1, x0.shape[0])
            return ops.gn_apply_split(x0, x1, folded, self.gamma, self.beta, self.eps, silu, sp, n_chunks=1)
        stats = ops.gn_stats(x0, x1, stats)
        return ops.gn_apply_split(x0, x1, stats, self.gamma, self.beta, self.eps, silu, sp)

    def __call__(self, x0, x1, silu, out_dtype, stats):
        t0 = getattr(x0, "_gn_tiles", None)
        t1 = getattr(x1, "_gn_tiles", None) if x1 is not None else None
        if t0 is not None and (x1 is None or t1 is not None):
            # statistics were produced by the epilogues of the convolutions that wrote x0 / x1: fold them per sample
            folded = ops.gn_fold_tiles(t0, t1, x0.shape[0])
            return ops.gn_apply(x0, x1, folded, self.gamma, self.beta, self.eps, silu, out_dtype, n_chunks=1)
        stats = ops.gn_stats(x0, x1, stats)
        return ops.gn_apply(x0, x1, stats, self.gamma, self.beta, self.eps, silu, out_dtype)


class StatsPool:
    """Scratch for the per-chunk GroupNorm statistics: one reusable double buffer, sized for the largest site
    (128 chunks x 32 groups x 2), since each site's statistics are consumed by the very next launch."""

    def __init__(self, n_sites, batch, device):
        self.buf = torch.empty((batch, 128, 32, 2), device=device, dtype=torch.float64)

    def next(self):
        return self.buf


class PackedResBlock:
    """GN-SiLU-conv3x3 (+emb) -> GN-SiLU-conv3x3 -> + skip(x)   (openaimodel.py:268-288, model.py:121-141)."""

    def __init__(self, norm1, conv1, norm2, conv2, skip, eps, prec, cin_split=None):
        self.n1, self.n2 = PackedNorm(norm1, eps), PackedNorm(norm2, eps)
        self.c1 = PackedConv(conv1.weight, conv1.bias, prec, cin_split=cin_split)
        self.c2 = PackedConv(conv2.weight, conv2.bias, prec)
        self.skip = None if skip is None else PackedConv(skip.weight, skip.bias, prec, cin_split=cin_split)
        self.fused_skip = False
        if self.skip is not None and prec.tc and FUSE_SKIP[0]:
            self.c2.fuse_skip(self.skip)          # the skip GEMM rides along the second 3x3 conv's K loop
            self.fused_skip = True
        self.prec = prec

    def _split_point(self, x0, x1):
        """conv1([h | skip]) may run as conv_lo(a[..., :sp]) + conv_hi(a[..., sp:]) with the second term computed once
        per DISTINCT skip sample: every GroupNorm group that lies wholly inside the skip half is normalised with
        statistics of the skip tensor alone, so those channels of the normalised concat are identical for the cond and
        uncond halves of a guided batch.  sp = the first K-slab boundary at or above the end of the group that straddles
        the h | skip boundary (= c0 when no group straddles).  Returns 0 when not applicable or not worth it: the saved
        FLOPs (9 * shared * 2 per output) must outweigh writing and twice re-reading the fp32 partial (12 B per
        output) at ~200 FLOP per HBM byte, with margin."""
        sp = self._shared_from(x0, x1)
        return sp if SPLIT_CONCAT[0] and sp and x0.shape[-1] + x1.shape[-1] - sp >= SPLIT_MIN_SHARED else 0

    def _shared_from(self, x0, x1):
        """First channel (a K-slab boundary) from which the normalised concat [x0 | x1] depends on x1 alone, for an x1 that
        holds fewer distinct samples than x0; 0 when there is no such channel or the tensor-core path does not run."""
        if not (self.c1.tc and x1 is not None and x1.shape[0] < x0.shape[0]):
            return 0
        c0, c1 = x0.shape[-1], x1.shape[-1]
        if c0 % PAD_TC or c1 % PAD_TC or (c0 + c1) % 32 or not self.c1.tc_ok(x0, x1):
            return 0
        cpg = (c0 + c1) // 32
        sp = _round_up(_round_up(c0, cpg), PAD_TC)
        return sp if sp < c0 + c1 else 0

    def _conv1(self, x0, x1, emb, pool):
        """conv1(SiLU(GN([x0 | x1]))) + emb, with the fold / split / fusion choices of the tensor-core path."""
        sp = self._split_point(x0, x1)
        c0 = x0.shape[-1]
        # GroupNorm + SiLU inside the convolution's operand path: needs the producers' tile statistics and a shape the
        # kernel takes (both launches of a split must qualify)
        coef = None
        if self.c1.tc and ops.GN_FUSION[0]:
            if sp:      # shared channels [sp, C) at the skip's batch, the rest at the full batch
                b, h_, w_, _ = x0.shape
                ok = (ops.conv_gn_fusable(x1.shape[0], h_, w_, c0 + x1.shape[-1] - sp, self.c1.cout) and
                      ops.conv_gn_fusable(b, h_, w_, sp, self.c1.cout))
            else:
                ok = self.c1.gn_fusable(x0, x1)
            coef = self.n1.coefs(x0, x1) if ok else None
        if coef is not None:
            if not sp:
                return self.c1(x0, x1, emb=emb, want_stats=True, gn=(coef, True))
            w_lo, w_hi = self.c1.split_weights(sp)
            part = ops.conv(x1[:, :, :, sp - c0:], w_hi, None, self.c1.cout, 3, out_dtype=torch.float32,
                            tensor_core=True, gn_coef=coef, gn_c_off=sp)
            tiles, meta = self.c1._tile_stats(x0, True)
            h = ops.conv(x0, w_lo, self.c1.bias, self.c1.cout, 3, x1=x1[:, :, :, :sp - c0] if sp > c0 else None, emb=emb,
                         residual=part, out_dtype=self.prec.act, tensor_core=True, stats_out=tiles, gn_coef=coef)
            h._gn_tiles = meta if getattr(h, "_stats_written", False) else None
            return h
        sg = self._shared_from(x0, x1) if SPLIT_GN[0] and x0.dtype == torch.bfloat16 and self.prec.act == torch.bfloat16 else 0
        if sg and not sp and (x0.shape[-1] + x1.shape[-1] - sg) * 4 < x0.shape[-1] + x1.shape[-1]:
            sg = 0    # under a quarter of the channels shared (640 = 512 + 128 at 64 x 64: 64 of them): measured no gain
        if sg:
            # the normalised skip-only channels [sg, C) once per distinct skip sample, the rest per sample (sg == sp when
            # the convolution is split too)
            a_lo, a_hi = self.n1.split(x0, x1, sg, True, pool.next())
            if not sp:
                return self.c1(a_lo, a_hi, emb=emb, want_stats=True)
        else:
            a = self.n1(x0, x1, True, self.prec.act, pool.next())
            if not sp:
                return self.c1(a, emb=emb, want_stats=True)
            a_lo, a_hi = a[..., :sp], a[:x1.shape[0], :, :, sp:]
        w_lo, w_hi = self.c1.split_weights(sp)
        # shared channels: one pass over the distinct skip samples (fp32 partial sums, no bias) ...
        part = ops.conv(a_hi, w_hi, None, self.c1.cout, 3, out_dtype=torch.float32, tensor_core=True)
        # ... the rest: + bias + embedding + the partial broadcast as b % bs, statistics of the sum for the next GN
        tiles, meta = self.c1._tile_stats(x0, True)
        h = ops.conv(a_lo, w_lo, self.c1.bias, self.c1.cout, 3, emb=emb, residual=part,
                     out_dtype=self.prec.act, tensor_core=True, stats_out=tiles)
        h._gn_tiles = meta if getattr(h, "_stats_written", False) else None
        return h

    def call_shared_input(self, x0, emb, pool, groups):
        """The block on ``groups`` copies of the SAME input x0 [B] that differ only in their embedding rows emb [groups*B]
        (ResBlockStyle under classifier-free guidance: the trunk is shared, the style vector is not): GroupNorm, SiLU and
        the first convolution run once at batch B (fp32 output, bias included), ``rows_add_emb`` expands that to
        groups*B samples while adding each sample's embedding — conv(a) + bias + emb, openaimodel.py:278-287 — and
        publishing the statistics of the sum; the rest runs at groups*B with x0 as a broadcast residual.  Saves a
        1024 -> 1024 convolution on B samples per guided step (1.5 % of its FLOPs) and the copy of the trunk output."""
        assert self.skip is None and self.c1.tc
        a = self.n1(x0, None, True, self.prec.act, pool.next())
        h_b = self.c1(a, out_dtype=torch.float32)
        h = ops.rows_add_emb(h_b, emb.contiguous(), groups, out_dtype=self.prec.act)
        a2 = self.n2(h, None, True, self.prec.act, pool.next())
        return self.c2(a2, residual=x0, want_stats=True)

    def __call__(self, x0, x1, emb, pool):
        h = self._conv1(x0, x1, emb, pool)
        fused = self.fused_skip and self.c2.tc_ok(h)
        skip = (x0, x1) if fused else None
        coef = None
        if self.c2.tc and ops.GN_FUSION[0] and self.c2.gn_fusable(h, None, skip):
            coef = self.n2.coefs(h, None)
        gn = None if coef is None else (coef, True)
        a = h if gn is not None else self.n2(h, None, True, self.prec.act, pool.next())
        if fused:
            return self.c2(a, want_stats=True, skip=skip, gn=gn)
        if self.skip is not None:
            xs = self.skip(x0, x1)
        else:
            assert x1 is None
            xs = x0
        return self.c2(a, residual=xs, want_stats=True, gn=gn)


class UNetRunner:
    """Executes UNetModel.forward (openaimodel.py:761-806) for the module tree built by
    stedm_b200.ldm.modules.diffusionmodules.openaimodel.UNetModel."""

    def __init__(self, unet, precision):
        self.prec = prec = Precision(precision)
        self.mc = unet.model_channels
        self.in_ch = unet.in_channels
        self.out_ch = unet.out_channels
        self.heads = unet.num_heads
        te = unet.time_embed
        self.te0 = (te[0].weight.detach().float().contiguous(), te[0].bias.detach().float().contiguous())
        self.te2 = (te[2].weight.detach().float().contiguous(), te[2].bias.detach().float().contiguous())
        emb_w, emb_b, self.emb_off = [], [], {}
        off = 0
        self.n_norms = 0

        def add_emb(lin, key):
            nonlocal off
            emb_w.append(lin.weight.detach().float())
            emb_b.append(lin.bias.detach().float())
            self.emb_off[key] = (off, lin.weight.shape[0])
            off += lin.weight.shape[0]

        def pack_res(rb, key, cin_split=None, style=False):
            skip = rb.skip_connection if isinstance(rb.skip_connection, torch.nn.Conv2d) else None
            blk = PackedResBlock(rb.in_layers[0], rb.in_layers[2], rb.out_layers[0], rb.out_layers[3], skip, 1e-5,
                                 prec, cin_split)
            self.n_norms += 2
            if style:
                self.style_emb = (rb.emb_layers[1].weight.detach().float().contiguous(),
                                  rb.emb_layers[1].bias.detach().float().contiguous())
            else:
                add_emb(rb.emb_layers[1], key)
            return blk

        # ---- encoder
        self.enc = []
        chans = []
        for i, blk in enumerate(unet.input_blocks):
            kind = blk.kind
            if kind == "stem":
                conv = blk[0]
                self.stem = PackedConv(conv.weight, conv.bias, prec)
                self.enc.append(("stem", None))
                chans.append(conv.weight.shape[0])
            elif kind == "down":
                conv = blk[0].op
                self.enc.append(("down", PackedConv(conv.weight, conv.bias, prec, stride=2)))
                chans.append(conv.weight.shape[0])
            else:
                self.enc.append(("res", pack_res(blk[0], ("in", i)), ("in", i)))
                chans.append(blk[0].out_channels)
        # ---- middle: ResBlock, ResBlockStyle, AttentionBlock, ResBlock
        mb = unet.middle_block
        self.mid0 = pack_res(mb[0], ("mid", 0))
        self.mid1 = pack_res(mb[1].block, None, style=True)
        att = mb[2]
        self.n_norms += 1
        self.spatial_transformer = None
        if hasattr(att, "transformer_blocks"):      # use_spatial_transformer=True (openaimodel.py:648-652)
            self.spatial_transformer = PackedSpatialTransformer(att, prec)
        else:
            self.att_norm = PackedNorm(att.norm, 1e-5)
            self.att_qkv = PackedConv(att.qkv.weight, att.qkv.bias, prec)
            self.att_proj = PackedConv(att.proj_out.weight, att.proj_out.bias, prec)
        self.mid3 = pack_res(mb[3], ("mid", 3))
        # ---- decoder
        self.dec = []
        ch = mb[3].out_channels
        for i, blk in enumerate(unet.output_blocks):
            skip_ch = chans.pop()
            rb = pack_res(blk[0], ("out", i), cin_split=(ch, skip_ch))
            ch = blk[0].out_channels
            up = None
            if len(blk) > 1:
                conv = blk[1].conv
                up = PackedConv(conv.weight, conv.bias, prec, fold_upsample=True)
            self.dec.append((rb, ("out", i), up))
        self.out_norm = PackedNorm(unet.out[0], 1e-5)
        self.n_norms += 1
        oc = unet.out[2]
        # head: N = 3 (zero-padded to one 16-wide UMMA tile in bf16 mode), written as NCHW fp32 eps directly by the
        # epilogue (fp32 for the CFG std)
        self.head = PackedConv(oc.weight, oc.bias, prec)
        self.emb_w = torch.cat(emb_w, 0).contiguous()
        self.emb_b = torch.cat(emb_b, 0).contiguous()

    # -- embedding path (K8): sinusoid -> time_embed MLP -> all 17 emb_layers in one stacked linear
    def embeddings(self, t, context, uniform_t=False):
        """``uniform_t``: the caller guarantees every sample has the same timestep (a DDIM step): the time-embedding
        MLP and all 17 emb_layers are then evaluated for ONE row and broadcast by the conv epilogues (row stride 0)."""
        if uniform_t:
            t = t[:1]
        temb = ops.timestep_embedding(t, self.mc)
        e = ops.linear(temb, *self.te0)
        e = ops.linear(e, *self.te2, silu_in=True)
        emb_all = ops.linear(e, self.emb_w, self.emb_b, silu_in=True)
        emb_style = ops.linear(context.float().contiguous(), *self.style_emb, silu_in=True)
        return emb_all, emb_style

    def time_embedding_row(self, t_value, device):
        """All 17 emb_layers outputs for ONE timestep value, cached: they depend on the weights and t only, so a DDIM
        loop (the same 50 timesteps for every batch) pays the three embedding launches once per timestep per model,
        not once per step.  The cache lives and dies with this runner (rebuilt whenever the weights change)."""
        cache = self.__dict__.setdefault("_temb_cache", {})
        row = cache.get(t_value)
        if row is None or row.device != device:
            if isinstance(t_value, float) and not t_value.is_integer():     # DPM-Solver's fractional model time
                t = torch.full((1,), t_value, dtype=torch.float32, device=device)
            else:
                t = torch.full((1,), int(t_value), dtype=torch.int64, device=device)
            e = ops.linear(ops.timestep_embedding(t, self.mc), *self.te0)
            e = ops.linear(e, *self.te2, silu_in=True)
            row = cache[t_value] = ops.linear(e, self.emb_w, self.emb_b, silu_in=True)
        return row

    def style_embedding(self, context):
        """ResBlockStyle's emb_layers on the style vectors (openaimodel.py:291-297): constant over a sampling loop."""
        return ops.linear(context.float().contiguous(), *self.style_emb, silu_in=True)

    def __call__(self, x, c_concat, t, context, uniform_t=False, emb=None):
        """x (B,3,L,L) and c_concat (B,3,L,L) NCHW fp32 (the 'hybrid' concat of ddpm.py:1414 is fused into the
        packing kernel), t (B,) int64, context (B,512) -> eps (B,3,L,L) NCHW fp32.

        Guided sampling with a shared encoder trunk: when ``context`` holds G*B rows (G = 2: [cond ; uncond]) for B
        inputs, the stem, all input blocks and middle_block[0] — which depend only on (x, c_concat, t), not on
        the style vector — run ONCE at batch B; from the ResBlockStyle on the network runs
        at batch G*B with the skip tensors broadcast (b % B) by the kernels.  Every op is per-sample, so eps is
        bit-identical to G separate passes while 24.8 % of a pass pair's FLOPs are not executed.  Returns
        (G*B,3,L,L)."""
        prec = self.prec
        B = x.shape[0]
        G = context.shape[0] // B
        assert context.shape[0] == G * B and G >= 1
        # emb = (time_embedding_row(t), style_embedding(context)) precomputed by the sampling loop
        emb_all, emb_style = emb if emb is not None else self.embeddings(t, context, uniform_t)
        pool = StatsPool(self.n_norms, G * B, x.device)
        h = ops.pack_nchw_to_nhwc(x.contiguous(), c_concat, self.stem.cin_pad, prec.act)
        hs = []
        for entry in self.enc:
            if entry[0] == "stem":
                h = self.stem(h, want_stats=True)
            elif entry[0] == "down":
                h = entry[1](h, want_stats=True)
            else:
                h = entry[1](h, None, self._emb_view(emb_all, entry[2]), pool)
            hs.append(h)
        h = self.mid0(h, None, self._emb_view(emb_all, ("mid", 0)), pool)
        shared_style = (G > 1 and prec.tc and SHARE_STYLE_CONV[0] and self.mid1.skip is None and self.mid1.c2.tc_ok(h)
                        and (h.shape[1] * h.shape[2]) % 128 == 0)
        if G > 1 and emb_all.shape[0] > 1:
            emb_all = torch.cat([emb_all] * G, 0)
        if shared_style:
            # cond and uncond still share their input here: the ResBlockStyle's first convolution runs once
            h = self.mid1.call_shared_input(h, emb_style, pool, G)
        else:
            if G > 1:
                tiles = getattr(h, "_gn_tiles", None)
                h = torch.cat([h] * G, 0)
                h._gn_tiles = tiles             # sample b of the copy owns the tile rows of sample b % B
            h = self.mid1(h, None, emb_style, pool)
        # TimestepEmbedSequential hands the context to StyleBlocks only (openaimodel.py:93-101): the transformer runs
        # without one, i.e. both of its attentions are self-attentions
        h = self.spatial_transformer(h, pool) if self.spatial_transformer is not None else self._attention(h, pool)
        h = self.mid3(h, None, self._emb_view(emb_all, ("mid", 3)), pool)
        for rb, key, up in self.dec:
            h = rb(h, hs.pop(), self._emb_view(emb_all, key), pool)
            if up is not None:
                h = up(h, upsample=True, want_stats=True)
        a = self.out_norm(h, None, True, prec.act, pool.next())
        return self.head(a, out_dtype=torch.float32, out_nchw=True)

    def shared_trunk_ok(self, batch, latent_hw):
        """The tensor-core kernel broadcasts the skip source per 128-pixel tile: the smallest skip map times the
        trunk batch must be a whole number of tiles (always true for even batches at latent >= 32)."""
        if not self.prec.tc:
            return True
        n_down = sum(1 for e in self.enc if e[0] == "down")
        smallest = (latent_hw >> n_down) ** 2
        return (batch * smallest) % 128 == 0

    def _emb_view(self, emb_all, key):
        off, n = self.emb_off[key]
        return emb_all[:, off:off + n]

    def _attention(self, x, pool):
        """AttentionBlock (openaimodel.py:340-346) with QKVAttentionLegacy's head-major qkv split (:378-394)."""
        prec = self.prec
        B, H, W, Cc = x.shape
        T, ch = H * W, Cc // self.heads
        a = self.att_norm(x, None, False, prec.act, pool.next())
        qkv = self.att_qkv(a)                                   # [B, H, W, 3C], channels [head][q|k|v][ch]
        scale = 1.0 / math.sqrt(ch)                             # (q*ch^-1/4).(k*ch^-1/4)
        if prec.tc and ops.attention_tc_supported(ch, T):
            o = ops.attention_tc(qkv, qkv, qkv, self.heads, ch, T, (T * 3 * Cc, 3 * ch, 3 * Cc), scale,
                                 q_off=0, k_off=ch, v_off=2 * ch)
        else:
            o = ops.attention_simt(qkv, qkv, qkv, self.heads, ch, T, 0, ch, 2 * ch, 3 * Cc, 3 * ch, scale, prec.act)
        return self.att_proj(o.view(B, H, W, Cc), residual=x, want_stats=True)


class DecoderRunner:
    """VQModelInterface.decode (autoencoder.py:274-282): VQ nearest code -> post_quant_conv -> Decoder
    (model.py:535-568).  Input z NCHW fp32, output image NCHW fp32."""

    def __init__(self, vq, precision):
        self.prec = prec = Precision(precision)
        self.codebook = vq.quantize.embedding.weight.detach().float().contiguous()
        d = vq.decoder
        pq = vq.post_quant_conv
        zc = pq.weight.shape[0]
        self.n_norms = 0
        self.conv_in = PackedConv(d.conv_in.weight, d.conv_in.bias, prec)
        # post_quant_conv: 1x1, 3 -> 3; output channel-padded with zero weights so it feeds conv_in directly
        self.post_quant = PackedConv(pq.weight, pq.bias, prec, force_simt=True, cout_pad=self.conv_in.cin_pad)
        self.zc = zc

        def res(rb):
            self.n_norms += 2
            skip = getattr(rb, "nin_shortcut", None)
            return PackedResBlock(rb.norm1, rb.conv1, rb.norm2, rb.conv2, skip, 1e-6, prec)

        self.mid1 = res(d.mid.block_1)
        at = d.mid.attn_1
        self.att_norm = PackedNorm(at.norm, 1e-6)
        self.n_norms += 1
        # q, k, v 1x1 convs stacked into one GEMM: output channels [q | k | v]
        wq = torch.cat([at.q.weight, at.k.weight, at.v.weight], 0)
        bq = torch.cat([at.q.bias, at.k.bias, at.v.bias], 0)
        self.att_qkv = PackedConv(wq, bq, prec)
        if prec.tc and not ops.attention_tc_supported(at.q.weight.shape[0], 0, heads=1):
            # widths the flash kernel does not take: separate dense q, k and (channel-major) v^T tensors feed two
            # tensor-core GEMMs per sample around a row softmax
            self.att_q = PackedConv(at.q.weight, at.q.bias, prec)
            self.att_k = PackedConv(at.k.weight, at.k.bias, prec)
            self.att_v = PackedConv(at.v.weight, at.v.bias, prec)
        self.att_proj = PackedConv(at.proj_out.weight, at.proj_out.bias, prec)
        self.att_c = at.q.weight.shape[0]
        self.mid2 = res(d.mid.block_2)
        self.levels = []
        for lvl in reversed(range(len(d.up))):
            up = d.up[lvl]
            blocks = [res(b) for b in up.block]
            upc = None
            if hasattr(up, "upsample"):
                upc = PackedConv(up.upsample.conv.weight, up.upsample.conv.bias, prec, fold_upsample=True)
            self.levels.append((blocks, upc))
        self.norm_out = PackedNorm(d.norm_out, 1e-6)
        self.n_norms += 1
        self.conv_out = PackedConv(d.conv_out.weight, d.conv_out.bias, prec)

    def __call__(self, z, force_not_quantize=False):
        prec = self.prec
        z = z.float().contiguous()
        if not force_not_quantize:
            z = ops.vq_nearest(z, self.codebook)
        B = z.shape[0]
        pool = StatsPool(self.n_norms, B, z.device)
        zin = ops.pack_nchw_to_nhwc(z, None, self.post_quant.cin_pad, torch.float32)
        h = self.post_quant(zin, out_dtype=prec.act)
        h = self.conv_in(h, want_stats=True)
        h = self.mid1(h, None, None, pool)
        h = self._attention(h, pool)
        h = self.mid2(h, None, None, pool)
        for blocks, upc in self.levels:
            for rb in blocks:
                h = rb(h, None, None, pool)
            if upc is not None:
                h = upc(h, upsample=True, want_stats=True)
        a = self.norm_out(h, None, True, prec.act, pool.next())
        return self.conv_out(a, out_dtype=torch.float32, out_nchw=True)

    def _attention(self, x, pool):
        """AttnBlock.forward (model.py:178-202): single head, d = C, scale C^-1/2."""
        prec = self.prec
        B, H, W, Cc = x.shape
        T = H * W
        a = self.att_norm(x, None, False, prec.act, pool.next())
        scale = float(Cc) ** -0.5
        if prec.tc and not (tc_geometry_ok(1, H, W) and T % 16 == 0):
            # a map the tensor-core kernels cannot tile (e.g. 96 x 96): materialised attention on CUDA cores
            qkv = self.att_qkv(a)
            outs = [ops.attention_simt(qkv[i:i + 1], qkv[i:i + 1], qkv[i:i + 1], 1, Cc, T, 0, Cc, 2 * Cc, 3 * Cc, Cc, scale,
                                       prec.act) for i in range(B)]
            o = outs[0] if B == 1 else torch.cat(outs, 0)
        elif prec.tc and ops.attention_tc_supported(Cc, T, heads=1):
            # flash attention for the single 512-wide head: q | k | v from ONE 1x1 convolution, no T x T tensor, one
            # launch for the whole batch (stedm_attention_tc's wide kernel)
            qkv = self.att_qkv(a).view(B, T, 3 * Cc)
            o = ops.attention_tc(qkv, qkv, qkv, 1, Cc, T, (T * 3 * Cc, Cc, 3 * Cc), scale, q_off=0, k_off=Cc, v_off=2 * Cc)
        elif prec.tc:
            # d = 512 is too wide for one CTA's TMEM (S + O accumulators), so the decoder attention runs as two
            # tensor-core GEMMs per sample on the implicit-GEMM kernel: S = Q K^T (K as the "weight" [T][C]) into a
            # reused fp32 T x T buffer, a scaled row softmax to bf16, and O = P V (V^T [C][T] written channel-major
            # by the v conv's epilogue as the "weight").
            q = self.att_q(a)
            k = self.att_k(a)
            vt = self.att_v(a, out_dtype=torch.bfloat16, out_nchw=True)          # [B, C, H, W]
            o = torch.empty((B, H, W, Cc), device=x.device, dtype=torch.bfloat16)
            s = torch.empty((1, H, W, T), device=x.device, dtype=torch.float32)
            pb = torch.empty((1, H, W, T), device=x.device, dtype=torch.bfloat16)
            for i in range(B):
                ops.conv(q[i:i + 1], k[i].view(T, Cc), None, T, 1, tensor_core=True, out=s)
                ops.softmax_rows(s.view(T, T), scale, out=pb.view(T, T))
                ops.conv(pb, vt[i].view(Cc, T), None, Cc, 1, tensor_core=True, out=o[i:i + 1])
        else:
            qkv = self.att_qkv(a)                               # [B, H, W, 3C] = [q | k | v]
            # bound the materialised score matrix (B x T x T fp32) to ~2 GiB per chunk
            chunk = max(1, min(B, (1 << 29) // (T * T)))
            outs = []
            for s in range(0, B, chunk):
                q = qkv[s:s + chunk]
                outs.append(ops.attention_simt(q, q, q, 1, Cc, T, 0, Cc, 2 * Cc, 3 * Cc, Cc, scale, prec.act))
            o = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return self.att_proj(o.view(B, H, W, Cc), residual=x, want_stats=True)


class PackedSpatialTransformer:
    """SpatialTransformer.forward (ldm/modules/attention.py:245-261) on an NHWC map: GroupNorm(eps 1e-6) -> proj_in
    (1x1) -> depth x BasicTransformerBlock._forward (:211-215: x = attn1(norm1(x)) + x; x = attn2(norm2(x), context) + x;
    x = ff(norm3(x)) + x with the GEGLU feed-forward) -> proj_out (1x1) + x_in.  Tokens are the pixels of the NHWC map,
    the transformer's residual stream is fp32, GEMM operands are the activation dtype."""

    def __init__(self, st, prec):
        from .style_engine import PackedLinear, PackedNormLN
        self.prec = prec
        self.norm = PackedNorm(st.norm, st.norm.eps)
        self.proj_in = PackedConv(st.proj_in.weight, st.proj_in.bias, prec)
        self.proj_out = PackedConv(st.proj_out.weight, st.proj_out.bias, prec)
        self.heads, self.d = st.n_heads, st.d_head
        self.inner = self.heads * self.d
        self.blocks = []
        for blk in st.transformer_blocks:
            a1, a2 = blk.attn1, blk.attn2
            ent = dict(
                n1=PackedNormLN(blk.norm1), n2=PackedNormLN(blk.norm2), n3=PackedNormLN(blk.norm3),
                qkv1=PackedLinear(torch.cat([a1.to_q.weight, a1.to_k.weight, a1.to_v.weight], 0), None, prec),
                out1=PackedLinear(a1.to_out[0].weight, a1.to_out[0].bias, prec),
                out2=PackedLinear(a2.to_out[0].weight, a2.to_out[0].bias, prec),
                ff1=PackedLinear(blk.ff.net[0].proj.weight, blk.ff.net[0].proj.bias, prec),
                ff2=PackedLinear(blk.ff.net[2].weight, blk.ff.net[2].bias, prec),
                scale1=a1.scale, scale2=a2.scale, ctx_dim=a2.to_k.weight.shape[1])
            if ent["ctx_dim"] == self.inner:     # usable without a context (self-attention): q, k, v in one GEMM
                ent["qkv2"] = PackedLinear(torch.cat([a2.to_q.weight, a2.to_k.weight, a2.to_v.weight], 0), None, prec)
            ent["q2"] = PackedLinear(a2.to_q.weight, None, prec)
            ent["kv2_w"] = torch.cat([a2.to_k.weight, a2.to_v.weight], 0).detach().float().contiguous()
            self.blocks.append(ent)

    def _ln(self, x, n):
        tc = self.prec.tc
        f32, b16 = ops.layernorm(x, None, n.gamma, n.beta, n.eps, want_f32=not tc, want_bf16=tc)
        return b16 if tc else f32

    def _self_attention(self, qkv, T, scale):
        B, inner, d = qkv.shape[0], self.inner, self.d
        q3 = qkv.view(B, T, 3 * inner)
        if self.prec.tc and ops.attention_tc_supported(d, T):
            return ops.attention_tc(q3, q3, q3, self.heads, d, T, (T * 3 * inner, d, 3 * inner), scale,
                                    q_off=0, k_off=inner, v_off=2 * inner)
        return ops.attention_simt(q3, q3, q3, self.heads, d, T, 0, inner, 2 * inner, 3 * inner, d, scale, self.prec.act)

    def _cross_attention(self, q, kv, T, N, scale):
        """q [B, T, inner]; kv [B, N, 2*inner] = [k | v] of the context tokens."""
        B, inner, d = q.shape[0], self.inner, self.d
        if self.prec.tc and ops.attention_tc_supported(d, T):
            return ops.attention_tc(q, kv, kv, self.heads, d, T, (T * inner, d, inner), scale, q_off=0, k_off=0,
                                    v_off=inner, tokens_kv=N, kv_strides=(N * 2 * inner, d, 2 * inner))
        return ops.attention_simt(q, kv, kv, self.heads, d, T, 0, 0, inner, inner, d, scale, self.prec.act,
                                  tokens_kv=N, kv_token_stride=2 * inner)

    def __call__(self, x_in, pool, context=None):
        prec = self.prec
        B, H, W, C = x_in.shape
        T = H * W
        a = self.norm(x_in, None, False, prec.act, pool.next())
        x = self.proj_in(a, out_dtype=torch.float32)                       # fp32 residual stream [B, H, W, inner]
        if context is not None and context.dim() == 2:
            context = context[:, None, :]
        for bi, e in enumerate(self.blocks):
            o = self._self_attention(e["qkv1"](self._ln(x, e["n1"])), T, e["scale1"])
            x = e["out1"](o.view(B, H, W, self.inner), residual=x, out_dtype=torch.float32)
            n2 = self._ln(x, e["n2"])
            if context is None:
                if "qkv2" not in e:
                    raise RuntimeError(f"SpatialTransformer called without a context but context_dim {e['ctx_dim']} != "
                                       f"{self.inner}: the reference fails here too (to_k applied to the tokens)")
                o = self._self_attention(e["qkv2"](n2), T, e["scale2"])
            else:
                N = context.shape[1]
                assert context.shape[0] == B and context.shape[2] == e["ctx_dim"], tuple(context.shape)
                kv = ops.linear(context.reshape(B * N, -1).float().contiguous(), e["kv2_w"], None)
                kv = kv.to(prec.act).view(B, N, 2 * self.inner)
                o = self._cross_attention(e["q2"](n2).view(B, T, self.inner), kv, T, N, e["scale2"])
            x = e["out2"](o.view(B, H, W, self.inner), residual=x, out_dtype=torch.float32)
            g = ops.geglu(e["ff1"](self._ln(x, e["n3"])))
            # the last block hands its tokens to proj_out: written in the GEMM operand dtype directly
            last = bi == len(self.blocks) - 1
            x = e["ff2"](g, residual=x, out_dtype=prec.act if last else torch.float32)
        return self.proj_out(x, residual=x_in, want_stats=True)
