"""Drop-in for the SpatialTransformer family of ``ldm.modules.attention`` (reference attention.py:37-261): GEGLU
feed-forward, CrossAttention, BasicTransformerBlock, SpatialTransformer.

The modules only OWN parameters under the reference's names (``norm``, ``proj_in``, ``transformer_blocks.N.attn1.{to_q,
to_k,to_v,to_out.0}``, ``.attn2.*``, ``.ff.net.0.proj``, ``.ff.net.2``, ``.norm{1,2,3}``, ``proj_out``); execution is
stedm_b200.engine.PackedSpatialTransformer on the C-ABI kernels: GroupNorm, tcgen05 GEMMs (q/k/v fused into one),
flash attention on ``stedm_attention_tc`` (self-attention, and cross-attention over a context of any length through its
separate key/value length and strides), LayerNorm, ``stedm_geglu``.

In the reference U-Net the block sits in ``middle_block[2]`` when ``use_spatial_transformer=True`` and is called
WITHOUT a context (TimestepEmbedSequential passes context only to StyleBlocks, openaimodel.py:93-101), so its second
attention is a self-attention too and ``context_dim`` must equal the block width; ``forward(x, context)`` with a
``(B, N, context_dim)`` context is the stand-alone module API of attention.py:245-261.
"""
import torch
from torch import nn


class GEGLU(nn.Module):                     # attention.py:37-44
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):               # attention.py:47-63 (glu=True as BasicTransformerBlock builds it)
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Identity(), nn.Linear(inner, dim))


class CrossAttention(nn.Module):            # attention.py:152-193
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Identity())


class BasicTransformerBlock(nn.Module):     # attention.py:196-215
    def __init__(self, dim, n_heads, d_head, context_dim=None):
        super().__init__()
        self.attn1 = CrossAttention(dim, heads=n_heads, dim_head=d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim=context_dim, heads=n_heads, dim_head=d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)


class SpatialTransformer(nn.Module):        # attention.py:218-261
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, precision="bf16"):
        super().__init__()
        self.in_channels, self.n_heads, self.d_head, self.context_dim = in_channels, n_heads, d_head, context_dim
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, n_heads, d_head, context_dim)
                                                 for _ in range(depth)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)
        nn.init.zeros_(self.proj_out.weight)                                   # zero_module, attention.py:241-245
        nn.init.zeros_(self.proj_out.bias)
        self.precision = precision
        self._runner = None

    def _apply(self, fn, *a, **k):
        self._runner = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._runner = None
        return super().load_state_dict(*a, **k)

    def set_precision(self, precision):
        if precision != self.precision:
            self.precision, self._runner = precision, None

    def runner(self):
        ver = tuple(p._version for p in self.parameters())
        if self._runner is None or self._runner_versions != ver:
            from ...engine import PackedSpatialTransformer, Precision
            if not self.proj_in.weight.is_cuda:
                raise RuntimeError("SpatialTransformer runs only on a CUDA (sm_100a) device — there is no CPU path")
            self._runner, self._runner_versions = PackedSpatialTransformer(self, Precision(self.precision)), ver
        return self._runner

    @torch.no_grad()
    def forward(self, x, context=None):
        """x (B, C, H, W) NCHW fp32, context None | (B, N, context_dim) | (B, context_dim) -> (B, C, H, W) fp32."""
        from ... import ops
        from ...engine import Precision, StatsPool
        r = self.runner()
        prec = Precision(self.precision)
        h = ops.pack_nchw_to_nhwc(x.float().contiguous(), None, x.shape[1], prec.act)
        out = r(h, StatsPool(1, x.shape[0], x.device), context)
        return ops.nhwc_to_nchw_f32(out)
