"""Native execution of the style encoder — torchvision's ``swin_v2_t`` with ``head = Linear(768, 512)``, as the
reference builds it in networks/s_zss_dm.py:19-20 and calls it from the aggregation blocks
(networks/agg_blocks.py:24-33, 47-54, 66-75).

The torchvision module keeps OWNING the parameters (reference checkpoints load by name:
``agg_block.embedder.features.1.0.attn.qkv.weight`` ...); this runner repacks them once and executes
SwinTransformer.forward (torchvision/models/swin_transformer.py) as C-ABI kernel launches:

  features.0            stedm_patch_embed_ln         (Conv2d 4x4/4 + LayerNorm, straight from the NHWC style images)
  SwinTransformerBlockV2
    attn.qkv / attn.proj, mlp.0 (+GELU) / mlp.3       stedm_conv_tc (tcgen05, ksize 1)  |  stedm_conv_simt in fp32 mode
    shifted-window cosine attention                   stedm_window_attention
    x = x + norm(.)                                   stedm_layernorm (fp32 residual stream + bf16 GEMM operand copy)
  PatchMergingV2        stedm_patch_merge_gather -> reduction GEMM -> stedm_layernorm
  norm, avgpool, head   stedm_ln_meanpool -> stedm_linear

What is input independent is folded at pack time: the continuous relative position bias
16*sigmoid(cpb_mlp(relative_coords_table))[relative_position_index] per block, exp(clamp(logit_scale, max=log 100)),
and the zeroed k-bias of the qkv Linear (ShiftedWindowAttentionV2.__init__ / shifted_window_attention).
"""
import math

import torch

from . import ops
from .engine import Precision


class PackedLinear:
    """nn.Linear on a token map [B, H, W, K] as a ksize-1 implicit GEMM.  tensor-core: bf16 [N][K]; SIMT: fp32 [K][N]."""

    def __init__(self, weight, bias, prec):
        w = weight.detach().float()
        self.n, self.k = w.shape
        self.tc = prec.tc
        self.weight = w.to(torch.bfloat16).contiguous() if self.tc else w.t().contiguous()
        self.bias = None if bias is None else bias.detach().float().contiguous()
        self.act_dtype = prec.act

    def __call__(self, x, act=ops.ACT_NONE):
        return ops.conv(x, self.weight, self.bias, self.n, 1, out_dtype=self.act_dtype, tensor_core=self.tc, act=act)


class PackedNormLN:
    def __init__(self, ln):
        self.gamma = ln.weight.detach().float().contiguous()
        self.beta = ln.bias.detach().float().contiguous()
        self.eps = ln.eps


def relative_position_bias(attn):
    """ShiftedWindowAttentionV2.get_relative_position_bias: [heads, N, N] fp32 (N = window area)."""
    table = attn.relative_coords_table.float()                                  # [1, 2Wh-1, 2Ww-1, 2]
    l0, l2 = attn.cpb_mlp[0], attn.cpb_mlp[2]
    hid = torch.relu(table @ l0.weight.detach().float().t() + l0.bias.detach().float())
    tbl = (hid @ l2.weight.detach().float().t()).view(-1, attn.num_heads)       # [(2Wh-1)(2Ww-1), heads]
    n = attn.window_size[0] * attn.window_size[1]
    bias = tbl[attn.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1)
    return (16.0 * torch.sigmoid(bias)).contiguous()


class PackedSwinBlock:
    def __init__(self, blk, prec):
        at = blk.attn
        assert list(at.window_size) == [8, 8], "stedm_window_attention is built for swin_v2_t's 8x8 windows"
        self.heads = at.num_heads
        self.shift = int(at.shift_size[0])
        assert at.shift_size[0] == at.shift_size[1]
        qb = at.qkv.bias.detach().float().clone()
        c = qb.numel() // 3
        qb[c:2 * c] = 0                                                         # swin_transformer.py: k bias zeroed
        self.qkv = PackedLinear(at.qkv.weight, qb, prec)
        self.qkv_bias = qb.contiguous()
        self.proj = PackedLinear(at.proj.weight, at.proj.bias, prec)
        self.logit_scale = torch.clamp(at.logit_scale.detach().float(), max=math.log(100.0)).exp().reshape(-1).contiguous()
        self.rel_bias = relative_position_bias(at)
        self.n1, self.n2 = PackedNormLN(blk.norm1), PackedNormLN(blk.norm2)
        self.fc1 = PackedLinear(blk.mlp[0].weight, blk.mlp[0].bias, prec)
        self.fc2 = PackedLinear(blk.mlp[3].weight, blk.mlp[3].bias, prec)
        self.tc = prec.tc

    def __call__(self, x, xb):
        """x: fp32 residual stream; xb: the GEMM operand copy (bf16 in bf16 mode, x itself in fp32 mode)."""
        tc = self.tc
        qkv = self.qkv(xb)
        a = ops.window_attention(qkv, self.logit_scale, self.rel_bias, self.qkv_bias, self.heads, self.shift)
        p = self.proj(a)
        x, b16 = ops.layernorm(p, x, self.n1.gamma, self.n1.beta, self.n1.eps, want_bf16=tc)
        xb = b16 if tc else x
        h = self.fc1(xb, act=ops.ACT_GELU)
        m = self.fc2(h)
        x, b16 = ops.layernorm(m, x, self.n2.gamma, self.n2.beta, self.n2.eps, want_bf16=tc)
        return x, (b16 if tc else x)


class PackedMerge:
    def __init__(self, pm, prec):
        self.reduction = PackedLinear(pm.reduction.weight, None, prec)
        self.norm = PackedNormLN(pm.norm)
        self.tc = prec.tc

    def __call__(self, xb):
        if xb.shape[1] % 2 or xb.shape[2] % 2:
            raise RuntimeError(f"style encoder: odd token map {tuple(xb.shape[1:3])} in PatchMergingV2 is unsupported")
        r = self.reduction(ops.patch_merge_gather(xb))
        x, b16 = ops.layernorm(r, None, self.norm.gamma, self.norm.beta, self.norm.eps, want_bf16=self.tc)
        return x, (b16 if self.tc else x)


class StyleEncoderRunner:
    """SwinTransformer.forward for the swin_v2_t module ``swin``; input NHWC fp32 images [B, P, P, 3] -> [B, 512]."""

    MAX_CHUNK_TOKENS = 1 << 20   # stage-1 tokens per pass (bounds the transient qkv / MLP activations to ~2 GB)

    def __init__(self, swin, precision):
        self.prec = prec = Precision(precision)
        pe = swin.features[0]
        conv, ln = pe[0], pe[2]
        assert tuple(conv.kernel_size) == (4, 4) and tuple(conv.stride) == (4, 4) and conv.in_channels == 3
        # OIHW [E, 3, 4, 4] -> [(dy*4+dx)*3 + c][E]: the order a 4x4 patch has in an NHWC image
        self.pe_w = conv.weight.detach().float().permute(2, 3, 1, 0).reshape(48, -1).contiguous()
        self.pe_b = conv.bias.detach().float().contiguous()
        self.pe_n = PackedNormLN(ln)
        self.stages = []
        for layer in list(swin.features)[1:]:
            if isinstance(layer, torch.nn.Sequential):
                self.stages.append(("blocks", [PackedSwinBlock(b, prec) for b in layer]))
            else:
                self.stages.append(("merge", PackedMerge(layer, prec)))
        self.norm = PackedNormLN(swin.norm)
        self.head_w = swin.head.weight.detach().float().contiguous()
        self.head_b = swin.head.bias.detach().float().contiguous()

    def _forward(self, imgs):
        tc = self.prec.tc
        x, b16 = ops.patch_embed_ln(imgs, self.pe_w, self.pe_b, self.pe_n.gamma, self.pe_n.beta, self.pe_n.eps,
                                    want_bf16=tc)
        xb = b16 if tc else x
        for kind, item in self.stages:
            if kind == "blocks":
                for blk in item:
                    x, xb = blk(x, xb)
            else:
                x, xb = item(xb)
        b, h, w, c = x.shape
        pooled = ops.ln_meanpool(x.view(b, h * w, c), self.norm.gamma, self.norm.beta, self.norm.eps)
        return ops.linear(pooled, self.head_w, self.head_b)

    def __call__(self, imgs):
        if imgs.dim() != 4 or imgs.shape[-1] != 3 or imgs.shape[1] != imgs.shape[2] or imgs.shape[1] % 4:
            raise RuntimeError(f"style encoder expects NHWC images [B, P, P, 3] with P % 4 == 0, got {tuple(imgs.shape)}")
        imgs = imgs.float().contiguous()
        per = (imgs.shape[1] // 4) ** 2
        chunk = max(1, self.MAX_CHUNK_TOKENS // per)
        if imgs.shape[0] <= chunk:
            return self._forward(imgs)
        return torch.cat([self._forward(imgs[i:i + chunk]) for i in range(0, imgs.shape[0], chunk)], 0)
