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

``SetViTRunner`` does the same for ``style_agg=svit`` (networks/vit_set.py sViT): stedm_spt_patchify -> LayerNorm ->
patch GEMM -> stedm_svit_assemble -> depth x [LayerNorm, qkv GEMM, LSA on stedm_attention_tc (masked diagonal, learned
temperature), to_out GEMM + residual, LayerNorm, GELU MLP + residual] -> token mean / cls -> LayerNorm -> Linear.

What is input independent is folded at pack time: the continuous relative position bias
16*sigmoid(cpb_mlp(relative_coords_table))[relative_position_index] per block, exp(clamp(logit_scale, max=log 100)),
and the zeroed k-bias of the qkv Linear (ShiftedWindowAttentionV2.__init__ / shifted_window_attention).
"""
import math

import torch

from . import ops
from .engine import Precision, tc_geometry_ok


class PackedLinear:
    """nn.Linear on a token map [B, H, W, K] as a ksize-1 implicit GEMM.  tensor-core: bf16 [N][K]; SIMT: fp32 [K][N]."""

    def __init__(self, weight, bias, prec):
        w = weight.detach().float()
        self.n, self.k = w.shape
        self.tc = prec.tc
        self.weight = w.to(torch.bfloat16).contiguous() if self.tc else w.t().contiguous()
        self.bias = None if bias is None else bias.detach().float().contiguous()
        self.act_dtype = prec.act
        self._w32 = w if self.tc else None       # source of the CUDA-core twin for row counts that do not tile
        self._simt_weight = None

    def __call__(self, x, act=ops.ACT_NONE, residual=None, out_dtype=None):
        od = out_dtype or self.act_dtype
        if self.tc:
            b, h, w_, c = x.shape
            if not tc_geometry_ok(b, h, w_):
                # a linear layer only needs ROWS: re-view the token matrix as whole 128-row tiles when it has them,
                # otherwise (e.g. a 6 x 6 map with an odd batch) run the same GEMM on CUDA cores
                rows = b * h * w_
                if rows % 128 == 0 or 128 % rows == 0:
                    shape = (1, rows // 128, 128) if rows % 128 == 0 else (1, 1, rows)
                    out = ops.conv(x.view(*shape, c), self.weight, self.bias, self.n, 1, out_dtype=od, tensor_core=True,
                                   split_k=False, act=act,
                                   residual=None if residual is None else residual.view(*shape, self.n))
                    return out.view(b, h, w_, self.n)
                if self._simt_weight is None:
                    self._simt_weight = self._w32.t().contiguous()
                return ops.conv(x, self._simt_weight, self.bias, self.n, 1, out_dtype=od, tensor_core=False, act=act,
                                residual=residual)
        # split_k=False: an image's style feature is bit-identical whatever batch / chunk it is embedded in
        return ops.conv(x, self.weight, self.bias, self.n, 1, out_dtype=od, tensor_core=self.tc, split_k=False, act=act,
                        residual=residual)


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


def _gemm_view(x2d_rows, c, t):
    """View a [rows, c] activation as the NHWC map [1, rows/128, 128, c] the implicit-GEMM kernel tiles over."""
    if x2d_rows % 128 == 0:
        return t.view(1, x2d_rows // 128, 128, c)
    if 128 % x2d_rows == 0:
        return t.view(1, 1, x2d_rows, c)
    return t.view(1, 1, x2d_rows, c)        # does not tile: PackedLinear runs it on the CUDA-core kernel


class SetViTRunner:
    """sViT.forward (vit_set.py:163-208, t_emb = None) for the parameter container stedm_b200.networks.vit_set.sViT;
    input 'b n h w c' fp32 style images -> [b, num_classes]."""

    def __init__(self, svit, precision):
        self.prec = prec = Precision(precision)
        if svit.pool not in ("mean", "cls"):
            raise NotImplementedError("native sViT: pool must be 'mean' or 'cls' (a [b, 512] context for ResBlockStyle)")
        self.pool, self.patch, self.ns, self.np = svit.pool, svit.patch_size, svit.ns, svit.np
        spt = svit.to_patch_embedding.to_patch_tokens
        self.spt_norm = PackedNormLN(spt[1])
        self.spt_lin = PackedLinear(spt[2].weight, spt[2].bias, prec)
        self.dim = spt[2].weight.shape[0]
        self.cls = svit.cls_token.detach().float().reshape(-1).contiguous()
        self.pos = svit.pos_embedding.detach().float().reshape(-1, self.dim).contiguous()
        self.layers = []
        for attn, ff in svit.transformer.layers:
            lsa = attn.fn
            if lsa.dim_head != 64:
                raise NotImplementedError("native sViT: dim_head must be 64 (the LSA default)")
            self.layers.append(dict(
                n1=PackedNormLN(attn.norm), qkv=PackedLinear(lsa.to_qkv.weight, None, prec), heads=lsa.heads,
                scale=float(lsa.temperature.detach().float().exp()),                      # vit_set.py:50
                out=PackedLinear(lsa.to_out[0].weight, lsa.to_out[0].bias, prec),
                n2=PackedNormLN(ff.norm), fc1=PackedLinear(ff.fn.net[0].weight, ff.fn.net[0].bias, prec),
                fc2=PackedLinear(ff.fn.net[3].weight, ff.fn.net[3].bias, prec)))
        self.head_norm = PackedNormLN(svit.mlp_head[0])
        self.head_w = svit.mlp_head[1].weight.detach().float().contiguous()
        self.head_b = svit.mlp_head[1].bias.detach().float().contiguous()

    def _ln(self, x, n):
        tc = self.prec.tc
        f32, b16 = ops.layernorm(x, None, n.gamma, n.beta, n.eps, want_f32=not tc, want_bf16=tc)
        return b16 if tc else f32

    def __call__(self, style_imgs):
        if style_imgs.dim() != 5 or style_imgs.shape[1] != self.ns or style_imgs.shape[-1] != 3:
            raise RuntimeError(f"sViT expects 'b n h w c' images with n = {self.ns}, got {tuple(style_imgs.shape)}")
        tc, dim = self.prec.tc, self.dim
        b = style_imgs.shape[0]
        tok = ops.spt_patchify(style_imgs.float().contiguous(), self.patch)       # fp32 [b, np, patch_dim]
        n_p, pd = tok.shape[1], tok.shape[2]
        if n_p + 2 > self.pos.shape[0]:
            raise RuntimeError("sViT: more patches than positional embeddings (image larger than image_size)")
        a = self._ln(tok, self.spt_norm)
        pe = self.spt_lin(_gemm_view(b * n_p, pd, a))                              # [.., dim]
        T = n_p + 2
        t_pad = (T + 127) // 128 * 128
        x = ops.svit_assemble(pe.view(b, n_p, dim), self.cls, self.pos, t_pad)     # fp32 residual stream [b, t_pad, dim]
        rows = b * t_pad
        for L in self.layers:
            inner = L["heads"] * 64
            qkv = L["qkv"](_gemm_view(rows, dim, self._ln(x, L["n1"]))).view(b, t_pad, 3 * inner)
            if tc:
                o = ops.attention_tc(qkv, qkv, qkv, L["heads"], 64, T, (t_pad * 3 * inner, 64, 3 * inner), L["scale"],
                                     q_off=0, k_off=inner, v_off=2 * inner, mask_diag=True, batch_tokens=t_pad)
            else:
                o = self._attention_fp32(qkv, L, T, t_pad, inner)
            x = L["out"](_gemm_view(rows, inner, o), residual=_gemm_view(rows, dim, x),
                         out_dtype=torch.float32).view(b, t_pad, dim)              # attn(x) + x
            h = L["fc1"](_gemm_view(rows, dim, self._ln(x, L["n2"])), act=ops.ACT_GELU)
            x = L["fc2"](h, residual=_gemm_view(rows, dim, x), out_dtype=torch.float32).view(b, t_pad, dim)
        pooled = ops.token_mean(x, T) if self.pool == "mean" else x[:, 0].contiguous()
        y, _ = ops.layernorm(pooled, None, self.head_norm.gamma, self.head_norm.beta, self.head_norm.eps)
        return ops.linear(y, self.head_w, self.head_b)

    @staticmethod
    def _attention_fp32(qkv, L, T, t_pad, inner):
        """Parity mode: materialised fp32 scores, at most ~1 GiB of them at a time."""
        b = qkv.shape[0]
        chunk = max(1, min(b, (1 << 28) // (L["heads"] * T * T)))
        outs = [ops.attention_simt(qkv[s:s + chunk], qkv[s:s + chunk], qkv[s:s + chunk], L["heads"], 64, T, 0, inner,
                                   2 * inner, 3 * inner, 64, L["scale"], torch.float32, mask_diag=True,
                                   batch_tokens=t_pad) for s in range(0, b, chunk)]
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)
