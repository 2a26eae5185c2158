"""A torch-on-CPU stand-in for stedm_b200.ops, used ONLY by tests/test_engine_glue.py to check the host-side engine
logic (block order, concat order, embedding offsets, weight repacking, padding) against the oracle without a GPU.
It honours the kernels' data contracts: NHWC activations, packed weight layouts, NCHW fp32 at the boundary."""
import math

import torch
import torch.nn.functional as F

LAUNCHES = [0]


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _bcast(x1, b):
    if x1 is None or x1.shape[0] == b:
        return x1
    return x1.repeat(b // x1.shape[0], 1, 1, 1)


def pack_nchw_to_nhwc(x0, x1, c_pad, out_dtype):
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    x = F.pad(x, (0, 0, 0, 0, 0, c_pad - x.shape[1]))
    return _nhwc(x).to(out_dtype)


def nhwc_to_nchw_f32(x):
    return _nchw(x).float().contiguous()


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    a = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(a), torch.sin(a)], -1)


def linear(x, w, b, silu_in=False, act_in=None, relu_out=False):
    x = F.silu(x) if silu_in else (F.relu(x) if act_in == "relu" else x)
    y = F.linear(x, w, b)
    return F.relu(y) if relu_out else y


def gn_stats(x0, x1, stats=None):
    return None


def rows_add_emb(src, emb, groups, out_dtype=torch.bfloat16, want_stats=True):
    out = (src.float().repeat(groups, 1, 1, 1) + emb[:, None, None, :]).to(out_dtype)
    out._gn_tiles = None
    return out


GN_FUSION = [True]


def conv_gn_fusable(batch, h, w, cin, cout, skip_c=0):
    """Mirror of the kernel's rule (stedm_conv_tc_plan with gn_coef set): 256-wide channel tiles, whole-row pixel tiles
    of one sample with room for the halo rows, more than one pixel tile."""
    if not GN_FUSION[0] or cout % 256 or w < 8 or w > 32 or 128 % w:
        return False
    th = 128 // w
    return th >= 2 and h % th == 0 and h >= th + 2 and batch * h * w > 128


def gn_fold_tiles(src0, src1, batch, out=None, coef_for=None):
    """The stand-in keeps no tile sums: the producers' outputs ride along on the tile buffers (conv sets ``_fake_src``)
    and the per-(sample, channel) coefficients are computed from them directly."""
    if coef_for is None:
        return None
    gamma, beta, eps, hw = coef_for
    x = src0[0]._fake_src.float()
    if src1 is not None:
        x = torch.cat([x, _bcast(src1[0]._fake_src.float(), x.shape[0])], -1)
    b, c = x.shape[0], x.shape[-1]
    assert b == batch and x.shape[1] * x.shape[2] == hw
    g = x.reshape(b, -1, 32, c // 32)
    mean = g.mean(dim=(1, 3))
    rstd = (g.var(dim=(1, 3), unbiased=False) + eps).rsqrt()
    a = rstd.repeat_interleave(c // 32, 1) * gamma[None]
    sh = beta[None] - mean.repeat_interleave(c // 32, 1) * a
    return torch.stack([a, sh], -1).contiguous()


def gn_apply(x0, x1, stats, gamma, beta, eps, silu, out_dtype, n_chunks=0):
    x = x0 if x1 is None else torch.cat([x0, _bcast(x1, x0.shape[0])], -1)
    y = F.group_norm(_nchw(x.float()), 32, gamma, beta, eps)
    return _nhwc(F.silu(y) if silu else y).to(out_dtype)


def gn_apply_split(x0, x1, stats, gamma, beta, eps, silu, split_c, n_chunks=0):
    full = gn_apply(x0, x1, stats, gamma, beta, eps, silu, x0.dtype, n_chunks)
    return full[..., :split_c].contiguous(), full[:x1.shape[0], :, :, split_c:].contiguous()


def upsample_nearest2x(x):
    return _nhwc(F.interpolate(_nchw(x.float()), scale_factor=2, mode="nearest")).to(x.dtype)


def im2col_3x3_s2(x):
    b, h, w, c = x.shape
    u = F.unfold(_nchw(x.float()), 3, padding=1, stride=2).reshape(b, c, 9, h // 2, w // 2)
    return u.permute(0, 3, 4, 2, 1).reshape(b, h // 2, w // 2, 9 * c).to(x.dtype)


def conv(x0, weight, bias, cout, ksize, *, x1=None, emb=None, residual=None, out_dtype=None, stride=1,
         upsample=False, out_nchw=False, tensor_core=True, out=None, cout_store=0, up_phase=None, stats_out=None,
         act=0, skip_x0=None, skip_x1=None, gn_coef=None, gn_c_off=0, gn_silu=True, split_k=True):
    y = _conv(x0, weight, bias, cout, ksize, x1=x1, emb=emb, residual=residual, out_dtype=out_dtype, stride=stride,
              upsample=upsample, out_nchw=out_nchw, tensor_core=tensor_core, out=out, cout_store=cout_store,
              up_phase=up_phase, act=act, skip_x0=skip_x0, skip_x1=skip_x1, gn_coef=gn_coef, gn_c_off=gn_c_off,
              gn_silu=gn_silu)
    if stats_out is not None and not out_nchw:
        stats_out._fake_src = y          # what the epilogue's tile statistics describe
        y._stats_written = True
    return y


def _conv(x0, weight, bias, cout, ksize, *, x1=None, emb=None, residual=None, out_dtype=None, stride=1,
          upsample=False, out_nchw=False, tensor_core=True, out=None, cout_store=0, up_phase=None,
          act=0, skip_x0=None, skip_x1=None, gn_coef=None, gn_c_off=0, gn_silu=True):
    x = x0 if x1 is None else torch.cat([x0, _bcast(x1, x0.shape[0])], -1)
    cin = x.shape[-1]
    if gn_coef is not None:   # GroupNorm (+ SiLU) in the operand path: fp32 math, operand rounded to the input dtype
        assert tensor_core and ksize == 3 and conv_gn_fusable(x.shape[0], x.shape[1], x.shape[2], cin, cout)
        co = gn_coef[:x.shape[0], gn_c_off:gn_c_off + cin]
        xn = x.float() * co[:, None, None, :, 0] + co[:, None, None, :, 1]
        x = (F.silu(xn) if gn_silu else xn).to(x0.dtype)
    y_skip = None
    if skip_x0 is not None:    # fused 1x1 skip conv: its weights are the trailing K columns
        sk = skip_x0 if skip_x1 is None else torch.cat([skip_x0, _bcast(skip_x1, skip_x0.shape[0])], -1)
        kmain = ksize * ksize * cin
        wfull = weight.float().reshape(cout, -1)
        y_skip = F.conv2d(_nchw(sk.float()), wfull[:, kmain:].reshape(cout, sk.shape[-1], 1, 1))
        weight = wfull[:, :kmain]
    if up_phase is not None:   # one 2x2 sub-pixel phase: taps (a,b) read (y+a-1+py, x+b-1+px); write (2y+py, 2x+px)
        py, px = up_phase >> 1, up_phase & 1
        w = weight.float().reshape(cout, 2, 2, cin).permute(0, 3, 1, 2)
        xin = F.pad(_nchw(x.float()), (1 - px, px, 1 - py, py))
        y = F.conv2d(xin, w, bias)
        out[:, py::2, px::2, :] = _nhwc(y).to(out.dtype)
        return out
    if tensor_core:   # [Cout][k*k*Cin]
        w = weight.float().reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2)
    else:             # [k*k*Cin][Cout]
        w = weight.t().reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2)
    xin = _nchw(x.float())
    if upsample:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    y = F.conv2d(xin, w, bias, stride=stride, padding=ksize // 2)
    if y_skip is not None:
        y = y + y_skip
    if emb is not None:
        y = y + emb[:, :, None, None]        # (1, C) broadcasts like the kernel's row stride 0
    if act == 1:
        y = F.gelu(y)
    if residual is not None:
        r = _nchw(residual.float())
        y = y + (r if r.shape[0] == y.shape[0] else r.repeat(y.shape[0] // r.shape[0], 1, 1, 1))
    if out_nchw:
        return y[:, :cout_store or cout].contiguous()
    return _nhwc(y).to(out_dtype or x0.dtype)


def attention_tc_supported(head_dim, tokens, heads=None):
    return False


def geglu(x):
    a, g = x.float().chunk(2, dim=-1)
    return (a * F.gelu(g)).to(x.dtype)


def attention_simt(q_src, k_src, v_src, heads, head_dim, tokens, q_off, k_off, v_off, token_stride, head_stride,
                   scale, out_dtype, mask_diag=False, batch_tokens=None, tokens_kv=None, kv_token_stride=None):
    b = q_src.shape[0]
    bt = batch_tokens or tokens
    tk = tokens_kv or tokens
    kts = kv_token_stride or token_stride
    flat = lambda t: t.reshape(b, bt, token_stride).float()[:, :tokens]
    flat_kv = (lambda t: t.reshape(b, tk, kts).float()) if tokens_kv else (lambda t: t.reshape(b, bt, kts).float()[:, :tk])
    outs = []
    for h in range(heads):
        q = flat(q_src)[:, :, q_off + h * head_stride: q_off + h * head_stride + head_dim]
        k = flat_kv(k_src)[:, :, k_off + h * head_stride: k_off + h * head_stride + head_dim]
        v = flat_kv(v_src)[:, :, v_off + h * head_stride: v_off + h * head_stride + head_dim]
        s = torch.einsum("btc,bsc->bts", q, k) * scale
        if mask_diag:
            s = s.masked_fill(torch.eye(tokens, dtype=torch.bool), float("-inf"))
        outs.append(torch.einsum("bts,bsc->btc", torch.softmax(s, -1), v))
    o = torch.cat(outs, -1)
    if bt != tokens:
        o = F.pad(o, (0, 0, 0, bt - tokens))
    return o.to(out_dtype)


def attention_tc(q, k, v, heads, head_dim, tokens, strides, scale, q_off=0, k_off=0, v_off=0, mask_diag=False,
                 batch_tokens=None, tokens_kv=0, kv_strides=(0, 0, 0)):
    assert strides[1] == head_dim and k is v
    return attention_simt(q, k, v, heads, head_dim, tokens, q_off, k_off, v_off, strides[2], strides[1], scale,
                          torch.bfloat16, mask_diag=mask_diag, batch_tokens=batch_tokens,
                          tokens_kv=tokens_kv or None, kv_token_stride=kv_strides[2] or None)


def vq_nearest(z, codebook, return_indices=False):
    b, c, h, w = z.shape
    zf = z.permute(0, 2, 3, 1).reshape(-1, c)
    d = (zf ** 2).sum(1, keepdim=True) + (codebook ** 2).sum(1) - 2.0 * zf @ codebook.t()
    idx = torch.argmin(d, 1)
    zq = codebook[idx].reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    return (zq, idx.int()) if return_indices else zq


# ---------------------------------------------------------------------------------------------- style encoder
ACT_NONE, ACT_GELU = 0, 1


def _two(y, want_f32, want_bf16):
    return (y.float() if want_f32 else None), (y.to(torch.bfloat16) if want_bf16 else None)


def patch_embed_ln(img, w, bias, gamma, beta, eps, want_f32=True, want_bf16=False):
    b, p = img.shape[0], img.shape[1]
    t = p // 4
    patches = img.reshape(b, t, 4, t, 4, 3).permute(0, 1, 3, 2, 4, 5).reshape(b, t, t, 48)     # (dy, dx, c) order
    y = F.layer_norm(patches @ w + bias, (w.shape[1],), gamma, beta, eps)
    return _two(y, want_f32, want_bf16)


def layernorm(x, residual, gamma, beta, eps, want_f32=True, want_bf16=False):
    y = F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps)
    if residual is not None:
        y = y + residual
    return _two(y, want_f32, want_bf16)


def window_attention(qkv, logit_scale, rel_bias, qkv_bias, heads, shift, window=8):
    """Independent restatement (roll / partition / mask built the torchvision way) of what the kernel folds into indexing."""
    b, h, w, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    x = qkv.float()
    ph, pw = -(-h // window) * window, -(-w // window) * window
    if (ph, pw) != (h, w):   # padded tokens carry the qkv bias (zero input through the Linear)
        full = qkv_bias.float().expand(b, ph, pw, c3).clone()
        full[:, :h, :w] = x
        x = full
    sy, sx = (0 if ph <= window else shift), (0 if pw <= window else shift)
    if sy or sx:
        x = torch.roll(x, (-sy, -sx), (1, 2))
    nw = (ph // window) * (pw // window)
    x = x.view(b, ph // window, window, pw // window, window, c3).permute(0, 1, 3, 2, 4, 5).reshape(b * nw, window * window, c3)
    q, k, v = x.reshape(b * nw, window * window, 3, heads, d).permute(2, 0, 3, 1, 4)
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1) * logit_scale.view(1, heads, 1, 1)
    attn = attn + rel_bias[None]
    if sy or sx:
        m = torch.zeros(ph, pw)
        cnt = 0
        for hs in ((0, -window), (-window, -sy), (-sy, None)):
            for ws_ in ((0, -window), (-window, -sx), (-sx, None)):
                m[hs[0]:hs[1], ws_[0]:ws_[1]] = cnt
                cnt += 1
        m = m.view(ph // window, window, pw // window, window).permute(0, 2, 1, 3).reshape(nw, window * window)
        m = m[:, None, :] - m[:, :, None]
        m = torch.where(m != 0, torch.tensor(-100.0), torch.tensor(0.0))
        attn = (attn.view(b, nw, heads, window * window, window * window) + m[None, :, None]).view(-1, heads, window * window, window * window)
    o = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(b * nw, window * window, c)
    o = o.view(b, ph // window, pw // window, window, window, c).permute(0, 1, 3, 2, 4, 5).reshape(b, ph, pw, c)
    if sy or sx:
        o = torch.roll(o, (sy, sx), (1, 2))
    return o[:, :h, :w].contiguous().to(qkv.dtype)


def patch_merge_gather(x):
    return torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).contiguous()


def ln_meanpool(x, gamma, beta, eps):
    return F.layer_norm(x, (x.shape[-1],), gamma, beta, eps).mean(1)


def set_reduce(x, mode):
    return x.mean(1) if mode == "mean" else x.max(1)[0]


def spt_patchify(style_imgs, patch):
    b, ns, p, _, _ = style_imgs.shape
    x = style_imgs.permute(0, 4, 1, 2, 3).reshape(b, 3 * ns, p, p)                      # channel c*ns + s
    g = p // patch
    return x.reshape(b, 3 * ns, g, patch, g, patch).permute(0, 2, 4, 3, 5, 1).reshape(b, g * g, patch * patch * 3 * ns).contiguous()


def svit_assemble(patches, cls, pos, t_pad):
    b, n, dim = patches.shape
    x = torch.cat([cls.view(1, 1, dim).expand(b, 1, dim), torch.zeros(b, 1, dim), patches.float()], 1) + pos[None, :n + 2]
    return F.pad(x, (0, 0, 0, t_pad - (n + 2))).contiguous()


def token_mean(x, tokens):
    return x[:, :tokens].mean(1)
