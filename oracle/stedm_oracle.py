"""CPU oracle for STEDM's sampling path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict in, tensors out) fp32 restatement of what the reference executes for
``predict_step``: conditioning -> classifier-free-guided DDIM loop over the eps U-Net -> VQ first-stage
decode -> uint8 images.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file; ``stedm_b200/`` never does.

Parity pinning: the reference ships NO tests, golden vectors or fixtures, so this oracle is
pinned against outputs of the reference's own code run in the build container through
``oracle/ref_shims.py`` — see ``oracle/make_golden.py`` (which also asserts oracle == reference) and the
committed ``tests/golden/*.npz``.  Third-party arithmetic: taming ``VectorQuantizer2`` (un-vendored, pinned
only to @master) is restated in ``vq_quantize``; torchvision ``swin_v2_t`` (0.18.1 pinned, 0.26 here) is
called as the library it is.

All arithmetic is float32 on the CPU (no TF32), tensors are NCHW like the reference.  Every function cites
the reference lines it restates (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

UNET = "model.diffusion_model."
VAE = "first_stage_model."


# --------------------------------------------------------------------------------------------------
# schedules
# --------------------------------------------------------------------------------------------------
def alphas_cumprod_linear(timesteps=1000, linear_start=0.0015, linear_end=0.0205):
    """float64 betas -> cumprod -> float32, ldm/modules/diffusionmodules/util.py:21-25 and
    ldm/models/diffusion/ddpm.py:120-140.  Returns (alphas_cumprod fp32 ndarray, float64 ndarray)."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
    ac64 = np.cumprod(1.0 - betas, axis=0)
    return ac64.astype(np.float32), ac64


def ddim_timesteps(S, T=1000):
    """uniform discretisation, util.py:46-60: c = T // S; range(0, T, c) + 1."""
    c = T // S
    return np.asarray(list(range(0, T, c))) + 1


def ddim_tables(S, eta=0.0, T=1000, linear_start=0.0015, linear_end=0.0205):
    """Per-index fp32 scalars used by p_sample_ddim (ddim.py:24-53, 195-198; util.py:63-74).

    ``alphas`` is a gather from the fp32 alphas_cumprod; ``alphas_prev`` = [ac[0]] + ac[ts[:-1]];
    sigma computed in float64 from fp32 alphas then rounded to fp32 by ``torch.full``.
    """
    ac32, _ = alphas_cumprod_linear(T, linear_start, linear_end)
    ts = ddim_timesteps(S, T)
    a = ac32[ts]                                                     # fp32
    a_prev = np.asarray([ac32[0]] + ac32[ts[:-1]].tolist())          # float64 holding fp32 values
    sig = eta * np.sqrt((1 - a_prev) / (1 - a.astype(np.float64)) * (1 - a.astype(np.float64) / a_prev))
    sqrt_1m = np.sqrt(1.0 - a)                                       # fp32 (np.sqrt on the fp32 tensor)
    return dict(timesteps=ts, a_t=a.astype(np.float32), a_prev=a_prev.astype(np.float32),
                sigma=sig.astype(np.float32), sqrt_one_minus_a=sqrt_1m.astype(np.float32))


def timestep_embedding(t, dim, max_period=10000):
    """[cos | sin] sinusoid, util.py:151-171."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# --------------------------------------------------------------------------------------------------
# U-Net (openaimodel.py)
# --------------------------------------------------------------------------------------------------
def _gn(x, sd, key, eps):
    return F.group_norm(x.float(), 32, sd[key + ".weight"], sd[key + ".bias"], eps)


def _conv(x, sd, key, stride=1, padding=1):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)


def resblock(x, emb, sd, p):
    """ResBlock._forward, openaimodel.py:268-288 (use_scale_shift_norm=False, no up/down, dropout 0)."""
    h = _conv(F.silu(_gn(x, sd, p + "in_layers.0", 1e-5)), sd, p + "in_layers.2")
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, p + "out_layers.0", 1e-5)), sd, p + "out_layers.3")
    if (p + "skip_connection.weight") in sd:
        x = _conv(x, sd, p + "skip_connection", padding=0)
    return x + h


def attention_block(x, sd, p, num_heads):
    """AttentionBlock._forward + QKVAttentionLegacy, openaimodel.py:340-346, 378-394.
    qkv channels are head-major [h0: q|k|v, h1: q|k|v, ...]; scale ch^-1/4 on q and k; fp32 softmax."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    n = F.group_norm(xf, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    qkv = F.conv1d(n, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    ch = c // num_heads
    q, k, v = qkv.reshape(b * num_heads, 3 * ch, -1).split(ch, dim=1)
    s = 1.0 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * s, k * s)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, c, -1)
    h = F.conv1d(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def spatial_transformer(x, sd, p, n_heads, context=None):
    """SpatialTransformer.forward, ldm/modules/attention.py:245-261, eval mode: GroupNorm(32, eps 1e-6) -> proj_in 1x1 ->
    'b c h w -> b (h w) c' -> BasicTransformerBlock._forward per block (:211-215) -> back -> proj_out 1x1 -> + x_in.
    CrossAttention (:169-193): q = to_q(x), k, v = to_k / to_v(context or x), heads split 'b n (h d)', softmax(q k^T *
    d^-1/2) v, to_out.  FeedForward with GEGLU (:37-63): proj -> x * gelu(gate) -> Linear."""
    b, c, hh, ww = x.shape
    g = lambda k: sd[p + k]

    def attn(y, q, ctx):
        cx = y if ctx is None else ctx
        qq, kk, vv = F.linear(y, g(q + "to_q.weight")), F.linear(cx, g(q + "to_k.weight")), F.linear(cx, g(q + "to_v.weight"))
        sp = lambda t: t.reshape(t.shape[0], t.shape[1], n_heads, -1).permute(0, 2, 1, 3)
        qq, kk, vv = sp(qq), sp(kk), sp(vv)
        sim = torch.softmax(qq @ kk.transpose(-1, -2) * (qq.shape[-1] ** -0.5), dim=-1)
        o = (sim @ vv).permute(0, 2, 1, 3).reshape(y.shape[0], y.shape[1], -1)
        return F.linear(o, g(q + "to_out.0.weight"), g(q + "to_out.0.bias"))

    h = F.group_norm(x, 32, g("norm.weight"), g("norm.bias"), 1e-6)
    h = F.conv2d(h, g("proj_in.weight"), g("proj_in.bias"))
    inner = h.shape[1]
    h = h.reshape(b, inner, -1).permute(0, 2, 1)
    if context is not None and context.dim() == 2:
        context = context[:, None, :]
    i = 0
    while (p + f"transformer_blocks.{i}.norm1.weight") in sd:
        q = f"transformer_blocks.{i}."
        ln = lambda y, n: F.layer_norm(y, (inner,), g(q + n + ".weight"), g(q + n + ".bias"), 1e-5)
        h = attn(ln(h, "norm1"), q + "attn1.", None) + h
        h = attn(ln(h, "norm2"), q + "attn2.", context) + h
        a, gate = F.linear(ln(h, "norm3"), g(q + "ff.net.0.proj.weight"), g(q + "ff.net.0.proj.bias")).chunk(2, dim=-1)
        h = F.linear(a * F.gelu(gate), g(q + "ff.net.2.weight"), g(q + "ff.net.2.bias")) + h
        i += 1
    h = h.permute(0, 2, 1).reshape(b, inner, hh, ww)
    return F.conv2d(h, g("proj_out.weight"), g("proj_out.bias")) + x


def unet_structure(sd, prefix=UNET):
    """Recover the block list from the state-dict keys (UNetModel.__init__, openaimodel.py:536-733)."""
    n_in = 1 + max(int(k[len(prefix) + 13:].split(".")[0]) for k in sd if k.startswith(prefix + "input_blocks."))
    n_out = 1 + max(int(k[len(prefix) + 14:].split(".")[0]) for k in sd if k.startswith(prefix + "output_blocks."))
    return n_in, n_out


def unet_forward(sd, x, t, context, num_heads=8, prefix=UNET):
    """UNetModel.forward, openaimodel.py:761-806, for the shipped config (17 ResBlock + 1 ResBlockStyle +
    1 AttentionBlock; the style vector replaces the timestep embedding in middle_block.1, :291-297)."""
    P = prefix
    mc = sd[P + "time_embed.0.weight"].shape[1]
    emb = timestep_embedding(t, mc)
    emb = F.linear(emb, sd[P + "time_embed.0.weight"], sd[P + "time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd[P + "time_embed.2.weight"], sd[P + "time_embed.2.bias"])
    n_in, n_out = unet_structure(sd, P)
    hs = []
    h = x.float()
    for i in range(n_in):
        p = f"{P}input_blocks.{i}."
        if (p + "0.weight") in sd:                       # stem conv (:539-545)
            h = _conv(h, sd, p + "0")
        elif (p + "0.op.weight") in sd:                  # Downsample: conv3x3 stride 2 pad 1 (:164-166)
            h = _conv(h, sd, p + "0.op", stride=2)
        else:
            h = resblock(h, emb, sd, p + "0.")
        hs.append(h)
    p = P + "middle_block."
    h = resblock(h, emb, sd, p + "0.")
    h = resblock(h, context.float(), sd, p + "1.block.")  # ResBlockStyle (:291-297; dispatch :93-101)
    if (p + "2.transformer_blocks.0.norm1.weight") in sd:
        # use_spatial_transformer=True (:648-652): TimestepEmbedSequential passes NO context to it (:93-101)
        h = spatial_transformer(h, sd, p + "2.", num_heads, None)
    else:
        h = attention_block(h, sd, p + "2.", num_heads)
    h = resblock(h, emb, sd, p + "3.")
    for i in range(n_out):
        p = f"{P}output_blocks.{i}."
        h = torch.cat([h, hs.pop()], dim=1)               # (:800)
        h = resblock(h, emb, sd, p + "0.")
        if (p + "1.conv.weight") in sd:                   # Upsample: nearest x2 then conv3x3 (:123-132)
            h = _conv(F.interpolate(h, scale_factor=2, mode="nearest"), sd, p + "1.conv")
    h = F.silu(_gn(h, sd, P + "out.0", 1e-5))
    return _conv(h, sd, P + "out.2")


def apply_model(sd, x, t, cond, num_heads=8):
    """LatentDiffusion.apply_model -> DiffusionWrapper.forward 'hybrid', ddpm.py:894-995, 1414-1417."""
    xc = torch.cat([x] + list(cond["c_concat"]), dim=1)
    cc = torch.cat(list(cond["c_crossattn"]), dim=1)
    return unet_forward(sd, xc, t, cc, num_heads)


# --------------------------------------------------------------------------------------------------
# DDIM step with STEDM's rescaled classifier-free guidance (ddim.py:164-210)
# --------------------------------------------------------------------------------------------------
def cfg_combine(e_c, e_u, scale, phi=0.7):
    """ddim.py:177-184.  std over dims (1, 2) = channels and HEIGHT only, unbiased, keepdim -> (B,1,1,W)."""
    e_w = e_u + scale * (e_c - e_u)
    dims = tuple(range(1, e_c.ndim - 1))
    resc = e_w * (e_c.std(dim=dims, keepdim=True) / e_w.std(dim=dims, keepdim=True))
    return resc * phi + (1.0 - phi) * e_c


def ddim_update(x, e, a_t, a_prev, sigma, sqrt_1m_at, noise=None):
    """ddim.py:195-209 with fp32 scalars (torch.full of python floats -> fp32 tensors)."""
    f = lambda v: torch.tensor(float(v), dtype=torch.float32)
    a_t, a_prev, sigma, sqrt_1m_at = f(a_t), f(a_prev), f(sigma), f(sqrt_1m_at)
    pred_x0 = (x - sqrt_1m_at * e) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma ** 2).sqrt() * e
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt
    if noise is not None:
        x_prev = x_prev + sigma * noise
    return x_prev, pred_x0


def ddim_sample(sd, cond, uncond, x_T, S=50, eta=0.0, cfg_scale=1.5, num_heads=8, log_every_t=1000,
                max_steps=None, noise_fn=None):
    """DDIMSampler.sample / ddim_sampling, ddim.py:55-162.  ``max_steps`` truncates the loop (tests)."""
    tab = ddim_tables(S, eta)
    ts = tab["timesteps"]
    total = ts.shape[0]
    img = x_T
    inter = {"x_inter": [img], "pred_x0": [img]}
    for i, step in enumerate(np.flip(ts)):
        if max_steps is not None and i >= max_steps:
            break
        index = total - i - 1
        t = torch.full((img.shape[0],), int(step), dtype=torch.long)
        e = apply_model(sd, img, t, cond, num_heads)
        if uncond is not None and cfg_scale != 1.0:
            e_u = apply_model(sd, img, t, uncond, num_heads)
            e = cfg_combine(e, e_u, cfg_scale)
        noise = noise_fn(img.shape) if (noise_fn is not None) else None
        img, pred_x0 = ddim_update(img, e, tab["a_t"][index], tab["a_prev"][index], tab["sigma"][index],
                                   tab["sqrt_one_minus_a"][index], noise)
        if index % log_every_t == 0 or index == total - 1:
            inter["x_inter"].append(img)
            inter["pred_x0"].append(pred_x0)
    return img, inter


def ddpm_ancestral_sample(sd, cond, x_T, timesteps, noises, quantize_denoised=False, clip_denoised=False, num_heads=8,
                          v_posterior=0.0, linear_start=0.0015, linear_end=0.0205, T=1000):
    """LatentDiffusion.sample -> p_sample_loop -> p_sample -> p_mean_variance (ddpm.py:1050-1110, 1168-1235) with the
    schedule buffers of register_schedule (ddpm.py:120-172: float64 numpy, stored fp32), predict_start_from_noise
    (:219-223) and q_posterior (:225-232).  ``noises[i]`` is the N(0,1) draw of loop iteration i (t = timesteps-1-i);
    the t == 0 draw is masked out.  ``quantize_denoised`` snaps x_recon to the VQ codebook (:1071-1072)."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=np.float64) ** 2
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    sqrt_recip, sqrt_recipm1 = f32(np.sqrt(1.0 / ac)), f32(np.sqrt(1.0 / ac - 1))
    pv = (1 - v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + v_posterior * betas
    logvar = f32(np.log(np.maximum(pv, 1e-20)))
    coef1 = f32(betas * np.sqrt(ac_prev) / (1.0 - ac))
    coef2 = f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac))
    img = x_T
    for i, t_i in enumerate(reversed(range(0, timesteps))):
        t = torch.full((img.shape[0],), t_i, dtype=torch.long)
        eps = apply_model(sd, img, t, cond, num_heads)
        x_recon = sqrt_recip[t_i] * img - sqrt_recipm1[t_i] * eps
        if clip_denoised:
            x_recon = x_recon.clamp(-1.0, 1.0)
        if quantize_denoised:
            x_recon, _ = vq_quantize(x_recon, sd[VAE + "quantize.embedding.weight"])
        mean = coef1[t_i] * x_recon + coef2[t_i] * img
        nonzero = 0.0 if t_i == 0 else 1.0
        img = mean + nonzero * (0.5 * logvar[t_i]).exp() * noises[i]
    return img


class DiscreteVPSchedule:
    """NoiseScheduleVP('discrete', alphas_cumprod=...), ldm/models/diffusion/dpm_solver/dpm_solver.py:7-160, restated on
    the host in float64: key points t_n = (n + 1) / N with log alpha(t_n) = 0.5 log(alphas_cumprod[n]) (:78-88),
    log alpha(t) piecewise linear in between (interpolate_fn, :1113-1151), sigma = sqrt(1 - alpha^2) (:126-130),
    lambda = log alpha - log sigma (:132-138)."""

    def __init__(self, alphas_cumprod):
        ac = np.asarray(alphas_cumprod, dtype=np.float64)
        self.N = ac.shape[0]
        self.t = np.arange(1, self.N + 1, dtype=np.float64) / self.N
        self.log_alpha_n = 0.5 * np.log(ac)

    def log_alpha(self, t):
        return float(np.interp(t, self.t, self.log_alpha_n))

    def alpha(self, t):
        return math.exp(self.log_alpha(t))

    def sigma(self, t):
        return math.sqrt(1.0 - math.exp(2.0 * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * math.log(1.0 - math.exp(2.0 * la))

    def model_time(self, t):
        """get_model_input_time, :246-255: continuous t in [1/N, 1] -> the U-Net's (fractional) timestep."""
        return (t - 1.0 / self.N) * 1000.0


def dpm_solver_sample(eps_fn, x_T, alphas_cumprod, S, cfg_scale=1.0, uncond_eps_fn=None):
    """DPMSolverSampler.sample (dpm_solver/sampler.py:27-95) as the reference configures it: DPM_Solver(predict_x0=True,
    thresholding=False).sample(steps=S, skip_type='time_uniform', method='multistep', order=2, lower_order_final=True)
    with plain classifier-free guidance (dpm_solver.py:302-320).  data prediction x0 = (x - sigma eps) / alpha (:361-374);
    first-order update x_t = (sigma_t / sigma_s) x - alpha_t expm1(-h) x0_s (:478-510); second-order multistep update
    with D1 = (x0_0 - x0_1) / r0, r0 = h_0 / h (:732-768); time steps linspace(1, 1/N, S + 1) (:402-403); the last step
    drops to first order when S < 15 (:1075-1078).  ``eps_fn(x, t_model)`` takes the FLOAT model time."""
    ns = DiscreteVPSchedule(alphas_cumprod)
    ts = np.linspace(1.0, 1.0 / ns.N, S + 1)
    b = x_T.shape[0]

    def x0_of(x, t):
        tm = torch.full((b,), ns.model_time(t), dtype=torch.float32)
        e = eps_fn(x, tm)
        if uncond_eps_fn is not None and cfg_scale != 1.0:
            e_u = uncond_eps_fn(x, tm)
            e = e_u + cfg_scale * (e - e_u)
        return (x - ns.sigma(t) * e) / ns.alpha(t)

    def first(x, s, t, m_s):
        h = ns.lam(t) - ns.lam(s)
        return (ns.sigma(t) / ns.sigma(s)) * x - (ns.alpha(t) * math.expm1(-h)) * m_s

    def second(x, t_p1, t_p0, t, m_p1, m_p0):
        h0, h = ns.lam(t_p0) - ns.lam(t_p1), ns.lam(t) - ns.lam(t_p0)
        d1 = (1.0 / (h0 / h)) * (m_p0 - m_p1)
        c = ns.alpha(t) * (math.exp(-h) - 1.0)
        return (ns.sigma(t) / ns.sigma(t_p0)) * x - c * m_p0 - 0.5 * c * d1

    x = x_T
    m_prev = [x0_of(x, ts[0])]
    t_prev = [ts[0]]
    x = first(x, ts[0], ts[1], m_prev[0])                       # init order 1
    m_prev.append(x0_of(x, ts[1]))
    t_prev.append(ts[1])
    for step in range(2, S + 1):
        order = min(2, S + 1 - step) if S < 15 else 2
        t = ts[step]
        x = second(x, t_prev[0], t_prev[1], t, m_prev[0], m_prev[1]) if order == 2 else first(x, t_prev[1], t, m_prev[1])
        t_prev, m_prev = [t_prev[1], t], [m_prev[1], None]
        if step < S:
            m_prev[1] = x0_of(x, t)
    return x


def plms_sample(eps_fn, x_T, S=50, cfg_scale=1.0, uncond_eps_fn=None, max_steps=None):
    """PLMSSampler.sample / plms_sampling / p_sample_plms, ldm/models/diffusion/plms.py:113-236, eta = 0:
    e_t = e_u + w (e_c - e_u) (plain guidance, :184); first step = pseudo improved Euler (:219-223), then
    Adams-Bashforth orders 2-4 over the kept eps history (:224-232); x_prev from e_t_prime (:198-216).
    ``eps_fn(x, t)`` / ``uncond_eps_fn(x, t)`` stand for apply_model with the (un)conditional conditioning."""
    tab = ddim_tables(S, 0.0)
    ts = tab["timesteps"]
    total = ts.shape[0]
    time_range = np.flip(ts)
    img = x_T
    old_eps = []

    def model_output(x, t):
        e = eps_fn(x, t)
        if uncond_eps_fn is not None and cfg_scale != 1.0:
            e_u = uncond_eps_fn(x, t)
            e = e_u + cfg_scale * (e - e_u)
        return e

    def x_prev_of(x, e, index):
        return ddim_update(x, e, tab["a_t"][index], tab["a_prev"][index], 0.0, tab["sqrt_one_minus_a"][index])

    for i, step in enumerate(time_range):
        if max_steps is not None and i >= max_steps:
            break
        index = total - i - 1
        t = torch.full((img.shape[0],), int(step), dtype=torch.long)
        t_next = torch.full((img.shape[0],), int(time_range[min(i + 1, len(time_range) - 1)]), dtype=torch.long)
        e_t = model_output(img, t)
        if len(old_eps) == 0:
            x_prev, _ = x_prev_of(img, e_t, index)
            e_prime = (e_t + model_output(x_prev, t_next)) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img, _ = x_prev_of(img, e_prime, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img


# --------------------------------------------------------------------------------------------------
# conditioning: SpatialRescaler (encoders/modules.py:104-133) and Agg_Mean (networks/agg_blocks.py:57-75)
# --------------------------------------------------------------------------------------------------
def spatial_rescaler(sd, seg_nchw, n_stages=2):
    x = seg_nchw
    for _ in range(n_stages):
        x = F.interpolate(x, scale_factor=0.5, mode="bilinear")
    return F.conv2d(x, sd["cond_stage_model.channel_mapper.weight"])


_SWIN_CACHE = {}


def swin_embedder(sd, prefix="agg_block.embedder."):
    """torchvision swin_v2_t with head = Linear(768, 512), networks/s_zss_dm.py:19-20 (library call)."""
    import torchvision
    key = id(sd)
    if key not in _SWIN_CACHE:
        m = torchvision.models.get_model("swin_v2_t")
        m.head = torch.nn.Linear(768, 512)
        sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        missing, unexpected = m.load_state_dict(sub, strict=False)
        # constructor-defined tensors (relative position tables/indices, logit_scale) may be absent from a
        # weights-only state dict; every learnable .weight/.bias must be present
        assert not unexpected and not [k for k in missing if k.endswith((".weight", ".bias"))], (missing, unexpected)
        _SWIN_CACHE.clear()
        _SWIN_CACHE[key] = m.eval()
    return _SWIN_CACHE[key]


@torch.no_grad()
def agg_mean(sd, style_bnhwc):
    b, n = style_bnhwc.shape[:2]
    imgs = style_bnhwc.permute(0, 1, 4, 2, 3).reshape(b * n, 3, *style_bnhwc.shape[2:4])
    f = swin_embedder(sd)(imgs)
    return f.reshape(b, n, -1).mean(dim=1)


@torch.no_grad()
def svit_aggregate(sd, style_bnhwc, heads, patch=8, pool="mean", prefix="agg_block."):
    """networks/vit_set.py sViT.forward (:163-208) with t_emb = None, c_old = None, eval mode (dropouts off), built by
    networks/s_zss_dm.py:31-38 for ``style_agg=svit``.  SPT (:82-107): the ns images of a sample are stacked along
    channels (channel c*ns + s), cut into patch x patch patches flattened as (p1 p2 c), LayerNorm, Linear.  Tokens
    = [cls, zero time token, patches] + pos_embedding (:176-188).  Transformer (:67-80): x = LSA(LN(x)) + x;
    x = FF(LN(x)) + x.  LSA (:44-60): softmax over (q k^T * exp(temperature)) with the diagonal masked out; dim_head 64.
    Pooling (:195-205) and mlp_head = LayerNorm + Linear (:147-150)."""
    g = lambda k: sd[prefix + k]
    x = style_bnhwc.float().permute(0, 1, 4, 2, 3)                                  # b ns c H W   (:165)
    b, ns, ch, H, W = x.shape
    x = x.permute(0, 2, 1, 3, 4).reshape(b, ch * ns, H, W)                          # (:104-105)
    hh, ww = H // patch, W // patch
    x = x.reshape(b, ch * ns, hh, patch, ww, patch).permute(0, 2, 4, 3, 5, 1).reshape(b, hh * ww, patch * patch * ch * ns)
    tp = "to_patch_embedding.to_patch_tokens."
    x = F.layer_norm(x, (x.shape[-1],), g(tp + "1.weight"), g(tp + "1.bias"), 1e-5)
    x = F.linear(x, g(tp + "2.weight"), g(tp + "2.bias"))
    n, dim = x.shape[1], x.shape[2]
    x = torch.cat([g("cls_token").expand(b, 1, dim), torch.zeros(b, 1, dim), x], 1) + g("pos_embedding")[:, :n + 2]
    depth = 0
    while (prefix + f"transformer.layers.{depth}.0.norm.weight") in sd:
        depth += 1
    for i in range(depth):
        p = f"transformer.layers.{i}."
        y = F.layer_norm(x, (dim,), g(p + "0.norm.weight"), g(p + "0.norm.bias"), 1e-5)
        q, k, v = F.linear(y, g(p + "0.fn.to_qkv.weight")).chunk(3, dim=-1)
        sp = lambda t: t.reshape(b, n + 2, heads, -1).permute(0, 2, 1, 3)
        q, k, v = sp(q), sp(k), sp(v)
        dots = q @ k.transpose(-1, -2) * g(p + "0.fn.temperature").exp()
        dots = dots.masked_fill(torch.eye(n + 2, dtype=torch.bool), -torch.finfo(dots.dtype).max)
        o = (torch.softmax(dots, -1) @ v).permute(0, 2, 1, 3).reshape(b, n + 2, -1)
        x = F.linear(o, g(p + "0.fn.to_out.0.weight"), g(p + "0.fn.to_out.0.bias")) + x
        y = F.layer_norm(x, (dim,), g(p + "1.norm.weight"), g(p + "1.norm.bias"), 1e-5)
        y = F.gelu(F.linear(y, g(p + "1.fn.net.0.weight"), g(p + "1.fn.net.0.bias")))
        x = F.linear(y, g(p + "1.fn.net.3.weight"), g(p + "1.fn.net.3.bias")) + x
    x = x.mean(1) if pool == "mean" else x[:, 0]
    x = F.layer_norm(x, (dim,), g("mlp_head.0.weight"), g("mlp_head.0.bias"), 1e-5)
    return F.linear(x, g("mlp_head.1.weight"), g("mlp_head.1.bias"))


def get_conditioning(sd, seg_bhwc, style_bnhwc):
    """S_ZSS_DM.get_input, networks/s_zss_dm.py:45-60, minus the discarded VAE encode of the image."""
    seg = seg_bhwc.permute(0, 3, 1, 2).contiguous().float()            # ddpm.py:332-338
    return {"c_concat": [spatial_rescaler(sd, seg)], "c_crossattn": [agg_mean(sd, style_bnhwc)]}


# --------------------------------------------------------------------------------------------------
# first stage decode (ldm/models/autoencoder.py:274-282; ldm/modules/diffusionmodules/model.py:462-568)
# --------------------------------------------------------------------------------------------------
def vq_quantize(z, codebook):
    """taming VectorQuantizer2.forward (see module docstring): nearest row of ``codebook`` per pixel."""
    b, c, h, w = z.shape
    zf = z.permute(0, 2, 3, 1).reshape(-1, c)
    d = (zf ** 2).sum(1, keepdim=True) + (codebook ** 2).sum(1) - 2.0 * zf @ codebook.t()
    idx = torch.argmin(d, dim=1)
    zq = codebook[idx].reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    return zq, idx


def _vae_resblock(x, sd, p):
    """ResnetBlock.forward with temb=None, model.py:121-141 (GroupNorm eps 1e-6, :38-39)."""
    h = _conv(F.silu(_gn(x, sd, p + "norm1", 1e-6)), sd, p + "conv1")
    h = _conv(F.silu(_gn(h, sd, p + "norm2", 1e-6)), sd, p + "conv2")
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(x, sd, p + "nin_shortcut", padding=0)
    return x + h


def _vae_attn(x, sd, p):
    """AttnBlock.forward, model.py:178-202: single head, d = C, scale C^-1/2."""
    b, c, hh, ww = x.shape
    n = _gn(x, sd, p + "norm", 1e-6)
    q = _conv(n, sd, p + "q", padding=0).reshape(b, c, -1).permute(0, 2, 1)
    k = _conv(n, sd, p + "k", padding=0).reshape(b, c, -1)
    v = _conv(n, sd, p + "v", padding=0).reshape(b, c, -1)
    w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, p + "proj_out", padding=0)


def vae_decode(sd, z, force_not_quantize=False, prefix=VAE):
    P = prefix
    if not force_not_quantize:
        z, _ = vq_quantize(z, sd[P + "quantize.embedding.weight"])
    h = _conv(z, sd, P + "post_quant_conv", padding=0)
    D = P + "decoder."
    h = _conv(h, sd, D + "conv_in")
    h = _vae_resblock(h, sd, D + "mid.block_1.")
    h = _vae_attn(h, sd, D + "mid.attn_1.")
    h = _vae_resblock(h, sd, D + "mid.block_2.")
    n_levels = 1 + max(int(k[len(D) + 3:].split(".")[0]) for k in sd if k.startswith(D + "up."))
    for lvl in reversed(range(n_levels)):
        i = 0
        while (f"{D}up.{lvl}.block.{i}.conv1.weight") in sd:
            h = _vae_resblock(h, sd, f"{D}up.{lvl}.block.{i}.")
            i += 1
        if (f"{D}up.{lvl}.upsample.conv.weight") in sd:
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, f"{D}up.{lvl}.upsample.conv")
    h = F.silu(_gn(h, sd, D + "norm_out", 1e-6))
    return _conv(h, sd, D + "conv_out")


def decode_first_stage(sd, z, force_not_quantize=False, scale_factor=1.0):
    """LatentDiffusion.decode_first_stage, ddpm.py:708-766 (no split_input_params)."""
    return vae_decode(sd, (1.0 / scale_factor) * z, force_not_quantize)


def to_uint8(img_nchw):
    """predict_step tail, modules/ldm_diffusion.py:94-96: clip, (x+1)*127.5, TRUNCATING uint8 cast, NHWC."""
    x = torch.clip(img_nchw, -1, 1)
    return ((x.permute(0, 2, 3, 1).cpu().numpy() + 1) * 127.5).astype(np.uint8)


# --------------------------------------------------------------------------------------------------
# whole path (modules/ldm_diffusion.py:76-96)
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def predict_images(sd, seg_bhwc, style_bnhwc, x_T, S=50, eta=0.0, cfg_scale=1.5, max_steps=None):
    cond = get_conditioning(sd, seg_bhwc, style_bnhwc)
    if cfg_scale == 1:
        uncond = None
    else:
        uncond = get_conditioning(sd, seg_bhwc, torch.zeros_like(style_bnhwc) - 2)
    z, _ = ddim_sample(sd, cond, uncond, x_T, S=S, eta=eta, cfg_scale=cfg_scale, max_steps=max_steps)
    img = decode_first_stage(sd, z)
    return to_uint8(img), img, z


def synthetic_batch(B, P, n_style=1, seed=0):
    """Synthetic predict batch: low-frequency 2-class layout, U(-1,1) style images."""
    g = torch.Generator().manual_seed(99 + seed)
    low = torch.rand(B, 1, 8, 8, generator=g)
    mask = (F.interpolate(low, size=(P, P), mode="bilinear") > 0.5).float()[:, 0]
    seg = torch.stack([1.0 - mask, mask], dim=-1)                               # (B,P,P,2) one-hot
    g2 = torch.Generator().manual_seed(3 + seed)
    style = torch.rand(B, n_style, P, P, 3, generator=g2) * 2 - 1
    L = P // 4
    x_T = torch.stack([torch.randn(3, L, L, generator=torch.Generator().manual_seed(1234 + seed * 100003 + i))
                       for i in range(B)])
    return seg, style, x_T
