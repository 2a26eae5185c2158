"""Drop-in for the sampling-path API of ``ldm.models.diffusion.ddpm`` (reference ddpm.py): ``DDPM`` schedule
buffers (:120-172), ``LatentDiffusion`` (:427-1250: ``apply_model``, ``sample``, ``sample_log``,
``decode_first_stage``, ``get_input``, ``get_learned_conditioning``, ancestral ``p_sample_loop``) and
``DiffusionWrapper`` (:1398-1424).  Same names, signatures, defaults and return structures; tensors crossing the
API are NCHW fp32 on the caller's CUDA device.  Training (losses, EMA, optimisers, logging) is outside the path.

What runs where: the U-Net and the VQ decoder execute on the native sm_100a engine; this file is host glue.
``precision`` ('bf16' throughput mode | 'fp32' parity mode) is the one constructor argument the reference lacks.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from ...util import default, instantiate_from_config
from ...modules.diffusionmodules.util import extract_into_tensor, make_beta_schedule, noise_like
from ..autoencoder import VQModelInterface
from .ddim import DDIMSampler

__conditioning_keys__ = {"concat": "c_concat", "crossattn": "c_crossattn", "adm": "y"}


class DiffusionWrapper(nn.Module):
    """ddpm.py:1398-1424.  'hybrid' = channel-concat c_concat onto x, pass cat(c_crossattn) as context; here the
    channel concat is fused into the U-Net's input packing kernel instead of a torch.cat."""

    def __init__(self, diff_model_config, conditioning_key, precision="bf16"):
        super().__init__()
        cfg = dict(diff_model_config)
        cfg["params"] = dict(cfg.get("params", {}), precision=precision)
        self.diffusion_model = instantiate_from_config(cfg)
        self.conditioning_key = conditioning_key
        assert self.conditioning_key in [None, "concat", "crossattn", "hybrid", "adm"]

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None):
        if self.conditioning_key != "hybrid":
            raise NotImplementedError(f"conditioning_key={self.conditioning_key!r}: STEDM samples with 'hybrid' "
                                      f"(conf/diffusion/ldm_based.yaml:13)")
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        xc = c_concat[0] if len(c_concat) == 1 else torch.cat(c_concat, 1)
        return self.diffusion_model.forward_split(x, xc, t, cc)


class DDPM(nn.Module):
    def __init__(self, unet_config, timesteps=1000, beta_schedule="linear", loss_type="l2", ckpt_path=None,
                 ignore_keys=(), load_only_unet=False, monitor="val/loss", use_ema=True, first_stage_key="image",
                 image_size=256, channels=3, log_every_t=100, clip_denoised=True, linear_start=1e-4,
                 linear_end=2e-2, cosine_s=8e-3, given_betas=None, original_elbo_weight=0., v_posterior=0.,
                 l_simple_weight=1., conditioning_key=None, parameterization="eps", scheduler_config=None,
                 use_positional_encodings=False, learn_logvar=False, logvar_init=0., precision="bf16"):
        super().__init__()
        assert parameterization in ["eps", "x0"], 'currently only supporting "eps" and "x0"'
        if parameterization != "eps" or use_positional_encodings or learn_logvar:
            raise NotImplementedError("only eps-prediction without positional encodings is on the sampling path")
        self.parameterization = parameterization
        self.cond_stage_model = None
        self.clip_denoised, self.log_every_t = clip_denoised, log_every_t
        self.first_stage_key, self.image_size, self.channels = first_stage_key, image_size, channels
        self.use_positional_encodings = use_positional_encodings
        self.precision = precision
        self.model = DiffusionWrapper(unet_config, conditioning_key, precision)
        self.use_ema = False  # predict_step samples the raw weights; model_ema.* keys are ignored
        self.v_posterior = v_posterior
        if monitor is not None:
            self.monitor = monitor
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=list(ignore_keys), only_model=load_only_unet)
        self.register_schedule(given_betas=given_betas, beta_schedule=beta_schedule, timesteps=timesteps,
                               linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        self.loss_type = loss_type
        self.register_buffer("logvar", torch.full(fill_value=logvar_init, size=(self.num_timesteps,)))

    @property
    def device(self):
        return self.betas.device

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """float64 numpy math, fp32 buffers — ddpm.py:120-172."""
        betas = given_betas if given_betas is not None else make_beta_schedule(
            beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1. - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1., ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        f32 = partial(torch.tensor, dtype=torch.float32)
        reg = self.register_buffer
        reg("betas", f32(betas))
        reg("alphas_cumprod", f32(ac))
        reg("alphas_cumprod_prev", f32(ac_prev))
        reg("sqrt_alphas_cumprod", f32(np.sqrt(ac)))
        reg("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1. - ac)))
        reg("log_one_minus_alphas_cumprod", f32(np.log(1. - ac)))
        reg("sqrt_recip_alphas_cumprod", f32(np.sqrt(1. / ac)))
        reg("sqrt_recipm1_alphas_cumprod", f32(np.sqrt(1. / ac - 1)))
        pv = (1 - self.v_posterior) * betas * (1. - ac_prev) / (1. - ac) + self.v_posterior * betas
        reg("posterior_variance", f32(pv))
        reg("posterior_log_variance_clipped", f32(np.log(np.maximum(pv, 1e-20))))
        reg("posterior_mean_coef1", f32(betas * np.sqrt(ac_prev) / (1. - ac)))
        reg("posterior_mean_coef2", f32((1. - ac_prev) * np.sqrt(alphas) / (1. - ac)))

    def init_from_ckpt(self, path, ignore_keys=(), only_model=False):
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd)
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        target = self.model if only_model else self
        missing, unexpected = target.load_state_dict(sd, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")

    def get_input(self, batch, k):
        """ddpm.py:332-338: 'b h w c -> b c h w', fp32, contiguous."""
        x = batch[k]
        if x.dim() == 3:
            x = x[..., None]
        return x.permute(0, 3, 1, 2).to(memory_format=torch.contiguous_format).float()

    # -- ancestral-sampler helpers (ddpm.py:219-232, 247-251)
    def predict_start_from_noise(self, x_t, t, noise):
        return (extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def q_posterior(self, x_start, x_t, t):
        mean = (extract_into_tensor(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + extract_into_tensor(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        return (mean, extract_into_tensor(self.posterior_variance, t, x_t.shape),
                extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape))

    def q_sample(self, x_start, t, noise=None):
        noise = default(noise, lambda: torch.randn_like(x_start))
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)


class LatentDiffusion(DDPM):
    """main class (ddpm.py:427-1250), sampling side."""

    def __init__(self, first_stage_config, cond_stage_config, num_timesteps_cond=None, cond_stage_key="image",
                 cond_stage_trainable=False, concat_mode=True, cond_stage_forward=None, conditioning_key=None,
                 scale_factor=1.0, scale_by_std=False, *args, **kwargs):
        self.num_timesteps_cond = default(num_timesteps_cond, 1)
        self.scale_by_std = scale_by_std
        assert self.num_timesteps_cond <= kwargs["timesteps"]
        if conditioning_key is None:
            conditioning_key = "concat" if concat_mode else "crossattn"
        if cond_stage_config == "__is_unconditional__":
            conditioning_key = None
        ckpt_path = kwargs.pop("ckpt_path", None)
        ignore_keys = kwargs.pop("ignore_keys", [])
        super().__init__(conditioning_key=conditioning_key, *args, **kwargs)
        self.concat_mode, self.cond_stage_trainable, self.cond_stage_key = concat_mode, cond_stage_trainable, cond_stage_key
        try:
            self.num_downs = len(first_stage_config["params"]["ddconfig"]["ch_mult"]) - 1
        except Exception:
            self.num_downs = 0
        if not scale_by_std:
            self.scale_factor = scale_factor
        else:
            self.register_buffer("scale_factor", torch.tensor(scale_factor))
        self.instantiate_first_stage(first_stage_config)
        self.instantiate_cond_stage(cond_stage_config)
        self.cond_stage_forward = cond_stage_forward
        self.clip_denoised = False
        self.shorten_cond_schedule = self.num_timesteps_cond > 1
        self.restarted_from_ckpt = False
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys)
            self.restarted_from_ckpt = True

    def instantiate_first_stage(self, config):
        cfg = dict(config)
        cfg["params"] = dict(cfg.get("params", {}), precision=self.precision)
        self.first_stage_model = instantiate_from_config(cfg).eval()
        for p in self.first_stage_model.parameters():
            p.requires_grad = False

    def instantiate_cond_stage(self, config):
        if config in ("__is_first_stage__", "__is_unconditional__"):
            raise NotImplementedError("STEDM conditions on a SpatialRescaler layout (cond_stage_config/spatial.yaml)")
        self.cond_stage_model = instantiate_from_config(config)
        if not self.cond_stage_trainable:
            self.cond_stage_model.eval()
            for p in self.cond_stage_model.parameters():
                p.requires_grad = False

    def set_precision(self, precision):
        """Switch every native component between 'bf16' (throughput) and 'fp32' (parity)."""
        self.precision = precision
        self.model.diffusion_model.set_precision(precision)
        self.first_stage_model.set_precision(precision)

    def get_learned_conditioning(self, c):
        if self.cond_stage_forward is None:
            enc = getattr(self.cond_stage_model, "encode", None)
            return enc(c) if callable(enc) else self.cond_stage_model(c)
        return getattr(self.cond_stage_model, self.cond_stage_forward)(c)

    @torch.no_grad()
    def get_input(self, batch, k, return_first_stage_outputs=False, force_c_encode=False, cond_key=None,
                  return_original_cond=False, bs=None):
        """ddpm.py:656-706.  The reference VAE-encodes batch[k] here and predict_step then uses only len(z)
        ("wasted work"): the encoder is not on the sampling path, so z is returned as zeros of the
        latent shape — same length, shape, dtype and device."""
        if return_first_stage_outputs:
            raise NotImplementedError("first-stage reconstructions need the encoder (outside the sampling path)")
        x = super().get_input(batch, k)
        if bs is not None:
            x = x[:bs]
        x = x.to(self.device)
        f = 2 ** self.num_downs
        z = torch.zeros((x.shape[0], self.channels, x.shape[2] // f, x.shape[3] // f), device=x.device)
        cond_key = cond_key or self.cond_stage_key
        xc = super().get_input(batch, cond_key).to(self.device) if cond_key != self.first_stage_key else x
        if not self.cond_stage_trainable or force_c_encode:
            c = self.get_learned_conditioning(xc)
        else:
            c = xc
        if bs is not None:
            c = c[:bs]
        out = [z, c]
        if return_original_cond:
            out.append(xc)
        return out

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        """ddpm.py:708-766 (no split_input_params: STEDM never sets it)."""
        if predict_cids:
            if z.dim() == 4:
                z = torch.argmax(z.exp(), dim=1).long()
            z = self.first_stage_model.quantize.get_codebook_entry(z, shape=None).permute(0, 3, 1, 2).contiguous()
        z = 1. / self.scale_factor * z
        if isinstance(self.first_stage_model, VQModelInterface):
            return self.first_stage_model.decode(z, force_not_quantize=predict_cids or force_not_quantize)
        return self.first_stage_model.decode(z)

    def apply_model(self, x_noisy, t, cond, return_ids=False):
        """ddpm.py:894-995: eps = U-Net(x, t, **cond)."""
        if not isinstance(cond, dict):
            if not isinstance(cond, list):
                cond = [cond]
            key = "c_concat" if self.model.conditioning_key == "concat" else "c_crossattn"
            cond = {key: cond}
        out = self.model(x_noisy, t, **cond)
        return out[0] if isinstance(out, tuple) and not return_ids else out

    # ---- DDPM ancestral sampler (ddpm.py:1050-1235); elementwise tail in torch, U-Net native -----------------
    def p_mean_variance(self, x, c, t, clip_denoised: bool, return_codebook_ids=False, quantize_denoised=False,
                        return_x0=False, score_corrector=None, corrector_kwargs=None):
        if return_codebook_ids or score_corrector is not None:
            raise NotImplementedError("codebook-id outputs / score correctors are not on STEDM's path")
        eps = self.apply_model(x, t, c)
        x_recon = self.predict_start_from_noise(x, t=t, noise=eps)
        if clip_denoised:
            x_recon.clamp_(-1., 1.)
        if quantize_denoised:           # ddpm.py:1071-1072: x_recon snapped to the first stage's VQ codebook
            from .... import ops
            x_recon = ops.vq_nearest(x_recon.float().contiguous(),
                                     self.first_stage_model.quantize.embedding.weight.detach().float().contiguous())
        mean, var, logvar = self.q_posterior(x_start=x_recon, x_t=x, t=t)
        return (mean, var, logvar, x_recon) if return_x0 else (mean, var, logvar)

    @torch.no_grad()
    def p_sample(self, x, c, t, clip_denoised=False, repeat_noise=False, return_codebook_ids=False,
                 quantize_denoised=False, return_x0=False, temperature=1., noise_dropout=0., score_corrector=None,
                 corrector_kwargs=None):
        b = x.shape[0]
        outs = self.p_mean_variance(x=x, c=c, t=t, clip_denoised=clip_denoised, return_codebook_ids=return_codebook_ids,
                                    quantize_denoised=quantize_denoised, return_x0=return_x0,
                                    score_corrector=score_corrector, corrector_kwargs=corrector_kwargs)
        mean, logvar = outs[0], outs[2]
        noise = noise_like(x.shape, x.device, repeat_noise) * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        mask = (1 - (t == 0).float()).reshape(b, *((1,) * (len(x.shape) - 1)))
        x_prev = mean + mask * (0.5 * logvar).exp() * noise
        return (x_prev, outs[3]) if return_x0 else x_prev

    @torch.no_grad()
    def p_sample_loop(self, cond, shape, return_intermediates=False, x_T=None, verbose=True, callback=None,
                      timesteps=None, quantize_denoised=False, mask=None, x0=None, img_callback=None, start_T=None,
                      log_every_t=None):
        log_every_t = log_every_t or self.log_every_t
        device = self.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        intermediates = [img]
        timesteps = self.num_timesteps if timesteps is None else timesteps
        if start_T is not None:
            timesteps = min(timesteps, start_T)
        for i in reversed(range(0, timesteps)):
            ts = torch.full((b,), i, device=device, dtype=torch.long)
            img = self.p_sample(img, cond, ts, clip_denoised=self.clip_denoised, quantize_denoised=quantize_denoised)
            if mask is not None:
                img = self.q_sample(x0, ts) * mask + (1. - mask) * img
            if i % log_every_t == 0 or i == timesteps - 1:
                intermediates.append(img)
            if callback:
                callback(i)
            if img_callback:
                img_callback(img, i)
        return (img, intermediates) if return_intermediates else img

    @torch.no_grad()
    def sample(self, cond, batch_size=16, return_intermediates=False, x_T=None, verbose=True, timesteps=None,
               quantize_denoised=False, mask=None, x0=None, shape=None, **kwargs):
        if shape is None:
            shape = (batch_size, self.channels, self.image_size, self.image_size)
        if cond is not None:
            if isinstance(cond, dict):
                cond = {k: cond[k][:batch_size] if not isinstance(cond[k], list)
                        else [x[:batch_size] for x in cond[k]] for k in cond}
            else:
                cond = [c[:batch_size] for c in cond] if isinstance(cond, list) else cond[:batch_size]
        return self.p_sample_loop(cond, shape, return_intermediates=return_intermediates, x_T=x_T, verbose=verbose,
                                  timesteps=timesteps, quantize_denoised=quantize_denoised, mask=mask, x0=x0)

    @torch.no_grad()
    def sample_log(self, cond, batch_size, ddim, ddim_steps, **kwargs):
        """ddpm.py:1237-1250."""
        if ddim:
            sampler = DDIMSampler(self)
            shape = (self.channels, self.image_size, self.image_size)
            return sampler.sample(ddim_steps, batch_size, shape, cond, verbose=False, **kwargs)
        return self.sample(cond=cond, batch_size=batch_size, return_intermediates=True, **kwargs)
