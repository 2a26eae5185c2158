"""Drop-in for ``ldm.models.diffusion.plms.PLMSSampler`` (reference plms.py:10-236) on the native engine: pseudo
linear multistep sampling around the same eps U-Net.

Same constructor, ``make_schedule`` / ``sample`` / ``plms_sampling`` / ``p_sample_plms`` signatures and return
structures.  The eps predictions come from the batched (cond | uncond) native U-Net pass shared with DDIMSampler
(ddim._GuidedStepper); guidance here is the PLAIN classifier-free combine of plms.py:184, e_u + w (e_c - e_u), not
DDIM's std-rescaled one.  The x_{t-1} / pred_x0 update of a (multistep-combined) eps is the unguided mode of the fused
``stedm_cfg_ddim_step`` kernel.

Note on the reference: its ``p_sample_plms`` concatenates the conditionings with ``torch.cat([uc, c])`` (plms.py:179),
which cannot take STEDM's dict conditioning, so the reference cannot run this sampler on its own model; the sampler
arithmetic is pinned against the reference class through a stand-in model instead (oracle/make_golden.py --only plms).
"""
import numpy as np
import torch

from .... import ops
from .ddim import DDIMSampler, _GuidedStepper, _f32
from ...modules.diffusionmodules.util import noise_like


class PLMSSampler(DDIMSampler):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")                       # plms.py:25-26
        return super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.plms_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def plms_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None):
        """plms.py:113-171."""
        if score_corrector is not None or ddim_use_original_steps or quantize_denoised:
            raise NotImplementedError("PLMS: score correctors / original steps / quantize_x0 are not supported natively")
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device).float()
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        stepper = _GuidedStepper(self, cond, unconditional_conditioning, unconditional_guidance_scale, shape)
        old_eps = []
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            ts_next = torch.full((b,), int(time_range[min(i + 1, len(time_range) - 1)]), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img = self.model.q_sample(x0, ts) * mask + (1. - mask) * img
            img, pred_x0, e_t = self._plms_step(stepper, img, ts, index, old_eps, ts_next, temperature, noise_dropout,
                                                False, t_values=(int(step), int(time_range[min(i + 1, len(time_range) - 1)])))
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_plms(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, old_eps=None, t_next=None):
        """plms.py:173-236, one step."""
        if score_corrector is not None or use_original_steps or quantize_denoised:
            raise NotImplementedError("PLMS: score correctors / original steps / quantize_x0 are not supported natively")
        stepper = _GuidedStepper(self, c, unconditional_conditioning, unconditional_guidance_scale, tuple(x.shape),
                                 allow_graph=False)
        return self._plms_step(stepper, x, t, index, old_eps or [], t_next, temperature, noise_dropout, repeat_noise)

    # -------------------------------------------------------------------------------------------------------------
    def _model_output(self, stepper, x, t, t_value=None):
        """get_model_output, plms.py:177-191: plain classifier-free guidance.  ``t_value``: the loop's python timestep
        (whole batch at one t: cached embeddings, no device sync); None: check the tensor."""
        uniform = True if t_value is not None else bool((t == t[0]).all())
        eps = stepper._eps(x.float().contiguous(), t, uniform_t=uniform, t_value=t_value)
        if not stepper.guided:
            return eps
        e_c, e_u = eps[:stepper.b], eps[stepper.b:]
        return torch.add(e_u, e_c - e_u, alpha=stepper.scale)

    def _x_prev(self, x, e, index, temperature, noise_dropout, repeat_noise):
        """get_x_prev_and_pred_x0, plms.py:198-216, on the fused step kernel in its unguided mode."""
        sigma = _f32(self.ddim_sigmas[index])
        noise = noise_like(x.shape, x.device, repeat_noise) * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        return ops.cfg_ddim_step(e.contiguous(), None, x.float().contiguous(), _f32(self.ddim_alphas[index]),
                                 _f32(self.ddim_alphas_prev[index]), sigma, _f32(self.ddim_sqrt_one_minus_alphas[index]),
                                 noise=noise.contiguous() if sigma != 0.0 else None)

    def _plms_step(self, stepper, x, t, index, old_eps, t_next, temperature, noise_dropout, repeat_noise,
                   t_values=(None, None)):
        e_t = self._model_output(stepper, x, t, t_values[0])
        if len(old_eps) == 0:       # pseudo improved Euler (2nd order), plms.py:219-223
            x_prev, _ = self._x_prev(x, e_t, index, temperature, noise_dropout, repeat_noise)
            e_t_next = self._model_output(stepper, x_prev, t_next, t_values[1])
            e_t_prime = (e_t + e_t_next) / 2
        elif len(old_eps) == 1:     # Adams-Bashforth 2nd .. 4th order, plms.py:224-232
            e_t_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_t_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_t_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        x_prev, pred_x0 = self._x_prev(x, e_t_prime, index, temperature, noise_dropout, repeat_noise)
        return x_prev, pred_x0, e_t
