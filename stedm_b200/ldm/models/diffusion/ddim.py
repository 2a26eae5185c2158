"""Drop-in for ``ldm.models.diffusion.ddim.DDIMSampler`` (reference ddim.py:11-210) — the hot loop of the path.

Same constructor, ``make_schedule``, ``sample``, ``ddim_sampling`` and ``p_sample_ddim`` signatures and return
structures ((x, {"x_inter": [...], "pred_x0": [...]})).  What changes is how a step is executed:

* the conditional and unconditional eps predictions of STEDM's guidance (two separate ``apply_model`` calls in the
  reference, ddim.py:177-178) run as ONE batched U-Net pass of 2B samples on the native engine — every op in the
  network is per-sample, so the results are those of the two separate calls;
* everything after them — guidance combine, the (C,H)-std rescale with phi = 0.7, pred_x0, direction and x_{t-1}
  (ddim.py:179-209, ~25 ATen kernels + 4 scalar fills in the reference) — is the single ``stedm_cfg_ddim_step``
  kernel;
* schedule scalars are Python floats rounded through fp32 exactly like the reference's ``torch.full`` (§A.4);
* with ``use_cuda_graph`` the U-Net pass is captured once per (batch, latent size) and replayed every step.
"""
import numpy as np
import torch

from ...modules.diffusionmodules.util import make_ddim_sampling_parameters, make_ddim_timesteps, noise_like
from .... import ops


def _f32(v):
    """Round a python/numpy scalar through float32 (what torch.full(..., device=cuda) does in the reference)."""
    return float(np.float32(v))


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", use_cuda_graph=None, share_trunk=None, **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.use_cuda_graph = getattr(model, "use_cuda_graph", False) if use_cuda_graph is None else use_cuda_graph
        self.share_trunk = getattr(model, "share_trunk", True) if share_trunk is None else share_trunk

    def register_buffer(self, name, attr):
        if isinstance(attr, torch.Tensor) and attr.device != self.model.device:
            attr = attr.to(self.model.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        """ddim.py:24-53 (same buffer names; numpy members stay numpy like the reference's)."""
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        ac = self.model.alphas_cumprod
        assert ac.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        to_torch = lambda x: torch.as_tensor(x).clone().detach().to(torch.float32).to(self.model.device)
        ac_cpu = ac.detach().cpu().numpy()           # fp32 ndarray: the reference's np.sqrt(alphas_cumprod.cpu())
        self.register_buffer("betas", to_torch(self.model.betas))
        self.register_buffer("alphas_cumprod", to_torch(ac))
        self.register_buffer("alphas_cumprod_prev", to_torch(self.model.alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(ac_cpu)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1. - ac_cpu)))
        self.register_buffer("log_one_minus_alphas_cumprod", to_torch(np.log(1. - ac_cpu)))
        self.register_buffer("sqrt_recip_alphas_cumprod", to_torch(np.sqrt(1. / ac_cpu)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", to_torch(np.sqrt(1. / ac_cpu - 1)))
        sig, a, a_prev = make_ddim_sampling_parameters(alphacums=ac_cpu, ddim_timesteps=self.ddim_timesteps,
                                                       eta=ddim_eta, verbose=verbose)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sig, a, a_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - a)                # fp32 ndarray (a is fp32)
        self.register_buffer("ddim_sigmas_for_original_num_steps", ddim_eta * torch.sqrt(
            (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) * (1 - self.alphas_cumprod / self.alphas_cumprod_prev)))

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None):
        """ddim.py:112-162."""
        if score_corrector is not None:
            raise NotImplementedError("score correctors are not on STEDM's path")
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device).float()
        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = list(reversed(range(0, timesteps))) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]
        stepper = _GuidedStepper(self, cond, unconditional_conditioning, unconditional_guidance_scale, shape)
        # The whole loop as ONE CUDA graph (ddim.py:139-160 unrolled: every U-Net pass, the fused guidance + DDIM update
        # and nothing else — no per-step copy, clone, fill or launch from Python): whenever a step does not depend on
        # host-side state, i.e. the plain eta = 0 sampler without callbacks, in-painting mask or codebook snapping.
        if (stepper.use_graph and mask is None and callback is None and img_callback is None and not quantize_denoised
                and not ddim_use_original_steps and noise_dropout == 0. and img.is_cuda
                and not np.any(np.asarray(self.ddim_sigmas[:total_steps]) != 0)):
            out = stepper.run_loop_graph(img.float().contiguous(), [int(t) for t in time_range], total_steps, log_every_t)
            if out is not None:
                img, logged = out
                for x_i, p_i in logged:
                    intermediates["x_inter"].append(x_i)
                    intermediates["pred_x0"].append(p_i)
                for _ in range(total_steps):      # the reference draws randn every step, also at sigma = 0 (ddim.py:206):
                    noise_like(shape, device, False)   # keep the generator where the reference leaves it
                return img, intermediates
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img = self.model.q_sample(x0, ts) * mask + (1. - mask) * img
            # ts is torch.full(step): one timestep for the whole batch -> embeddings are computed for one row
            img, pred_x0 = stepper.step(img, ts, index, temperature=temperature, noise_dropout=noise_dropout,
                                        uniform_t=True, use_original_steps=ddim_use_original_steps,
                                        quantize_denoised=quantize_denoised, t_value=int(step))
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, rescale_phi=0.7):
        """ddim.py:164-210, one step (kept for callers that drive the loop themselves)."""
        if score_corrector is not None:
            raise NotImplementedError("score correctors are not on STEDM's path")
        stepper = _GuidedStepper(self, c, unconditional_conditioning, unconditional_guidance_scale, tuple(x.shape),
                                 rescale_phi=rescale_phi, allow_graph=False)
        return stepper.step(x, t, index, temperature=temperature, noise_dropout=noise_dropout,
                            repeat_noise=repeat_noise, use_original_steps=use_original_steps,
                            quantize_denoised=quantize_denoised)


class _GuidedStepper:
    """One DDIM step = batched (cond ‖ uncond) U-Net pass + the fused K11 kernel."""

    def __init__(self, sampler, cond, uncond, scale, shape, rescale_phi=0.7, allow_graph=True):
        self.s = sampler
        self.model = sampler.model
        self.scale, self.phi = float(scale), float(rescale_phi)
        self.guided = not (uncond is None or scale == 1.)
        self.b = shape[0]
        unet = self.model.model.diffusion_model
        self.unet = unet
        cc = lambda c: c["c_concat"][0] if len(c["c_concat"]) == 1 else torch.cat(c["c_concat"], 1)
        ca = lambda c: c["c_crossattn"][0] if len(c["c_crossattn"]) == 1 else torch.cat(c["c_crossattn"], 1)
        if not isinstance(cond, dict):
            raise NotImplementedError("STEDM's hybrid conditioning is a dict {'c_concat': [...], 'c_crossattn': [...]}")
        self.shared = False
        if self.guided:
            self.context = torch.cat([ca(cond), ca(uncond)], 0).float().contiguous()
            # cond and uncond differ only in the style vector when their layouts agree (always so in predict_step):
            # then the encoder trunk is evaluated once for both (bit-identical)
            self.shared = (getattr(sampler, "share_trunk", True) and cc(cond).shape == cc(uncond).shape
                           and bool(torch.equal(cc(cond), cc(uncond))) and unet.shared_trunk_ok(self.b, shape[-1]))
            if self.shared:
                self.c_concat = cc(cond).float().contiguous()
            else:
                self.c_concat = torch.cat([cc(cond), cc(uncond)], 0).float().contiguous()
        else:
            self.c_concat = cc(cond).float().contiguous()
            self.context = ca(cond).float().contiguous()
        self.use_graph = allow_graph and sampler.use_cuda_graph
        self._graph_inputs_stale = True
        self._emb_style = None

    def _eps(self, x, t, uniform_t=False, t_value=None):
        """eps for the (cond ‖ uncond) batch.  x (B,3,L,L), t (B,).  ``t_value``: the loop's python timestep when the
        whole batch shares it: the timestep embeddings then come from the runner's per-timestep cache and the style
        embedding is computed once per loop — no embedding launch is left in the step (or in its CUDA graph)."""
        if self.guided and not self.shared:
            x2 = torch.cat([x, x], 0)
            t2 = torch.cat([t, t], 0)
        else:
            x2, t2 = x, t
        emb = None
        if uniform_t and t_value is not None:
            runner = self.unet.runner()
            if self._emb_style is None:
                self._emb_style = runner.style_embedding(self.context)
            emb = (runner.time_embedding_row(t_value, x.device), self._emb_style)
        if not self.use_graph:
            return self.unet.forward_split(x2, self.c_concat, t2, self.context, uniform_t=uniform_t, emb=emb)
        # CUDA graph of the U-Net pass (worth it only when a pass is launch-bound, i.e. small batches): captured
        # once per shape signature and cached on the U-Net module, with static input buffers, so later sampler
        # objects (sample_log builds a new DDIMSampler per call, like the reference) replay instead of re-capturing.
        cache = self.unet.__dict__.setdefault("_graph_cache", {})
        key = (tuple(x2.shape), tuple(self.c_concat.shape), tuple(self.context.shape), self.unet.precision, uniform_t,
               emb is not None)
        ent = cache.get(key)
        if ent is None:
            warm = cache.setdefault(("warm",) + key, [0])
            if warm[0] < 1:   # one eager pass first: lazy kernel attribute setup must not happen under capture
                warm[0] += 1
                return self.unet.forward_split(x2, self.c_concat, t2, self.context, uniform_t=uniform_t, emb=emb)
            ent = {"x": x2.clone(), "t": t2.clone(), "cc": self.c_concat.clone(), "ctx": self.context.clone(),
                   "graph": torch.cuda.CUDAGraph()}
            if emb is not None:
                ent["emb"] = (emb[0].clone(), emb[1].clone())
            n0 = ops.LAUNCHES[0]
            with torch.cuda.graph(ent["graph"]):
                ent["eps"] = self.unet.forward_split(ent["x"], ent["cc"], ent["t"], ent["ctx"], uniform_t=uniform_t,
                                                     emb=ent.get("emb"))
            ent["launches"] = ops.LAUNCHES[0] - n0
            ops.LAUNCHES[0] = n0                      # capture enqueued nothing; replays are counted below
            cache[key] = ent
        if self._graph_inputs_stale:
            ent["cc"].copy_(self.c_concat)
            ent["ctx"].copy_(self.context)
            if emb is not None:
                ent["emb"][1].copy_(emb[1])
            self._graph_inputs_stale = False
        ent["x"].copy_(x2)
        if emb is not None:
            ent["emb"][0].copy_(emb[0])
        else:
            ent["t"].copy_(t2)
        ent["graph"].replay()
        ops.LAUNCHES[0] += ent["launches"]
        return ent["eps"].clone()

    def run_loop_graph(self, x_T, steps, total_steps, log_every_t):
        """All ``len(steps)`` guided steps captured in one CUDA graph, cached on the U-Net module per (shapes, schedule,
        guidance) signature and replayed with fresh x_T / conditioning.  Returns (x_0, [(x, pred_x0) logged steps]) as
        fresh tensors, or None on the first call for a signature (the caller runs that loop eagerly: lazy kernel
        attribute setup and the per-timestep embedding cache must not happen under capture)."""
        s, unet = self.s, self.unet
        sched = tuple((_f32(s.ddim_alphas[i]), _f32(s.ddim_alphas_prev[i]), _f32(s.ddim_sqrt_one_minus_alphas[i]))
                      for i in range(total_steps))
        key = ("loop", tuple(x_T.shape), tuple(self.c_concat.shape), tuple(self.context.shape), unet.precision,
               self.guided, self.shared, self.scale, self.phi, tuple(steps), sched, log_every_t)
        cache = unet.__dict__.setdefault("_graph_cache", {})
        ent = cache.get(key)
        if ent is None:
            warm = cache.setdefault(("warm",) + key, [0])
            if warm[0] < 1:
                warm[0] += 1
                return None
            runner = unet.runner()
            rows = [runner.time_embedding_row(t, x_T.device) for t in steps]     # cached by the eager warm-up loop
            ent = {"x": x_T.clone(), "cc": self.c_concat.clone(), "ctx": self.context.clone(),
                   "t": torch.zeros((x_T.shape[0] * (2 if self.guided and not self.shared else 1),), dtype=torch.long,
                                    device=x_T.device),
                   "graph": torch.cuda.CUDAGraph()}
            n0 = ops.LAUNCHES[0]
            with torch.cuda.graph(ent["graph"]):
                emb_style = runner.style_embedding(ent["ctx"])
                x, logged = ent["x"], []
                for i, row in enumerate(rows):
                    index = total_steps - i - 1
                    x2 = torch.cat([x, x], 0) if (self.guided and not self.shared) else x
                    eps = unet.forward_split(x2, ent["cc"], ent["t"], ent["ctx"], uniform_t=True, emb=(row, emb_style))
                    e_c, e_u = (eps[:self.b], eps[self.b:]) if self.guided else (eps, None)
                    a_t, a_prev, sq1m = sched[index]
                    x, pred_x0 = ops.cfg_ddim_step(e_c, e_u, x, a_t, a_prev, 0.0, sq1m, cfg_scale=self.scale, phi=self.phi)
                    if index % log_every_t == 0 or index == total_steps - 1:
                        logged.append((x, pred_x0))
                ent["out"], ent["logged"] = x, logged
            ent["launches"] = ops.LAUNCHES[0] - n0
            ops.LAUNCHES[0] = n0                      # capture enqueued nothing; replays are counted below
            cache[key] = ent
        ent["x"].copy_(x_T)
        ent["cc"].copy_(self.c_concat)
        ent["ctx"].copy_(self.context)
        ent["graph"].replay()
        ops.LAUNCHES[0] += ent["launches"]
        return ent["out"].clone(), [(a.clone(), b.clone()) for a, b in ent["logged"]]

    def step(self, x, t, index, temperature=1., noise_dropout=0., repeat_noise=False, uniform_t=False,
             use_original_steps=False, quantize_denoised=False, t_value=None):
        s = self.s
        x = x.float().contiguous()
        eps = self._eps(x, t, uniform_t, t_value)
        e_c, e_u = (eps[:self.b], eps[self.b:]) if self.guided else (eps, None)
        if use_original_steps:          # the 1000-step DDPM tables instead of the DDIM subsequence (ddim.py:188-191)
            m = self.model
            a_t, a_prev = float(m.alphas_cumprod[index]), float(m.alphas_cumprod_prev[index])
            sigma = float(s.ddim_sigmas_for_original_num_steps[index])
            sq1m = float(m.sqrt_one_minus_alphas_cumprod[index])
        else:
            a_t, a_prev = _f32(s.ddim_alphas[index]), _f32(s.ddim_alphas_prev[index])
            sigma, sq1m = _f32(s.ddim_sigmas[index]), _f32(s.ddim_sqrt_one_minus_alphas[index])
        # the reference draws randn every step, also when sigma == 0 (ddim.py:206): keep the RNG stream aligned
        noise = noise_like(x.shape, x.device, repeat_noise) * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        x_prev, pred_x0 = ops.cfg_ddim_step(e_c, e_u, x, a_t, a_prev, sigma, sq1m, cfg_scale=self.scale, phi=self.phi,
                                            noise=noise.contiguous() if sigma != 0.0 else None)
        if quantize_denoised:           # ddim.py:200-201: pred_x0 snapped to the VQ codebook before it enters x_prev
            pq = ops.vq_nearest(pred_x0, self.model.first_stage_model.quantize.embedding.weight.detach().float().contiguous())
            x_prev = x_prev + (a_prev ** 0.5) * (pq - pred_x0)     # x_prev = sqrt(a_prev) pred_x0 + dir_xt + noise
            pred_x0 = pq
        return x_prev, pred_x0
