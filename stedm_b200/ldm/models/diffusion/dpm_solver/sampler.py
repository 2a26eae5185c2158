"""Drop-in for ``ldm.models.diffusion.dpm_solver.sampler.DPMSolverSampler`` (reference dpm_solver/sampler.py:12-95) on the
native engine, in the one configuration the reference uses: ``DPM_Solver(predict_x0=True, thresholding=False).sample(
steps=S, skip_type="time_uniform", method="multistep", order=2, lower_order_final=True)`` over the discrete VP schedule
built from ``alphas_cumprod``, with plain classifier-free guidance (dpm_solver.py:302-320 — this sampler DOES take STEDM's
dict conditioning in the reference, unlike PLMS).

Per model evaluation: one batched (cond | uncond) native U-Net pass at the FRACTIONAL model time
t = (t_continuous - 1/N) * 1000 (``stedm_timestep_embedding_f32`` through the runner's per-timestep cache), then a handful
of elementwise updates (data prediction + first / second order multistep update; S evaluations in total, so they are not
fused).  Schedule scalars (log alpha interpolation, sigma, lambda, h, expm1) are float64 on the host, rounded to fp32 as
multipliers — the reference evaluates the same expressions in fp32 tensors.
"""
import math

import numpy as np
import torch

from ..ddim import _GuidedStepper


class _DiscreteVP:
    """NoiseScheduleVP('discrete'), dpm_solver.py:78-88, 106-138."""

    def __init__(self, alphas_cumprod):
        ac = alphas_cumprod.detach().double().cpu().numpy()
        self.N = ac.shape[0]
        self.t = np.arange(1, self.N + 1, dtype=np.float64) / self.N
        self.la = 0.5 * np.log(ac)

    def log_alpha(self, t):
        return float(np.interp(t, self.t, self.la))

    def alpha(self, t):
        return math.exp(self.log_alpha(t))

    def sigma(self, t):
        return math.sqrt(1.0 - math.exp(2.0 * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * math.log(1.0 - math.exp(2.0 * la))


class DPMSolverSampler(object):
    def __init__(self, model, device=None, use_cuda_graph=None, share_trunk=None, **kwargs):
        super().__init__()
        self.model = model
        self.device = device
        self.use_cuda_graph = getattr(model, "use_cuda_graph", False) if use_cuda_graph is None else use_cuda_graph
        self.share_trunk = getattr(model, "share_trunk", True) if share_trunk is None else share_trunk
        self.alphas_cumprod = model.alphas_cumprod.detach().float()

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        C, H, W = shape
        size = (batch_size, C, H, W)
        device = self.model.betas.device
        x = torch.randn(size, device=device) if x_T is None else x_T.to(device).float()
        ns = _DiscreteVP(self.alphas_cumprod)
        stepper = _GuidedStepper(self, conditioning, unconditional_conditioning, unconditional_guidance_scale, size)
        t_dummy = torch.zeros((batch_size,), dtype=torch.long, device=device)
        ts = np.linspace(1.0, 1.0 / ns.N, S + 1)                      # get_time_steps('time_uniform'), :402-403

        def x0_of(x, t):                                              # data_prediction_fn, dpm_solver.py:361-374
            eps = stepper._eps(x.contiguous(), t_dummy, uniform_t=True, t_value=float((t - 1.0 / ns.N) * 1000.0))
            if stepper.guided:
                e_c, e_u = eps[:batch_size], eps[batch_size:]
                eps = torch.add(e_u, e_c - e_u, alpha=stepper.scale)
            return (x - ns.sigma(t) * eps) / ns.alpha(t)

        def first(x, s, t, m_s):                                      # dpm_solver_first_update, :478-510
            h = ns.lam(t) - ns.lam(s)
            return (ns.sigma(t) / ns.sigma(s)) * x - (ns.alpha(t) * math.expm1(-h)) * m_s

        def second(x, t_p1, t_p0, t, m_p1, m_p0):                     # multistep_dpm_solver_second_update, :732-768
            h0, h = ns.lam(t_p0) - ns.lam(t_p1), ns.lam(t) - ns.lam(t_p0)
            d1 = (1.0 / (h0 / h)) * (m_p0 - m_p1)
            c = ns.alpha(t) * (math.exp(-h) - 1.0)
            return (ns.sigma(t) / ns.sigma(t_p0)) * x - c * m_p0 - 0.5 * c * d1

        assert S >= 2, "multistep DPM-Solver of order 2 needs at least 2 steps"
        m_prev, t_prev = [x0_of(x, ts[0])], [ts[0]]
        x = first(x, ts[0], ts[1], m_prev[0])
        m_prev.append(x0_of(x, ts[1]))
        t_prev.append(ts[1])
        for step in range(2, S + 1):                                  # sample(method='multistep'), :1058-1085
            order = min(2, S + 1 - step) if S < 15 else 2
            t = ts[step]
            x = second(x, t_prev[0], t_prev[1], t, m_prev[0], m_prev[1]) if order == 2 else first(x, t_prev[1], t, m_prev[1])
            t_prev, m_prev = [t_prev[1], t], [m_prev[1], None]
            if step < S:
                m_prev[1] = x0_of(x, t)
        return x.to(device), None
