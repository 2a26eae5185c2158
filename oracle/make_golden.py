"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, via oracle/ref_shims.py)
on fixture weights, and assert that oracle/stedm_oracle.py reproduces every tensor.

Run once in the build container (the reference cannot travel to the GPU box):

    python -m oracle.make_golden [--only small|c1|sched|svit|plms|st|dpm|ancestral]

TEST INFRASTRUCTURE ONLY.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

from oracle import ref_shims
from oracle import stedm_oracle as O
from stedm_b200.utils.fixture import apply_fixture_weights

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def canonical_sd(model):
    """Reference state dict reduced to the names the oracle / engine use (fp32, CPU)."""
    out = {}
    for k, v in model.state_dict().items():
        if k.startswith("_agg_block.") or k.startswith("model_ema."):
            continue
        k = k.replace("agg_block._embedder.", "agg_block.embedder.")
        out[k] = v.detach().float() if v.is_floating_point() else v.detach()
    return out


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def ref_batch(seg, style):
    P = seg.shape[1]
    B = seg.shape[0]
    return {"image": torch.zeros(B, P, P, 3), "segmentation": seg, "style_imgs": style}


@torch.no_grad()
def gen_case(name, B, L, n_style, S, full_steps, seed, store_f16_image):
    t0 = time.time()
    torch.set_num_threads(os.cpu_count())
    P = 4 * L
    model = ref_shims.build_reference_model(latent_size=L, style_sampling="mp" if n_style > 1 else "augmented",
                                            num_patches=n_style)
    apply_fixture_weights(model, seed=0)
    sd = canonical_sd(model)
    seg, style, x_T = O.synthetic_batch(B, P, n_style, seed)
    out = {"B": B, "L": L, "n_style": n_style, "S": S, "seed": seed, "cfg_scale": 1.5, "eta": 0.0}

    # --- conditioning through the reference's own get_input (networks/s_zss_dm.py:45-60)
    z, c = model.get_input(ref_batch(seg, style), "image")
    _, cu = model.get_input(ref_batch(seg, torch.zeros_like(style) - 2), "image")
    oc = O.get_conditioning(sd, seg, style)
    ou = O.get_conditioning(sd, seg, torch.zeros_like(style) - 2)
    for a, b, n in [(c["c_concat"][0], oc["c_concat"][0], "c_concat"), (c["c_crossattn"][0], oc["c_crossattn"][0], "c_crossattn"),
                    (cu["c_crossattn"][0], ou["c_crossattn"][0], "uc_crossattn")]:
        d = maxdiff(a, b)
        print(f"[{name}] oracle vs reference {n}: max|d| = {d:.3e}")
        assert d < 1e-5, n
    out["c_concat"] = c["c_concat"][0].numpy()
    out["c_crossattn"] = c["c_crossattn"][0].numpy()
    out["uc_crossattn"] = cu["c_crossattn"][0].numpy()

    # --- per-step eps at three timesteps (ddpm.py:894 apply_model)
    for t in (981, 481, 1):
        tt = torch.full((B,), t, dtype=torch.long)
        e_c = model.apply_model(x_T, tt, c)
        e_u = model.apply_model(x_T, tt, cu)
        o_c = O.apply_model(sd, x_T, tt, oc)
        d = maxdiff(e_c, o_c)
        print(f"[{name}] oracle vs reference eps_c t={t}: max|d| = {d:.3e}  (|eps|max {float(e_c.abs().max()):.3f}, std {float(e_c.std()):.3f})")
        assert d < 2e-4
        out[f"eps_c_{t}"] = e_c.numpy()
        out[f"eps_u_{t}"] = e_u.numpy()

    # --- DDIM loop through the reference's sample_log (ddpm.py:1237-1250, ddim.py)
    from ldm.models.diffusion.ddim import DDIMSampler
    sampler = DDIMSampler(model)
    sampler.make_schedule(ddim_num_steps=S, ddim_eta=0.0, verbose=False)
    xs = {}
    img = x_T
    total = sampler.ddim_timesteps.shape[0]
    n_run = total if full_steps else 3
    for i, step in enumerate(np.flip(sampler.ddim_timesteps)[:n_run]):
        index = total - i - 1
        ts = torch.full((B,), int(step), dtype=torch.long)
        img, pred_x0 = sampler.p_sample_ddim(img, c, ts, index=index, unconditional_guidance_scale=1.5,
                                             unconditional_conditioning=cu)
        if i in (0, 1, 2, 9, 24, total - 1):
            xs[i] = img.clone()
            if i == 0:
                out["pred_x0_step0"] = pred_x0.numpy()
    for i, v in xs.items():
        out[f"x_after_{i + 1}"] = v.numpy()
    print(f"[{name}] reference DDIM ran {n_run} steps, |x| max {float(img.abs().max()):.3f} std {float(img.std()):.3f}  ({time.time() - t0:.0f}s)")
    # oracle trajectory for the first 3 steps
    zo, _ = O.ddim_sample(sd, oc, ou, x_T, S=S, cfg_scale=1.5, max_steps=3)
    d = maxdiff(zo, xs[2])
    print(f"[{name}] oracle vs reference x after 3 steps: max|d| = {d:.3e}")
    assert d < 1e-3
    if full_steps:
        # the reference's own top-level entry must agree with the manual loop above
        z_ref, _ = model.sample_log(c, batch_size=B, ddim=True, ddim_steps=S, eta=0.0, log_every_t=1000, x_T=x_T,
                                    unconditional_conditioning=cu, unconditional_guidance_scale=1.5)
        assert maxdiff(z_ref, img) == 0.0
    z_fin = img
    out["z_final"] = z_fin.numpy()
    out["n_steps_run"] = n_run

    # --- first-stage decode (ddpm.py:708-766) with quantisation on and off
    dec_q = model.decode_first_stage(z_fin)
    dec_n = model.decode_first_stage(z_fin, force_not_quantize=True)
    o_q = O.decode_first_stage(sd, z_fin)
    o_n = O.decode_first_stage(sd, z_fin, force_not_quantize=True)
    print(f"[{name}] oracle vs reference decode: quant {maxdiff(dec_q, o_q):.3e}  noquant {maxdiff(dec_n, o_n):.3e}"
          f"  |img| max {float(dec_q.abs().max()):.3f}")
    assert maxdiff(dec_q, o_q) < 1e-3 and maxdiff(dec_n, o_n) < 1e-3
    _, idx = O.vq_quantize(z_fin, sd["first_stage_model.quantize.embedding.weight"])
    out["vq_idx"] = idx.numpy().astype(np.int32)
    out["n_codes_used"] = int(idx.unique().numel())
    if store_f16_image:
        out["dec_quant"] = dec_q.numpy().astype(np.float16)
        out["dec_noquant"] = dec_n.numpy().astype(np.float16)
    else:
        out["dec_quant"] = dec_q.numpy()
        out["dec_noquant"] = dec_n.numpy()
    out["img_u8"] = O.to_uint8(dec_q)
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)
    print(f"[{name}] wrote {name}.npz  codes used {out['n_codes_used']}  total {time.time() - t0:.0f}s")


def gen_sched():
    """Known answers of the schedule, from the reference's own functions."""
    ref_shims.install()
    from ldm.modules.diffusionmodules.util import (make_beta_schedule, make_ddim_timesteps,
                                                   make_ddim_sampling_parameters, timestep_embedding)
    betas = make_beta_schedule("linear", 1000, linear_start=0.0015, linear_end=0.0205)
    ac = torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)
    out = {"alphas_cumprod": ac.numpy()}
    for S in (50, 128, 20):
        ts = make_ddim_timesteps("uniform", S, 1000, verbose=False)
        sig, a, ap = make_ddim_sampling_parameters(ac, ts, 0.0, verbose=False)
        out[f"ts_{S}"] = ts
        out[f"a_{S}"] = a.numpy()
        out[f"a_prev_{S}"] = np.asarray(ap, dtype=np.float64)
        out[f"sqrt1m_{S}"] = np.sqrt(1.0 - a).numpy()
        tab = O.ddim_tables(S)
        assert (tab["timesteps"] == ts).all()
        assert (tab["a_t"] == a.numpy()).all() and (tab["a_prev"] == np.asarray(ap, dtype=np.float32)).all()
        assert (tab["sqrt_one_minus_a"] == out[f"sqrt1m_{S}"]).all()
    sig, a, ap = make_ddim_sampling_parameters(ac, out["ts_50"], 0.5, verbose=False)
    out["sigma_50_eta05"] = np.asarray(sig, dtype=np.float64)
    assert np.allclose(O.ddim_tables(50, eta=0.5)["sigma"], out["sigma_50_eta05"].astype(np.float32), rtol=1e-6)
    t = torch.tensor([981, 481, 1])
    out["temb_128"] = timestep_embedding(t, 128).numpy()
    assert maxdiff(torch.from_numpy(out["temb_128"]), O.timestep_embedding(t, 128)) == 0.0
    np.savez_compressed(os.path.join(GOLD, "sched.npz"), **out)
    print("[sched] wrote sched.npz; len(ts_128) =", len(out["ts_128"]))


SVIT_CASES = {  # name -> (image size, style images per sample, pool, seed); conf/style_agg/svit.yaml otherwise
    "svit_p64_n2_mean": (64, 2, "mean", 11),
    "svit_p128_n1_cls": (128, 1, "cls", 12),
}
SVIT_KW = dict(patch_size=8, num_classes=512, dim=256, depth=6, heads=12, mlp_dim=256, channels=3, dropout=0.1,
               emb_dropout=0.1, t_dim=256)


@torch.no_grad()
def gen_svit():
    """style_agg=svit: the reference's networks/vit_set.py sViT (as networks/s_zss_dm.py:31-38 builds it) on fixture
    weights; asserts oracle.svit_aggregate == reference and stores the [B, 512] style vectors."""
    ref_shims.install()
    from networks.vit_set import sViT
    out = {}
    for name, (P, ns, pool, seed) in SVIT_CASES.items():
        holder = torch.nn.Module()
        holder.agg_block = sViT(image_size=P, ns=ns, pool=pool, **SVIT_KW).eval()
        apply_fixture_weights(holder, seed=0)
        sd = {k: v.detach().float() for k, v in holder.state_dict().items()}
        _, style, _ = O.synthetic_batch(2, P, ns, seed)
        want = holder.agg_block(style)
        got = O.svit_aggregate(sd, style, heads=SVIT_KW["heads"], patch=SVIT_KW["patch_size"], pool=pool)
        d = maxdiff(want, got)
        print(f"[{name}] oracle vs reference sViT: max|d| = {d:.3e} (|out|max {float(want.abs().max()):.3f})")
        assert d < 1e-5
        out[name] = want.numpy()
    np.savez_compressed(os.path.join(GOLD, "svit.npz"), **out)
    print("[svit] wrote svit.npz")


ST_UNET_KW = dict(image_size=32, in_channels=6, out_channels=3, model_channels=128, attention_resolutions=[32, 16, 8],
                  num_res_blocks=2, channel_mult=[1, 4, 8], num_heads=8, use_spatial_transformer=True, context_dim=1024)


@torch.no_grad()
def gen_spatial_transformer():
    """use_spatial_transformer=True: (1) the reference UNetModel with a SpatialTransformer in middle_block[2]
    (openaimodel.py:648-652; context_dim = block width, the only value the reference can run, since the block is called
    without a context) and (2) the stand-alone reference SpatialTransformer with a (B, N, 512) context — cross-attention
    over style tokens (attention.py:245-261).  Asserts oracle == reference; stores eps / outputs."""
    ref_shims.install()
    from ldm.modules.attention import SpatialTransformer
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    out = {}
    holder = torch.nn.Module()
    holder.model = torch.nn.Module()
    holder.model.diffusion_model = UNetModel(**ST_UNET_KW).eval()
    apply_fixture_weights(holder, seed=0)
    sd = {k: v.detach().float() for k, v in holder.state_dict().items()}
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 6, 32, 32, generator=g)
    ctx = torch.randn(2, 512, generator=g) * 0.5
    for t in (981, 1):
        tt = torch.full((2,), t, dtype=torch.long)
        want = holder.model.diffusion_model(x, tt, ctx)
        got = O.unet_forward(sd, x, tt, ctx)
        d = maxdiff(want, got)
        print(f"[st_unet t={t}] oracle vs reference: max|d| = {d:.3e} (|eps|max {float(want.abs().max()):.3f})")
        assert d < 2e-4
        out[f"unet_eps_{t}"] = want.numpy()
    h2 = torch.nn.Module()
    h2.st = SpatialTransformer(256, 4, 64, depth=2, context_dim=512).eval()
    apply_fixture_weights(h2, seed=0)
    sd2 = {k: v.detach().float() for k, v in h2.state_dict().items()}
    xs = torch.randn(2, 256, 16, 16, generator=g)
    for name, n_tok in (("st_ctx10", 10), ("st_ctx1", 1)):
        c = torch.randn(2, n_tok, 512, generator=g)
        want = h2.st(xs, c)
        got = O.spatial_transformer(xs, sd2, "st.", 4, c)
        d = maxdiff(want, got)
        print(f"[{name}] oracle vs reference SpatialTransformer: max|d| = {d:.3e} (|out|max {float(want.abs().max()):.3f})")
        assert d < 1e-4
        out[name] = want.numpy()
    np.savez_compressed(os.path.join(GOLD, "spatial_transformer.npz"), **out)
    print("[st] wrote spatial_transformer.npz")


@torch.no_grad()
def gen_dpm():
    """DPM-Solver on the REAL reference model (it takes STEDM's dict conditioning, dpm_solver.py:306-316): the reference's
    DPMSolverSampler on S_ZSS_DM with fixture weights, S = 12 (second-order multistep, first-order final step) and
    S = 16; asserts oracle.dpm_solver_sample == reference and stores the final latents."""
    import contextlib
    import io
    model = ref_shims.build_reference_model(latent_size=32, style_sampling="mp", num_patches=2)
    apply_fixture_weights(model, seed=0)
    sd = canonical_sd(model)
    from ldm.models.diffusion.dpm_solver.sampler import DPMSolverSampler
    g = load_small_cond()
    cond = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["c_crossattn"])]}
    unc = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["uc_crossattn"])]}
    _, _, x_T = O.synthetic_batch(2, 128, 2, 0)
    out = {}
    for S in (12, 16):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            want, _ = DPMSolverSampler(model, device=torch.device("cpu")).sample(
                S, 2, (3, 32, 32), conditioning=cond, verbose=False, x_T=x_T, unconditional_guidance_scale=1.5,
                unconditional_conditioning=unc)
        got = O.dpm_solver_sample(lambda x, t: O.apply_model(sd, x, t, cond), x_T, model.alphas_cumprod.numpy(), S,
                                  cfg_scale=1.5, uncond_eps_fn=lambda x, t: O.apply_model(sd, x, t, unc))
        d = maxdiff(want, got)
        print(f"[dpm S={S}] oracle vs reference DPMSolverSampler: max|d| = {d:.3e} (|x|max {float(want.abs().max()):.3f})")
        assert d < 1e-3 * float(want.abs().max())
        out[f"dpm_s{S}"] = want.numpy()
    np.savez_compressed(os.path.join(GOLD, "dpm_solver.npz"), **out)
    print("[dpm] wrote dpm_solver.npz")


@torch.no_grad()
def gen_ancestral():
    """DDPM ancestral sampler (kept API): the reference's LatentDiffusion.sample(cond, timesteps=5) on
    fixture weights (ddpm.py:1050-1110, 1168-1235), plain and with quantize_denoised=True.  The per-step N(0,1) draws
    are fixed: the reference's noise_like is replaced by a reader of pre-drawn tensors (seed 7), stored in the golden so
    the GPU test can feed the very same noise.  Asserts oracle.ddpm_ancestral_sample == reference."""
    model = ref_shims.build_reference_model(latent_size=32, style_sampling="mp", num_patches=2)
    apply_fixture_weights(model, seed=0)
    sd = canonical_sd(model)
    import ldm.models.diffusion.ddpm as ref_ddpm
    g = load_small_cond()
    cond = {"c_concat": [torch.from_numpy(g["c_concat"])], "c_crossattn": [torch.from_numpy(g["c_crossattn"])]}
    _, _, x_T = O.synthetic_batch(2, 128, 2, 0)
    T = 5
    noises = torch.randn(T, 2, 3, 32, 32, generator=torch.Generator().manual_seed(7))
    out = {"timesteps": T, "noises": noises.numpy()}
    orig = ref_ddpm.noise_like
    try:
        for name, scale, quant in (("ancestral_t5", 1.0, False), ("ancestral_t5_quant", 40.0, True)):
            it = iter(noises)
            ref_ddpm.noise_like = lambda shape, device, repeat=False: next(it).clone()
            want = model.sample(cond, batch_size=2, x_T=x_T * scale, timesteps=T, verbose=False, quantize_denoised=quant)
            got = O.ddpm_ancestral_sample(sd, cond, x_T * scale, T, noises, quantize_denoised=quant)
            d = maxdiff(want, got)
            print(f"[{name}] oracle vs reference LatentDiffusion.sample: max|d| = {d:.3e} (|x|max {float(want.abs().max()):.3f})")
            assert d < 1e-4 * max(1.0, float(want.abs().max()))
            out[name] = want.numpy()
    finally:
        ref_ddpm.noise_like = orig
    np.savez_compressed(os.path.join(GOLD, "ancestral.npz"), **out)
    print("[ancestral] wrote ancestral.npz")


def load_small_cond():
    z = np.load(os.path.join(GOLD, "small_b2_l32.npz"))
    return {k: z[k] for k in ("c_concat", "c_crossattn", "uc_crossattn")}


class _StandInModel:
    """The minimum a reference sampler touches (plms.py:12-56, 177-191), with a deterministic non-linear eps so that the
    sampler arithmetic — timestep sequence, multistep coefficients, guidance combine, x_prev update — can be run through
    the REFERENCE PLMSSampler, which cannot take STEDM's dict conditioning (torch.cat of conditionings, plms.py:179)."""

    def __init__(self):
        ac, _ = O.alphas_cumprod_linear()
        self.num_timesteps = 1000
        self.alphas_cumprod = torch.from_numpy(ac)
        self.alphas_cumprod_prev = torch.cat([torch.ones(1), self.alphas_cumprod[:-1]])
        self.betas = torch.from_numpy(np.linspace(0.0015 ** 0.5, 0.0205 ** 0.5, 1000, dtype=np.float64) ** 2).float()
        self.device = torch.device("cpu")

    @staticmethod
    def eps(x, t, c):
        return torch.tanh(0.3 * x + c) * (1.0 + t.float().view(-1, 1, 1, 1) / 1000.0)

    def apply_model(self, x, t, c):
        return self.eps(x, t, c)


@torch.no_grad()
def gen_plms():
    """PLMS sampler arithmetic: reference PLMSSampler on the stand-in model == oracle.plms_sample; stores the result."""
    ref_shims.install()
    import contextlib
    import io
    from ldm.models.diffusion.plms import PLMSSampler
    PLMSSampler.register_buffer = lambda self, n, a: setattr(self, n, a)        # the original hard-codes .to("cuda")
    model = _StandInModel()
    g = torch.Generator().manual_seed(21)
    x_T = torch.randn(2, 3, 8, 8, generator=g)
    c, uc = torch.randn(2, 3, 8, 8, generator=g) * 0.5, torch.zeros(2, 3, 8, 8)
    out = {}
    for name, scale in (("plms_s20_cfg1", 1.0), ("plms_s20_cfg3", 3.0)):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            want, _ = PLMSSampler(model).sample(20, 2, (3, 8, 8), conditioning=c, verbose=False, x_T=x_T, eta=0.0,
                                                unconditional_guidance_scale=scale,
                                                unconditional_conditioning=None if scale == 1.0 else uc)
        got = O.plms_sample(lambda x, t: model.eps(x, t, c), x_T, S=20, cfg_scale=scale,
                            uncond_eps_fn=None if scale == 1.0 else (lambda x, t: model.eps(x, t, uc)))
        d = maxdiff(want, got)
        print(f"[{name}] oracle vs reference PLMSSampler: max|d| = {d:.3e}")
        assert d < 1e-5
        out[name] = want.numpy()
    np.savez_compressed(os.path.join(GOLD, "plms.npz"), **out)
    print("[plms] wrote plms.npz")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if a.only in ("all", "sched"):
        gen_sched()
    if a.only in ("all", "svit"):
        gen_svit()
    if a.only in ("all", "plms"):
        gen_plms()
    if a.only in ("all", "st"):
        gen_spatial_transformer()
    if a.only in ("all", "dpm"):
        gen_dpm()
    if a.only in ("all", "ancestral"):
        gen_ancestral()
    if a.only in ("all", "small"):
        # B=2, latent 32 (128^2 image), two style images per sample (exercises Agg_Mean), full DDIM-50
        gen_case("small_b2_l32", B=2, L=32, n_style=2, S=50, full_steps=True, seed=0, store_f16_image=False)
    if a.only in ("all", "c1"):
        # BASELINE config[0]: batch 4, 256^2, DDIM-50, cfg 1.5
        gen_case("c1_b4_l64", B=4, L=64, n_style=1, S=50, full_steps=True, seed=1, store_f16_image=True)
