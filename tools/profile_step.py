"""Profiling driver: one guided DDIM step (or one VQ decode) at the bench workload, bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly that step.

    python tools/profile_step.py [--batch 64] [--latent 64] [--what unet|decode|style] [--precision bf16]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--what", default="unet")
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    m = bench.build_model(a.latent, 1, a.precision).to(dev).eval()
    model = m._model
    img, seg_oh, style, x_T = [t.to(dev) for t in bench.synthetic_batch(a.batch, 4 * a.latent, 1, 0)]
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    with torch.no_grad():
        batch = m.prepare_batch((img, seg_oh, None, style, None))
        _, c = model.get_input(batch, "image")
        _, cu = model.get_input(dict(batch, style_imgs=torch.zeros_like(batch["style_imgs"]) - 2), "image")
        sampler = DDIMSampler(model, use_cuda_graph=False)
        sampler.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
        ts = torch.full((a.batch,), 481, device=dev, dtype=torch.long)
        step = lambda: sampler.p_sample_ddim(x_T, c, ts, index=24, unconditional_guidance_scale=1.5,
                                             unconditional_conditioning=cu)
        dec = lambda: model.decode_first_stage(x_T * 60)
        sty_imgs = (torch.rand(a.batch, 1, 4 * a.latent, 4 * a.latent, 3, device=dev) * 2 - 1)
        sty = lambda: model._agg_block(sty_imgs)      # the native style encoder + aggregation on `batch` images
        fn = {"unet": step, "decode": dec, "style": sty}[a.what]
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.cudart().cudaProfilerStart()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(f"{a.what} B={a.batch} L={a.latent} {a.precision}: {e0.elapsed_time(e1):.3f} ms")


if __name__ == "__main__":
    main()
