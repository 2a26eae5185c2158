"""Per-launch table of the tensor-core convolutions of ONE guided DDIM step at the bench workload: shape, CUDA-event
time, TFLOP/s (executed FLOPs).  Shows which layers pull the average below the BN=256 figure.
    python tools/conv_table.py [--batch 64] [--latent 64]"""
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
    ap.add_argument("--what", default="unet", help="unet (one guided DDIM step) | decode (one VQ decode)")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    from stedm_b200 import ops
    from stedm_b200.ldm.models.diffusion.ddim import DDIMSampler
    dev = torch.device("cuda", 0)
    m = bench.build_model(a.latent, 1, "bf16").to(dev).eval()
    model = m._model
    img, seg_oh, style, x_T = [t.to(dev) for t in bench.synthetic_batch(a.batch, 4 * a.latent, 1, 0)]
    with torch.no_grad():
        batch = m.prepare_batch((img, seg_oh, None, style, None))
        _, c = model.get_input(batch, "image")
        _, cu = model.get_input(dict(batch, style_imgs=torch.zeros_like(batch["style_imgs"]) - 2), "image")
        s = DDIMSampler(model, use_cuda_graph=False)
        s.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
        ts = torch.full((a.batch,), 481, device=dev, dtype=torch.long)
        step = lambda: s.p_sample_ddim(x_T, c, ts, index=24, unconditional_guidance_scale=1.5, unconditional_conditioning=cu)
        if a.what == "decode":
            step = lambda: model.decode_first_stage(x_T * 60)
        for _ in range(3):
            step()
        rec, orig = [], ops.conv

        def timed(x0, weight, bias, cout, ksize, **kw):
            if not kw.get("tensor_core", True):
                return orig(x0, weight, bias, cout, ksize, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig(x0, weight, bias, cout, ksize, **kw)
            e1.record()
            b, h, w, c0 = x0.shape
            st = kw.get("stride", 1)
            h, w = h // st, w // st                                      # stride 2: GEMM rows = output pixels
            c1 = 0 if kw.get("x1") is None else kw["x1"].shape[-1]
            taps = 4 if kw.get("up_phase") is not None else ksize * ksize
            sk = kw.get("skip_x0")
            skc = 0 if sk is None else sk.shape[-1] + (0 if kw.get("skip_x1") is None else kw["skip_x1"].shape[-1])
            rec.append((e0, e1, b, h, w, c0, c1, cout, taps, kw.get("residual") is not None, kw.get("stats_out") is not None, skc,
                        kw.get("gn_coef") is not None))
            return out

        ops.conv = timed
        clocks = bench.ClockSampler(0)
        clocks.start()
        try:
            reps = a.reps
            for _ in range(reps):
                step()
            torch.cuda.synchronize()
        finally:
            ops.conv = orig
        clk = clocks.stop()
    n = len(rec) // reps
    print(f"{'#':>3s} {'B':>4s} {'HxW':>9s} {'Cin':>10s} {'Cout':>5s} {'taps':>4s} res stats gn {'ms':>8s} {'TFLOP/s':>8s} {'GFLOP':>8s}")
    tot_ms = tot_fl = 0.0
    for i in range(n):
        ms = sum(rec[r * n + i][0].elapsed_time(rec[r * n + i][1]) for r in range(reps)) / reps
        _, _, b, h, w, c0, c1, cout, taps, res, st, skc, gn = rec[i]
        fl = 2.0 * b * h * w * cout * (taps * (c0 + c1) + skc)
        tot_ms += ms
        tot_fl += fl
        cin = (f"{c0}+{c1}" if c1 else f"{c0}") + (f"|s{skc}" if skc else "")
        print(f"{i:3d} {b:4d} {h:4d}x{w:<4d} {cin:>10s} {cout:5d} {taps:4d} {int(res):3d} {int(st):5d} {int(gn):2d} {ms:8.3f} {fl / ms / 1e9:8.1f} {fl / 1e9:8.1f}")
    print(f"total {tot_ms:.3f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s over {n} launches; clocks {clk}")


if __name__ == "__main__":
    main()
