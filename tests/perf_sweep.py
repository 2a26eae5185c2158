"""BASELINE configs[4]: single U-Net eps-step microbench sweep (batch 1-256, 256^2 / 512^2) of the native engine
against the reference's PyTorch ops on the same B200.

The reference package cannot be imported on the GPU box (no pytorch_lightning / taming there, and /root/reference
does not travel), so "reference PyTorch" = the oracle's functional restatement of UNetModel.forward
(oracle/stedm_oracle.py — pinned bit-equal to the reference on CPU) executed by torch 2.11 + cuDNN/cuBLAS on the GPU,
with the reference's own settings (TF32, predict_diff.py:68) and again under bf16 autocast.  Test infrastructure.

    python tests/perf_sweep.py [--latents 64 128] [--batches 1 4 16 64 256]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import stedm_oracle as O
from tests.util import build_model, oracle_state_dict

GFLOP_L64 = 217.29


def time_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latents", type=int, nargs="+", default=[64, 128])
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 4, 16, 64, 256])
    ap.add_argument("--no-torch", action="store_true", help="native columns only")
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    print(f"{'latent':>6} {'batch':>5} | {'native bf16 ms':>14} {'graph ms':>8} {'TFLOP/s':>8} | {'torch TF32 ms':>13} "
          f"{'torch bf16 ms':>13} | {'x TF32':>6} {'x bf16':>6}   (x = torch / native-graph)", flush=True)
    m = build_model(64, n_style=1, precision="bf16")
    unet = m._model.model.diffusion_model
    sd = {k: v.cuda() for k, v in oracle_state_dict(m._model).items() if k.startswith(O.UNET)}
    for L in a.latents:
        for B in a.batches:
            if B * (L / 64) ** 2 > 300:
                continue
            g = torch.Generator().manual_seed(B + L)
            x = torch.randn(B, 3, L, L, generator=g).cuda()
            cc = torch.randn(B, 3, L, L, generator=g).cuda()
            ctx = torch.randn(B, 512, generator=g).cuda()
            t = torch.full((B,), 481, dtype=torch.long, device="cuda")
            reps = 3 if B * (L / 64) ** 2 >= 64 else 10
            from stedm_b200 import ops as _ops
            with torch.no_grad():
                native = time_ms(lambda: unet.forward_split(x, cc, t, ctx), reps)
                graph = torch.cuda.CUDAGraph()                  # the sampler replays the pass from a cached graph
                with torch.cuda.graph(graph):
                    unet.forward_split(x, cc, t, ctx)
                replay = time_ms(graph.replay, reps)
                del graph
                _ops.enable_split_k(False)                      # single-pass K loops (bit-identical across batch sizes)
                graph = torch.cuda.CUDAGraph()
                unet.forward_split(x, cc, t, ctx)
                with torch.cuda.graph(graph):
                    unet.forward_split(x, cc, t, ctx)
                splitk = time_ms(graph.replay, reps)
                _ops.enable_split_k(True)
                del graph
                xc = torch.cat([x, cc], 1)
                ref32 = ref16 = float("nan")
                if not a.no_torch:
                    for _ in range(8):                              # let cuDNN's algorithm choice settle
                        O.unet_forward(sd, xc, t, ctx)
                    ref32 = time_ms(lambda: O.unet_forward(sd, xc, t, ctx), reps)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        for _ in range(8):
                            O.unet_forward(sd, xc, t, ctx)
                        ref16 = time_ms(lambda: O.unet_forward(sd, xc, t, ctx), reps)
            tf = B * GFLOP_L64 * (L / 64) ** 2 / replay
            print(f"{L:6d} {B:5d} | {native:14.3f} {replay:8.3f} {tf:8.1f} | {ref32:13.3f} {ref16:13.3f} | "
                  f"{ref32 / replay:6.2f} {ref16 / replay:6.2f}   single-pass-K graph {splitk:8.3f} ms", flush=True)


if __name__ == "__main__":
    main()
