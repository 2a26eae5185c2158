"""GroupNorm kernel bandwidth at the U-Net's shapes (CUDA events, L2-cold by rotating over distinct buffers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stedm_b200 import ops

SHAPES = [(64, 64, 128, 0), (128, 16, 1024, 0), (128, 32, 512, 0), (128, 64, 512, 128), (128, 16, 1024, 1024),
          (128, 64, 128, 0)]


def main():
    for B, hw, c0, c1 in SHAPES:
        n_buf = max(2, int(600e6 // (B * hw * hw * (c0 + c1) * 2)) + 1)
        xs = [torch.randn(B, hw, hw, c0, device="cuda").to(torch.bfloat16) for _ in range(n_buf)]
        x1 = torch.randn(B, hw, hw, c1, device="cuda").to(torch.bfloat16) if c1 else None
        gamma, beta = torch.ones(c0 + c1, device="cuda"), torch.zeros(c0 + c1, device="cuda")
        stats = ops.gn_stats(xs[0], x1)
        out = ops.gn_apply(xs[0], x1, stats, gamma, beta, 1e-5, True, torch.bfloat16)
        torch.cuda.synchronize()
        res = []
        for which in ("stats", "apply"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3 * n_buf
            e0.record()
            for i in range(reps):
                if which == "stats":
                    ops.gn_stats(xs[i % n_buf], x1, stats)
                else:
                    ops.gn_apply(xs[i % n_buf], x1, stats, gamma, beta, 1e-5, True, torch.bfloat16)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            nbytes = B * hw * hw * (c0 + c1) * 2 * (1 if which == "stats" else 2)
            res.append(f"{which} {us:7.1f} us {nbytes / us / 1e6:6.2f} TB/s")
        print(f"B={B:3d} {hw:2d}x{hw:<2d} C={c0}+{c1}: " + "   ".join(res), flush=True)


if __name__ == "__main__":
    main()
