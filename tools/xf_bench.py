"""A/B of one 3x3 convolution with its GroupNorm + SiLU: (gn_apply kernel + plain conv) vs the conv with the
normalisation in its operand path, interleaved on one box.   python tools/xf_bench.py [--reps 30]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

SHAPES = [  # B, H, W, C, cout
    (128, 16, 16, 1024, 1024),
    (64, 16, 16, 1024, 1024),
    (128, 32, 32, 512, 512),
    (64, 32, 32, 512, 512),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    from stedm_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    for B, H, W, C, co in SHAPES:
        x = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(co, 9 * C, generator=g) / (9 * C) ** 0.5).to(torch.bfloat16).to(dev)
        gamma, beta = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
        t = x.float().reshape(B * H * W // 128, 128, C)
        tiles = torch.stack([t.sum(1), (t * t).sum(1)], -1).contiguous()
        src = (tiles, C, 1, tiles.shape[0], H * W // 128, B)
        coef = ops.gn_fold_tiles(src, None, B, coef_for=(gamma, beta, 1e-5, H * W))
        folded = ops.gn_fold_tiles(src, None, B)
        stats = torch.empty((B * H * W // 128, co, 2), device=dev)

        def unfused():
            n = ops.gn_apply(x, None, folded, gamma, beta, 1e-5, True, torch.bfloat16, n_chunks=1)
            return ops.conv(n, w, None, co, 3, out_dtype=torch.bfloat16, tensor_core=True, stats_out=stats)

        def plain():
            return ops.conv(x, w, None, co, 3, out_dtype=torch.bfloat16, tensor_core=True, stats_out=stats)

        def fused():
            return ops.conv(x, w, None, co, 3, out_dtype=torch.bfloat16, tensor_core=True, stats_out=stats, gn_coef=coef)

        res = {}
        for name, fn in (("apply+conv", unfused), ("conv only", plain), ("fused", fused)) * 2:
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res.setdefault(name, []).append(e0.elapsed_time(e1) / a.reps)
        fl = 2.0 * B * H * W * co * 9 * C
        print(f"B={B:3d} {H}x{W} {C}->{co}: " + "  ".join(
            f"{k} {min(v):.3f} ms ({fl / min(v) / 1e9:.0f} TF)" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
