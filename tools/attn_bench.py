"""The decoder's single-head d = 512 attention at the bench shapes: wide flash kernel, TFLOP/s of useful work.
    python tools/attn_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    from stedm_b200 import ops
    dev = torch.device("cuda", 0)
    for B, T in ((64, 4096), (8, 16384), (16, 4096)):
        C = 512
        qkv = (torch.randn(B, T, 3 * C, device=dev) * 0.5).to(torch.bfloat16)
        fn = lambda: ops.attention_tc(qkv, qkv, qkv, 1, C, T, (T * 3 * C, C, 3 * C), C ** -0.5, q_off=0, k_off=C, v_off=2 * C)
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 5
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 4.0 * B * T * T * C
        print(f"B={B} T={T}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s useful ({1.5 * fl / ms / 1e9:.0f} executed)", flush=True)


if __name__ == "__main__":
    main()
