"""Style encoder timing on one B200: native StyleEncoderRunner (bf16 / fp32) vs torchvision swin_v2_t eager
(TF32 defaults, and bf16 autocast) on the same images.  CUDA events, 3 warm-up + 10 timed.
    python tools/style_bench.py [--images 64 640] [--size 256]"""
import argparse
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, nargs="+", default=[64, 640])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--per-kernel", action="store_true")
    a = ap.parse_args()
    from stedm_b200 import ops
    from stedm_b200.style_engine import StyleEncoderRunner
    torch.manual_seed(0)
    m = torchvision.models.get_model("swin_v2_t")
    m.head = torch.nn.Linear(768, 512)
    m = m.cuda().eval()
    gf = 11.9 * (a.size / 256) ** 2
    for n in a.images:
        imgs = (torch.rand(n, a.size, a.size, 3, device="cuda") * 2 - 1)
        nchw = imgs.permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            for prec in ("bf16", "fp32"):
                if prec == "fp32" and n > 64:
                    continue
                r = StyleEncoderRunner(m, prec)
                l0 = ops.LAUNCHES[0]
                r(imgs)
                launches = ops.LAUNCHES[0] - l0
                ms = timed(lambda: r(imgs))
                print(f"images={n} size={a.size} native {prec}: {ms:8.2f} ms  {n * gf / ms:8.1f} TFLOP/s-equivalent  "
                      f"({launches} launches)", flush=True)
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
            ms = timed(lambda: m(nchw))
            print(f"images={n} size={a.size} torchvision eager TF32: {ms:8.2f} ms", flush=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ms = timed(lambda: m(nchw))
            print(f"images={n} size={a.size} torchvision eager bf16 autocast: {ms:8.2f} ms", flush=True)
        del imgs, nchw
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
