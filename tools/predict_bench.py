"""predict_step throughput INCLUDING the output tail (device->host copy + PNG files, reference
modules/ldm_diffusion.py:96-107): synchronous PIL loop as in the reference vs AsyncImageWriter (pinned D2H on a side
stream + threaded encode overlapping the next batch).  Wall clock, batches of 64 at 256^2, DDIM-50 cfg 1.5.
    python tools/predict_bench.py [--batches 3] [--batch 64]"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    from stedm_b200.utils.image_writer import AsyncImageWriter
    dev = torch.device("cuda", 0)
    m = bench.build_model(64, 1, "bf16").to(dev).eval()
    B = a.batch
    img, seg_oh, style, _ = bench.synthetic_batch(B, 256, 1, 0)
    tup = lambda k: (img.to(dev), seg_oh.to(dev), None, style.to(dev), torch.arange(k * B, (k + 1) * B))
    with torch.no_grad():
        m.predict_dir = tempfile.mkdtemp()
        m.predict_step(tup(0), 0)                       # warm-up (graph capture, packed weights)
        for mode in ("sync", "async"):
            out = tempfile.mkdtemp()
            m.predict_dir = out
            m.writer = AsyncImageWriter(out) if mode == "async" else None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(a.batches):
                m.predict_step(tup(k), k)
            if m.writer is not None:
                m.writer.close()
                m.writer = None
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            n = len([f for f in os.listdir(out) if f.startswith("img_")])
            print(f"{mode:5s} output tail: {a.batches * B / dt:7.2f} images/s  ({dt / a.batches * 1e3:7.1f} ms per batch of {B}, "
                  f"{n} img + {n} seg PNGs written, {os.cpu_count()} host cores)", flush=True)
            shutil.rmtree(out, ignore_errors=True)


if __name__ == "__main__":
    main()
