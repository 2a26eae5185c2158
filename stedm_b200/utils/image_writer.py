"""Output plumbing of predict_step (reference modules/ldm_diffusion.py:94-107): device->host copy of the uint8 images and
PNG encoding, taken OFF the sampling critical path.

The reference converts and writes every image synchronously inside predict_step (one PIL encode after another on the
main thread, ~10 ms per 256x256 PNG, two files per sample) — at B200 sampling rates that is as long as the sampling
itself.  Here a batch's uint8 images / label maps are copied into pinned host buffers on a side stream (ordered after
the producing stream by an event), and a small thread pool waits for the copy and encodes the PNGs (zlib releases the
GIL) while the GPU already samples the next batch.  File names and contents are the reference's:
``img_XXXXX.png`` (RGB uint8) and ``seg_XXXXX.png`` (uint8 labels) keyed by the zero-padded dataset index.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


class AsyncImageWriter:
    def __init__(self, out_dir, workers=None, max_pending=4):
        self.out_dir = out_dir
        os.makedirs(out_dir, exist_ok=True)
        self.workers = workers or min(16, os.cpu_count() or 4)
        self.pool = ThreadPoolExecutor(max_workers=self.workers)
        self.max_pending = max_pending            # batches in flight (bounds pinned memory)
        self._pending = []
        self._copy_stream = None
        self.written = 0

    def _to_host(self, t):
        """Asynchronous D2H into pinned memory on the copy stream; returns (host tensor, event | None)."""
        if not t.is_cuda:
            return t.contiguous(), None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=t.device)
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            host.copy_(t, non_blocking=True)
            t.record_stream(self._copy_stream)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return host, done

    def submit(self, images_u8, labels_u8, indices):
        """images_u8 (B,P,P,3) uint8, labels_u8 (B,P,P) uint8 (device or host), indices: B dataset indices."""
        imgs, e1 = self._to_host(images_u8)
        segs, e2 = self._to_host(labels_u8)
        nums = [int(i) for i in (indices.tolist() if torch.is_tensor(indices) else indices)]
        futs = [self.pool.submit(self._encode, imgs, segs, e1, e2, k, num) for k, num in enumerate(nums)]
        self._pending.append(futs)
        while len(self._pending) > self.max_pending:
            self._drain_one()
        return imgs

    def _encode(self, imgs, segs, e1, e2, k, num):
        from PIL import Image
        for e in (e1, e2):
            if e is not None:
                e.synchronize()
        name = str(num).zfill(5)                                               # ldm_diffusion.py:103
        Image.fromarray(np.asarray(imgs[k])).save(os.path.join(self.out_dir, f"img_{name}.png"))
        Image.fromarray(np.asarray(segs[k])).save(os.path.join(self.out_dir, f"seg_{name}.png"))
        return 1

    def _drain_one(self):
        for f in self._pending.pop(0):
            self.written += f.result()

    def flush(self):
        while self._pending:
            self._drain_one()
        return self.written

    def close(self):
        n = self.flush()
        self.pool.shutdown(wait=True)
        return n

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
