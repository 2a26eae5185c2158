"""AsyncImageWriter (predict_step's output tail, reference modules/ldm_diffusion.py:96-107): same file names and pixel
content as the reference's synchronous PIL loop; host-only here, the CUDA copy path is exercised by the predict test."""
import os

import numpy as np
import torch
from PIL import Image


def test_async_writer_matches_sync_pil(tmp_path):
    from stedm_b200.utils.image_writer import AsyncImageWriter
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (5, 32, 32, 3), generator=g, dtype=torch.uint8)
    segs = torch.randint(0, 6, (5, 32, 32), generator=g, dtype=torch.uint8)
    idx = torch.tensor([3, 17, 101, 4, 99999])
    with AsyncImageWriter(str(tmp_path), workers=3, max_pending=1) as w:
        w.submit(imgs[:3], segs[:3], idx[:3])
        w.submit(imgs[3:], segs[3:], idx[3:])        # exceeds max_pending: the first batch is drained first
    assert w.written == 5
    for k, num in enumerate(idx.tolist()):
        name = str(num).zfill(5)
        assert np.array_equal(np.asarray(Image.open(os.path.join(tmp_path, f"img_{name}.png"))), imgs[k].numpy())
        assert np.array_equal(np.asarray(Image.open(os.path.join(tmp_path, f"seg_{name}.png"))), segs[k].numpy())
