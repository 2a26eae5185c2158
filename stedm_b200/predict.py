"""Prediction entry point — the B200 counterpart of the reference's ``predict_diff.py`` (predict_diff.py:34-89).

    python -m stedm_b200.predict [hydra-style overrides] [+ckpt_path=... +predict_dir=out] [+synthetic=N]
    torchrun --nproc-per-node 8 -m stedm_b200.predict location=b200x8 ...

Same config tree and override syntax (``style_agg=mean ddim_steps=50 diffusion.image_size=64 data.patch_size=256``);
the Lightning Trainer / DDPStrategy of the reference is replaced by one process per GPU: each rank takes the
dataset indices ``rank, rank + world, ...`` exactly like Lightning's distributed predict sampler and writes its own
``img_XXXXX.png`` / ``seg_XXXXX.png`` keyed by dataset index (modules/ldm_diffusion.py:99-107).  There is no
per-step collective.  The reference's datasets need private WSI data; ``+synthetic=N`` generates N samples of the
configured shape instead (low-frequency layout masks, U(-1,1) style images), and user code can pass any iterable of
the reference's batch tuples ``(img, seg_onehot, seg, style, index)`` to :func:`run`.
"""
import os
import sys

import torch

from . import parallel
from .config import load_config
from .modules.ldm_diffusion import LDM_Diffusion


class SyntheticPredictSet(torch.utils.data.Dataset):
    """Stand-in for data/ds.py Predict_DS: (img, seg one-hot, seg, style images, dataset index)."""

    def __init__(self, n, patch, n_classes, n_style):
        self.n, self.p, self.k, self.ns = n, patch, n_classes, n_style

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(99 + i)
        low = torch.rand(1, self.k, 8, 8, generator=g)
        lab = torch.nn.functional.interpolate(low, size=(self.p, self.p), mode="bilinear")[0].argmax(0)
        seg_oh = torch.nn.functional.one_hot(lab, self.k).permute(2, 0, 1).float()
        style = torch.rand(self.ns, 3, self.p, self.p, generator=g) * 2 - 1
        return torch.zeros(3, self.p, self.p), seg_oh, lab, style, i


def run(cfg, batches, predict_dir, ckpt_path=None, precision=None, device=None, async_io=True):
    """Generate and save images for every batch tuple of ``batches`` on this rank's GPU."""
    rank, world, local = parallel.env_rank_world()
    device = device or torch.device("cuda", local)
    torch.cuda.set_device(device)
    torch.set_float32_matmul_precision("high")                          # predict_diff.py:68
    module = LDM_Diffusion(cfg, precision=precision or getattr(cfg, "precision", "bf16"),
                           load_first_stage_ckpt=ckpt_path is not None and getattr(cfg, "load_first_stage", False))
    if ckpt_path is not None:
        sd = torch.load(ckpt_path, map_location="cpu")
        sd = sd.get("state_dict", sd)
        missing, unexpected = module.load_state_dict(sd, strict=False)   # predict_diff.py:48
        print(f"[rank {rank}] restored {ckpt_path}: {len(missing)} missing, {len(unexpected)} unexpected keys")
    else:
        from .utils.fixture import apply_fixture_weights
        apply_fixture_weights(module._model, seed=0)
        print(f"[rank {rank}] no checkpoint given: deterministic random-init (fixture) weights")
    module = module.to(device).eval()
    os.makedirs(predict_dir, exist_ok=True)
    module.predict_dir = predict_dir
    if async_io:
        from .utils.image_writer import AsyncImageWriter
        module.writer = AsyncImageWriter(predict_dir)
    n = 0
    for idx, batch in enumerate(batches):
        batch = tuple(t.to(device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)
        n += len(module.predict_step(batch, idx))
    if module.writer is not None:
        module.writer.close()
        module.writer = None
    return n


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    cfg = load_config(argv)
    rank, world, local = parallel.init_distributed()
    batch_size = cfg.data.batch_base * cfg.location.batch_mul              # predict_diff.py:37
    n = int(cfg.get("synthetic", 2 * batch_size * max(world, 1)))
    ds = SyntheticPredictSet(n, cfg.data.patch_size, cfg.data.num_classes,
                             cfg.style_sampling.get("num_patches", 1) if cfg.style_sampling.name == "mp" else 1)
    mine = torch.utils.data.Subset(ds, list(range(rank, n, world)))          # Lightning's distributed predict sampler
    # input side of the plumbing: worker processes build the next batches into pinned memory while the GPU samples
    # (the reference's datamodule does the same through Lightning, data/dm.py:82-87)
    workers = int(cfg.get("num_workers", min(4, os.cpu_count() or 1)))
    loader = torch.utils.data.DataLoader(mine, batch_size=batch_size, shuffle=False, num_workers=workers,
                                         pin_memory=True, persistent_workers=workers > 0)
    out = cfg.get("predict_dir", os.path.join(os.getcwd(), "stedm_predict"))
    done = run(cfg, loader, out, ckpt_path=cfg.get("ckpt_path_full"))
    print(f"[rank {rank}/{world}] wrote {done} images to {out}")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
