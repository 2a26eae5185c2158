"""Prediction entry point — the B200 counterpart of the reference's ``predict_diff.py`` (predict_diff.py:34-89).

    python -m stedm_b200.predict [hydra-style overrides] [+ckpt_path=/abs/file.ckpt | +ckpt_name=x.ckpt] [+predict_dir=out]
                                 [+synthetic=N] [+fixture_weights=true]
    torchrun --nproc-per-node 8 -m stedm_b200.predict location=b200x8 ...

Checkpoint: ``+ckpt_path=`` (a file) or, like predict_diff.py:39-44, ``+ckpt_name=`` / the default
``Diff_<data>_<samples>_<sampling>_last.ckpt`` under ``location.result_dir/checkpoints``.  A missing checkpoint is an
error, as in the reference (``load_from_checkpoint``) — unless ``+fixture_weights=true`` (or ``+synthetic=N``, the
self-test mode) explicitly asks for the deterministic random-init weights the tests and the benchmark use.

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


def run(cfg, batches, predict_dir, ckpt_path=None, precision=None, device=None, async_io=True,
        allow_fixture_weights=False):
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
        for name, keys in (("missing", missing), ("unexpected", unexpected)):
            if keys:
                print(f"[rank {rank}]   {name}: {', '.join(keys[:12])}{' ...' if len(keys) > 12 else ''}")
        # the sampled networks must come from the checkpoint: a file without them would sample random weights
        # (encoder.* / loss.* / model_ema.* are outside the sampling path and may be absent)
        vital = [k for k in missing if ".diffusion_model." in k or ".first_stage_model.decoder." in k
                 or ".first_stage_model.quantize." in k or ".first_stage_model.post_quant_conv." in k
                 or ".cond_stage_model." in k or ".agg_block." in k]
        if vital:
            raise RuntimeError(f"checkpoint {ckpt_path} lacks {len(vital)} weights of the sampling path "
                               f"(first: {vital[0]}); refusing to sample partly random weights")
    elif not allow_fixture_weights:
        raise FileNotFoundError("no checkpoint: pass +ckpt_path=/abs/file.ckpt or +ckpt_name=... (looked up under "
                                "location.result_dir/checkpoints), or +fixture_weights=true for random-init weights")
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


def resolve_checkpoint(cfg):
    """``+ckpt_path=`` (a file; ``ckpt_path_full`` is the older spelling) or predict_diff.py:39-44's rule:
    ``location.result_dir/checkpoints/<ckpt_name | Diff_<data>_<samples>_<sampling>_last.ckpt>``.  Returns None when no
    checkpoint was named and the default file does not exist (run() then raises unless fixture weights are allowed)."""
    explicit = cfg.get("ckpt_path") or cfg.get("ckpt_path_full")
    if explicit:
        if not os.path.isfile(str(explicit)):
            raise FileNotFoundError(f"checkpoint {explicit} does not exist")
        return str(explicit)
    named = cfg.get("ckpt_name")
    name = named or (f"Diff_{cfg.data.name}_{cfg.data.get('class_train_samples', '')}_{cfg.style_sampling.name}_last.ckpt")
    path = os.path.join(str(cfg.location.result_dir), "checkpoints", str(name))
    if os.path.isfile(path):
        return path
    if named:
        raise FileNotFoundError(f"checkpoint {path} does not exist")
    return None


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
    done = run(cfg, loader, out, ckpt_path=resolve_checkpoint(cfg),
               allow_fixture_weights=bool(cfg.get("fixture_weights", False)) or cfg.get("synthetic") is not None)
    print(f"[rank {rank}/{world}] wrote {done} images to {out}")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
