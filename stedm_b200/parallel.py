"""Multi-GPU plumbing for the sampling path: one process per GPU, generation batches sharded by sample.

The path needs no per-step collective: every op is per-sample and the weights are replicated.
torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in CPU tests) is used only for
  * the final gather of uint8 images, and
  * an optional broadcast of precomputed style features (a "style bank") from rank 0.
Per-sample noise is keyed by GLOBAL sample index, so results do not depend on the number of ranks.
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_range(n_samples, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`; shards differ in size by at most one sample."""
    base, rem = divmod(n_samples, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def noise_for_samples(first_index, count, shape, seed=1234, device="cpu"):
    """x_T for global samples [first_index, first_index+count): one generator per sample index."""
    out = torch.stack([torch.randn(shape, generator=torch.Generator().manual_seed(seed + first_index + i))
                       for i in range(count)]) if count else torch.empty((0, *shape))
    return out.to(device)


def gather_images(images_u8, n_total=None):
    """All-gather the per-rank uint8 image shards (B_r, P, P, 3) into global order on every rank.
    Shards may differ by one sample; they are padded to the largest shard for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return images_u8
    world = dist.get_world_size()
    counts = torch.tensor([images_u8.shape[0]], device=images_u8.device, dtype=torch.int64)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    sizes = [int(c.item()) for c in all_counts]
    mx = max(sizes)
    padded = images_u8
    if images_u8.shape[0] < mx:
        pad = torch.zeros((mx - images_u8.shape[0], *images_u8.shape[1:]), dtype=images_u8.dtype, device=images_u8.device)
        padded = torch.cat([images_u8, pad], 0)
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded.contiguous())
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def broadcast_style_bank(features, src=0):
    """Broadcast precomputed style features (N_bank, 512) fp32 from rank `src` (KBs over NVLink)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(features, src=src)
    return features
