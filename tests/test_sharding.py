"""Multi-rank host logic on CPU (gloo, world_size 2): shard ranges, per-sample noise keyed by global index, image
gather in global order, style-bank broadcast.  A 2-rank run must reproduce the 1-rank result exactly."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stedm_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_generate(x_T):
    """Stand-in for the per-sample generation: any per-sample deterministic map to uint8 'images'."""
    return ((torch.tanh(x_T).permute(0, 2, 3, 1) + 1) * 127.5).to(torch.uint8).contiguous()


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(n_total, r, w)
    x_T = parallel.noise_for_samples(lo, hi - lo, (3, 8, 8))
    imgs = parallel.gather_images(_fake_generate(x_T))
    bank = torch.arange(12, dtype=torch.float32).reshape(3, 4) if r == 0 else torch.zeros(3, 4)
    bank = parallel.broadcast_style_bank(bank)
    torch.save({"imgs": imgs, "bank": bank}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_run_equals_single(tmp_path):
    n_total = 7                                                     # uneven shards: 4 + 3
    want = _fake_generate(parallel.noise_for_samples(0, n_total, (3, 8, 8)))
    mp.spawn(_worker, args=(2, _free_port(), n_total, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = torch.load(os.path.join(tmp_path, f"rank{r}.pt"))
        assert torch.equal(got["imgs"], want)
        assert torch.equal(got["bank"], torch.arange(12, dtype=torch.float32).reshape(3, 4))
