"""N > 1 host logic on CPU: shard ranges tile the batch, the gloo world_size-2 path reduces timings with MAX and gathers
results in index order. (The compute itself needs a GPU; here each rank 'computes' a deterministic function of its
indices so that the assembled result can be checked.)"""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from rsruckig_b200.sharding import gather_host, max_over_ranks, shard_range


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (7, 2), (64, 8), (1_000_003, 8), (8 << 20, 4)])
def test_shards_tile_the_batch(n, world):
    ranges = [shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    idx = np.arange(lo, hi, dtype=np.float64)
    local = np.stack([idx * 2.0, idx + 0.5])          # a [2][n_local] "result"
    dist.barrier()
    t = max_over_ranks(10.0 + rank)                   # per-rank "device time"
    full = gather_host(local, n, rank, world)
    if rank == 0:
        out_q.put((t, full))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world_size_2_gather_and_max():
    n, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert t == 11.0
    ref = np.arange(n, dtype=np.float64)
    assert full.shape == (2, n)
    assert np.array_equal(full[0], ref * 2.0) and np.array_equal(full[1], ref + 0.5)
