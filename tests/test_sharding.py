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


def test_shard_range_matches_the_c_abi():
    """rk_shard_range (the split rk_calculate_batch_multi uses) is the same function as sharding.shard_range."""
    import ctypes as C
    from rsruckig_b200 import _cabi as cabi
    lib = cabi.load()
    for n, parts in ((0, 1), (1, 2), (7, 2), (64, 8), (1_000_003, 8), (8 << 20, 3)):
        for g in range(parts):
            lo, cnt = C.c_int64(-1), C.c_int64(-1)
            lib.rk_shard_range(n, parts, g, C.byref(lo), C.byref(cnt))
            assert (lo.value, lo.value + cnt.value) == shard_range(n, g, parts)


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [1, 2, 3])
def test_multi_handle_batch_equals_single_handle_bitwise(parts):
    """rk_calculate_batch_multi over every visible device (and, to exercise several parts on a one-GPU box, several handles per
    device): everything it returns and everything the parts hold equals one rk_calculate_batch bit for bit."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsruckig_b200 import BatchRuckig, MultiBatchRuckig, ThrowErrorHandler, workloads, _cabi as cabi
    n, dof = 300_001, 7
    f = workloads.bench_inputs(n, dof, seed=51)
    f["minimum_duration"] = np.where(np.arange(n) % 5 == 0, 6.0, np.nan)
    f["enabled"] = np.ascontiguousarray((np.random.default_rng(52).random((dof, n)) > 0.03).astype(np.uint8))
    one = BatchRuckig(dof, 0.005, ThrowErrorHandler)
    st1, du1, im1 = np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros((dof, n))
    one.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST, status=st1, duration=du1, independent_min_durations=im1)
    devices = [g % torch.cuda.device_count() for g in range(max(parts, torch.cuda.device_count()))]
    multi = MultiBatchRuckig(dof, 0.005, ThrowErrorHandler, devices=devices)
    st, du, im = np.full(n, 77, dtype=np.int32), np.full(n, -1.0), np.full((dof, n), -1.0)
    multi.calculate(f, n=n, status=st, duration=du, independent_min_durations=im)
    assert np.array_equal(st, st1) and np.array_equal(du, du1) and np.array_equal(im, im1)
    ok = st1 == 0
    for fld in (cabi.RK_FIELD_T, cabi.RK_FIELD_P, cabi.RK_FIELD_CODES, cabi.RK_FIELD_LIMITING_DOF, cabi.RK_FIELD_BRAKE_T):
        a, b = multi.read(fld), one.read(fld)
        assert np.array_equal(a[..., ok], b[..., ok]), fld
    multi.close()
