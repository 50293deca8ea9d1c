"""Pinned host -> device copy bandwidth of this box (context for the e2e number), one process per GPU, all GPUs copying at the
same time: `python -m torch.distributed.run --nproc-per-node N tools/pcie_bw.py [--bind]` (or plain `python tools/pcie_bw.py`).
--bind pins each rank to the CPU cores / NUMA node of its GPU (bench.py's bind_to_gpu_cores) before the buffer is allocated."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
binding = None
if "--bind" in sys.argv:
    from bench import bind_to_gpu_cores
    binding = bind_to_gpu_cores(local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = 8 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
lo = t.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"H2D pinned, {world} GPU(s) concurrently{' (ranks bound: ' + str(binding) + ')' if binding else ''}: aggregate {t.item():.1f} GB/s, slowest rank {lo.item():.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
