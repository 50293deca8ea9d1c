#!/usr/bin/env python3
"""Host time of one rk_calculate_batch call (kernel launches only, no sync) next to its device time (run under gpurun)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rsruckig_b200 import BatchRuckig, ThrowErrorHandler, workloads, _cabi as cabi  # noqa: E402

dev = torch.device("cuda:0")
for n in [int(x) for x in (sys.argv[1:] or ["1024", "16384", "65536"])]:
    f = {k: torch.from_numpy(v).to(dev) for k, v in workloads.bench_inputs(n, 7, seed=1).items()}
    otg = BatchRuckig(7, 0.005, ThrowErrorHandler, device=0)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    du = torch.zeros(n, dtype=torch.float64, device=dev)
    run = lambda: otg.calculate(f, n=n, memory_space=cabi.RK_MEM_DEVICE, status=st, duration=du)
    for _ in range(5):
        run()
    otg.sync()
    host, total = [], []
    for _ in range(50):
        otg.sync()
        t0 = time.perf_counter()
        run()
        t1 = time.perf_counter()
        otg.sync()
        t2 = time.perf_counter()
        host.append(t1 - t0)
        total.append(t2 - t0)
    host.sort(); total.sort()
    print(f"n={n}: host launch p50 {host[25]*1e3:.3f} ms, launch+sync p50 {total[25]*1e3:.3f} ms, launches per call {otg.kernel_launches() // 56}")
