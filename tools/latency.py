#!/usr/bin/env python3
"""p50 latency of one rk_calculate_batch on device-resident buffers for a few batch sizes (run under gpurun).
RK_B200_LIB selects a library variant (tools/build_variant.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rsruckig_b200 import BatchRuckig, ThrowErrorHandler, workloads, _cabi as cabi  # noqa: E402

dev = torch.device("cuda:0")
sizes = [int(x) for x in (sys.argv[1:] or ["16384", "65536", "262144"])]
out = []
for n in sizes:
    f = {k: torch.from_numpy(v).to(dev) for k, v in workloads.bench_inputs(n, 7, seed=1).items()}
    otg = BatchRuckig(7, 0.005, ThrowErrorHandler, device=0)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    du = torch.zeros(n, dtype=torch.float64, device=dev)
    run = lambda: otg.calculate(f, n=n, memory_space=cabi.RK_MEM_DEVICE, status=st, duration=du)
    for _ in range(5):
        run()
    otg.sync()
    stream = torch.cuda.ExternalStream(otg._engine.lib.rk_stream(otg.handle), device=dev)
    times = []
    for _ in range(60):
        otg.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if stream is not None:
            a.record(stream)
            run()
            b.record(stream)
        else:
            a.record()
            run()
            b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    out.append(f"n={n}: p50 {times[len(times)//2]:.3f} ms ({n/times[len(times)//2]/1e3:.1f} M/s) p99 {times[-1]:.3f}")
print(" | ".join(out))
