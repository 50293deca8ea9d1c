#!/usr/bin/env python3
"""configs[3] in isolation (run under gpurun, optionally under ncu -k regex:k_at_time): 64 Ki x 7-DoF calculate, then a few
1000-cycle rollouts of rk_at_time_batch (position, velocity, acceleration) on device buffers.
usage: tools/at_time_run.py [compact|full] [K] [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rsruckig_b200 import BatchRuckig, ThrowErrorHandler, workloads, _cabi as cabi  # noqa: E402

layout = cabi.RK_RESULT_COMPACT if (sys.argv[1] if len(sys.argv) > 1 else "compact") == "compact" else cabi.RK_RESULT_FULL
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n, dof = 65536, 7
dev = torch.device("cuda:0")
f = {k: torch.from_numpy(v).to(dev) for k, v in workloads.bench_inputs(n, dof, seed=1).items()}
otg = BatchRuckig(dof, 0.005, ThrowErrorHandler, result_layout=layout)
otg.calculate(f, n=n, memory_space=cabi.RK_MEM_DEVICE)
pos = torch.empty((K, dof, n), dtype=torch.float64, device=dev)
vel, acc = torch.empty_like(pos), torch.empty_like(pos)
stream = torch.cuda.ExternalStream(otg._engine.lib.rk_stream(otg.handle), device=dev)
run = lambda: otg.at_time(time_start=0.0, delta_time=0.005, steps=K, memory_space=cabi.RK_MEM_DEVICE, position=pos, velocity=vel, acceleration=acc)
for _ in range(2):
    run()
otg.sync()
with torch.cuda.stream(stream):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        run()
    b.record(stream)
    b.synchronize()
ms = a.elapsed_time(b) / reps
gb = 3 * 8 * dof * n * K / 1e9
print(f"at_time {sys.argv[1] if len(sys.argv) > 1 else 'compact'} K={K}: {ms:.3f} ms per rollout, {gb / ms:.3f} TB/s written")
