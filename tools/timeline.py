#!/usr/bin/env python3
"""Kernel timeline of one rk_calculate_batch on device-resident buffers (run under gpurun): torch.profiler (CUPTI) records every
kernel of the process with its stream, start and duration; the table shows, per kernel name, its summed duration as it ran
(alone or beside the other chunk's kernels), and the time during which NO kernel was running.
usage: tools/timeline.py [n] [out.csv]; RK_B200_LIB selects a library variant."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from rsruckig_b200 import BatchRuckig, ThrowErrorHandler, workloads, _cabi as cabi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2097152
out_csv = sys.argv[2] if len(sys.argv) > 2 else None
dev = torch.device("cuda:0")
f = {k: torch.from_numpy(v).to(dev) for k, v in workloads.bench_inputs(n, 7, seed=1).items()}
otg = BatchRuckig(7, 0.005, ThrowErrorHandler, result_layout=cabi.RK_RESULT_COMPACT)
for _ in range(3):
    otg.calculate(f, n=n, memory_space=cabi.RK_MEM_DEVICE)
otg.sync()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    otg.calculate(f, n=n, memory_space=cabi.RK_MEM_DEVICE)
    otg.sync()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and ("rk::" in e.name or "rk_lazy::" in e.name)]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
t1 = max(e.time_range.end for e in ev)
rows = [(e.name.split("(")[0].replace("void ", ""), e.time_range.start - t0, e.time_range.end - e.time_range.start) for e in ev]
if out_csv:
    with open(out_csv, "w") as fh:
        fh.write("kernel,start_us,duration_us\n")
        for r in rows:
            fh.write(f"{r[0]},{r[1]:.1f},{r[2]:.1f}\n")
# union of the busy intervals
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
# time with exactly one / two or more kernels running
pts = sorted([(s, 1) for s, _ in iv] + [(e, -1) for _, e in iv])
depth, last, single, multi = 0, pts[0][0], 0.0, 0.0
for t, d in pts:
    if depth == 1:
        single += t - last
    elif depth >= 2:
        multi += t - last
    depth += d
    last = t
total = t1 - t0
by = {}
for name, _, dur in rows:
    a = by.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += dur
print(f"{n} x 7-DoF: {len(rows)} kernels, wall {total / 1e3:.3f} ms = {n / total:.2f} M traj/s; some kernel running {100 * busy / total:.1f} % "
      f"(exactly one {100 * single / total:.1f} %, two or more {100 * multi / total:.1f} %), idle {100 * (total - busy) / total:.1f} %")
ssum = sum(a[1] for a in by.values())
print(f"sum of kernel durations {ssum / 1e3:.3f} ms = {ssum / total:.2f} x wall")
for name, (cnt, dur) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"  {dur / 1e3:9.3f} ms  {100 * dur / ssum:5.1f} %  x{cnt:<4d} {name}")
