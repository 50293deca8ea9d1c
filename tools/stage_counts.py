#!/usr/bin/env python3
"""How many (DoF, trajectory) items enter each kernel of the pipeline (run under gpurun): one chunk of the bench distribution,
then the work-list counters the kernels left behind (rk_debug_counters)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rsruckig_b200 import BatchRuckig, ThrowErrorHandler, workloads, _cabi as cabi  # noqa: E402

n, dof = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 7   # below 192 Ki items: a single chunk on scratch set 0
f = {k: torch.from_numpy(v).cuda() for k, v in workloads.bench_inputs(n, dof, seed=1).items()}
otg = BatchRuckig(dof, 0.005, ThrowErrorHandler)
otg.calculate(f, n=n, memory_space=cabi.RK_MEM_DEVICE)
c = (C.c_uint32 * 112)()
lib = otg._engine.lib
lib.rk_debug_counters.argtypes = [C.c_void_p, C.c_void_p]
cabi.check(lib.rk_debug_counters(otg.handle, c))
c = np.array(c[:])
items = n * dof
names = ["stage0 ACC0_ACC1_VEL", "VEL/UDDU", "VEL/UDUD", "ACC0_VEL", "ACC1_VEL", "stage0 (2nd dir)", "VEL/UDDU (2nd)", "VEL/UDUD (2nd)",
         "ACC0_VEL (2nd)", "ACC1_VEL (2nd)", "tail (8 variants)"]
print(f"{items} items ({n} x {dof})")
for s, name in enumerate(names):
    print(f"  enter {name:24s} {c[s]:9d}  {100 * c[s] / items:6.2f} % of items")
for t in range(3):
    print(f"  quartic roots queued, family {t}: {c[20 + t]:9d}  {c[20 + t] / items:5.2f} per item")
print(f"  closed-form candidates queued: {c[23]:9d}  {c[23] / items:5.2f} per item; rescue items: {c[24]}")
for step in (1, 6):
    print(f"  VEL/UDDU step {step}: pending items {c[32 + 4 * step]}, polish tasks {c[33 + 4 * step]}")
