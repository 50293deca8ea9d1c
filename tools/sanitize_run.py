"""A small tour through every kernel for compute-sanitizer (run under gpurun): calculate on edge cases of several DoF
counts and settings, at_time (both instantiations, sliced), query helpers, closed-loop cycles."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rsruckig_b200 import (BatchRollout, BatchRuckig, ControlInterface, DurationDiscretization, IgnoreErrorHandler, Synchronization,  # noqa: E402
                           _cabi as cabi, workloads)

for dof, n in ((1, 300), (4, 1500), (7, 3000)):
    f = workloads.edge_case_inputs(n, dof, seed=dof)
    for handler in (None, IgnoreErrorHandler):
        otg = BatchRuckig(dof, 0.005, handler) if handler else BatchRuckig(dof, 0.005)
        otg.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST)
        for sync in (Synchronization.Phase, Synchronization.None_, Synchronization.TimeIfNecessary):
            otg.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST, synchronization=sync, duration_discretization=DurationDiscretization.Discrete)
    otg.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST, control_interface=ControlInterface.Velocity)
    K = 40
    p, v, a, j = (np.zeros((K, dof, n)) for _ in range(4))
    s = np.zeros((K, n), dtype=np.int32)
    otg.at_time(time_start=0.0, delta_time=0.05, steps=K, memory_space=cabi.RK_MEM_HOST, position=p, velocity=v, acceleration=a, jerk=j, section=s)
    otg.at_time(time_start=0.3, delta_time=0.05, steps=K, memory_space=cabi.RK_MEM_HOST, position=p, velocity=v, acceleration=a)
    otg.position_extrema()
    otg.first_time_at_position(f["target_position"])
f2 = workloads.bench_inputs(2000, 3, seed=3)
f2["max_jerk"][:] = np.inf
f2["current_acceleration"][:] = 0.0
f2["target_acceleration"][:] = 0.0
BatchRuckig(3, 0.005).calculate(f2, n=2000, memory_space=cabi.RK_MEM_HOST)
f2["max_acceleration"][:] = np.inf
f2["current_velocity"][:] = 0.0
f2["target_velocity"][:] = 0.0
BatchRuckig(3, 0.005).calculate(f2, n=2000, memory_space=cabi.RK_MEM_HOST)
dev = torch.device("cuda", 0)
n, dof = 2048, 3
d = {k: torch.from_numpy(x).to(dev) for k, x in workloads.bench_inputs(n, dof, seed=9).items()}
ro = BatchRollout(n, dof, 0.01)
newp = torch.zeros((dof, n), dtype=torch.float64, device=dev)
res = torch.zeros(n, dtype=torch.int32, device=dev)
for c in range(12):
    if c == 5:
        d["target_position"][:, ::3] += 1.0
    ro.update(d, new_position=newp, result=res)
ro.sync()
torch.cuda.synchronize()
print("sanitize tour ok")
