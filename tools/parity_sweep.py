"""Wide parity sweep (run under gpurun): CUDA path vs the portable oracle, every stored double compared, over many seeds of
the bench distribution and the other workload generators. Complements tests/test_parity.py (fixed seeds, smaller)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from rsruckig_b200 import BatchRuckig, workloads  # noqa: E402
from test_parity import assert_exact, run_both  # noqa: E402

total = 0
t0 = time.time()
for seed in range(10, 10 + int(sys.argv[1]) if len(sys.argv) > 1 else 18):
    n = 300_000
    f = workloads.bench_inputs(n, 7, seed=seed)
    otg, ob = run_both(BatchRuckig, f, n, 7, threads=16)
    total += assert_exact(otg, ob, label=f"7dof seed {seed}")
    del otg, ob
    # states ON trajectories (closed-loop distribution: velocities / accelerations riding their limits, sometimes one ulp beyond):
    # re-plan from the state each trajectory of a smaller batch passes through at about 0.6 of its duration
    n = 100_000
    f = workloads.bench_inputs(n, 7, seed=300 + seed)
    otg, ob = run_both(BatchRuckig, f, n, 7, threads=16)
    from rsruckig_b200 import _cabi as cabi
    T, ok = ob.read(cabi.RK_FIELD_DURATION), ob.read(cabi.RK_FIELD_STATUS) == 0
    K = 32
    dt = float(np.median(T[ok])) / 12
    pos, vel, acc, _, _ = ob.at_time(time_start=0.0, delta_time=dt, steps=K, threads=16, want_jerk=False)
    k = np.searchsorted(np.cumsum(np.full(K, dt)), 0.6 * T, side="right") - 1
    idx = np.nonzero(ok & (k >= 0))[0]
    for name, arr in (("current_position", pos), ("current_velocity", vel), ("current_acceleration", acc)):
        f[name][:, idx] = arr[k[idx], :, idx].T
    del otg, ob, pos, vel, acc
    otg, ob = run_both(BatchRuckig, f, n, 7, threads=16)
    total += assert_exact(otg, ob, label=f"re-planned states seed {seed}")
    del otg, ob
    n = 200_000
    f = workloads.config2_inputs(n, seed=seed)
    otg, ob = run_both(BatchRuckig, f, n, 6, threads=16)
    total += assert_exact(otg, ob, label=f"config2 seed {seed}")
    del otg, ob
    for dof in (1, 3):
        f = workloads.bench_inputs(n, dof, seed=seed)
        otg, ob = run_both(BatchRuckig, f, n, dof, threads=16)
        total += assert_exact(otg, ob, label=f"{dof}dof seed {seed}")
        del otg, ob
    f = workloads.edge_case_inputs(50_000, 4, seed=seed)
    otg, ob = run_both(BatchRuckig, f, 50_000, 4, threads=16)
    total += assert_exact(otg, ob, label=f"edge seed {seed}")
    del otg, ob
    # configs[2]: velocity interface + Phase, minimum_duration on a third, continuous and discrete durations
    from rsruckig_b200 import ControlInterface, DurationDiscretization, Synchronization
    n = 100_000
    f = workloads.config3_inputs(n, seed=seed)
    md = np.full(n, np.nan)
    rng = np.random.default_rng(seed)
    third = np.arange(n) % 3 == 1
    md[third] = rng.uniform(0.5, 8.0, int(third.sum()))
    f["minimum_duration"] = md
    for disc in (DurationDiscretization.Continuous, DurationDiscretization.Discrete):
        otg, ob = run_both(BatchRuckig, f, n, 3, threads=16, control_interface=ControlInterface.Velocity, synchronization=Synchronization.Phase,
                           duration_discretization=disc)
        total += assert_exact(otg, ob, label=f"config3 disc={int(disc)} seed {seed}")
        del otg, ob
    # position interface: Phase with a collinear quarter, mixed per-DoF settings, edge cases with Phase / both handlers
    f = workloads.bench_inputs(n, 3, seed=seed)
    col = np.arange(n) % 4 == 0
    scale = rng.uniform(0.25, 2.0, (3, n))
    scale[0] = 1.0
    pd = f["target_position"] - f["current_position"]
    f["target_position"][:, col] = (f["current_position"] + pd[0:1] * scale)[:, col]
    for k in ("current_velocity", "current_acceleration", "target_velocity", "target_acceleration", "max_velocity", "max_acceleration", "max_jerk"):
        f[k][:, col] = (f[k][0:1] * scale)[:, col]
    otg, ob = run_both(BatchRuckig, f, n, 3, threads=16, synchronization=Synchronization.Phase)
    total += assert_exact(otg, ob, label=f"phase seed {seed}")
    del otg, ob
    otg, ob = run_both(BatchRuckig, f, n, 3, threads=16,
                       per_dof_synchronization=[Synchronization.Phase, Synchronization.None_, Synchronization.TimeIfNecessary],
                       per_dof_control_interface=[ControlInterface.Position, ControlInterface.Velocity, ControlInterface.Position])
    total += assert_exact(otg, ob, label=f"mixed seed {seed}")
    del otg, ob
    f = workloads.edge_case_inputs(50_000, 5, seed=100 + seed)
    for sync in (Synchronization.Phase, Synchronization.TimeIfNecessary, Synchronization.None_):
        otg, ob = run_both(BatchRuckig, f, 50_000, 5, threads=16, synchronization=sync, error_mode=1)
        total += assert_exact(otg, ob, label=f"edge sync={int(sync)} ignore-errors seed {seed}")
        del otg, ob
    # second and first order
    f = workloads.bench_inputs(n, 3, seed=200 + seed)
    f["max_jerk"][:] = np.inf
    f["current_acceleration"][:] = 0.0
    f["target_acceleration"][:] = 0.0
    for ci in (ControlInterface.Position, ControlInterface.Velocity):
        otg, ob = run_both(BatchRuckig, f, n, 3, threads=16, control_interface=ci)
        total += assert_exact(otg, ob, label=f"second order ci={int(ci)} seed {seed}")
        del otg, ob
    f["max_acceleration"][:] = np.inf
    f["current_velocity"][:] = 0.0
    f["target_velocity"][:] = 0.0
    otg, ob = run_both(BatchRuckig, f, n, 3, threads=16)
    total += assert_exact(otg, ob, label=f"first order seed {seed}")
    del otg, ob
    print(f"seed {seed} ok, {total} Working trajectories compared so far, {time.time() - t0:.0f} s", flush=True)
print(f"PARITY SWEEP OK: {total} trajectories bit-identical (status, duration, limiting DoF, codes, all profile arrays)")
