"""RK_RESULT_COMPACT (20 doubles per DoF instead of the 59 of the reference's Profile): every reader expands the
record with the operations the full layout is filled with, so everything a caller can read — all Profile fields through
rk_trajectories_read, Trajectory::at_time samples, the query helpers — must equal the oracle bit for bit exactly as in the
full layout (profile.rs:76-118, trajectory.rs:52-246)."""
import numpy as np
import pytest

from rsruckig_b200 import _cabi as cabi
from rsruckig_b200 import workloads
from test_parity import assert_exact, run_both

pytestmark = pytest.mark.gpu
COMPACT = cabi.RK_RESULT_COMPACT


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsruckig_b200 import BatchRuckig
    return BatchRuckig


def test_compact_bench_and_edge_cases_exact(gpu):
    n = 100_000
    otg, ob = run_both(gpu, workloads.bench_inputs(n, 7, seed=31), n, 7, result_layout=COMPACT)
    assert assert_exact(otg, ob, label="compact 7dof") > 0.9 * n
    f = workloads.edge_case_inputs(60_000, 4, seed=32)  # brakes, disabled DoFs, error statuses, asymmetric limits
    for mode in (cabi.RK_ERRORS_THROW, cabi.RK_ERRORS_IGNORE):
        otg, ob = run_both(gpu, f, 60_000, 4, error_mode=mode, result_layout=COMPACT)
        assert_exact(otg, ob, label=f"compact edge mode {mode}")


def test_compact_other_interfaces_orders_and_synchronisations_exact(gpu):
    from rsruckig_b200 import ControlInterface, DurationDiscretization, Synchronization
    n = 40_000
    f = workloads.config3_inputs(n, seed=33)
    md = np.full(n, np.nan)
    md[1::3] = np.random.default_rng(34).uniform(0.5, 8.0, md[1::3].size)
    f["minimum_duration"] = md
    for disc in (DurationDiscretization.Continuous, DurationDiscretization.Discrete):
        otg, ob = run_both(gpu, f, n, 3, control_interface=ControlInterface.Velocity, synchronization=Synchronization.Phase,
                           duration_discretization=disc, result_layout=COMPACT)
        assert_exact(otg, ob, label=f"compact config3 disc={int(disc)}")
    assert ((otg.read(cabi.RK_FIELD_CODES) >> 24) == 3).sum() > 0, "no phase-synchronised profile in the batch"
    # position interface, Phase with a collinear quarter (phase profiles are stored with the limiting DoF's timing), mixed settings
    f = workloads.bench_inputs(n, 3, seed=35)
    col = np.arange(n) % 4 == 0
    scale = np.random.default_rng(36).uniform(0.25, 2.0, (3, n))
    scale[0] = 1.0
    pd = f["target_position"] - f["current_position"]
    f["target_position"][:, col] = (f["current_position"] + pd[0:1] * scale)[:, col]
    for k in ("current_velocity", "current_acceleration", "target_velocity", "target_acceleration", "max_velocity", "max_acceleration", "max_jerk"):
        f[k][:, col] = (f[k][0:1] * scale)[:, col]
    otg, ob = run_both(gpu, f, n, 3, synchronization=Synchronization.Phase, result_layout=COMPACT)
    assert_exact(otg, ob, label="compact phase")
    otg, ob = run_both(gpu, f, n, 3, result_layout=COMPACT,
                       per_dof_synchronization=[Synchronization.Phase, Synchronization.None_, Synchronization.TimeIfNecessary],
                       per_dof_control_interface=[ControlInterface.Position, ControlInterface.Velocity, ControlInterface.Position])
    assert_exact(otg, ob, label="compact mixed")
    # second and first order (with brake pre-trajectories of the second-order kind: velocities beyond the limits)
    f = workloads.edge_case_inputs(n, 3, seed=37)
    f["max_jerk"][:] = np.inf
    f["current_acceleration"][:] = 0.0
    f["target_acceleration"][:] = 0.0
    for ci in (ControlInterface.Position, ControlInterface.Velocity):
        otg, ob = run_both(gpu, f, n, 3, control_interface=ci, result_layout=COMPACT)
        assert_exact(otg, ob, label=f"compact second order ci={int(ci)}")
    f["max_acceleration"][:] = np.inf
    f["current_velocity"][:] = 0.0
    f["target_velocity"][:] = 0.0
    otg, ob = run_both(gpu, f, n, 3, result_layout=COMPACT)
    assert_exact(otg, ob, label="compact first order")


def test_compact_at_time_and_queries_exact(gpu):
    n, dof = 8192, 7
    f = workloads.edge_case_inputs(n, dof, seed=38)
    otg, ob = run_both(gpu, f, n, dof, result_layout=COMPACT)
    ok = ob.read(cabi.RK_FIELD_STATUS) == 0
    for K, dt, t_start, all_outputs in ((200, 0.02, 0.0, True), (300, 0.05, 0.0, False), (64, 0.013, 0.731, False), (50, -0.07, 6.0, True)):
        pos, vel, acc, jerk = (np.zeros((K, dof, n)) for _ in range(4))
        sec = np.zeros((K, n), dtype=np.int32)
        extra = dict(jerk=jerk, section=sec) if all_outputs else {}
        otg.at_time(time_start=t_start, delta_time=dt, steps=K, memory_space=cabi.RK_MEM_HOST, position=pos, velocity=vel, acceleration=acc, **extra)
        o = ob.at_time(time_start=t_start, delta_time=dt, steps=K, threads=8)
        names = ("position", "velocity", "acceleration") + (("jerk",) if all_outputs else ())
        for a, b, name in zip((pos, vel, acc, jerk), o[:4], names):
            assert np.array_equal(a[:, :, ok], b[:, :, ok]), f"{name} K={K} dt={dt}"
        if all_outputs:
            assert np.array_equal(sec[:, ok], o[4][:, ok])
    assert np.array_equal(otg.position_extrema()[:, :, ok], ob.position_extrema()[:, :, ok])
    q = 0.5 * (f["current_position"] + f["target_position"])
    a, b = otg.first_time_at_position(q)[:, ok], ob.first_time_at_position(q)[:, ok]
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def test_backwards_sample_times_exact_in_the_full_layout(gpu):
    """Trajectory::at_time is order-independent (trajectory.rs:52-189): a negative cycle time samples backwards."""
    n, dof, K = 4096, 5, 120
    f = workloads.edge_case_inputs(n, dof, seed=39)
    otg, ob = run_both(gpu, f, n, dof)
    ok = ob.read(cabi.RK_FIELD_STATUS) == 0
    pos, vel, acc, jerk = (np.zeros((K, dof, n)) for _ in range(4))
    sec = np.zeros((K, n), dtype=np.int32)
    otg.at_time(time_start=9.0, delta_time=-0.08, steps=K, memory_space=cabi.RK_MEM_HOST, position=pos, velocity=vel, acceleration=acc,
                jerk=jerk, section=sec)
    o = ob.at_time(time_start=9.0, delta_time=-0.08, steps=K, threads=8)
    for a, b, name in zip((pos, vel, acc, jerk), o[:4], ("position", "velocity", "acceleration", "jerk")):
        assert np.array_equal(a[:, :, ok], b[:, :, ok]), name
    assert np.array_equal(sec[:, ok], o[4][:, ok])


def test_compact_memory_footprint(gpu):
    """What the layout is for: 64 Mi x 7-DoF results must fit one B200 next to inputs and scratch — 20 instead of 59 doubles."""
    import torch
    from rsruckig_b200 import ThrowErrorHandler
    n, dof = 1 << 20, 7
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    otg = gpu(dof, 0.005, ThrowErrorHandler, result_layout=COMPACT)
    otg._ensure(n)
    used_compact = free0 - torch.cuda.mem_get_info()[0]
    otg.close()
    otg = gpu(dof, 0.005, ThrowErrorHandler, result_layout=cabi.RK_RESULT_FULL)
    otg._ensure(n)
    used_full = free0 - torch.cuda.mem_get_info()[0]
    otg.close()
    assert used_compact < 0.45 * used_full, (used_compact, used_full)
