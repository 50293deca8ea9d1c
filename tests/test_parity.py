"""Parity of the CUDA path with the oracle on seeded random batches (-m gpu), through the C ABI.

Bar (north_star): RuckigResult codes, profile type (limits + control signs) and direction exact; durations and sampled
p/v/a within 1e-9 relative. The device restates the oracle's *portable* math flavour bit for bit, so against that
flavour the tests demand exact equality of every stored double; against the glibc flavour (what a Linux build of the
Rust reference links) they apply the 1e-9 bar and require zero code mismatches.
"""
import numpy as np
import pytest

from rsruckig_b200 import _cabi as cabi
from rsruckig_b200 import workloads

pytestmark = pytest.mark.gpu

PROFILE_FIELDS = (cabi.RK_FIELD_T, cabi.RK_FIELD_T_SUM, cabi.RK_FIELD_J, cabi.RK_FIELD_A, cabi.RK_FIELD_V, cabi.RK_FIELD_P,
                  cabi.RK_FIELD_BRAKE_DURATION, cabi.RK_FIELD_BRAKE_T, cabi.RK_FIELD_BRAKE_J, cabi.RK_FIELD_BRAKE_A,
                  cabi.RK_FIELD_BRAKE_V, cabi.RK_FIELD_BRAKE_P, cabi.RK_FIELD_PF_VF_AF)


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsruckig_b200 import BatchRuckig
    return BatchRuckig


def run_both(BatchRuckig, fields, n, dof, *, delta_time=0.005, flavour="portable", threads=8, **kw):
    from oracle.oracle import OracleBatch
    from rsruckig_b200 import IgnoreErrorHandler, ThrowErrorHandler
    error_mode = kw.pop("error_mode", cabi.RK_ERRORS_THROW)
    otg = BatchRuckig(dof, delta_time, ThrowErrorHandler if error_mode == cabi.RK_ERRORS_THROW else IgnoreErrorHandler,
                      result_layout=kw.pop("result_layout", None))
    otg.calculate(fields, n=n, memory_space=cabi.RK_MEM_HOST, **kw)
    ob = OracleBatch(n, dof, flavour)
    ob.calculate(fields, delta_time=delta_time, threads=threads, error_mode=error_mode, **kw)
    return otg, ob


def assert_exact(otg, ob, *, label=""):
    st_g, st_o = otg.read(cabi.RK_FIELD_STATUS), ob.read(cabi.RK_FIELD_STATUS)
    bad = np.nonzero(st_g != st_o)[0]
    assert bad.size == 0, f"{label}: {bad.size} status mismatches, first at {bad[:5]}: gpu {st_g[bad[:5]]} oracle {st_o[bad[:5]]}"
    ok = st_o == 0
    du_g, du_o = otg.read(cabi.RK_FIELD_DURATION), ob.read(cabi.RK_FIELD_DURATION)
    keep = ok | (st_o == -101)
    bad = np.nonzero((du_g != du_o) & keep)[0]
    assert bad.size == 0, f"{label}: {bad.size} duration mismatches, first at {bad[:5]}: {du_g[bad[:5]]} vs {du_o[bad[:5]]}"
    assert np.array_equal(otg.read(cabi.RK_FIELD_LIMITING_DOF)[ok], ob.read(cabi.RK_FIELD_LIMITING_DOF)[ok]), f"{label}: limiting DoF"
    im_g, im_o = otg.read(cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS), ob.read(cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS)
    assert np.array_equal(im_g[:, ok], im_o[:, ok]), f"{label}: independent_min_durations"
    c_g, c_o = otg.read(cabi.RK_FIELD_CODES), ob.read(cabi.RK_FIELD_CODES)
    bad = np.argwhere((c_g != c_o) & ok[None, :])
    assert bad.shape[0] == 0, f"{label}: {bad.shape[0]} profile-code mismatches, first {bad[:5].tolist()}: {[hex(c_g[tuple(b)]) for b in bad[:5]]} vs {[hex(c_o[tuple(b)]) for b in bad[:5]]}"
    for f in PROFILE_FIELDS:
        a, b = otg.read(f)[:, :, ok], ob.read(f)[:, :, ok]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        bad = np.argwhere(~same)
        assert bad.shape[0] == 0, f"{label}: field {f}: {bad.shape[0]} differing doubles, first {bad[:5].tolist()}"
    return int(ok.sum())


def test_division_paths_are_ieee_exact(gpu):
    """The kernels divide through their own routines (shared refined reciprocal, zero-safe division; rk_math.cuh). Both
    must return the compiler's correctly rounded IEEE quotient bit for bit, special values included."""
    import ctypes as C
    otg = gpu(1, 0.005)
    lib = otg._engine.lib
    bad = C.c_int64(-1)
    for seed in (1, 2, 3):
        cabi.check(lib.rk_selftest_division(otg.handle, 200_000_000, seed * 0x9E3779B97F4A7C15 % (1 << 64), C.byref(bad)))
        assert bad.value == 0, f"{bad.value} division results differ from IEEE a / b (seed {seed})"


def test_bench_distribution_7dof_exact(gpu):
    n, dof = 100_000, 7
    f = workloads.bench_inputs(n, dof, seed=1)
    otg, ob = run_both(gpu, f, n, dof)
    n_ok = assert_exact(otg, ob, label="7dof")
    assert n_ok > 0.9 * n


def test_config2_6dof_exact(gpu):
    n = 100_000
    f = workloads.config2_inputs(n, seed=2)
    otg, ob = run_both(gpu, f, n, 6)
    assert_exact(otg, ob, label="6dof")


@pytest.mark.parametrize("dof", [1, 2, 3, 8, 38])
def test_other_dof_counts_exact(gpu, dof):
    n = 20_000 if dof < 38 else 3_000
    f = workloads.bench_inputs(n, dof, seed=10 + dof)
    otg, ob = run_both(gpu, f, n, dof)
    assert_exact(otg, ob, label=f"dof{dof}")


def test_edge_cases_exact(gpu):
    """Start states outside the limits (brake), rest-to-rest, zero displacement, tiny / huge limits, asymmetric minimum
    limits, disabled DoFs — under both error handlers."""
    n, dof = 60_000, 4
    f = workloads.edge_case_inputs(n, dof, seed=3)
    for mode in (cabi.RK_ERRORS_THROW, cabi.RK_ERRORS_IGNORE):
        otg, ob = run_both(gpu, f, n, dof, error_mode=mode)
        assert_exact(otg, ob, label=f"edge mode {mode}")


def test_disabled_dofs_in_zero_duration_and_single_dof_trajectories_exact(gpu):
    """calculator_target.rs:475-481 and :535-541 copy blocks[dof].p_min into the trajectory for EVERY DoF, disabled ones
    included, when the trajectory has a single DoF or zero duration — a disabled DoF then holds Profile::default() instead
    of the current state that :313-323 had put there (found by tools/parity_sweep.py, seed 16)."""
    n, dof = 4096, 4
    f = workloads.edge_case_inputs(n, dof, seed=16)
    f["enabled"][:, ::3] = 0                      # every third trajectory: all DoFs disabled -> duration 0
    at_target = slice(1, None, 3)                 # every third: already at the target, one DoF disabled -> duration 0
    for k in ("velocity", "acceleration"):
        f[f"current_{k}"][:, at_target] = 0.0
        f[f"target_{k}"][:, at_target] = 0.0
    f["target_position"][:, at_target] = f["current_position"][:, at_target]
    f["enabled"][:, at_target] = 1
    f["enabled"][2, at_target] = 0
    otg, ob = run_both(gpu, f, n, dof)
    assert assert_exact(otg, ob, label="zero duration, disabled DoFs") > n // 2
    assert np.all(ob.read(cabi.RK_FIELD_DURATION)[::3] == 0.0)
    f1 = workloads.edge_case_inputs(n, 1, seed=17)
    f1["enabled"][0, ::2] = 0
    otg, ob = run_both(gpu, f1, n, 1)
    assert_exact(otg, ob, label="single DoF, disabled")


def test_config3_velocity_phase_minimum_duration_discrete_exact(gpu):
    """BASELINE.json configs[2]: 3-DoF velocity interface + Phase synchronisation, thirds of the batch with no extra /
    minimum_duration / (separately) Discrete durations."""
    from rsruckig_b200 import ControlInterface, DurationDiscretization, Synchronization
    n = 60_000
    f = workloads.config3_inputs(n, seed=4)
    md = np.full(n, np.nan)
    rng = np.random.default_rng(5)
    third = np.arange(n) % 3 == 1
    md[third] = rng.uniform(0.5, 8.0, int(third.sum()))
    f["minimum_duration"] = md
    for disc in (DurationDiscretization.Continuous, DurationDiscretization.Discrete):
        otg, ob = run_both(gpu, f, n, 3, control_interface=ControlInterface.Velocity, synchronization=Synchronization.Phase,
                           duration_discretization=disc)
        assert_exact(otg, ob, label=f"config3 disc={int(disc)}")
    src = otg.read(cabi.RK_FIELD_CODES) >> 24
    assert (src == 3).sum() > 0, "no trajectory took the phase-synchronised path"


def test_position_phase_and_mixed_sync_exact(gpu):
    from rsruckig_b200 import ControlInterface, Synchronization
    n, dof = 40_000, 3
    f = workloads.bench_inputs(n, dof, seed=6)
    # collinear quarter: scaled copies of DoF 0 so that phase synchronisation succeeds
    col = np.arange(n) % 4 == 0
    rng = np.random.default_rng(7)
    scale = rng.uniform(0.25, 2.0, (dof, n))
    scale[0] = 1.0
    pd = f["target_position"] - f["current_position"]
    f["target_position"][:, col] = (f["current_position"] + pd[0:1] * scale)[:, col]
    for k in ("current_velocity", "current_acceleration", "target_velocity", "target_acceleration", "max_velocity", "max_acceleration", "max_jerk"):
        f[k][:, col] = (f[k][0:1] * scale)[:, col]
    otg, ob = run_both(gpu, f, n, dof, synchronization=Synchronization.Phase)
    assert_exact(otg, ob, label="phase")
    src = otg.read(cabi.RK_FIELD_CODES) >> 24
    assert (src == 3).sum() > 0, "no trajectory took the phase-synchronised path"
    otg, ob = run_both(gpu, f, n, dof, per_dof_synchronization=[Synchronization.Phase, Synchronization.None_, Synchronization.TimeIfNecessary],
                       per_dof_control_interface=[ControlInterface.Position, ControlInterface.Velocity, ControlInterface.Position])
    assert_exact(otg, ob, label="mixed")


def test_second_and_first_order_exact(gpu):
    """max_jerk = inf (acceleration-limited) and max_acceleration = inf as well (velocity-limited)."""
    from rsruckig_b200 import ControlInterface
    n, dof = 30_000, 3
    f = workloads.bench_inputs(n, dof, seed=8)
    f["max_jerk"][:] = np.inf
    f["current_acceleration"][:] = 0.0
    f["target_acceleration"][:] = 0.0
    otg, ob = run_both(gpu, f, n, dof)
    assert_exact(otg, ob, label="second order position")
    otg, ob = run_both(gpu, f, n, dof, control_interface=ControlInterface.Velocity)
    assert_exact(otg, ob, label="second order velocity")
    f["max_acceleration"][:] = np.inf
    f["current_velocity"][:] = 0.0
    f["target_velocity"][:] = 0.0
    otg, ob = run_both(gpu, f, n, dof)
    assert_exact(otg, ob, label="first order position")


def test_at_time_sampling_exact(gpu):
    """BASELINE.json configs[3] in small: calculate + K control cycles of Trajectory::at_time."""
    n, dof, K = 4096, 7, 200
    f = workloads.edge_case_inputs(n, dof, seed=9)
    otg, ob = run_both(gpu, f, n, dof)
    ok = ob.read(cabi.RK_FIELD_STATUS) == 0
    pos, vel, acc, jerk = (np.zeros((K, dof, n)) for _ in range(4))
    sec = np.zeros((K, n), dtype=np.int32)
    otg.at_time(time_start=0.0, delta_time=0.02, steps=K, memory_space=cabi.RK_MEM_HOST, position=pos, velocity=vel, acceleration=acc,
                jerk=jerk, section=sec)
    opos, ovel, oacc, ojerk, osec = ob.at_time(time_start=0.0, delta_time=0.02, steps=K, threads=8)
    for a, b, name in ((pos, opos, "position"), (vel, ovel, "velocity"), (acc, oacc, "acceleration"), (jerk, ojerk, "jerk")):
        assert np.array_equal(a[:, :, ok], b[:, :, ok]), name
    assert np.array_equal(sec[:, ok], osec[:, ok])
    # the position / velocity / acceleration-only form (its own kernel instantiation, the one the RL rollout uses), a longer
    # horizon with a coarser cycle so that every segment kind (brake, phases, own end, trajectory end) is crossed, and a
    # sliced rollout (time_start > 0 continues an earlier one)
    for K2, dt2, t_start in ((300, 0.05, 0.0), (64, 0.013, 0.731)):
        p2, v2, a2 = (np.zeros((K2, dof, n)) for _ in range(3))
        otg.at_time(time_start=t_start, delta_time=dt2, steps=K2, memory_space=cabi.RK_MEM_HOST, position=p2, velocity=v2, acceleration=a2)
        opos, ovel, oacc, _, _ = ob.at_time(time_start=t_start, delta_time=dt2, steps=K2, threads=8)
        for a, b, name in ((p2, opos, "position"), (v2, ovel, "velocity"), (a2, oacc, "acceleration")):
            assert np.array_equal(a[:, :, ok], b[:, :, ok]), f"{name} (p/v/a only, K={K2})"


def test_against_glibc_flavour_within_tolerance(gpu):
    """The 1e-9 bar of north_star against the reference-faithful (glibc libm) oracle build: zero code mismatches."""
    n, dof = 200_000, 7
    f = workloads.bench_inputs(n, dof, seed=11)
    otg, ob = run_both(gpu, f, n, dof, flavour="glibc")
    st_g, st_o = otg.read(cabi.RK_FIELD_STATUS), ob.read(cabi.RK_FIELD_STATUS)
    assert np.array_equal(st_g, st_o)
    ok = st_o == 0
    assert np.array_equal((otg.read(cabi.RK_FIELD_CODES) & 0xFFFFFF)[:, ok], (ob.read(cabi.RK_FIELD_CODES) & 0xFFFFFF)[:, ok])
    du_g, du_o = otg.read(cabi.RK_FIELD_DURATION)[ok], ob.read(cabi.RK_FIELD_DURATION)[ok]
    assert np.max(np.abs(du_g - du_o) / np.maximum(np.abs(du_o), 1e-12)) <= 1e-9
    t_g, t_o = otg.read(cabi.RK_FIELD_T)[:, :, ok], ob.read(cabi.RK_FIELD_T)[:, :, ok]
    assert np.max(np.abs(t_g - t_o) - 1e-9 * np.abs(t_o)) <= 1e-12


def test_device_buffers_and_full_size_properties(gpu):
    """BASELINE.json configs[1] at full size (1M x 6-DoF) on device buffers, checked through size-independent
    properties: every Working trajectory reaches its target state (|dp|,|dv| <= 1e-8, |da| <= 1e-10, the reference's
    guarantees, README.md:418-423) at t = duration, all DoFs end together, and duration >= every independent minimum."""
    import torch
    n, dof = 1 << 20, 6
    f = workloads.config2_inputs(n, seed=12)
    d = {k: torch.from_numpy(v).cuda() for k, v in f.items()}
    from rsruckig_b200 import ThrowErrorHandler
    otg = gpu(dof, 0.005, ThrowErrorHandler)
    status = torch.zeros(n, dtype=torch.int32, device="cuda")
    duration = torch.zeros(n, dtype=torch.float64, device="cuda")
    imd = torch.zeros((dof, n), dtype=torch.float64, device="cuda")
    otg.calculate(d, n=n, memory_space=cabi.RK_MEM_DEVICE, status=status, duration=duration, independent_min_durations=imd)
    otg.sync()
    ok = status == 0
    assert ok.float().mean() > 0.9
    assert set(torch.unique(status).tolist()) <= {0, -100, -101, -110, -111}
    assert bool((duration[ok][None, :] >= imd[:, ok] - 1e-12).all())
    P = torch.from_numpy(otg.read(cabi.RK_FIELD_P)).cuda()
    V = torch.from_numpy(otg.read(cabi.RK_FIELD_V)).cuda()
    A = torch.from_numpy(otg.read(cabi.RK_FIELD_A)).cuda()
    assert float((P[7] - d["target_position"]).abs()[:, ok].max()) <= 1e-8
    assert float((V[7] - d["target_velocity"]).abs()[:, ok].max()) <= 1e-8
    assert float((A[7] - d["target_acceleration"]).abs()[:, ok].max()) <= 1e-10
    ts6 = torch.from_numpy(otg.read(cabi.RK_FIELD_T_SUM)).cuda()[6] + torch.from_numpy(otg.read(cabi.RK_FIELD_BRAKE_DURATION)).cuda()[0]
    assert float((ts6 - duration[None, :]).abs()[:, ok].max()) <= 1e-9 * float(duration[ok].max())


@pytest.mark.gpu
def test_host_buffers_over_several_chunks_return_results_chunk_by_chunk(gpu):
    """A host-buffer batch of more than one chunk: the first chunk is computed in pieces while later inputs are still being
    copied, and status / duration / independent_min_durations come back chunk by chunk on their own stream. 20 DoFs make a
    chunk 209 715 trajectories, so 450 000 trajectories are three chunks (the first in four pieces). Everything the call
    returns equals what the trajectories object holds, and both equal the oracle bit for bit."""
    n, dof = 450_000, 20
    f = workloads.bench_inputs(n, dof, seed=21)
    from oracle.oracle import OracleBatch
    from rsruckig_b200 import ThrowErrorHandler
    otg = gpu(dof, 0.005, ThrowErrorHandler)
    status = np.full(n, 12345, dtype=np.int32)
    duration = np.full(n, -1.0)
    imd = np.full((dof, n), -1.0)
    for _ in range(2):  # the second call re-uses the staging buffers and events of the first
        status[:] = 12345
        otg.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST, status=status, duration=duration, independent_min_durations=imd)
        assert np.array_equal(status, otg.read(cabi.RK_FIELD_STATUS))
        assert np.array_equal(duration, otg.read(cabi.RK_FIELD_DURATION))
        assert np.array_equal(imd, otg.read(cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS))
    ob = OracleBatch(n, dof, "portable")
    ob.calculate(f, threads=16)
    st_o = ob.read(cabi.RK_FIELD_STATUS)
    assert np.array_equal(status, st_o)
    ok = st_o == 0
    assert ok.mean() > 0.8
    assert np.array_equal(duration[ok], ob.read(cabi.RK_FIELD_DURATION)[ok])
    assert np.array_equal(imd[:, ok], ob.read(cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS)[:, ok])
    for fld in (cabi.RK_FIELD_T, cabi.RK_FIELD_P):
        assert np.array_equal(otg.read(fld)[:, :, ok], ob.read(fld)[:, :, ok])


@pytest.mark.parametrize("n,dof", [(1, 7), (2, 7), (31, 7), (33, 3), (127, 7), (129, 1), (4097, 7),
                                   (28087, 7),      # 196 609 items: the smallest 7-DoF batch that is split in two (ragged) halves
                                   (524289, 8)])    # one trajectory more than a full chunk of 4 Mi items: two ragged chunks
def test_ragged_and_boundary_sizes_exact(gpu, n, dof):
    """Batch sizes that are not multiples of the warp / block / chunk size, and the two sizes at which the chunking changes."""
    f = workloads.edge_case_inputs(n, dof, seed=n % 97) if n < 5000 else workloads.bench_inputs(n, dof, seed=n % 97)
    otg, ob = run_both(gpu, f, n, dof, threads=16)
    assert_exact(otg, ob, label=f"n={n} dof={dof}")


def test_empty_batch_is_a_no_op(gpu):
    from rsruckig_b200 import ThrowErrorHandler
    otg = gpu(7, 0.005, ThrowErrorHandler)
    f = {k: v[:, :0].copy() for k, v in workloads.bench_inputs(4, 7, seed=1).items()}
    status = np.zeros(0, dtype=np.int32)
    duration = np.zeros(0)
    before = otg.kernel_launches()  # the handle of a device is shared by the calculators of a process
    otg.calculate(f, n=0, memory_space=cabi.RK_MEM_HOST, status=status, duration=duration)
    otg.sync()
    assert otg.kernel_launches() == before


def _same_results(a, b, *, label):
    """Two BatchRuckig objects hold the same results bit for bit (status, duration, codes, every profile field)."""
    for f in (cabi.RK_FIELD_STATUS, cabi.RK_FIELD_DURATION, cabi.RK_FIELD_LIMITING_DOF, cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS,
              cabi.RK_FIELD_CODES, cabi.RK_FIELD_ERROR_DETAIL) + PROFILE_FIELDS:
        x, y = a.read(f), b.read(f)
        same = (x == y) | ((x != x) & (y != y))
        assert same.all(), f"{label}: field {f} differs in {int((~same).sum())} places"


def test_limits_per_dof_layout(gpu):
    """RK_LIMITS_PER_DOF: the five limit arrays as [dof] (a fleet of identical robots) give the same bits as the same values
    spelled out per trajectory — host buffers staged in several chunks, device buffers, asymmetric minimum limits, validation,
    and the closed loop (rk_update_batch)."""
    import torch
    from rsruckig_b200 import BatchRollout, ThrowErrorHandler
    dof = 7
    rng = np.random.default_rng(5)
    lim = {"max_velocity": rng.uniform(2.0, 12.0, dof), "max_acceleration": rng.uniform(2.0, 12.0, dof), "max_jerk": rng.uniform(1.0, 12.0, dof),
           "min_velocity": -rng.uniform(2.0, 12.0, dof), "min_acceleration": -rng.uniform(2.0, 12.0, dof)}
    for n, space in ((700_001, cabi.RK_MEM_HOST), (50_000, cabi.RK_MEM_DEVICE)):  # 700 001 x 7 items = two chunks of scratch
        f = workloads.bench_inputs(n, dof, seed=21)
        # every 50th target velocity lies outside the shared limits: valid and invalid problems in both layouts
        f["target_velocity"] = f["target_velocity"].copy()
        f["target_velocity"][3, ::50] = 20.0
        full = dict(f)
        for k, v in lim.items():
            full[k] = np.ascontiguousarray(np.repeat(v[:, None], n, axis=1))
        shared = dict(f)
        shared.update({k: np.ascontiguousarray(v) for k, v in lim.items()})
        if space == cabi.RK_MEM_DEVICE:
            full = {k: torch.from_numpy(v).cuda() for k, v in full.items()}
            shared = {k: torch.from_numpy(v).cuda() for k, v in shared.items()}
        a, b = gpu(dof, 0.005, ThrowErrorHandler), gpu(dof, 0.005, ThrowErrorHandler)
        a.calculate(full, n=n, memory_space=space)
        b.calculate(shared, n=n, memory_space=space, limits_layout=cabi.RK_LIMITS_PER_DOF)
        st = a.read(cabi.RK_FIELD_STATUS)
        assert int((st == 0).sum()) > 0.3 * n, f"only {int((st == 0).sum())} of {n} problems are valid"
        assert int((st == -100).sum()) >= n // 50, "the invalid problems were not reported"
        _same_results(a, b, label=f"limits per DoF, n = {n}, memory space {space}")
    # validate_input
    n = 20_000
    f = workloads.bench_inputs(n, dof, seed=22)
    full, shared = dict(f), dict(f)
    for k, v in lim.items():
        full[k] = np.ascontiguousarray(np.repeat(v[:, None], n, axis=1))
        shared[k] = np.ascontiguousarray(v)
    otg = gpu(dof, 0.005, ThrowErrorHandler)
    for cc, ct in ((True, True), (False, True)):
        va, ra = otg.validate(full, n=n, check_current=cc, check_target=ct)
        vb, rb = otg.validate(shared, n=n, check_current=cc, check_target=ct, limits_layout=cabi.RK_LIMITS_PER_DOF)
        assert np.array_equal(va, vb) and np.array_equal(ra, rb)
    # closed loop: ten cycles, every cycle's state equal
    n = 4096
    f = workloads.bench_inputs(n, dof, seed=23)
    states = []
    for layout in (cabi.RK_LIMITS_PER_TRAJECTORY, cabi.RK_LIMITS_PER_DOF):
        g = {k: torch.from_numpy(v.copy()).cuda() for k, v in f.items()}
        for k, v in lim.items():
            g[k] = torch.from_numpy(np.repeat(v[:, None], n, axis=1).copy() if layout == cabi.RK_LIMITS_PER_TRAJECTORY else v.copy()).cuda()
        roll = BatchRollout(n, dof, 0.01, ThrowErrorHandler)
        pos = torch.zeros((dof, n), dtype=torch.float64, device="cuda")
        res = torch.zeros(n, dtype=torch.int32, device="cuda")
        seq = []
        for _ in range(10):
            roll.update(g, new_position=pos, result=res, pass_to_input=True, limits_layout=layout)
            roll.sync()  # the library works on its own stream
            seq.append((pos.cpu().numpy().copy(), res.cpu().numpy().copy()))
        states.append(seq)
    for cycle, ((pa, ra), (pb, rb)) in enumerate(zip(*states)):
        bad = np.nonzero(ra != rb)[0]
        assert bad.size == 0, f"cycle {cycle}: {bad.size} results differ, first at {bad[:5]}: {ra[bad[:5]]} vs {rb[bad[:5]]}"
        bad = np.argwhere(~((pa == pb) | ((pa != pa) & (pb != pb))))
        assert bad.shape[0] == 0, f"cycle {cycle}: {bad.shape[0]} positions differ, first at {bad[:5].tolist()}"


def test_lazy_kernels_exact_redo_path(gpu):
    """k1_closed_cands and two chain stages run from the lazy instantiation (rk_math.cuh: a `Den` division outside its fast path's
    domain raises a per-thread flag, the kernel then recomputes the work item through the exact instantiation). Ordinary data
    never raises the flag, so RK_B200_FORCE_REDO=1 (read when the library first builds a parameter block, hence a process of its
    own) sends every fourth work item through the redo path: the results must not change by a bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RK_B200_FORCE_REDO="1")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "quick_parity.py")], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "compared" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
