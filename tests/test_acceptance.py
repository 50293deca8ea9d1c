"""north_star's acceptance bar, at north_star's sizes, in the driver-run GPU suite (-m gpu), through the C ABI.

  (a) 10 M random 7-DoF bench inputs (20 seeds x 500 k): against the reference-faithful oracle build (glibc libm) zero
      RuckigResult / profile-type (limits + control signs) / direction / limiting-DoF mismatches, durations and sampled
      p/v/a within 1e-9 relative; against the portable build every compared double bit for bit.
  (b) the closed-loop state distribution: >= 1 M problems re-planned from states ON their trajectories (velocities and
      accelerations riding a limit, sometimes one ulp beyond: brake pre-trajectories and the Step-1 rescues decide).
  (c) rk_validate_batch (Ruckig::validate_input, ruckig.rs:150-171) for all four flag pairs, reason texts included.
  (d) BASELINE.json configs[3] at full size: 64 Ki x 7-DoF calculate + 1000 control cycles of Trajectory::at_time, compared
      with the oracle through one exact integer checksum per (cycle, DoF) row of position / velocity / acceleration.

Reference: calculator_target.rs:278-833, ruckig.rs:150-171, :293, trajectory.rs:52-189, tests_known.rs:15-27.
"""
import os

import numpy as np
import pytest

from rsruckig_b200 import _cabi as cabi
from rsruckig_b200 import workloads

pytestmark = pytest.mark.gpu

THREADS = max(4, min(32, os.cpu_count() or 4))
REL = 1e-9  # north_star: "durations and sampled p/v/a within 1e-9 relative (fp64)"


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsruckig_b200 import BatchRuckig
    return BatchRuckig


def rel_err(a, b):
    """|a - b| relative to max(|b|, 1): positions / velocities / accelerations pass through zero, durations do not."""
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


def test_ten_million_bench_inputs_match_the_reference_oracle(gpu):
    from oracle.oracle import OracleBatch
    from rsruckig_b200 import ThrowErrorHandler
    n, dof, seeds = 500_000, 7, 20
    K, dt = 4, 0.9  # sampled states at t = 0.9, 1.8, 2.7, 3.6 s (the median duration of this distribution is ~4 s)
    otg = gpu(dof, 0.005, ThrowErrorHandler)
    ob_p, ob_g = OracleBatch(n, dof, "portable"), OracleBatch(n, dof, "glibc")
    total = working = 0
    worst_dur = worst_pva = 0.0
    pos, vel, acc = (np.zeros((K, dof, n)) for _ in range(3))
    for seed in range(100, 100 + seeds):
        f = workloads.bench_inputs(n, dof, seed=seed)
        otg.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST)
        ob_p.calculate(f, threads=THREADS)
        ob_g.calculate(f, threads=THREADS)
        st = otg.read(cabi.RK_FIELD_STATUS)
        du = otg.read(cabi.RK_FIELD_DURATION)
        lim = otg.read(cabi.RK_FIELD_LIMITING_DOF)
        codes = otg.read(cabi.RK_FIELD_CODES)
        t = otg.read(cabi.RK_FIELD_T)
        otg.at_time(time_start=0.0, delta_time=dt, steps=K, memory_space=cabi.RK_MEM_HOST, position=pos, velocity=vel, acceleration=acc)

        # portable flavour: bit for bit
        st_p = ob_p.read(cabi.RK_FIELD_STATUS)
        assert np.array_equal(st, st_p), f"seed {seed}: {(st != st_p).sum()} status mismatches (portable)"
        ok = st_p == 0
        assert np.array_equal(du[ok], ob_p.read(cabi.RK_FIELD_DURATION)[ok]), f"seed {seed}: durations differ (portable)"
        assert np.array_equal(lim[ok], ob_p.read(cabi.RK_FIELD_LIMITING_DOF)[ok]), f"seed {seed}: limiting DoF (portable)"
        assert np.array_equal(codes[:, ok], ob_p.read(cabi.RK_FIELD_CODES)[:, ok]), f"seed {seed}: profile codes (portable)"
        assert np.array_equal(t[:, :, ok], ob_p.read(cabi.RK_FIELD_T)[:, :, ok]), f"seed {seed}: Profile::t (portable)"
        assert np.array_equal(otg.read(cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS)[:, ok],
                              ob_p.read(cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS)[:, ok]), f"seed {seed}: independent_min_durations"
        o = ob_p.at_time(time_start=0.0, delta_time=dt, steps=K, threads=THREADS, want_jerk=False)
        for a, b, name in ((pos, o[0], "position"), (vel, o[1], "velocity"), (acc, o[2], "acceleration")):
            assert np.array_equal(a[:, :, ok], b[:, :, ok]), f"seed {seed}: sampled {name} (portable)"

        # glibc flavour (what a Linux build of the Rust crate links): codes exact, doubles within 1e-9 relative
        st_g = ob_g.read(cabi.RK_FIELD_STATUS)
        assert np.array_equal(st, st_g), f"seed {seed}: {(st != st_g).sum()} RuckigResult mismatches (glibc)"
        ok = st_g == 0
        c_g = ob_g.read(cabi.RK_FIELD_CODES)
        bad = ((codes ^ c_g) & 0xFFFFFF)[:, ok]  # limits | control signs << 8 | direction << 16
        assert not bad.any(), f"seed {seed}: {np.count_nonzero(bad)} profile type / direction mismatches (glibc)"
        assert np.array_equal(lim[ok], ob_g.read(cabi.RK_FIELD_LIMITING_DOF)[ok]), f"seed {seed}: limiting DoF (glibc)"
        du_g = ob_g.read(cabi.RK_FIELD_DURATION)
        worst_dur = max(worst_dur, float(np.max(np.abs(du[ok] - du_g[ok]) / np.abs(du_g[ok]).clip(1e-300))) if ok.any() else 0.0)
        o = ob_g.at_time(time_start=0.0, delta_time=dt, steps=K, threads=THREADS, want_jerk=False)
        for a, b in ((pos, o[0]), (vel, o[1]), (acc, o[2])):
            worst_pva = max(worst_pva, float(rel_err(a[:, :, ok], b[:, :, ok]).max()))
        total += n
        working += int(ok.sum())
    assert total == 10_000_000
    assert working > 0.9 * total
    assert worst_dur <= REL, f"max relative duration error {worst_dur:.3e}"
    assert worst_pva <= REL, f"max relative sampled p/v/a error {worst_pva:.3e}"
    print(f"\n[acceptance] {total} trajectories ({working} Working): 0 code mismatches, max rel. duration error {worst_dur:.2e}, "
          f"max rel. sampled p/v/a error {worst_pva:.2e} vs glibc oracle; bitwise vs portable oracle")


def test_reference_benchmark_stream_first_million(gpu):
    """The reference benchmark's own inputs (Pcg64Mcg seeds 42 / 43 / 44 + rand_distr, benchmark_target.rs:115-153; restated in
    csrc/refstream.c and pinned on rand_pcg's known answers): the first 2^20 problems, same bar as the 10 M run."""
    from test_parity import assert_exact, run_both
    n, dof = 1 << 20, 7
    f = workloads.reference_bench_inputs(n, dof)
    otg, ob = run_both(gpu, f, n, dof, threads=THREADS, result_layout=cabi.RK_RESULT_COMPACT)
    working = assert_exact(otg, ob, label="reference stream (portable)")
    assert working > 0.9 * n
    from oracle.oracle import OracleBatch
    og = OracleBatch(n, dof, "glibc")
    og.calculate(f, threads=THREADS)
    st = og.read(cabi.RK_FIELD_STATUS)
    assert np.array_equal(otg.read(cabi.RK_FIELD_STATUS), st)
    ok = st == 0
    assert not (((otg.read(cabi.RK_FIELD_CODES) ^ og.read(cabi.RK_FIELD_CODES)) & 0xFFFFFF)[:, ok]).any()
    du, du_g = otg.read(cabi.RK_FIELD_DURATION)[ok], og.read(cabi.RK_FIELD_DURATION)[ok]
    assert float(np.max(np.abs(du - du_g) / du_g.clip(1e-300))) <= REL


def replanned_states(f, ob, n, frac):
    """Inputs whose current state is the state each (Working) trajectory of `ob` passes through at about frac * duration,
    on a grid of 32 sample times accumulated by repeated addition like a control loop would."""
    T, ok = ob.read(cabi.RK_FIELD_DURATION), ob.read(cabi.RK_FIELD_STATUS) == 0
    K = 32
    dt = float(np.median(T[ok])) / 12
    p, v, a, _, _ = ob.at_time(time_start=0.0, delta_time=dt, steps=K, threads=THREADS, want_jerk=False)
    grid = np.cumsum(np.full(K, dt))
    k = np.searchsorted(grid, frac * T, side="right") - 1
    idx = np.nonzero(ok & (k >= 0))[0]
    g = {name: arr.copy() for name, arr in f.items()}
    for name, arr in (("current_position", p), ("current_velocity", v), ("current_acceleration", a)):
        g[name][:, idx] = arr[k[idx], :, idx].T
    return g, idx.size


def test_one_million_replanned_closed_loop_states_exact(gpu):
    from test_parity import assert_exact, run_both
    n, dof = 150_000, 7
    replanned = 0
    hit_brake = hit_error = 0
    for seed, frac in ((300, 0.6), (301, 0.3), (302, 0.8), (303, 0.5), (304, 0.25), (305, 0.7), (306, 0.4), (307, 0.9), (308, 0.55)):
        f = workloads.bench_inputs(n, dof, seed=seed)
        from oracle.oracle import OracleBatch
        ob0 = OracleBatch(n, dof, "portable")
        ob0.calculate(f, threads=THREADS)
        g, m = replanned_states(f, ob0, n, frac)
        del ob0
        otg, ob = run_both(gpu, g, n, dof, threads=THREADS)
        assert_exact(otg, ob, label=f"re-planned states seed {seed}")
        replanned += m
        hit_brake += int((ob.read(cabi.RK_FIELD_BRAKE_DURATION)[0] > 0).sum())
        hit_error += int((ob.read(cabi.RK_FIELD_STATUS) != 0).sum())
        del otg, ob
    assert replanned >= 1_000_000, replanned
    assert hit_brake > 0, "no re-planned state needed a brake pre-trajectory: the distribution is not the closed-loop one"
    print(f"\n[acceptance] {replanned} problems re-planned from states on their trajectories, {hit_brake} brake pre-trajectories, "
          f"{hit_error} error statuses: every stored double identical to the oracle")


def oracle_text(reason):
    """The oracle's (abridged) text for a GPU reason word: check id << 8 | DoF."""
    from rsruckig_b200.api import _VALIDATION_MESSAGES
    cid, d = reason >> 8, reason & 0xFF
    if cid == 24:
        return ""  # the delta_time rule of Ruckig::validate_input has no per-DoF text in the oracle's batch entry
    return _VALIDATION_MESSAGES[cid].replace(" of DoF {dof}", "").replace("DoF {dof} ", "") + f" (DoF {d})"


@pytest.mark.parametrize("check_current,check_target", [(False, False), (False, True), (True, False), (True, True)])
def test_validate_batch_all_flag_pairs_match_the_oracle(gpu, check_current, check_target):
    from oracle.oracle import OracleBatch
    from rsruckig_b200 import ControlInterface, DurationDiscretization
    n, dof = 200_000, 5
    f = workloads.edge_case_inputs(n, dof, seed=40)
    rng = np.random.default_rng(41)
    # salt with every failing check of input_parameter.rs:251-420: NaNs, negative / wrong-sign limits, states beyond limits
    for name in ("max_jerk", "max_acceleration", "max_velocity", "current_acceleration", "target_acceleration", "current_velocity",
                 "target_velocity", "current_position", "target_position", "min_velocity", "min_acceleration"):
        m = rng.random((dof, n)) < 0.004
        f[name][m] = np.nan
    for name in ("max_jerk", "max_acceleration", "max_velocity"):
        m = rng.random((dof, n)) < 0.004
        f[name][m] = -np.abs(f[name][m]) - 0.1
    for name in ("min_velocity", "min_acceleration"):
        m = rng.random((dof, n)) < 0.004
        f[name][m] = np.abs(f[name][m]) + 0.1
    m = rng.random((dof, n)) < 0.02
    f["target_velocity"][m] = (f["max_velocity"] * rng.uniform(0.9, 1.3, (dof, n)) * rng.choice([-1.0, 1.0], (dof, n)))[m]
    m = rng.random((dof, n)) < 0.02
    f["target_acceleration"][m] = (f["max_acceleration"] * rng.uniform(0.9, 1.3, (dof, n)) * rng.choice([-1.0, 1.0], (dof, n)))[m]
    seen = set()
    for kw in (dict(),
               dict(per_dof_control_interface=[ControlInterface.Position, ControlInterface.Velocity, ControlInterface.Position,
                                               ControlInterface.Velocity, ControlInterface.Position]),
               dict(duration_discretization=DurationDiscretization.Discrete, delta_time=0.0)):
        delta_time = kw.pop("delta_time", 0.005)
        otg = gpu(dof, delta_time)
        valid, reason = otg.validate(f, n=n, check_current=check_current, check_target=check_target, **kw)
        ob = OracleBatch(n, dof, "portable")
        o_valid, o_msg = ob.validate_messages(f, check_current=check_current, check_target=check_target, delta_time=delta_time, **kw)
        assert np.array_equal(valid, o_valid), f"{(valid != o_valid).sum()} validity mismatches"
        assert np.array_equal(valid == 1, reason == 0)
        for r in np.unique(reason[reason != 0]):
            texts = np.unique(o_msg[reason == r])
            assert texts.size == 1 and texts[0].decode() == oracle_text(int(r)), (hex(int(r)), texts[:3], oracle_text(int(r)))
            seen.add(int(r) >> 8)
        assert 0.02 < 1.0 - valid.mean() < 0.98 or delta_time == 0.0
    expected = {1, 2, 3, 4, 5, 10, 11, 12, 13, 14, 15, 24}
    if check_current:
        expected |= {6, 7, 16, 17, 20, 21}
    if check_target:
        expected |= {8, 9, 18, 19, 22, 23}
    assert expected <= seen, f"checks never triggered: {sorted(expected - seen)}"


def test_config4_full_size_rollout_checksums(gpu):
    """64 Ki x 7-DoF x 1000 control cycles (11 GB of samples): the GPU writes the whole rollout into device memory with one
    call; both sides reduce every (cycle, DoF) row of position / velocity / acceleration over the Working trajectories to one
    wrapping int64 sum of the doubles' bit patterns — equal sums for all 21 000 rows."""
    import torch
    from oracle.oracle import OracleBatch
    from rsruckig_b200 import ThrowErrorHandler
    n, dof, K, dt = 65_536, 7, 1000, 0.005
    f = workloads.bench_inputs(n, dof, seed=77)
    otg = gpu(dof, dt, ThrowErrorHandler)
    otg.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST)
    ob = OracleBatch(n, dof, "portable")
    ob.calculate(f, threads=THREADS)
    st = ob.read(cabi.RK_FIELD_STATUS)
    assert np.array_equal(otg.read(cabi.RK_FIELD_STATUS), st)
    ok = st == 0
    assert ok.mean() > 0.9
    ok_idx = np.nonzero(ok)[0]
    dev = [torch.empty((K, dof, n), dtype=torch.float64, device="cuda") for _ in range(3)]
    otg.at_time(time_start=0.0, delta_time=dt, steps=K, memory_space=cabi.RK_MEM_DEVICE, position=dev[0], velocity=dev[1],
                acceleration=dev[2])
    otg.sync()
    ok_dev = torch.from_numpy(ok_idx).cuda()
    sums = [x.view(torch.int64).index_select(2, ok_dev).sum(dim=2).cpu().numpy() for x in dev]
    del dev
    # the oracle in slices of 50 cycles; the start time of a slice is the sum the loop of ruckig.rs:293 has reached by then
    t_start, step = 0.0, 50
    for k0 in range(0, K, step):
        o = ob.at_time(time_start=t_start, delta_time=dt, steps=step, threads=THREADS, want_jerk=False)
        for which, name in enumerate(("position", "velocity", "acceleration")):
            want = np.ascontiguousarray(o[which][:, :, ok_idx]).view(np.int64).sum(axis=2)
            bad = np.argwhere(want != sums[which][k0:k0 + step])
            assert bad.shape[0] == 0, f"{name}: {bad.shape[0]} (cycle, DoF) rows differ in cycles {k0}..{k0 + step}, first {bad[:3].tolist()}"
        for _ in range(step):
            t_start += dt
    durations = ob.read(cabi.RK_FIELD_DURATION)[ok]
    assert (durations < K * dt).mean() > 0.1 and (durations > K * dt).mean() > 0.1, "the horizon should cut through the durations"
