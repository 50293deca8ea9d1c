"""The guarantees the reference documents for every trajectory it returns (README.md:418-423) checked on the oracle itself
(CPU): SURVEY.md 8c safeguard (1). The reference's own tests pin durations to 1e-4; these invariants hold the restatement to
1e-8 / 1e-8 / 1e-10 on random inputs, independently of any golden vector."""
import numpy as np
import pytest

from rsruckig_b200 import _cabi as cabi
from rsruckig_b200 import workloads


@pytest.mark.parametrize("flavour", ["glibc", "portable"])
@pytest.mark.parametrize("gen,dof", [(workloads.bench_inputs, 7), (workloads.edge_case_inputs, 4)])
def test_working_trajectories_reach_the_target_inside_the_limits(flavour, gen, dof):
    from oracle.oracle import OracleBatch
    n = 20_000
    f = gen(n, dof, seed=77)
    ob = OracleBatch(n, dof, flavour)
    ob.calculate(f, delta_time=0.005, threads=4)
    status = ob.read(cabi.RK_FIELD_STATUS)
    ok = status == 0
    assert ok.mean() > 0.8
    assert set(np.unique(status).tolist()) <= {0, -100, -101, -104, -110, -111}
    en = f["enabled"].astype(bool) if "enabled" in f else np.ones((dof, n), dtype=bool)
    sel = en & ok[None, :]
    P, V, A = (ob.read(x) for x in (cabi.RK_FIELD_P, cabi.RK_FIELD_V, cabi.RK_FIELD_A))
    # final state (profile.rs:472-474)
    assert np.abs(P[7] - f["target_position"])[sel].max() <= 1e-8
    assert np.abs(V[7] - f["target_velocity"])[sel].max() <= 1e-8
    assert np.abs(A[7] - f["target_acceleration"])[sel].max() <= 1e-10
    # every enabled DoF ends with the trajectory (Time synchronisation), and not before its own minimum
    duration = ob.read(cabi.RK_FIELD_DURATION)
    end = ob.read(cabi.RK_FIELD_T_SUM)[6] + ob.read(cabi.RK_FIELD_BRAKE_DURATION)[0]
    assert np.abs(end - duration[None, :])[sel].max() <= 1e-9 * duration[ok].max()
    imd = ob.read(cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS)
    assert np.all((duration[None, :] >= imd - 1e-12)[sel])
    # kinematic limits at the phase boundaries after the brake pre-trajectory (profile.rs:475-481)
    v_max = f["max_velocity"]
    v_min = f.get("min_velocity", -v_max)
    a_max = f["max_acceleration"]
    a_min = f.get("min_acceleration", -a_max)
    for k in (3, 4, 5, 6):
        assert np.all((V[k] <= v_max + 1e-9)[sel]) and np.all((V[k] >= v_min - 1e-9)[sel]), f"velocity at boundary {k}"
    for k in (1, 3, 5):
        assert np.all((A[k] <= a_max + 1e-9)[sel]) and np.all((A[k] >= a_min - 1e-9)[sel]), f"acceleration at boundary {k}"
    # sampled states stay inside the limits too
    K = 64
    pos, vel, acc, _, _ = ob.at_time(time_start=0.0, delta_time=float(np.median(duration[ok])) / K, steps=2 * K, threads=4)
    brake = ob.read(cabi.RK_FIELD_BRAKE_DURATION)[0]
    t = (np.arange(1, 2 * K + 1) * (float(np.median(duration[ok])) / K))[:, None, None]
    # ... for the problems whose CURRENT state is inside the limits as well (validate_input with check_current = true): a start
    # beyond a limit is only brought back as fast as the other limits allow
    inside = ob.validate(f, check_current=True, check_target=True).astype(bool)
    after_brake = (t >= brake[None]) & (t <= duration[None, None, :]) & (sel & inside[None, :])[None]  # past the end at_time extrapolates
    assert after_brake.sum() > 500_000
    assert np.all((vel <= v_max[None] + 1e-7)[after_brake]) and np.all((vel >= v_min[None] - 1e-7)[after_brake])
    assert np.all((acc <= a_max[None] + 1e-7)[after_brake]) and np.all((acc >= a_min[None] - 1e-7)[after_brake])


@pytest.mark.parametrize("flavour", ["glibc", "portable"])
def test_replanning_from_a_state_on_the_trajectory_never_finds_a_shorter_one(flavour):
    """SURVEY.md 8c safeguard (2), the random-input form of `check_full_duration` (tests_known.rs:15-27): take the state a
    trajectory of duration T passes through at time t and plan again from there. The rest of the old trajectory is a valid
    answer, and the old one was time-optimal, so the new duration can never be shorter than T - t; it is exactly T - t
    unless T - t now sits on the edge of a blocked interval (then the next feasible duration is taken: longer). A restatement
    that returned non-optimal profiles, or profiles that do not pass through the states it reports, would break this."""
    from oracle.oracle import OracleBatch
    n, dof, K = 6000, 3, 128
    f = workloads.bench_inputs(n, dof, seed=5)
    ob = OracleBatch(n, dof, flavour)
    ob.calculate(f, threads=4)
    ok = ob.read(cabi.RK_FIELD_STATUS) == 0
    T = ob.read(cabi.RK_FIELD_DURATION)
    dt = float(np.median(T[ok])) / 48
    pos, vel, acc, _, _ = ob.at_time(time_start=0.0, delta_time=dt, steps=K, threads=4)
    tk = np.cumsum(np.full(K, dt))  # repeated addition, as at_time accumulates its sample times
    k = np.searchsorted(tk, 0.6 * T, side="right") - 1  # the last sample not later than 0.6 T
    use = ok & (k >= 0)
    idx = np.nonzero(use)[0]
    assert idx.size > 0.9 * n
    g = {key: v.copy() for key, v in f.items()}
    g["current_position"][:, idx] = pos[k[idx], :, idx].T
    g["current_velocity"][:, idx] = vel[k[idx], :, idx].T
    g["current_acceleration"][:, idx] = acc[k[idx], :, idx].T
    ob2 = OracleBatch(n, dof, flavour)
    ob2.calculate(g, threads=4)
    st2, T2 = ob2.read(cabi.RK_FIELD_STATUS), ob2.read(cabi.RK_FIELD_DURATION)
    # a sampled state can sit one ulp beyond a limit it was riding on; Step 1 then finds no profile for a handful of them
    assert (st2[use] != 0).mean() < 0.005
    both = use & (st2 == 0)
    d = (T2 - (T - tk[np.clip(k, 0, K - 1)]))[both]
    assert d.min() > -1e-9, d.min()
    assert (np.abs(d) < 1e-9).mean() > 0.8
