"""Trajectory query helpers: Trajectory::get_position_extrema and get_first_time_at_position
(trajectory.rs:211-246, profile.rs:754-895). The oracle is pinned on the reference's own assertions (test_suite/tests/tests.rs:
190-237, abs <= 1e-4); the CUDA path must then equal the oracle bit for bit on random batches (-m gpu)."""
import numpy as np
import pytest

from rsruckig_b200 import workloads

TOL = 1e-4


def secondary_inputs():
    """tests.rs:81-105: the 3-DoF problem of test_secondary."""
    col = lambda *x: np.array(x, dtype=np.float64).reshape(3, 1)
    return {"current_position": col(0.0, -2.0, 0.0), "current_velocity": col(0.0, 0.0, 0.0), "current_acceleration": col(0.0, 0.0, 0.0),
            "target_position": col(1.0, -3.0, 2.0), "target_velocity": col(0.0, 0.3, 0.0), "target_acceleration": col(0.0, 0.0, 0.0),
            "max_velocity": col(1.0, 1.0, 1.0), "max_acceleration": col(1.0, 1.0, 1.0), "max_jerk": col(1.0, 1.0, 1.0)}


def check_reference_assertions(extrema, first_time):
    e = extrema[:, :, 0]  # [4][dof] = min, max, t_min, t_max
    assert e[3, 0] == pytest.approx(4.0, abs=TOL) and e[1, 0] == pytest.approx(1.0, abs=TOL)            # tests.rs:191-194
    assert e[2, 0] == pytest.approx(0.0, abs=TOL) and e[0, 0] == pytest.approx(0.0, abs=TOL)
    assert e[3, 1] == pytest.approx(0.0, abs=TOL) and e[1, 1] == pytest.approx(-2.0, abs=TOL)           # :196-199
    assert e[2, 1] == pytest.approx(3.2254033308, abs=TOL) and e[0, 1] == pytest.approx(-3.1549193338, abs=TOL)
    assert e[3, 2] == pytest.approx(4.0, abs=TOL) and e[1, 2] == pytest.approx(2.0, abs=TOL)            # :201-204
    assert e[2, 2] == pytest.approx(0.0, abs=TOL) and e[0, 2] == pytest.approx(0.0, abs=TOL)
    q = lambda d, pos: float(first_time(d, pos))
    assert q(0, 0.0) == pytest.approx(0.0, abs=TOL)            # :206-208
    assert q(0, 0.5) == pytest.approx(2.0, abs=TOL)            # :210-212
    assert q(0, 1.0) == pytest.approx(4.0, abs=TOL)            # :214-216
    assert q(1, -3.0) == pytest.approx(2.6004877902, abs=TOL)  # :218-220
    assert q(1, -3.1) == pytest.approx(2.8644154489, abs=TOL)  # :222-224
    assert q(2, 0.05) == pytest.approx(0.6694329501, abs=TOL)  # :226-228
    assert np.isnan(q(0, -1.0))                                # :230-231
    assert np.isnan(q(1, -3.4))                                # :233-234


@pytest.mark.parametrize("flavour", ["glibc", "portable"])
def test_oracle_matches_the_reference_assertions(flavour):
    from oracle.oracle import OracleBatch
    f = secondary_inputs()
    ob = OracleBatch(1, 3, flavour)
    ob.calculate(f, delta_time=0.005)

    def first_time(d, pos):
        q = np.zeros((3, 1))
        q[d, 0] = pos
        return ob.first_time_at_position(q)[d, 0]
    check_reference_assertions(ob.position_extrema(), first_time)


@pytest.mark.gpu
def test_cuda_matches_the_reference_assertions():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsruckig_b200 import BatchRuckig, _cabi as cabi
    f = secondary_inputs()
    otg = BatchRuckig(3, 0.005)
    otg.calculate(f, n=1, memory_space=cabi.RK_MEM_HOST)

    def first_time(d, pos):
        q = np.zeros((3, 1))
        q[d, 0] = pos
        return otg.first_time_at_position(q)[d, 0]
    check_reference_assertions(otg.position_extrema(), first_time)


@pytest.mark.gpu
def test_cuda_equals_oracle_on_random_batches():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.oracle import OracleBatch
    from rsruckig_b200 import BatchRuckig, _cabi as cabi
    for dof, gen, seed in ((7, workloads.bench_inputs, 41), (4, workloads.edge_case_inputs, 42)):
        n = 50_000
        f = gen(n, dof, seed=seed)
        otg = BatchRuckig(dof, 0.005)
        otg.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST)
        ob = OracleBatch(n, dof, "portable")
        ob.calculate(f, delta_time=0.005, threads=8)
        ok = ob.read(cabi.RK_FIELD_STATUS) == 0
        eg, eo = otg.position_extrema(), ob.position_extrema()
        assert np.array_equal(eg[:, :, ok], eo[:, :, ok]), "position extrema"
        # query positions: somewhere along the way, exactly at a phase boundary, exactly at the target, and out of reach
        rng = np.random.default_rng(seed)
        lo, hi = eo[0], eo[1]
        for kind in range(4):
            if kind == 0:
                q = lo + (hi - lo) * rng.random((dof, n))
            elif kind == 1:
                q = ob.read(cabi.RK_FIELD_P)[3]
            elif kind == 2:
                q = f["target_position"].copy()
            else:
                q = hi + 1.0 + rng.random((dof, n))
            tg, to = otg.first_time_at_position(q), ob.first_time_at_position(q)
            same = (tg == to) | (np.isnan(tg) & np.isnan(to))
            assert np.all(same[:, ok]), f"first time at position, query kind {kind}: {np.argwhere(~same[:, ok])[:5].tolist()}"
