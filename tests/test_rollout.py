"""Closed-loop control cycles on the device (rk_update_batch = the batched Ruckig::update, ruckig.rs:266-321) against the
oracle's stateful update, one OracleRuckig per environment (-m gpu)."""
import numpy as np
import pytest

from rsruckig_b200 import _cabi as cabi
from rsruckig_b200 import workloads

pytestmark = pytest.mark.gpu


def _oracle_envs(f, n, dof, dt):
    from oracle.oracle import OracleRuckig
    from rsruckig_b200 import InputParameter, OutputParameter
    envs = []
    for i in range(n):
        otg = OracleRuckig(dof, dt)
        inp, out = InputParameter(dof), OutputParameter(dof)
        for k in workloads.FIELDS:
            getattr(inp, k)[:] = f[k][:, i]
        envs.append((otg, inp, out))
    return envs


@pytest.mark.parametrize("reset_time", [False, True])
def test_update_cycles_match_the_oracle(reset_time):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsruckig_b200 import BatchRollout, RuckigResult
    n, dof, dt, cycles = 192, 3, 0.01, 60
    f = workloads.bench_inputs(n, dof, seed=31)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    ro = BatchRollout(n, dof, dt)
    envs = _oracle_envs(f, n, dof, dt)
    new_p, new_v, new_a, new_j = (torch.zeros((dof, n), dtype=torch.float64, device=dev) for _ in range(4))
    result = torch.zeros(n, dtype=torch.int32, device=dev)
    new_calc = torch.zeros(n, dtype=torch.uint8, device=dev)
    time = torch.zeros(n, dtype=torch.float64, device=dev)
    rng = np.random.default_rng(5)
    n_recalc = 0
    for c in range(cycles):
        if c in (7, 20, 21, 45):  # new targets for a random third of the environments, mid-motion
            pick = np.nonzero(rng.random(n) < 0.33)[0]
            tp = rng.normal(0.0, 4.0, (dof, pick.size))
            d["target_position"][:, torch.from_numpy(pick).to(dev)] = torch.from_numpy(tp).to(dev)
            for jj, i in enumerate(pick):
                envs[i][1].target_position[:] = tp[:, jj]
                if reset_time:
                    # upstream Ruckig resets output.time on a new calculation, the reference does not (Q1); the option is
                    # checked against a fresh calculator + fresh OutputParameter continuing from the current state
                    from oracle.oracle import OracleRuckig
                    from rsruckig_b200 import OutputParameter
                    envs[i] = (OracleRuckig(dof, dt), envs[i][1], OutputParameter(dof))
        m = ro.update(d, new_position=new_p, new_velocity=new_v, new_acceleration=new_a, new_jerk=new_j, result=result,
                      new_calculation=new_calc, time=time, pass_to_input=True, reset_time_on_new_calculation=reset_time)
        ro.sync()
        gp, gv, ga, gj = (t.cpu().numpy() for t in (new_p, new_v, new_a, new_j))
        gres, gnew, gtime = result.cpu().numpy(), new_calc.cpu().numpy(), time.cpu().numpy()
        exp_new = 0
        for i, (otg, inp, out) in enumerate(envs):
            res = otg.update(inp, out)
            exp_new += int(out.new_calculation)
            assert int(gres[i]) == int(res), (c, i)
            assert bool(gnew[i]) == bool(out.new_calculation), (c, i)
            assert gtime[i] == out.time, (c, i)
            assert np.array_equal(gp[:, i], np.asarray(out.new_position)), (c, i)
            assert np.array_equal(gv[:, i], np.asarray(out.new_velocity)), (c, i)
            assert np.array_equal(ga[:, i], np.asarray(out.new_acceleration)), (c, i)
            assert np.array_equal(gj[:, i], np.asarray(out.new_jerk)), (c, i)
            out.pass_to_input(inp)
        assert m == exp_new, (c, m, exp_new)
        n_recalc += m
        # the caller's current_* arrays now hold the new state (pass_to_input)
        assert torch.equal(d["current_position"], new_p) and torch.equal(d["current_velocity"], new_v)
    assert n_recalc > n, "no mid-motion recalculation happened"
    assert (gres == int(RuckigResult.Working)).any() or (gres == int(RuckigResult.Finished)).any()


@pytest.mark.parametrize("throw", [True, False])
def test_update_with_invalid_inputs_follows_the_error_handler(throw):
    """ruckig.rs:285: under ThrowErrorHandler a validation error is returned before the output is touched (and the environment
    is solved again next cycle); IgnoreErrorHandler skips validation and carries on with whatever the calculation yields."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.oracle import OracleRuckig
    from rsruckig_b200 import (BatchRollout, IgnoreErrorHandler, InputParameter, OutputParameter, ThrowErrorHandler, ValidationError)
    handler = ThrowErrorHandler if throw else IgnoreErrorHandler
    n, dof, dt, cycles = 64, 2, 0.01, 8
    f = workloads.bench_inputs(n, dof, seed=33)
    bad = np.arange(n) % 4 == 1
    f["target_velocity"][0, bad] = f["max_velocity"][0, bad] * 1.5   # target velocity beyond the limit: fails validate(false, true)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    ro = BatchRollout(n, dof, dt, handler)
    envs = []
    for i in range(n):
        inp, out = InputParameter(dof), OutputParameter(dof)
        for k in workloads.FIELDS:
            getattr(inp, k)[:] = f[k][:, i]
        envs.append((OracleRuckig(dof, dt, handler), inp, out))
    new_p = torch.zeros((dof, n), dtype=torch.float64, device=dev)
    result = torch.zeros(n, dtype=torch.int32, device=dev)
    time = torch.zeros(n, dtype=torch.float64, device=dev)
    for c in range(cycles):
        m = ro.update(d, new_position=new_p, result=result, time=time, pass_to_input=True)
        ro.sync()
        gp, gres, gtime = new_p.cpu().numpy(), result.cpu().numpy(), time.cpu().numpy()
        status = ro.read(cabi.RK_FIELD_STATUS)
        state = [d[k].cpu().numpy() for k in ("current_position", "current_velocity", "current_acceleration")]
        for i, (otg, inp, out) in enumerate(envs):
            try:
                res = int(otg.update(inp, out))
            except ValidationError:
                assert throw and bad[i]
                assert gres[i] == -100 and gtime[i] == 0.0, (c, i, gres[i], gtime[i])
                continue
            assert int(gres[i]) == res, (c, i)
            assert gtime[i] == out.time, (c, i)
            if status[i] == 0:
                assert np.array_equal(gp[:, i], np.asarray(out.new_position)), (c, i)
                out.pass_to_input(inp)
            else:
                # Ok(Error*) of the calculation is dropped (Q1) and the reference samples whatever the aborted calculation left
                # in the trajectory; the batched path specifies only status / duration for such a trajectory and
                # holds the current state. Keep both sides on the same input for the next cycle.
                assert not throw and bad[i]
                for k, name in enumerate(("current_position", "current_velocity", "current_acceleration")):
                    getattr(inp, name)[:] = state[k][:, i]
        if throw:
            assert m == (n if c == 0 else int(bad.sum())), (c, m)  # the failing environments are tried again every cycle
