"""Ports of the reference's integration tests (test_suite/tests/tests.rs), replayed against the oracle (CPU, both math
flavours — this pins the oracle) and against the CUDA path (-m gpu). Tolerances are the reference's (abs <= 1e-4).

The `get_position_extrema` / `get_first_time_at_position` assertions of test_secondary (tests.rs:190-237) live in
tests/test_queries.py (batched helpers).
"""
import numpy as np
import pytest

from rsruckig_b200 import (CalculatorError, ControlInterface, DurationDiscretization, InputParameter, OutputParameter,
                           RuckigError, Synchronization, ThrowErrorHandler, Trajectory, ValidationError)

A = lambda *x: np.array(x, dtype=np.float64)
TOL = 1e-4


def close(a, b, eps=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.all(np.abs(a - b) <= eps), (a, b)
    return True


def t_of(traj, dof):
    return traj.get_profiles()[0][dof].t


def test_at_time(backend):  # tests.rs:33-77
    otg = backend.ruckig(1, 0.005, ThrowErrorHandler)
    input = InputParameter(1)
    input.current_position = A(0.0)
    input.target_position = A(1.0)
    input.max_velocity = A(1.0)
    input.max_acceleration = A(1.0)
    input.max_jerk = A(1.0)

    traj = Trajectory(1)
    otg.calculate(input, traj)
    output = OutputParameter(1)
    otg.update(input, output)
    assert traj.get_duration() == pytest.approx(3.1748, abs=TOL)

    p, v, a, _, _ = traj.at_time(0.0)
    assert p[0] == pytest.approx(input.current_position[0], abs=TOL)
    p, v, a, _, _ = traj.at_time(3.1748 / 2.0)
    assert p[0] == pytest.approx(0.5, abs=TOL)


def test_secondary(backend):  # tests.rs:81-303
    otg = backend.ruckig(3, 0.005, ThrowErrorHandler)
    input = InputParameter(3)
    output = OutputParameter(3)
    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.target_velocity = A(0.0, 0.3, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)

    traj = Trajectory(3)
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(4.0, abs=TOL)

    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(4.0, abs=TOL)

    p, v, a, _, _ = output.trajectory.at_time(0.0)
    assert np.array_equal(p, input.current_position)
    assert np.array_equal(v, input.current_velocity)
    assert np.array_equal(a, input.current_acceleration)

    p, v, a, _, _ = output.trajectory.at_time(output.trajectory.get_duration())
    close(p, input.target_position)
    close(v, input.target_velocity)
    close(a, input.target_acceleration)

    p, v, a, j, section = output.trajectory.at_time(2.0)
    close(p, [0.5, -2.6871268303003437, 1.0])
    assert np.array_equal(j, A(0.0, 0.0, -1.0))
    assert section == 0

    p, v, a, j, section = output.trajectory.at_time(5.0)
    assert np.array_equal(j, A(0.0, 0.0, 0.0))
    assert section == 1

    imd = output.trajectory.get_independent_min_durations()
    assert imd[0] == pytest.approx(3.1748021039, abs=TOL)
    assert imd[1] == pytest.approx(3.6860977315, abs=TOL)
    assert imd[2] == pytest.approx(output.trajectory.get_duration(), abs=TOL)

    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.target_velocity = A(2.0, 0.3, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)

    with pytest.raises(ValidationError, match="exceeds its maximum velocity limit"):
        otg.update(input, output)
    assert not output.new_calculation

    input.target_velocity = A(0.2, -0.3, 0.8)
    otg.update(input, output)
    assert output.new_calculation

    input.minimum_duration = 12.0
    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(12.0, abs=TOL)

    input.current_position = A(1300.0, 0.0, 0.02)
    input.current_velocity = A(1200.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(1400.0, 0.0, 0.02)
    input.target_velocity = A(0.0, 0.0, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(800.0, 1.0, 1.0)
    input.max_acceleration = A(40000.0, 1.0, 1.0)
    input.max_jerk = A(200000.0, 1.0, 1.0)
    input.minimum_duration = None

    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(0.167347, abs=TOL)
    imd = output.trajectory.get_independent_min_durations()
    assert imd[0] == pytest.approx(output.trajectory.get_duration(), abs=TOL)
    assert imd[1] == pytest.approx(0.0, abs=TOL)
    assert imd[2] == pytest.approx(0.0, abs=TOL)


def test_enabled(backend):  # tests.rs:306-404
    otg = backend.ruckig(3, 0.005, ThrowErrorHandler)
    input = InputParameter(3)
    output = OutputParameter(3)
    input.enabled = np.array([True, False, False])
    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.0, 0.1, 0.0)
    input.current_acceleration = A(0.0, 0.0, -0.2)
    input.target_position = A(1.0, -3.0, 2.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)

    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(3.1748021039, abs=TOL)

    p, v, a, _, _ = output.trajectory.at_time(0.0)
    close(p, input.current_position)
    close(v, input.current_velocity)
    close(a, input.current_acceleration)

    p, v, a, _, _ = output.trajectory.at_time(output.trajectory.get_duration())
    close(p, [input.target_position[0], -1.6825197896, -1.0079368399])

    # Make sure that disabled DoFs overwrite prior blocks
    input.enabled = np.array([True, True, True])
    input.current_position = A(0.0, 0.0, 0.0)
    input.target_position = A(100.0, -3000.0, 2000.0)
    input.target_velocity = A(1.0, 1.0, 1.0)
    try:
        otg.update(input, output)
    except RuckigError:
        pass

    input.enabled = np.array([False, False, True])
    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.0, 0.2, 0.0)
    input.current_acceleration = A(0.0, 0.2, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.target_velocity = A(0.0, 0.0, 0.2)
    input.target_acceleration = A(0.0, 0.0, -0.1)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)

    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(3.6578610221, abs=TOL)


def test_phase_synchronization(backend):  # tests.rs:407-667
    otg = backend.ruckig(3, 0.005, ThrowErrorHandler)
    input = InputParameter(3)
    output = OutputParameter(3)
    traj = Trajectory(3)

    input.current_position = A(0.0, -2.0, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)
    input.synchronization = Synchronization.Phase

    otg.calculate(input, traj)
    close(t_of(traj, 0), t_of(traj, 1))
    close(t_of(traj, 0), t_of(traj, 2))

    otg.update(input, output)
    p, _, _, _, _ = output.trajectory.at_time(1.0)
    assert output.trajectory.get_duration() == pytest.approx(4.0, abs=TOL)
    close(p, [0.0833333333, -2.0833333333, 0.1666666667])
    close(t_of(output.trajectory, 0), t_of(output.trajectory, 1))
    close(t_of(output.trajectory, 0), t_of(output.trajectory, 2))

    input.current_position = A(0.0, -2.0, 0.0)
    input.target_position = A(10.0, -3.0, 2.0)
    input.max_velocity = A(10.0, 2.0, 1.0)
    input.max_acceleration = A(10.0, 2.0, 1.0)
    input.max_jerk = A(10.0, 2.0, 1.0)

    otg.update(input, output)
    p, _, _, _, _ = output.trajectory.at_time(1.0)
    assert output.trajectory.get_duration() == pytest.approx(4.0, abs=TOL)
    close(p, [0.8333333333, -2.0833333333, 0.1666666667])
    close(t_of(output.trajectory, 0), t_of(output.trajectory, 1))
    close(t_of(output.trajectory, 0), t_of(output.trajectory, 2))

    # Test equal start and target state
    input.current_position = A(1.0, -2.0, 3.0)
    input.target_position = A(1.0, -2.0, 3.0)

    otg.update(input, output)
    p, _, _, _, _ = output.trajectory.at_time(0.0)
    assert output.trajectory.get_duration() == pytest.approx(0.0, abs=TOL)
    close(p, [1.0, -2.0, 3.0])
    close(t_of(output.trajectory, 0), t_of(output.trajectory, 1))
    close(t_of(output.trajectory, 0), t_of(output.trajectory, 2))

    input.current_position = A(0.0, 0.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(0.0, 0.0, 0.0)
    input.target_velocity = A(0.2, 0.3, 0.4)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)

    otg.calculate(input, traj)
    close(t_of(traj, 0), t_of(traj, 1))
    close(t_of(traj, 0), t_of(traj, 2))

    input.current_position = A(0.0, 0.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(0.0, 0.0, 0.01)
    input.target_velocity = A(0.2, 0.3, 0.4)
    input.target_acceleration = A(0.0, 0.0, 0.0)

    otg.calculate(input, traj)
    assert not np.array_equal(t_of(traj, 0), t_of(traj, 1))
    assert not np.array_equal(t_of(traj, 0), t_of(traj, 2))

    input.current_position = A(0.0, 0.0, 0.0)
    input.current_velocity = A(0.4, 0.15, 0.2)
    input.current_acceleration = A(0.8, 0.3, 0.4)
    input.target_position = A(0.0, 0.0, 0.0)
    input.target_velocity = A(0.0, 0.0, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)

    otg.calculate(input, traj)
    close(t_of(traj, 0), t_of(traj, 1))
    close(t_of(traj, 0), t_of(traj, 2))

    input.max_velocity = A(1.0, 0.2, 1.0)

    otg.calculate(input, traj)
    assert not np.array_equal(t_of(traj, 0), t_of(traj, 1))
    assert not np.array_equal(t_of(traj, 0), t_of(traj, 2))

    input.current_position = A(0.0, 0.02, 1.0)
    input.current_velocity = A(-0.2, 0.15, 0.2)
    input.current_acceleration = A(-0.4, 0.3, 0.4)
    input.target_position = A(0.03, 0.0, 0.0)
    input.target_velocity = A(-0.02, 0.015, 0.02)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)
    input.control_interface = ControlInterface.Velocity

    otg.calculate(input, traj)
    close(t_of(traj, 0), t_of(traj, 1))
    close(t_of(traj, 0), t_of(traj, 2))

    input.max_jerk = A(1.0, 0.1, 1.0)

    otg.calculate(input, traj)
    close(t_of(traj, 0), t_of(traj, 1))
    close(t_of(traj, 0), t_of(traj, 2))

    input.target_acceleration = A(0.01, 0.0, 0.0)

    otg.calculate(input, traj)
    assert not np.array_equal(t_of(traj, 0), t_of(traj, 1))
    assert not np.array_equal(t_of(traj, 0), t_of(traj, 2))

    input.current_position = A(0.0, 0.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(0.0, 0.0, 0.0)
    input.target_velocity = A(0.0, 0.0, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)

    otg.calculate(input, traj)


def test_discretion(backend):  # tests.rs:670-707
    otg = backend.ruckig(3, 0.01, ThrowErrorHandler)
    input = InputParameter(3)
    output = OutputParameter(3)
    traj = Trajectory(3)
    input.current_position = A(0.0, 0.0, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.target_velocity = A(0.2, 0.2, 0.2)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(2.0, 2.0, 2.0)
    input.max_jerk = A(1.8, 2.4, 2.0)
    input.duration_discretization = DurationDiscretization.Discrete

    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(4.5, abs=TOL)

    otg.update(input, output)
    p, _, _, _, _ = output.trajectory.at_time(4.5)
    close(p, [1.0, -3.0, 2.0])


def test_per_dof_setting(backend):  # tests.rs:710-965
    otg = backend.ruckig(3, 0.005, ThrowErrorHandler)
    input = InputParameter(3)
    traj = Trajectory(3)

    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.target_velocity = A(0.0, 0.3, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)

    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(4.0, abs=TOL)
    p, _, _, _, _ = traj.at_time(2.0)
    close(p, [0.5, -2.6871268303, 1.0])

    input.control_interface = ControlInterface.Velocity
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(1.095445115, abs=TOL)
    p, _, _, _, _ = traj.at_time(1.0)
    close(p, [0.0, -1.8641718534, 0.0])

    input.per_dof_control_interface = [ControlInterface.Position, ControlInterface.Velocity, ControlInterface.Position]
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(4.0, abs=TOL)

    input.per_dof_synchronization = [Synchronization.Time, Synchronization.None_, Synchronization.Time]
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(4.0, abs=TOL)
    p, _, _, _, _ = traj.at_time(2.0)
    close(p, [0.5, -1.5643167673, 1.0])

    input.control_interface = ControlInterface.Position
    input.per_dof_control_interface = None
    input.per_dof_synchronization = [Synchronization.None_, Synchronization.Time, Synchronization.Time]
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(4.0, abs=TOL)
    p, _, _, _, _ = traj.at_time(2.0)
    close(p, [0.7482143874, -2.6871268303, 1.0])

    imd = traj.get_independent_min_durations()
    p, _, _, _, _ = traj.at_time(imd[0])
    assert p[0] == pytest.approx(input.target_position[0], abs=TOL)
    p, _, _, _, _ = traj.at_time(imd[1])
    assert p[1] == pytest.approx(-3.0890156397, abs=TOL)
    p, _, _, _, _ = traj.at_time(imd[2])
    assert p[2] == pytest.approx(input.target_position[2], abs=TOL)

    input.current_position = A(0.0, 0.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(35.0, 35.0, 35.0)
    input.target_velocity = A(125.0, 125.0, 100.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(125.0, 125.0, 100.0)
    input.max_acceleration = A(2000.0, 2000.0, 2000.0)
    input.max_jerk = A(20000.0, 20000.0, 20000.0)
    input.per_dof_synchronization = [Synchronization.Time, Synchronization.Time, Synchronization.None_]
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(0.4207106781, abs=TOL)

    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.0, 0.2, 0.0)
    input.current_acceleration = A(0.0, 0.2, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.target_velocity = A(0.0, 0.0, 0.2)
    input.target_acceleration = A(0.0, 0.0, -0.1)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)
    input.per_dof_synchronization = [Synchronization.None_, Synchronization.None_, Synchronization.Time]
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(3.7885667284, abs=TOL)

    input.per_dof_synchronization = [Synchronization.None_, Synchronization.Time, Synchronization.None_]
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(3.7885667284, abs=TOL)

    input.enabled = np.array([True, False, True])
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(3.6578610221, abs=TOL)

    input.current_position = A(0.0, 0.0, 0.0)
    input.current_velocity = A(0.2, 0.0, -0.1)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(1.0, -0.2, -0.5)
    input.target_velocity = A(0.0, 0.0, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)
    input.per_dof_synchronization = [Synchronization.Phase, Synchronization.None_, Synchronization.Phase]
    input.enabled = np.array([True, True, True])
    otg.calculate(input, traj)
    assert traj.get_duration() == pytest.approx(2.848387279, abs=TOL)
    assert not np.array_equal(t_of(traj, 0), t_of(traj, 1))
    close(t_of(traj, 0), t_of(traj, 2))


def test_dynamic_dofs(backend):  # tests.rs:968-1017 (heap variant: same code path here, dof is always a run-time value)
    otg = backend.ruckig(3, 0.005, ThrowErrorHandler)
    input = InputParameter(3)
    output = OutputParameter(3)
    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.0, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(1.0, -3.0, 2.0)
    input.target_velocity = A(0.0, 0.3, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(1.0, 1.0, 1.0)
    input.max_jerk = A(1.0, 1.0, 1.0)

    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(4.0, abs=TOL)
    p, v, a, _, _ = output.trajectory.at_time(0.0)
    close(p, input.current_position)
    close(v, input.current_velocity)
    close(a, input.current_acceleration)


def test_zero_limits(backend):  # tests.rs:1020-1125
    otg = backend.ruckig(3, 0.005, ThrowErrorHandler)
    input = InputParameter(3)
    output = OutputParameter(3)
    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(0.2, 0.0, 0.0)
    input.current_acceleration = A(0.0, 0.0, 0.0)
    input.target_position = A(1.0, -3.0, 0.0)
    input.target_velocity = A(0.2, 0.0, 0.0)
    input.target_acceleration = A(0.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 1.0, 1.0)
    input.max_acceleration = A(0.0, 1.0, 0.0)
    input.max_jerk = A(0.0, 1.0, 0.0)

    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(5.0, abs=TOL)

    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(-0.2, 0.0, 0.0)
    input.current_acceleration = A(1.0, 0.0, 0.0)
    input.target_position = A(0.4, -3.0, 0.0)
    input.target_velocity = A(0.8, 0.0, 0.0)
    input.target_acceleration = A(1.0, 0.0, 0.0)
    input.max_velocity = A(1.0, 200.0, 0.0)
    input.max_acceleration = A(1.0, 200.0, 0.0)
    input.max_jerk = A(0.0, 200.0, 0.0)

    with pytest.raises(CalculatorError, match="zero limits conflict in step 1"):
        otg.update(input, output)

    input.target_position = A(0.3, -3.0, 0.0)
    input.max_velocity = A(1.0, 2.0, 1.0)
    input.max_acceleration = A(1.0, 2.0, 0.0)
    input.max_jerk = A(0.0, 2.0, 0.0)

    with pytest.raises(CalculatorError, match="zero limits conflict with other"):
        otg.update(input, output)

    input.control_interface = ControlInterface.Velocity
    input.current_position = A(0.0, -2.0, 0.0)
    input.current_velocity = A(-0.2, 0.0, 0.0)
    input.current_acceleration = A(1.0, 0.0, 0.2)
    input.target_position = A(0.4, -3.0, 0.0)
    input.target_velocity = A(0.9, 0.5, 0.4)
    input.target_acceleration = A(1.0, 0.0, 0.2)
    input.max_velocity = A(1.0, 2.0, 1.0)
    input.max_acceleration = A(1.0, 2.0, 6.0)
    input.max_jerk = A(0.0, 2.0, 0.0)

    with pytest.raises(CalculatorError, match="zero limits conflict with other"):
        otg.update(input, output)

    input.max_jerk = A(1.0, 2.0, 0.0)
    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(2.0, abs=TOL)

    input.max_jerk = A(0.0, 2.0, 20.0)
    otg.update(input, output)
    assert output.trajectory.get_duration() == pytest.approx(1.1, abs=TOL)


def test_min_duration(backend):  # tests.rs:1128-1160
    otg = backend.ruckig(1, 0.01, ThrowErrorHandler)
    input = InputParameter(1)
    input.current_position[0] = 0.0
    input.current_velocity[0] = 0.0
    input.current_acceleration[0] = 0.0
    input.target_position[0] = 1.0
    input.target_velocity[0] = 1.0
    input.target_acceleration[0] = 1.0
    input.max_velocity[0] = 1.0
    input.max_acceleration[0] = 2.0
    input.max_jerk[0] = 3.0
    input.minimum_duration = None
    input.synchronization = Synchronization.None_

    trajectory = Trajectory(1)
    otg.calculate(input, trajectory)
    duration = trajectory.get_duration()

    trajectory_min_duration = Trajectory(1)
    input.minimum_duration = 5.0
    otg.calculate(input, trajectory_min_duration)
    new_duration = trajectory_min_duration.get_duration()
    assert new_duration > duration
    assert new_duration == pytest.approx(5.0, abs=TOL)


@pytest.mark.parametrize("start,target", [((0.0, 0.0), (1.0, 2.0)), ((1.0, 0.0), (0.0, 2.0))])
def test_signs_phase_sync(backend, start, target):  # tests.rs:1163-1193 (matched), :1196-1226 (mixed)
    otg = backend.ruckig(2, 0.01, ThrowErrorHandler)
    input = InputParameter(2)
    input.synchronization = Synchronization.Phase
    input.current_position = A(*start)
    input.target_position = A(*target)
    input.current_velocity = A(0.0, 0.0)
    input.target_velocity = A(0.0, 0.0)
    input.max_velocity = A(1.0, 1000.0)
    input.max_acceleration = A(10.0, 1000.0)  # max_jerk stays +inf: second-order path

    trajectory = Trajectory(2)
    otg.calculate(input, trajectory)
    assert np.array_equal(t_of(trajectory, 0), t_of(trajectory, 1))
