"""Replay of the reference's known-answer test (test_suite/tests/tests_known.rs, 44 duration checks, abs <= 1e-4).

The op list in tests/golden/known_trajectories.json was extracted by tests/golden/make_known.py. The same replay runs
against the oracle (both math flavours, CPU) — which is how the oracle is pinned — and against the CUDA path (-m gpu).
"""
import json
import os

import numpy as np
import pytest

from rsruckig_b200 import (ControlInterface, DurationDiscretization, IgnoreErrorHandler, InputParameter, OutputParameter,
                           RuckigResult, Synchronization, ThrowErrorHandler, Trajectory)

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_trajectories.json")))
HANDLERS = {"IgnoreErrorHandler": IgnoreErrorHandler, "ThrowErrorHandler": ThrowErrorHandler}
ENUMS = {"ControlInterface": ControlInterface, "DurationDiscretization": DurationDiscretization, "Synchronization": Synchronization}


def check_duration(otg, input, duration):  # tests_known.rs:5-13
    output = OutputParameter(otg.degrees_of_freedom)
    try:
        otg.update(input, output)
    except Exception:
        pass  # `let _result = otg.update(..)`: an Err is ignored
    assert output.trajectory.get_duration() == pytest.approx(duration, abs=1e-4)


def check_full_duration(otg, input, duration):  # tests_known.rs:15-27
    output = OutputParameter(otg.degrees_of_freedom)
    steps = 0
    while otg.update(input, output) == RuckigResult.Working:
        input.current_position = output.new_position.copy()
        input.current_velocity = output.new_velocity.copy()
        input.current_acceleration = output.new_acceleration.copy()
        steps += 1
        assert steps < 100000
    assert output.trajectory.get_duration() == pytest.approx(duration, abs=1e-4)


def test_known_trajectory(backend):
    env = {}
    last_dof = None
    half = None
    n_checks = 0
    for op in GOLD["ops"]:
        kind = op["op"]
        if kind == "new_otg":
            env[op["var"]] = backend.ruckig(op["dof"], op["delta_time"], HANDLERS[op["handler"]])
            last_dof = op["dof"]
        elif kind == "new_input":
            env[op["var"]] = InputParameter(last_dof)
        elif kind == "set":
            v = op["value"]
            if isinstance(v, list):
                v = np.array(v, dtype=np.float64)
            setattr(env[op["var"]], op["field"], v)
        elif kind == "set_enum":
            name = op["value"] if op["value"] != "None" else "None_"
            setattr(env[op["var"]], op["field"], getattr(ENUMS[op["enum"]], name))
        elif kind == "check_duration":
            check_duration(env[op["otg"]], env[op["input"]], op["duration"])
            n_checks += 1
        elif kind == "check_full_duration":
            check_full_duration(env[op["otg"]], env[op["input"]], op["duration"])
            n_checks += 1
        elif kind == "new_trajectory":
            env[op["var"]] = Trajectory(op["dof"])
        elif kind == "calculate":
            env[op["otg"]].calculate(env[op["input"]], env[op["traj"]])
        elif kind == "at_half_duration":
            traj = env[op["traj"]]
            half = traj.at_time(traj.get_duration() / 2.0)
        elif kind == "assert_new_position":
            assert half[0][op["index"]] == pytest.approx(op["value"], abs=op["abs"])
        else:
            raise AssertionError(kind)
    assert n_checks == GOLD["checks"] == 44
