"""Host-side mirror of the reference's public API on top of the C ABI.

Names, argument meaning and error behaviour follow rsruckig v2.1.2:
`Ruckig::{new, reset, validate_input, calculate, update}` (ruckig.rs:110-321),
`InputParameter` (input_parameter.rs:147-243), `OutputParameter`
(output_parameter.rs:36-120), `Trajectory::{at_time, get_duration, ...}`
(trajectory.rs:158-209), `RuckigResult` (result.rs:34-61), the two error
handlers (error.rs:98-122) — plus the batched entry point `BatchRuckig`.

Everything is computed by the CUDA library; this module only moves buffers.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _cabi as cabi


class ControlInterface(enum.IntEnum):  # input_parameter.rs:35-45
    Position = 0
    Velocity = 1
    Acceleration = 2


class Synchronization(enum.IntEnum):  # input_parameter.rs:72-85
    Time = 0
    TimeIfNecessary = 1
    Phase = 2
    None_ = 3


class DurationDiscretization(enum.IntEnum):  # input_parameter.rs:110-117
    Continuous = 0
    Discrete = 1


class RuckigResult(enum.IntEnum):  # result.rs:34-61
    Working = 0
    Finished = 1
    Error = -1
    ErrorInvalidInput = -100
    ErrorTrajectoryDuration = -101
    ErrorPositionalLimits = -102
    ErrorZeroLimits = -104
    ErrorExecutionTimeCalculation = -110
    ErrorSynchronizationCalculation = -111


class ReachedLimits(enum.IntEnum):  # profile.rs:25-34
    Acc0Acc1Vel = 0
    Vel = 1
    Acc0 = 2
    Acc1 = 3
    Acc0Acc1 = 4
    Acc0Vel = 5
    Acc1Vel = 6
    None_ = 7


class Direction(enum.IntEnum):  # profile.rs:37-41
    UP = 0
    DOWN = 1


class ControlSigns(enum.IntEnum):  # profile.rs:44-49
    UDDU = 0
    UDUD = 1


class RuckigError(Exception):  # error.rs:15-34
    pass


class ValidationError(RuckigError):
    pass


class CalculatorError(RuckigError):
    pass


class ThrowErrorHandler:  # error.rs:101-111
    throws = True


class IgnoreErrorHandler:  # error.rs:113-122
    throws = False


_CALC_CODES = (RuckigResult.ErrorZeroLimits, RuckigResult.ErrorExecutionTimeCalculation, RuckigResult.ErrorSynchronizationCalculation)

# texts of InputParameter::validate (input_parameter.rs:251-420), keyed by the RK_INVALID_* check id the GPU reports
_VALIDATION_MESSAGES = {
    1: "Maximum jerk limit of DoF {dof} should be larger than or equal to zero.",
    2: "maximum acceleration limit of DoF {dof} should be larger than or equal to zero.",
    3: "minimum acceleration limit of DoF {dof} should be smaller than or equal to zero.",
    4: "current acceleration of DoF {dof} should be a valid number.",
    5: "target acceleration of DoF {dof} should be a valid number.",
    6: "current acceleration of DoF {dof} exceeds its maximum acceleration limit.",
    7: "current acceleration of DoF {dof} undercuts its minimum acceleration limit.",
    8: "target acceleration of DoF {dof} exceeds its maximum acceleration limit.",
    9: "target acceleration of DoF {dof} undercuts its minimum acceleration limit.",
    10: "current velocity of DoF {dof} should be a valid number.",
    11: "target velocity of DoF {dof} should be a valid number.",
    12: "current position of DoF {dof} should be a valid number.",
    13: "target position of DoF {dof} should be a valid number.",
    14: "maximum velocity limit of DoF {dof} should be larger than or equal to zero.",
    15: "minimum velocity limit of DoF {dof} should be smaller than or equal to zero.",
    16: "current velocity of DoF {dof} exceeds its maximum velocity limit.",
    17: "current velocity of DoF {dof} undercuts its minimum velocity limit.",
    18: "target velocity of DoF {dof} exceeds its maximum velocity limit.",
    19: "target velocity of DoF {dof} undercuts its minimum velocity limit.",
    20: "DoF {dof} will inevitably reach a velocity from the current kinematic state that will exceed its maximum velocity limit.",
    21: "DoF {dof} will inevitably reach a velocity from the current kinematic state that will undercut its minimum velocity limit.",
    22: "DoF {dof} will inevitably have reached a velocity from the target kinematic state that will exceed its maximum velocity limit.",
    23: "DoF {dof} will inevitably have reached a velocity from the target kinematic state that will undercut its minimum velocity limit.",
    24: "delta time (control rate) parameter should be larger than zero.",
}


def validation_message(reason: int) -> str:
    return _VALIDATION_MESSAGES.get(reason >> 8, "invalid input").format(dof=reason & 0xFF)


def calculator_message(code: int, detail: int) -> str:
    """Texts of calculator_target.rs:459-470, :509-517, :820-825 rebuilt from the status and the error-detail word."""
    stage, dof = detail >> 8, detail & 0xFF
    if stage == 1:
        return (f"zero limits conflict in step 1, dof: {dof}" if code == RuckigResult.ErrorZeroLimits else f"error in step 1, dof: {dof}")
    if stage == 2:
        return ("zero limits conflict with other degrees of freedom in time synchronization" if code == RuckigResult.ErrorZeroLimits
                else "error in time synchronization")
    return f"error in step 2, dof: {dof}"


class InputParameter:
    """input_parameter.rs:147-243 (same defaults: max_acceleration = max_jerk = +inf, all DoFs enabled)."""

    _ARRAYS = ("current_position", "current_velocity", "current_acceleration", "target_position", "target_velocity",
               "target_acceleration", "max_velocity", "max_acceleration", "max_jerk")

    def __init__(self, dofs: int):
        self.degrees_of_freedom = int(dofs)
        self.control_interface = ControlInterface.Position
        self.synchronization = Synchronization.Time
        self.duration_discretization = DurationDiscretization.Continuous
        z = lambda v=0.0: np.full(dofs, v, dtype=np.float64)
        self.current_position, self.current_velocity, self.current_acceleration = z(), z(), z()
        self.target_position, self.target_velocity, self.target_acceleration = z(), z(), z()
        self.max_velocity, self.max_acceleration, self.max_jerk = z(), z(np.inf), z(np.inf)
        self.min_velocity = None
        self.min_acceleration = None
        self.enabled = np.ones(dofs, dtype=bool)
        self.per_dof_control_interface = None
        self.per_dof_synchronization = None
        self.minimum_duration = None
        self.interrupt_calculation_duration = None

    def _key(self):
        opt = lambda a: None if a is None else tuple(np.asarray(a, dtype=np.float64).tolist())
        opti = lambda a: None if a is None else tuple(int(x) for x in a)
        return (tuple(tuple(np.asarray(getattr(self, f), dtype=np.float64).tolist()) for f in self._ARRAYS),
                tuple(bool(x) for x in self.enabled), self.minimum_duration, opt(self.min_velocity), opt(self.min_acceleration),
                int(self.control_interface), int(self.synchronization), int(self.duration_discretization),
                opti(self.per_dof_control_interface), opti(self.per_dof_synchronization))

    def __eq__(self, other):  # input_parameter.rs:190-211
        return isinstance(other, InputParameter) and self._key() == other._key()

    def copy(self) -> "InputParameter":
        o = InputParameter(self.degrees_of_freedom)
        for f in self._ARRAYS:
            setattr(o, f, np.array(getattr(self, f), dtype=np.float64))
        o.enabled = np.array(self.enabled, dtype=bool)
        for f in ("min_velocity", "min_acceleration"):
            v = getattr(self, f)
            setattr(o, f, None if v is None else np.array(v, dtype=np.float64))
        for f in ("per_dof_control_interface", "per_dof_synchronization"):
            v = getattr(self, f)
            setattr(o, f, None if v is None else list(v))
        o.control_interface, o.synchronization = self.control_interface, self.synchronization
        o.duration_discretization, o.minimum_duration = self.duration_discretization, self.minimum_duration
        return o

    # -- n = 1 view for the C ABI ------------------------------------------------------------
    def as_batch(self, delta_time: float, error_mode: int):
        """(desc, rk_batch_in, keepalive) describing this single problem as a batch of one."""
        D = self.degrees_of_freedom
        keep = []
        fields = {}
        for f in self._ARRAYS:
            a = np.ascontiguousarray(np.asarray(getattr(self, f), dtype=np.float64).reshape(D))
            keep.append(a)
            fields[f] = a
        for f in ("min_velocity", "min_acceleration"):
            v = getattr(self, f)
            if v is not None:
                a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(D))
                keep.append(a)
                fields[f] = a
        en = np.ascontiguousarray(np.asarray(self.enabled, dtype=np.uint8).reshape(D))
        keep.append(en)
        fields["enabled"] = en
        if self.minimum_duration is not None:
            md = np.array([float(self.minimum_duration)], dtype=np.float64)
            keep.append(md)
            fields["minimum_duration"] = md
        desc, k2 = cabi.make_desc(1, D, control_interface=self.control_interface, synchronization=self.synchronization,
                                  duration_discretization=self.duration_discretization, error_mode=error_mode,
                                  memory_space=cabi.RK_MEM_HOST, delta_time=delta_time,
                                  per_dof_control_interface=self.per_dof_control_interface,
                                  per_dof_synchronization=self.per_dof_synchronization)
        keep += k2
        return desc, cabi.make_in(fields), keep


class Profile:
    """Host copy of one DoF's Profile (profile.rs:76-118)."""

    def __init__(self):
        self.t = np.zeros(7); self.t_sum = np.zeros(7); self.j = np.zeros(7)
        self.a = np.zeros(8); self.v = np.zeros(8); self.p = np.zeros(8)
        self.brake_duration = 0.0
        self.brake_t = np.zeros(2)
        self.limits = ReachedLimits.None_
        self.direction = Direction.UP
        self.control_signs = ControlSigns.UDDU


class Trajectory:
    """trajectory.rs:14-21. The profiles live on the GPU (an rk_trajectories of one); fields are read on demand."""

    def __init__(self, dofs: int):
        self.degrees_of_freedom = int(dofs)
        self.duration = 0.0
        self.independent_min_durations = np.zeros(dofs)
        self._engine = None  # set by Ruckig.calculate
        self._dev = None

    def close(self):
        """Release the device-side trajectory (also done when the object is collected)."""
        dev, eng = self._dev, self._engine
        self._dev = None
        if dev is not None and eng is not None:
            try:
                eng.free_trajectories(dev)
            except Exception:  # interpreter shutdown: the library may be gone already
                pass

    def __del__(self):
        self.close()

    def get_duration(self) -> float:  # trajectory.rs:197-199
        return self.duration

    def get_independent_min_durations(self):  # trajectory.rs:207-209
        return self.independent_min_durations

    def get_profiles(self):  # trajectory.rs:192-194: one section of `dof` profiles
        assert self._dev is not None, "calculate() has not filled this trajectory"
        eng, D = self._engine, self.degrees_of_freedom
        rd = lambda f: eng.read(self._dev, f, 1, D)
        t, ts, j, a, v, p = (rd(f) for f in (cabi.RK_FIELD_T, cabi.RK_FIELD_T_SUM, cabi.RK_FIELD_J, cabi.RK_FIELD_A,
                                             cabi.RK_FIELD_V, cabi.RK_FIELD_P))
        bd, bt, codes = rd(cabi.RK_FIELD_BRAKE_DURATION), rd(cabi.RK_FIELD_BRAKE_T), rd(cabi.RK_FIELD_CODES)
        out = []
        for d in range(D):
            q = Profile()
            q.t, q.t_sum, q.j = t[:, d, 0].copy(), ts[:, d, 0].copy(), j[:, d, 0].copy()
            q.a, q.v, q.p = a[:, d, 0].copy(), v[:, d, 0].copy(), p[:, d, 0].copy()
            q.brake_duration, q.brake_t = float(bd[0, d, 0]), bt[:, d, 0].copy()
            c = int(codes[d, 0])
            q.limits, q.control_signs, q.direction = ReachedLimits(c & 0xFF), ControlSigns((c >> 8) & 0xFF), Direction((c >> 16) & 0xFF)
            out.append(q)
        return [out]

    def at_time(self, time: float):
        """Trajectory::at_time (trajectory.rs:158-189): returns (position, velocity, acceleration, jerk, section)."""
        assert self._dev is not None, "calculate() has not filled this trajectory"
        return self._engine.at_time(self._dev, self.degrees_of_freedom, time)


class OutputParameter:
    """output_parameter.rs:36-61."""

    def __init__(self, dofs: int):
        self.degrees_of_freedom = int(dofs)
        self.trajectory = Trajectory(dofs)
        self.new_position = np.zeros(dofs)
        self.new_velocity = np.zeros(dofs)
        self.new_acceleration = np.zeros(dofs)
        self.new_jerk = np.zeros(dofs)
        self.time = 0.0
        self.new_section = 0
        self.did_section_change = False
        self.new_calculation = False
        self.was_calculation_interrupted = False
        self.calculation_duration = 0.0

    def pass_to_input(self, input: InputParameter):  # output_parameter.rs:110-120
        input.current_position = self.new_position.copy()
        input.current_velocity = self.new_velocity.copy()
        input.current_acceleration = self.new_acceleration.copy()


class _Engine:
    """One rk_handle per process/device, shared by the n = 1 façade objects."""

    _instances: dict = {}

    @classmethod
    def get(cls, device: int = 0) -> "_Engine":
        if device not in cls._instances:
            cls._instances[device] = _Engine(device)
        return cls._instances[device]

    def __init__(self, device: int):
        self.lib = cabi.load()
        self.h = C.c_void_p()
        cabi.check(self.lib.rk_create(device, C.byref(self.h)))
        self.device = device

    def new_trajectories(self, n: int, dof: int, layout: int = cabi.RK_RESULT_FULL) -> C.c_void_p:
        t = C.c_void_p()
        cabi.check(self.lib.rk_trajectories_create_ex(self.h, n, dof, int(layout), C.byref(t)))
        return t

    def free_trajectories(self, t):
        if t is not None and self.lib is not None:
            self.lib.rk_trajectories_destroy(t)

    def read(self, traj, field: int, n: int, dof: int) -> np.ndarray:
        if field in (cabi.RK_FIELD_STATUS, cabi.RK_FIELD_LIMITING_DOF, cabi.RK_FIELD_ERROR_DETAIL):
            out = np.empty(n, dtype=np.int32)
        elif field == cabi.RK_FIELD_DURATION:
            out = np.empty(n, dtype=np.float64)
        elif field == cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS:
            out = np.empty((dof, n), dtype=np.float64)
        elif field == cabi.RK_FIELD_CODES:
            out = np.empty((dof, n), dtype=np.uint32)
        else:
            out = np.empty((cabi.FIELD_SLOTS[field], dof, n), dtype=np.float64)
        cabi.check(self.lib.rk_trajectories_read(self.h, traj, field, out.ctypes.data, cabi.RK_MEM_HOST))
        return out

    def at_time(self, traj, dof: int, time: float):
        # time_start + delta_time with delta_time = 0 would give K identical samples; one sample at `time`
        # is "time_start = time, then add 0.0": the kernel adds delta_time once before the first sample.
        sd = cabi.RkSampleDesc(float(time), 0.0, 1, cabi.RK_MEM_HOST)
        pos, vel, acc, jerk = (np.empty((1, dof, 1)) for _ in range(4))
        sec = np.empty((1, 1), dtype=np.int32)
        so = cabi.RkSampleOut(pos.ctypes.data, vel.ctypes.data, acc.ctypes.data, jerk.ctypes.data, sec.ctypes.data)
        cabi.check(self.lib.rk_at_time_batch(self.h, traj, C.byref(sd), C.byref(so)))
        return pos[0, :, 0].copy(), vel[0, :, 0].copy(), acc[0, :, 0].copy(), jerk[0, :, 0].copy(), int(sec[0, 0])


class Ruckig:
    """ruckig.rs:76-321 for a single problem; every calculation runs on the GPU as a batch of one."""

    def __init__(self, dofs: int, delta_time: float, error_handler=ThrowErrorHandler, device: int = 0):
        self.degrees_of_freedom = int(dofs)
        self.delta_time = float(delta_time)
        self.error_handler = error_handler
        self.current_input = InputParameter(dofs)
        self.current_input_initialized = False
        self._engine = _Engine.get(device)

    def reset(self):  # ruckig.rs:125-127
        self.current_input_initialized = False

    def validate_input(self, input: InputParameter, check_current_state_within_limits: bool,
                       check_target_state_within_limits: bool) -> bool:
        """ruckig.rs:150-171. Raises ValidationError under ThrowErrorHandler; returns validity otherwise."""
        desc, bin_, keep = input.as_batch(self.delta_time, cabi.RK_ERRORS_THROW)
        valid = np.zeros(1, dtype=np.uint8)
        reason = np.zeros(1, dtype=np.int32)
        cabi.check(self._engine.lib.rk_validate_batch(self._engine.h, C.byref(desc), C.byref(bin_),
                                                      int(check_current_state_within_limits),
                                                      int(check_target_state_within_limits), valid.ctypes.data, reason.ctypes.data))
        if not valid[0] and self.error_handler.throws:
            raise ValidationError(validation_message(int(reason[0])))
        return bool(valid[0])

    def calculate(self, input: InputParameter, traj: Trajectory) -> RuckigResult:
        """ruckig.rs:210-218."""
        eng = self._engine
        throws = self.error_handler.throws
        desc, bin_, keep = input.as_batch(self.delta_time, cabi.RK_ERRORS_THROW if throws else cabi.RK_ERRORS_IGNORE)
        if traj._dev is None or traj._engine is not eng:
            traj._engine = eng
            traj._dev = eng.new_trajectories(1, self.degrees_of_freedom)
        status = np.zeros(1, dtype=np.int32)
        duration = np.zeros(1)
        imd = np.zeros((self.degrees_of_freedom, 1))
        out = cabi.RkBatchOut(status.ctypes.data, duration.ctypes.data, imd.ctypes.data)
        cabi.check(eng.lib.rk_calculate_batch(eng.h, C.byref(desc), C.byref(bin_), traj._dev, C.byref(out)))
        code = RuckigResult(int(status[0]))
        if throws and code == RuckigResult.ErrorInvalidInput:
            detail = int(eng.read(traj._dev, cabi.RK_FIELD_ERROR_DETAIL, 1, self.degrees_of_freedom)[0])
            raise ValidationError(validation_message(detail))
        traj.duration = float(duration[0])
        traj.independent_min_durations = imd[:, 0].copy()
        if throws and code in _CALC_CODES:
            detail = int(eng.read(traj._dev, cabi.RK_FIELD_ERROR_DETAIL, 1, self.degrees_of_freedom)[0])
            raise CalculatorError(calculator_message(code, detail))
        return code

    def update(self, input: InputParameter, output: OutputParameter) -> RuckigResult:
        """ruckig.rs:266-321 (output.time is not reset on a new calculation, as in the reference)."""
        output.new_calculation = False
        if not self.current_input_initialized or input != self.current_input:
            self.calculate(input, output.trajectory)  # Ok(Error*) codes are dropped, ruckig.rs:285
            output.new_calculation = True
            self.current_input = input.copy()
            self.current_input_initialized = True

        output.time += self.delta_time
        p, v, a, j, _section = output.trajectory.at_time(output.time)
        output.new_position, output.new_velocity, output.new_acceleration, output.new_jerk = p, v, a, j
        output.did_section_change = False  # ruckig.rs:300-302: new_section is passed by value
        output.pass_to_input(self.current_input)
        if output.time > output.trajectory.get_duration():
            return RuckigResult.Finished
        return RuckigResult.Working


class BatchRuckig:
    """The batched entry point: Ruckig::calculate / Trajectory::at_time for n independent problems.

    Arrays are SoA `[dof][n]` float64 (numpy = host buffers, torch CUDA tensors = device buffers).
    """

    def __init__(self, dofs: int, delta_time: float, error_handler=ThrowErrorHandler, device: int = 0,
                 result_layout: int | None = None):
        """result_layout: cabi.RK_RESULT_FULL (the reference's Profile verbatim, 59 doubles per DoF) or
        cabi.RK_RESULT_COMPACT (20 doubles per DoF, expanded on the fly by every reader; same results bit for bit).
        Default: RK_B200_RESULT_LAYOUT from the environment ("compact" / "full"), else full."""
        self.degrees_of_freedom = int(dofs)
        self.delta_time = float(delta_time)
        self.error_handler = error_handler
        self._engine = _Engine.get(device)
        self._traj = None
        self._n = 0
        if result_layout is None:
            import os
            result_layout = cabi.RK_RESULT_COMPACT if os.environ.get("RK_B200_RESULT_LAYOUT", "full") == "compact" else cabi.RK_RESULT_FULL
        self.result_layout = int(result_layout)

    def close(self):
        """Release the device-side trajectories (also done when the object is collected)."""
        traj = self._traj
        self._traj = None
        self._n = 0
        if traj is not None:
            try:
                self._engine.free_trajectories(traj)
            except Exception:  # interpreter shutdown: the library may be gone already
                pass

    def __del__(self):
        self.close()

    @property
    def handle(self):
        return self._engine.h

    def _ensure(self, n: int):
        if self._traj is None or self._n != n:
            if self._traj is not None:
                self._engine.free_trajectories(self._traj)
            self._traj = self._engine.new_trajectories(n, self.degrees_of_freedom, self.result_layout)
            self._n = n

    def calculate(self, fields: dict, *, n: int, memory_space: int, control_interface=ControlInterface.Position,
                  synchronization=Synchronization.Time, duration_discretization=DurationDiscretization.Continuous,
                  per_dof_control_interface=None, per_dof_synchronization=None, status=None, duration=None,
                  independent_min_durations=None, limits_layout=cabi.RK_LIMITS_PER_TRAJECTORY):
        """fields: name -> buffer for the rk_batch_in members. Outputs (optional) are caller-owned buffers.
        limits_layout=cabi.RK_LIMITS_PER_DOF: the five limit arrays are `[dof]`, shared by the whole batch."""
        self._ensure(n)
        desc, keep = cabi.make_desc(n, self.degrees_of_freedom, control_interface=control_interface,
                                    synchronization=synchronization, duration_discretization=duration_discretization,
                                    error_mode=cabi.RK_ERRORS_THROW if self.error_handler.throws else cabi.RK_ERRORS_IGNORE,
                                    memory_space=memory_space, delta_time=self.delta_time,
                                    per_dof_control_interface=per_dof_control_interface,
                                    per_dof_synchronization=per_dof_synchronization, limits_layout=limits_layout)
        bin_ = cabi.make_in(fields)
        out = cabi.RkBatchOut(cabi.ptr(status), cabi.ptr(duration), cabi.ptr(independent_min_durations))
        eng = self._engine
        cabi.check(eng.lib.rk_calculate_batch(eng.h, C.byref(desc), C.byref(bin_), self._traj, C.byref(out)))

    def validate(self, fields: dict, *, n: int, check_current: bool, check_target: bool, **kw) -> np.ndarray:
        desc, keep = cabi.make_desc(n, self.degrees_of_freedom, memory_space=cabi.RK_MEM_HOST, delta_time=self.delta_time, **kw)
        bin_ = cabi.make_in(fields)
        valid = np.zeros(n, dtype=np.uint8)
        reason = np.zeros(n, dtype=np.int32)
        eng = self._engine
        cabi.check(eng.lib.rk_validate_batch(eng.h, C.byref(desc), C.byref(bin_), int(check_current), int(check_target),
                                             valid.ctypes.data, reason.ctypes.data))
        return valid, reason

    def read(self, field: int) -> np.ndarray:
        return self._engine.read(self._traj, field, self._n, self.degrees_of_freedom)

    def at_time(self, *, time_start: float, delta_time: float, steps: int, memory_space: int, position=None,
                velocity=None, acceleration=None, jerk=None, section=None):
        sd = cabi.RkSampleDesc(float(time_start), float(delta_time), int(steps), int(memory_space))
        so = cabi.RkSampleOut(cabi.ptr(position), cabi.ptr(velocity), cabi.ptr(acceleration), cabi.ptr(jerk), cabi.ptr(section))
        eng = self._engine
        cabi.check(eng.lib.rk_at_time_batch(eng.h, self._traj, C.byref(sd), C.byref(so)))

    def position_extrema(self) -> np.ndarray:
        """Trajectory::get_position_extrema for every (DoF, trajectory): [4][dof][n] = min, max, t_min, t_max (host array)."""
        out = np.zeros((4, self.degrees_of_freedom, self._n))
        eng = self._engine
        cabi.check(eng.lib.rk_position_extrema_batch(eng.h, self._traj, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data,
                                                     out[3].ctypes.data, cabi.RK_MEM_HOST))
        return out

    def first_time_at_position(self, positions: np.ndarray) -> np.ndarray:
        """Trajectory::get_first_time_at_position for one query position per (DoF, trajectory); NaN where the reference
        returns None."""
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        assert positions.shape == (self.degrees_of_freedom, self._n)
        out = np.zeros_like(positions)
        eng = self._engine
        cabi.check(eng.lib.rk_first_time_at_position_batch(eng.h, self._traj, positions.ctypes.data, out.ctypes.data, cabi.RK_MEM_HOST))
        return out

    def sync(self):
        cabi.check(self._engine.lib.rk_sync(self._engine.h))

    def kernel_launches(self) -> int:
        return int(self._engine.lib.rk_kernel_launches(self._engine.h))


class MultiBatchRuckig:
    """One host batch over several handles — one per GPU of the box by default (rk_calculate_batch_multi): contiguous index
    ranges (sharding.shard_range == rk_shard_range), one host thread per part, status / duration / independent minimum
    durations gathered into the caller's host arrays. `devices` may name a device more than once (several handles on one GPU)."""

    def __init__(self, dofs: int, delta_time: float, error_handler=ThrowErrorHandler, devices=None, result_layout: int = cabi.RK_RESULT_FULL):
        self.degrees_of_freedom, self.delta_time, self.error_handler = int(dofs), float(delta_time), error_handler
        self.lib = cabi.load()
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        self.devices = [int(d) for d in devices]
        self.result_layout = int(result_layout)
        self._handles, self._trajs, self._n = [], [], -1
        for d in self.devices:
            h = C.c_void_p()
            cabi.check(self.lib.rk_create(d, C.byref(h)))
            self._handles.append(h)

    def close(self):
        try:
            for t in self._trajs:
                self.lib.rk_trajectories_destroy(t)
            for h in self._handles:
                self.lib.rk_destroy(h)
        except Exception:
            pass
        self._trajs, self._handles = [], []

    def __del__(self):
        self.close()

    def part_range(self, n: int, g: int):
        from .sharding import shard_range
        return shard_range(n, g, len(self.devices))

    def _ensure(self, n: int):
        if n == self._n:
            return
        for t in self._trajs:
            self.lib.rk_trajectories_destroy(t)
        self._trajs = []
        for g, h in enumerate(self._handles):
            lo, hi = self.part_range(n, g)
            t = C.c_void_p()
            cabi.check(self.lib.rk_trajectories_create_ex(h, hi - lo, self.degrees_of_freedom, self.result_layout, C.byref(t)))
            self._trajs.append(t)
        self._n = n

    def calculate(self, fields: dict, *, n: int, status=None, duration=None, independent_min_durations=None, **kw):
        """fields / outputs: host (numpy or pinned torch) arrays [dof][n] / [n]."""
        self._ensure(n)
        desc, keep = cabi.make_desc(n, self.degrees_of_freedom, memory_space=cabi.RK_MEM_HOST, delta_time=self.delta_time,
                                    error_mode=cabi.RK_ERRORS_THROW if self.error_handler.throws else cabi.RK_ERRORS_IGNORE, **kw)
        bin_ = cabi.make_in(fields)
        out = cabi.RkBatchOut(cabi.ptr(status), cabi.ptr(duration), cabi.ptr(independent_min_durations))
        G = len(self._handles)
        hs = (C.c_void_p * G)(*[h.value for h in self._handles])
        ts = (C.c_void_p * G)(*[t.value for t in self._trajs])
        cabi.check(self.lib.rk_calculate_batch_multi(hs, ts, G, C.byref(desc), C.byref(bin_), C.byref(out)))

    def read(self, field: int) -> np.ndarray:
        """One field of all parts, concatenated in index order (host array)."""
        parts = []
        for g, (h, t) in enumerate(zip(self._handles, self._trajs)):
            lo, hi = self.part_range(self._n, g)
            m, dof = hi - lo, self.degrees_of_freedom
            if m == 0:
                continue
            if field in (cabi.RK_FIELD_STATUS, cabi.RK_FIELD_LIMITING_DOF, cabi.RK_FIELD_ERROR_DETAIL):
                o = np.empty(m, dtype=np.int32)
            elif field == cabi.RK_FIELD_DURATION:
                o = np.empty(m)
            elif field == cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS:
                o = np.empty((dof, m))
            elif field == cabi.RK_FIELD_CODES:
                o = np.empty((dof, m), dtype=np.uint32)
            else:
                o = np.empty((cabi.FIELD_SLOTS[field], dof, m))
            cabi.check(self.lib.rk_trajectories_read(h, t, field, o.ctypes.data, cabi.RK_MEM_HOST))
            parts.append(o)
        return np.concatenate(parts, axis=-1)


class BatchRollout:
    """n independent closed control loops on the device: the batched `Ruckig::update` (ruckig.rs:266-321).

    `update(fields, ...)` is one control cycle for every environment: those whose input changed (or that were never
    calculated) are re-solved, all advance their time by `delta_time`, sample their trajectory and — with
    `pass_to_input=True` — get the new state written into `fields["current_*"]`, ready for the next cycle
    (`OutputParameter::pass_to_input`). All buffers are device tensors `[dof][n]` / `[n]`.
    """

    def __init__(self, n: int, dofs: int, delta_time: float, error_handler=ThrowErrorHandler, device: int = 0):
        self.n, self.degrees_of_freedom, self.delta_time = int(n), int(dofs), float(delta_time)
        self.error_handler = error_handler
        self._engine = _Engine.get(device)
        r = C.c_void_p()
        cabi.check(self._engine.lib.rk_rollout_create(self._engine.h, self.n, self.degrees_of_freedom, C.byref(r)))
        self._r = r

    def __del__(self):
        try:
            if getattr(self, "_r", None):
                self._engine.lib.rk_rollout_destroy(self._r)
                self._r = None
        except Exception:
            pass

    def reset(self):
        cabi.check(self._engine.lib.rk_rollout_reset(self._engine.h, self._r))

    def update(self, fields: dict, *, control_interface=ControlInterface.Position, synchronization=Synchronization.Time,
               duration_discretization=DurationDiscretization.Continuous, per_dof_control_interface=None, per_dof_synchronization=None,
               new_position=None, new_velocity=None, new_acceleration=None, new_jerk=None, result=None, new_calculation=None,
               time=None, pass_to_input=True, reset_time_on_new_calculation=False,
               limits_layout=cabi.RK_LIMITS_PER_TRAJECTORY) -> int:
        """One control cycle; returns the number of environments that were re-solved."""
        desc, keep = cabi.make_desc(self.n, self.degrees_of_freedom, control_interface=control_interface,
                                    synchronization=synchronization, duration_discretization=duration_discretization,
                                    error_mode=cabi.RK_ERRORS_THROW if self.error_handler.throws else cabi.RK_ERRORS_IGNORE,
                                    memory_space=cabi.RK_MEM_DEVICE, delta_time=self.delta_time,
                                    per_dof_control_interface=per_dof_control_interface,
                                    per_dof_synchronization=per_dof_synchronization, limits_layout=limits_layout)
        bin_ = cabi.make_in(fields)
        out = cabi.RkUpdateOut(cabi.ptr(new_position), cabi.ptr(new_velocity), cabi.ptr(new_acceleration), cabi.ptr(new_jerk),
                               cabi.ptr(result), cabi.ptr(new_calculation), cabi.ptr(time), int(bool(pass_to_input)),
                               int(bool(reset_time_on_new_calculation)))
        m = C.c_int64(0)
        eng = self._engine
        cabi.check(eng.lib.rk_update_batch(eng.h, C.byref(desc), C.byref(bin_), self._r, C.byref(out), C.byref(m)))
        return int(m.value)

    def read(self, field: int) -> np.ndarray:
        return self._engine.read(self._engine.lib.rk_rollout_trajectories(self._r), field, self.n, self.degrees_of_freedom)

    def sync(self):
        cabi.check(self._engine.lib.rk_sync(self._engine.h))

