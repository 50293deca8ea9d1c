"""ctypes binding of the parity oracle (oracle/_build/liboracle_{glibc,portable}.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
`OracleRuckig` offers the same surface as rsruckig_b200.Ruckig so that the ports of the reference's tests run
unchanged against either side; `OracleBatch` is the batch checker.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rsruckig_b200 import _cabi as cabi
from rsruckig_b200.api import (CalculatorError, ControlSigns, Direction, InputParameter, OutputParameter, Profile,
                               ReachedLimits, RuckigResult, ThrowErrorHandler, Trajectory, ValidationError)

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build():
    subprocess.run(["make", "-C", HERE, "-j2"], check=True, capture_output=True)


def load(flavour: str = "portable"):
    """flavour: 'glibc' (libm of the platform, what the Rust reference links) or 'portable' (fdlibm restatement)."""
    if flavour in _LIBS:
        return _LIBS[flavour]
    path = os.path.join(HERE, "_build", f"liboracle_{flavour}.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.rko_math_flavour.restype = C.c_char_p
    lib.rko_batch_new.argtypes = [C.c_int64, C.c_int32]
    lib.rko_batch_new.restype = vp
    lib.rko_batch_free.argtypes = [vp]
    lib.rko_batch_calculate.argtypes = [vp, C.POINTER(cabi.RkBatchDesc), C.POINTER(cabi.RkBatchIn), C.c_int32]
    lib.rko_batch_calculate.restype = C.c_double
    lib.rko_bench_calculate.argtypes = [C.POINTER(cabi.RkBatchDesc), C.POINTER(cabi.RkBatchIn), C.c_int32, vp, vp]
    lib.rko_bench_calculate.restype = C.c_double
    lib.rko_batch_validate.argtypes = [C.POINTER(cabi.RkBatchDesc), C.POINTER(cabi.RkBatchIn), C.c_int32, C.c_int32, vp]
    lib.rko_batch_validate_messages.argtypes = [C.POINTER(cabi.RkBatchDesc), C.POINTER(cabi.RkBatchIn), C.c_int32, C.c_int32, vp, vp, C.c_int32]
    lib.rko_batch_read.argtypes = [vp, C.c_int32, vp]
    lib.rko_batch_read.restype = C.c_int32
    lib.rko_batch_position_extrema.argtypes = [vp, vp]
    lib.rko_batch_first_time_at_position.argtypes = [vp, vp, vp]
    lib.rko_batch_at_time.argtypes = [vp, C.POINTER(cabi.RkSampleDesc), C.POINTER(cabi.RkSampleOut), C.c_int32]
    lib.rko_batch_at_time.restype = C.c_double
    lib.rko_otg_new.argtypes = [C.c_int32, C.c_double, C.c_int32]
    lib.rko_otg_new.restype = vp
    for f in ("rko_otg_free", "rko_otg_reset", "rko_otg_new_output"):
        getattr(lib, f).argtypes = [vp]
    lib.rko_otg_set_input.argtypes = [vp, C.POINTER(cabi.RkBatchDesc), C.POINTER(cabi.RkBatchIn)]
    for f in ("rko_otg_update", "rko_otg_calculate", "rko_otg_last_code", "rko_otg_new_calculation"):
        getattr(lib, f).argtypes = [vp]
        getattr(lib, f).restype = C.c_int32
    lib.rko_otg_validate.argtypes = [vp, C.c_int32, C.c_int32]
    lib.rko_otg_validate.restype = C.c_int32
    lib.rko_otg_last_message.argtypes = [vp]
    lib.rko_otg_last_message.restype = C.c_char_p
    lib.rko_otg_duration.argtypes = [vp]
    lib.rko_otg_duration.restype = C.c_double
    lib.rko_otg_time.argtypes = [vp]
    lib.rko_otg_time.restype = C.c_double
    lib.rko_otg_get_output.argtypes = [vp, C.c_int32, vp]
    lib.rko_otg_at_time.argtypes = [vp, C.c_double, vp, vp, vp, vp, vp]
    lib.rko_otg_profile.argtypes = [vp, C.c_int32, C.c_int32, vp]
    lib.rko_otg_profile.restype = C.c_int32
    _LIBS[flavour] = lib
    return lib


class _OracleTrajectoryView:
    """What `output.trajectory` / `traj` resolve to for the oracle: reads go to the C++ object."""

    def __init__(self, owner: "OracleRuckig"):
        self._o = owner

    def get_duration(self):
        return self._o.lib.rko_otg_duration(self._o.h)

    @property
    def duration(self):
        return self.get_duration()

    def get_independent_min_durations(self):
        out = np.zeros(self._o.degrees_of_freedom)
        self._o.lib.rko_otg_get_output(self._o.h, 4, out.ctypes.data)
        return out

    def at_time(self, time: float):
        D = self._o.degrees_of_freedom
        p, v, a, j = (np.zeros(D) for _ in range(4))
        sec = C.c_int32(0)
        self._o.lib.rko_otg_at_time(self._o.h, float(time), p.ctypes.data, v.ctypes.data, a.ctypes.data, j.ctypes.data, C.byref(sec))
        return p, v, a, j, int(sec.value)

    def get_profiles(self):
        out = []
        buf = np.zeros(8)
        for d in range(self._o.degrees_of_freedom):
            q = Profile()
            for name, f in (("t", cabi.RK_FIELD_T), ("t_sum", cabi.RK_FIELD_T_SUM), ("j", cabi.RK_FIELD_J),
                            ("a", cabi.RK_FIELD_A), ("v", cabi.RK_FIELD_V), ("p", cabi.RK_FIELD_P)):
                k = self._o.lib.rko_otg_profile(self._o.h, d, f, buf.ctypes.data)
                setattr(q, name, buf[:k].copy())
            self._o.lib.rko_otg_profile(self._o.h, d, cabi.RK_FIELD_BRAKE_DURATION, buf.ctypes.data)
            q.brake_duration = float(buf[0])
            self._o.lib.rko_otg_profile(self._o.h, d, cabi.RK_FIELD_BRAKE_T, buf.ctypes.data)
            q.brake_t = buf[:2].copy()
            self._o.lib.rko_otg_profile(self._o.h, d, cabi.RK_FIELD_CODES, buf.ctypes.data)
            q.limits, q.control_signs, q.direction = ReachedLimits(int(buf[0])), ControlSigns(int(buf[1])), Direction(int(buf[2]))
            out.append(q)
        return [out]


class OracleRuckig:
    """Same surface as rsruckig_b200.Ruckig, computed by the C++ restatement (stateful like the reference)."""

    def __init__(self, dofs: int, delta_time: float, error_handler=ThrowErrorHandler, flavour: str = "portable"):
        self.lib = load(flavour)
        self.degrees_of_freedom = int(dofs)
        self.delta_time = float(delta_time)
        self.error_handler = error_handler
        self.h = self.lib.rko_otg_new(dofs, delta_time, 1 if error_handler.throws else 0)
        self._bound = None  # the python OutputParameter / Trajectory the C++ output currently stands for

    def __del__(self):
        try:
            self.lib.rko_otg_free(self.h)
        except Exception:
            pass

    def reset(self):
        self.lib.rko_otg_reset(self.h)

    def _bind(self, obj):
        if self._bound is not obj:
            self.lib.rko_otg_new_output(self.h)  # a fresh OutputParameter / Trajectory on the reference side
            self._bound = obj

    def _set(self, input: InputParameter):
        desc, bin_, keep = input.as_batch(self.delta_time, cabi.RK_ERRORS_THROW)
        self.lib.rko_otg_set_input(self.h, C.byref(desc), C.byref(bin_))

    def _result(self, r: int):
        if r == -1000:
            raise ValidationError(self.lib.rko_otg_last_message(self.h).decode())
        if r == -2000:
            raise CalculatorError(self.lib.rko_otg_last_message(self.h).decode())
        return RuckigResult(r)

    def validate_input(self, input, check_current_state_within_limits, check_target_state_within_limits):
        self._set(input)
        # validity itself is reported through the batch checker; here only the Err/Ok behaviour matters
        r = self.lib.rko_otg_validate(self.h, int(check_current_state_within_limits), int(check_target_state_within_limits))
        self._result(r)
        desc, bin_, keep = input.as_batch(self.delta_time, cabi.RK_ERRORS_THROW)
        valid = np.zeros(1, dtype=np.uint8)
        self.lib.rko_batch_validate(C.byref(desc), C.byref(bin_), int(check_current_state_within_limits),
                                    int(check_target_state_within_limits), valid.ctypes.data)
        return bool(valid[0])

    def calculate(self, input: InputParameter, traj: Trajectory):
        self._bind(traj)
        self._set(input)
        r = self._result(self.lib.rko_otg_calculate(self.h))
        view = _OracleTrajectoryView(self)
        traj.duration = view.get_duration()
        traj.independent_min_durations = view.get_independent_min_durations()
        traj.get_profiles = view.get_profiles
        traj.at_time = view.at_time
        return r

    def update(self, input: InputParameter, output: OutputParameter):
        self._bind(output)
        self._set(input)
        code = self.lib.rko_otg_update(self.h)
        output.new_calculation = bool(self.lib.rko_otg_new_calculation(self.h))
        r = self._result(code)
        D = self.degrees_of_freedom
        for which, name in enumerate(("new_position", "new_velocity", "new_acceleration", "new_jerk")):
            buf = np.zeros(D)
            self.lib.rko_otg_get_output(self.h, which, buf.ctypes.data)
            setattr(output, name, buf)
        output.time = self.lib.rko_otg_time(self.h)
        output.new_calculation = bool(self.lib.rko_otg_new_calculation(self.h))
        view = _OracleTrajectoryView(self)
        output.trajectory.duration = view.get_duration()
        output.trajectory.independent_min_durations = view.get_independent_min_durations()
        output.trajectory.get_profiles = view.get_profiles
        output.trajectory.at_time = view.at_time
        return r


class OracleBatch:
    """Batch checker with the product's buffer conventions (host numpy arrays, SoA [dof][n])."""

    def __init__(self, n: int, dof: int, flavour: str = "portable"):
        self.lib = load(flavour)
        self.n, self.dof = int(n), int(dof)
        self.h = self.lib.rko_batch_new(n, dof)

    def __del__(self):
        try:
            self.lib.rko_batch_free(self.h)
        except Exception:
            pass

    def calculate(self, fields: dict, *, delta_time=0.005, threads=1, error_mode=cabi.RK_ERRORS_THROW, **kw) -> float:
        desc, keep = cabi.make_desc(self.n, self.dof, delta_time=delta_time, error_mode=error_mode, memory_space=cabi.RK_MEM_HOST, **kw)
        bin_ = cabi.make_in(fields)
        return self.lib.rko_batch_calculate(self.h, C.byref(desc), C.byref(bin_), int(threads))

    def validate(self, fields: dict, *, check_current: bool, check_target: bool, delta_time=0.005, **kw) -> np.ndarray:
        desc, keep = cabi.make_desc(self.n, self.dof, delta_time=delta_time, memory_space=cabi.RK_MEM_HOST, **kw)
        bin_ = cabi.make_in(fields)
        valid = np.zeros(self.n, dtype=np.uint8)
        self.lib.rko_batch_validate(C.byref(desc), C.byref(bin_), int(check_current), int(check_target), valid.ctypes.data)
        return valid

    def validate_messages(self, fields: dict, *, check_current: bool, check_target: bool, delta_time=0.005, **kw):
        """(valid[n], messages[n]): the oracle's text of the first failing check ("<check> (DoF d)"), "" when none."""
        desc, keep = cabi.make_desc(self.n, self.dof, delta_time=delta_time, memory_space=cabi.RK_MEM_HOST, **kw)
        bin_ = cabi.make_in(fields)
        valid = np.zeros(self.n, dtype=np.uint8)
        stride = 192
        msgs = np.zeros((self.n, stride), dtype=np.uint8)
        self.lib.rko_batch_validate_messages(C.byref(desc), C.byref(bin_), int(check_current), int(check_target), valid.ctypes.data,
                                             msgs.ctypes.data, stride)
        return valid, msgs.view(f"S{stride}")[:, 0]

    def read(self, field: int) -> np.ndarray:
        n, dof = self.n, self.dof
        if field in (cabi.RK_FIELD_STATUS, cabi.RK_FIELD_LIMITING_DOF):
            out = np.empty(n, dtype=np.int32)
        elif field == cabi.RK_FIELD_DURATION:
            out = np.empty(n, dtype=np.float64)
        elif field == cabi.RK_FIELD_INDEPENDENT_MIN_DURATIONS:
            out = np.empty((dof, n), dtype=np.float64)
        elif field == cabi.RK_FIELD_CODES:
            out = np.empty((dof, n), dtype=np.uint32)
        else:
            out = np.empty((cabi.FIELD_SLOTS[field], dof, n), dtype=np.float64)
        assert self.lib.rko_batch_read(self.h, field, out.ctypes.data) == 0
        return out

    def position_extrema(self) -> np.ndarray:
        """Trajectory::get_position_extrema: [4][dof][n] = min, max, t_min, t_max."""
        out = np.zeros((4, self.dof, self.n))
        self.lib.rko_batch_position_extrema(self.h, out.ctypes.data)
        return out

    def first_time_at_position(self, positions: np.ndarray) -> np.ndarray:
        """Trajectory::get_first_time_at_position for one query per (dof, trajectory); NaN = None."""
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        out = np.zeros((self.dof, self.n))
        self.lib.rko_batch_first_time_at_position(self.h, positions.ctypes.data, out.ctypes.data)
        return out

    def at_time(self, *, time_start: float, delta_time: float, steps: int, threads=1, want_jerk=True):
        K, D, n = int(steps), self.dof, self.n
        pos, vel, acc = (np.empty((K, D, n)) for _ in range(3))
        jerk = np.empty((K, D, n)) if want_jerk else None
        sec = np.empty((K, n), dtype=np.int32)
        sd = cabi.RkSampleDesc(float(time_start), float(delta_time), K, cabi.RK_MEM_HOST)
        so = cabi.RkSampleOut(pos.ctypes.data, vel.ctypes.data, acc.ctypes.data, cabi.ptr(jerk), sec.ctypes.data)
        self.lib.rko_batch_at_time(self.h, C.byref(sd), C.byref(so), int(threads))
        return pos, vel, acc, jerk, sec


def bench_calculate(fields: dict, *, n: int, dof: int, threads: int, flavour="glibc", delta_time=0.005, want_results=False, **kw):
    """Reference-bench style loop (reused calculator per thread). Returns (seconds, status, duration)."""
    lib = load(flavour)
    desc, keep = cabi.make_desc(n, dof, delta_time=delta_time, memory_space=cabi.RK_MEM_HOST, **kw)
    bin_ = cabi.make_in(fields)
    status = np.zeros(n, dtype=np.int32) if want_results else None
    duration = np.zeros(n) if want_results else None
    sec = lib.rko_bench_calculate(C.byref(desc), C.byref(bin_), int(threads), cabi.ptr(status), cabi.ptr(duration))
    return sec, status, duration
