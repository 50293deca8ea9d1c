"""ctypes view of include/ruckig_b200.h.

The struct layouts here mirror the header one to one; `load()` opens the CUDA
shared library built in-tree (rsruckig_b200/csrc/libruckig_b200.so). There is
no CPU fallback: if the library is missing `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

RK_OK = 0
RK_POSITION, RK_VELOCITY, RK_ACCELERATION = 0, 1, 2
RK_SYNC_TIME, RK_SYNC_TIME_IF_NECESSARY, RK_SYNC_PHASE, RK_SYNC_NONE = 0, 1, 2, 3
RK_CONTINUOUS, RK_DISCRETE = 0, 1
RK_ERRORS_THROW, RK_ERRORS_IGNORE = 0, 1
RK_MEM_HOST, RK_MEM_DEVICE = 0, 1
RK_LIMITS_PER_TRAJECTORY, RK_LIMITS_PER_DOF = 0, 1
RK_RESULT_FULL, RK_RESULT_COMPACT = 0, 1
RK_MAX_DOF = 64

(RK_FIELD_STATUS, RK_FIELD_DURATION, RK_FIELD_INDEPENDENT_MIN_DURATIONS, RK_FIELD_LIMITING_DOF, RK_FIELD_CODES,
 RK_FIELD_T, RK_FIELD_T_SUM, RK_FIELD_J, RK_FIELD_A, RK_FIELD_V, RK_FIELD_P, RK_FIELD_BRAKE_DURATION,
 RK_FIELD_BRAKE_T, RK_FIELD_BRAKE_J, RK_FIELD_BRAKE_A, RK_FIELD_BRAKE_V, RK_FIELD_BRAKE_P, RK_FIELD_PF_VF_AF,
 RK_FIELD_ERROR_DETAIL, RK_FIELD_COUNT) = range(20)

# slots per (dof, trajectory) of each profile field; 0 = per-trajectory field
FIELD_SLOTS = {
    RK_FIELD_T: 7, RK_FIELD_T_SUM: 7, RK_FIELD_J: 7, RK_FIELD_A: 8, RK_FIELD_V: 8, RK_FIELD_P: 8,
    RK_FIELD_BRAKE_DURATION: 1, RK_FIELD_BRAKE_T: 2, RK_FIELD_BRAKE_J: 2, RK_FIELD_BRAKE_A: 2,
    RK_FIELD_BRAKE_V: 2, RK_FIELD_BRAKE_P: 2, RK_FIELD_PF_VF_AF: 3,
}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class RkBatchDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("dof", C.c_int32),
        ("control_interface", C.c_int32),
        ("synchronization", C.c_int32),
        ("duration_discretization", C.c_int32),
        ("error_mode", C.c_int32),
        ("memory_space", C.c_int32),
        ("limits_layout", C.c_int32),
        ("delta_time", C.c_double),
        ("per_dof_control_interface", c_int32_p),
        ("per_dof_synchronization", c_int32_p),
    ]


IN_FIELDS = ("current_position", "current_velocity", "current_acceleration", "target_position", "target_velocity",
             "target_acceleration", "max_velocity", "max_acceleration", "max_jerk", "min_velocity", "min_acceleration")


class RkBatchIn(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in IN_FIELDS] + [("enabled", C.c_void_p), ("minimum_duration", C.c_void_p)]


class RkBatchOut(C.Structure):
    _fields_ = [("status", C.c_void_p), ("duration", C.c_void_p), ("independent_min_durations", C.c_void_p)]


class RkSampleDesc(C.Structure):
    _fields_ = [("time_start", C.c_double), ("delta_time", C.c_double), ("steps", C.c_int32), ("memory_space", C.c_int32)]


class RkSampleOut(C.Structure):
    _fields_ = [("position", C.c_void_p), ("velocity", C.c_void_p), ("acceleration", C.c_void_p), ("jerk", C.c_void_p),
                ("section", C.c_void_p)]


class RkUpdateOut(C.Structure):
    _fields_ = [("new_position", C.c_void_p), ("new_velocity", C.c_void_p), ("new_acceleration", C.c_void_p), ("new_jerk", C.c_void_p),
                ("result", C.c_void_p), ("new_calculation", C.c_void_p), ("time", C.c_void_p), ("pass_to_input", C.c_int32),
                ("reset_time_on_new_calculation", C.c_int32)]


def ptr(a) -> int | None:
    """Address of a numpy array (host) / torch tensor (host or device) / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "buffers handed to the C ABI must be contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def make_desc(n, dof, *, control_interface=RK_POSITION, synchronization=RK_SYNC_TIME,
              duration_discretization=RK_CONTINUOUS, error_mode=RK_ERRORS_THROW, memory_space=RK_MEM_HOST,
              delta_time=0.005, per_dof_control_interface=None, per_dof_synchronization=None, limits_layout=0):
    """Build an rk_batch_desc; returns (desc, keepalive) — keep `keepalive` referenced while desc is in use."""
    keep = []
    d = RkBatchDesc()
    d.n, d.dof = int(n), int(dof)
    d.control_interface, d.synchronization = int(control_interface), int(synchronization)
    d.duration_discretization, d.error_mode, d.memory_space = int(duration_discretization), int(error_mode), int(memory_space)
    d.delta_time = float(delta_time)
    d.limits_layout = int(limits_layout)
    for name, val in (("per_dof_control_interface", per_dof_control_interface), ("per_dof_synchronization", per_dof_synchronization)):
        if val is not None:
            arr = np.ascontiguousarray(np.asarray([int(x) for x in val], dtype=np.int32))
            assert arr.shape == (dof,)
            keep.append(arr)
            setattr(d, name, arr.ctypes.data_as(c_int32_p))
    return d, keep


def make_in(fields: dict) -> RkBatchIn:
    """Build an rk_batch_in from a dict name -> array/tensor/None (missing optional fields = NULL)."""
    s = RkBatchIn()
    for f in IN_FIELDS + ("enabled", "minimum_duration"):
        setattr(s, f, ptr(fields.get(f)))
    return s


_LIB = None
LIB_PATH = os.environ.get("RK_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libruckig_b200.so")

EXPORTS = (
    "rk_create", "rk_destroy", "rk_trajectories_create", "rk_trajectories_create_ex", "rk_trajectories_destroy", "rk_calculate_batch",
    "rk_validate_batch", "rk_at_time_batch", "rk_trajectories_read", "rk_sync", "rk_stream", "rk_kernel_launches",
    "rk_status_string", "rk_last_error", "rk_version", "rk_measure_fp64_peak", "rk_selftest_division",
    "rk_position_extrema_batch", "rk_first_time_at_position_batch",
    "rk_debug_counters", "rk_shard_range", "rk_calculate_batch_multi", "rk_rollout_create", "rk_rollout_destroy", "rk_rollout_reset", "rk_rollout_trajectories", "rk_update_batch",
)


def load():
    """Open libruckig_b200.so and declare prototypes. Raises if the CUDA library was not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.rk_create.argtypes = [C.c_int32, C.POINTER(vp)]
    lib.rk_destroy.argtypes = [vp]
    lib.rk_trajectories_create.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(vp)]
    lib.rk_trajectories_create_ex.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.rk_trajectories_destroy.argtypes = [vp]
    lib.rk_calculate_batch.argtypes = [vp, C.POINTER(RkBatchDesc), C.POINTER(RkBatchIn), vp, C.POINTER(RkBatchOut)]
    lib.rk_calculate_batch_multi.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int32, C.POINTER(RkBatchDesc), C.POINTER(RkBatchIn), C.POINTER(RkBatchOut)]
    lib.rk_shard_range.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.rk_shard_range.restype = None
    lib.rk_debug_counters.argtypes = [vp, vp]
    lib.rk_validate_batch.argtypes = [vp, C.POINTER(RkBatchDesc), C.POINTER(RkBatchIn), C.c_int32, C.c_int32, vp, vp]
    lib.rk_at_time_batch.argtypes = [vp, vp, C.POINTER(RkSampleDesc), C.POINTER(RkSampleOut)]
    lib.rk_trajectories_read.argtypes = [vp, vp, C.c_int32, vp, C.c_int32]
    lib.rk_sync.argtypes = [vp]
    lib.rk_stream.argtypes = [vp]
    lib.rk_stream.restype = vp
    lib.rk_kernel_launches.argtypes = [vp]
    lib.rk_kernel_launches.restype = C.c_int64
    lib.rk_measure_fp64_peak.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
    lib.rk_measure_fp64_peak.restype = C.c_int32
    lib.rk_selftest_division.argtypes = [vp, C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    lib.rk_selftest_division.restype = C.c_int32
    lib.rk_status_string.argtypes = [C.c_int32]
    lib.rk_status_string.restype = C.c_char_p
    lib.rk_last_error.restype = C.c_char_p
    lib.rk_version.restype = C.c_char_p
    lib.rk_position_extrema_batch.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int32]
    lib.rk_first_time_at_position_batch.argtypes = [vp, vp, vp, vp, C.c_int32]
    lib.rk_rollout_create.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(vp)]
    lib.rk_rollout_destroy.argtypes = [vp]
    lib.rk_rollout_reset.argtypes = [vp, vp]
    lib.rk_rollout_trajectories.argtypes = [vp]
    lib.rk_rollout_trajectories.restype = vp
    lib.rk_update_batch.argtypes = [vp, C.POINTER(RkBatchDesc), C.POINTER(RkBatchIn), vp, C.POINTER(RkUpdateOut), C.POINTER(C.c_int64)]
    for f in ("rk_create", "rk_destroy", "rk_trajectories_create", "rk_trajectories_create_ex", "rk_trajectories_destroy", "rk_calculate_batch",
              "rk_validate_batch", "rk_at_time_batch", "rk_trajectories_read", "rk_sync", "rk_debug_counters", "rk_calculate_batch_multi",
              "rk_rollout_create", "rk_rollout_destroy", "rk_rollout_reset", "rk_update_batch", "rk_position_extrema_batch", "rk_first_time_at_position_batch"):
        getattr(lib, f).restype = C.c_int32
    _LIB = lib
    return lib


class RkError(RuntimeError):
    pass


def check(status: int):
    if status != RK_OK:
        lib = load()
        raise RkError(f"{lib.rk_status_string(status).decode()} ({status}): {lib.rk_last_error().decode()}")
