"""Stable binary files for a batch of problems and for its results.

Purpose: a third party with a Rust toolchain can run the real rsruckig on exactly the inputs used here and diff the
outcome (`tools/rsruckig_dump/`, `tools/compare_results.py`) — the one check this repository cannot run itself.

Both files are little-endian and start with a 64-byte header:

    offset  type      inputs file ("RKB2IN\\0\\0")                      results file ("RKB2RS\\0\\0")
    0       char[8]   magic                                            magic
    8       u32       format version (1)                               format version (1)
    12      i32       dof                                              dof
    16      i64       n                                                n
    24      i32       control_interface  (0 Position, 1 Velocity)      reserved (0)
    28      i32       synchronization    (0 Time, 1 TimeIfNecessary, 2 Phase, 3 None)   reserved (0)
    32      i32       duration_discretization (0 Continuous, 1 Discrete)                reserved (0)
    36      u32       bit mask of optional arrays present (inputs)     reserved (0)
    40      f64       delta_time                                       reserved (0)
    48      16 bytes  reserved (0)

Inputs body: the nine mandatory arrays in the order of `workloads.FIELDS`, each f64 [dof][n] (trajectory index fastest), then
the optional ones whose bit is set: bit 0 min_velocity f64 [dof][n], bit 1 min_acceleration f64 [dof][n], bit 2 enabled
u8 [dof][n], bit 3 minimum_duration f64 [n] (NaN = None).

Results body: status i32 [n] (RuckigResult value; -100 = Err(ValidationError)), duration f64 [n], codes u32 [dof][n]
(limits | control_signs << 8 | direction << 16, the enums' declaration order), t f64 [7][dof][n] (Profile::t).
"""
from __future__ import annotations

import struct

import numpy as np

from .workloads import FIELDS

MAGIC_IN, MAGIC_RS, VERSION = b"RKB2IN\0\0", b"RKB2RS\0\0", 1
OPTIONAL = (("min_velocity", np.float64, True), ("min_acceleration", np.float64, True), ("enabled", np.uint8, True),
            ("minimum_duration", np.float64, False))
_HEADER = struct.Struct("<8sIiqiiiId16x")
assert _HEADER.size == 64


def write_inputs(path: str, fields: dict, *, n: int, dof: int, delta_time: float, control_interface: int = 0,
                 synchronization: int = 0, duration_discretization: int = 0) -> None:
    mask = sum(1 << b for b, (name, _, _) in enumerate(OPTIONAL) if fields.get(name) is not None)
    with open(path, "wb") as fh:
        fh.write(_HEADER.pack(MAGIC_IN, VERSION, dof, n, int(control_interface), int(synchronization), int(duration_discretization),
                              mask, float(delta_time)))
        for name in FIELDS:
            a = np.ascontiguousarray(fields[name], dtype="<f8")
            assert a.shape == (dof, n), (name, a.shape)
            fh.write(a.tobytes())
        for name, dtype, per_dof in OPTIONAL:
            if fields.get(name) is not None:
                a = np.ascontiguousarray(fields[name], dtype=np.dtype(dtype).newbyteorder("<"))
                assert a.shape == ((dof, n) if per_dof else (n,)), (name, a.shape)
                fh.write(a.tobytes())


def read_inputs(path: str):
    """Returns (fields, meta) with meta = dict(n, dof, delta_time, control_interface, synchronization, duration_discretization)."""
    with open(path, "rb") as fh:
        magic, version, dof, n, ci, sync, disc, mask, dt = _HEADER.unpack(fh.read(_HEADER.size))
        if magic != MAGIC_IN or version != VERSION:
            raise ValueError(f"{path}: not a version-{VERSION} batch input file")
        fields = {}
        for name in FIELDS:
            fields[name] = np.frombuffer(fh.read(8 * dof * n), dtype="<f8").reshape(dof, n).copy()
        for b, (name, dtype, per_dof) in enumerate(OPTIONAL):
            if mask >> b & 1:
                count = dof * n if per_dof else n
                a = np.frombuffer(fh.read(count * np.dtype(dtype).itemsize), dtype=np.dtype(dtype).newbyteorder("<"))
                fields[name] = (a.reshape(dof, n) if per_dof else a).copy()
        if fh.read(1):
            raise ValueError(f"{path}: trailing bytes")
    meta = dict(n=n, dof=dof, delta_time=dt, control_interface=ci, synchronization=sync, duration_discretization=disc)
    return fields, meta


def write_results(path: str, *, status, duration, codes, t) -> None:
    status = np.ascontiguousarray(status, dtype="<i4")
    n = status.shape[0]
    codes = np.ascontiguousarray(codes, dtype="<u4") & np.uint32(0x00FFFFFF)  # the producing-stage byte is ours, not the reference's
    dof = codes.shape[0]
    t = np.ascontiguousarray(t, dtype="<f8")
    assert codes.shape == (dof, n) and t.shape == (7, dof, n) and np.shape(duration) == (n,)
    with open(path, "wb") as fh:
        fh.write(_HEADER.pack(MAGIC_RS, VERSION, dof, n, 0, 0, 0, 0, 0.0))
        fh.write(status.tobytes())
        fh.write(np.ascontiguousarray(duration, dtype="<f8").tobytes())
        fh.write(codes.tobytes())
        fh.write(t.tobytes())


def read_results(path: str) -> dict:
    with open(path, "rb") as fh:
        magic, version, dof, n, *_ = _HEADER.unpack(fh.read(_HEADER.size))
        if magic != MAGIC_RS or version != VERSION:
            raise ValueError(f"{path}: not a version-{VERSION} batch results file")
        status = np.frombuffer(fh.read(4 * n), dtype="<i4").copy()
        duration = np.frombuffer(fh.read(8 * n), dtype="<f8").copy()
        codes = np.frombuffer(fh.read(4 * dof * n), dtype="<u4").reshape(dof, n).copy()
        t = np.frombuffer(fh.read(8 * 7 * dof * n), dtype="<f8").reshape(7, dof, n).copy()
        if fh.read(1):
            raise ValueError(f"{path}: trailing bytes")
    return dict(n=n, dof=dof, status=status, duration=duration, codes=codes, t=t)


def compare_results(a: dict, b: dict, *, rtol: float = 1e-9) -> dict:
    """north_star's bar between two result sets: statuses and profile codes exact, durations and t[] within rtol (relative,
    over the trajectories both sides report as Working)."""
    assert a["n"] == b["n"] and a["dof"] == b["dof"]
    ok = (a["status"] == 0) & (b["status"] == 0)
    rel = lambda x, y: np.abs(x - y) / np.maximum(np.maximum(np.abs(x), np.abs(y)), 1e-300)
    dur = rel(a["duration"][ok], b["duration"][ok])
    tt = np.abs(a["t"][:, :, ok] - b["t"][:, :, ok]) / np.maximum(a["duration"][ok], 1e-300)
    return {"n": int(a["n"]), "working_both": int(ok.sum()), "status_mismatches": int((a["status"] != b["status"]).sum()),
            "code_mismatches": int((a["codes"][:, ok] != b["codes"][:, ok]).sum()),
            "max_rel_duration_diff": float(dur.max()) if dur.size else 0.0,
            "max_phase_time_diff_rel_to_duration": float(tt.max()) if tt.size else 0.0,
            "pass": bool((a["status"] == b["status"]).all() and (a["codes"][:, ok] == b["codes"][:, ok]).all()
                         and (dur.size == 0 or dur.max() <= rtol))}
