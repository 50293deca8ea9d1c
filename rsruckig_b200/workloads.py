"""Synthetic workloads: the reference bench's input distribution (bench/src/benchmark_target.rs:115-153),
generalised to any DoF count and to the BASELINE.json configurations.

Three independent seeded streams, as in the reference (seeds 42 / 43 / 44): positions ~ N(0, 4); velocities /
accelerations ~ N(0, 0.8) or exactly 0 (current v: p=0.9, current a: 0.8, target v: 0.7, target a: 0.6);
max_velocity = U(0.1, 12) + |target_velocity|; max_acceleration = U(0.1, 12) + |target_acceleration|;
max_jerk = U(0.1, 12). `bench_inputs` draws from numpy PCG64 streams (any seed, vectorised: what the tests and the
benchmark use); `reference_bench_inputs` is the reference's OWN stream — Pcg64Mcg seeds 42 / 43 / 44 through rand_distr's
ziggurat Normal and rand's Uniform in the draw order of benchmark_target.rs:137-153 (csrc/refstream.c) — so that a Rust
build of rsruckig can be run on bit-identical problems (tools/rsruckig_dump). Layout: SoA [dof][n] float64.
"""
from __future__ import annotations

import numpy as np

FIELDS = ("current_position", "current_velocity", "current_acceleration", "target_position", "target_velocity",
          "target_acceleration", "max_velocity", "max_acceleration", "max_jerk")


def bench_inputs(n: int, dof: int, *, seed: int = 0, p_target_velocity: float = 0.7, p_target_acceleration: float = 0.6) -> dict:
    """The reference bench distribution for n trajectories x dof; returns name -> [dof][n] float64 arrays."""
    pos = np.random.Generator(np.random.PCG64(42 + 1000 * seed))
    dyn = np.random.Generator(np.random.PCG64(43 + 1000 * seed))
    lim = np.random.Generator(np.random.PCG64(44 + 1000 * seed))
    shape = (dof, n)

    def fill_or_zero(p):
        keep = dyn.random(shape) < p
        vals = dyn.normal(0.0, 0.8, shape)
        return np.where(keep, vals, 0.0)

    f = {}
    f["current_position"] = pos.normal(0.0, 4.0, shape)
    f["current_velocity"] = fill_or_zero(0.9)
    f["current_acceleration"] = fill_or_zero(0.8)
    f["target_position"] = pos.normal(0.0, 4.0, shape)
    f["target_velocity"] = fill_or_zero(p_target_velocity)
    f["target_acceleration"] = fill_or_zero(p_target_acceleration)
    f["max_velocity"] = lim.uniform(0.1, 12.0, shape) + np.abs(f["target_velocity"])
    f["max_acceleration"] = lim.uniform(0.1, 12.0, shape) + np.abs(f["target_acceleration"])
    f["max_jerk"] = lim.uniform(0.1, 12.0, shape)
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in f.items()}


def config2_inputs(n: int, *, seed: int = 0) -> dict:
    """BASELINE.json configs[1]: 6-DoF position interface, nonzero target velocity / acceleration, Time sync."""
    return bench_inputs(n, 6, seed=seed, p_target_velocity=1.0, p_target_acceleration=1.0)


def config3_inputs(n: int, *, seed: int = 0) -> dict:
    """BASELINE.json configs[2]: 3-DoF velocity interface + Phase synchronisation. Every 4th trajectory is made
    collinear (DoFs 1, 2 are scaled copies of DoF 0) so that phase synchronisation actually succeeds there; the
    caller adds minimum_duration / Discrete per third of the batch."""
    f = bench_inputs(n, 3, seed=seed)
    rng = np.random.Generator(np.random.PCG64(45 + 1000 * seed))
    col = np.arange(n) % 4 == 0
    scale = rng.uniform(0.25, 2.0, (3, n))
    scale[0] = 1.0
    for k in ("current_velocity", "current_acceleration", "target_velocity", "target_acceleration"):
        f[k][:, col] = (f[k][0:1, :] * scale)[:, col]
    # limits scaled alike keep the scaled profile feasible
    for k in ("max_velocity", "max_acceleration", "max_jerk"):
        f[k][:, col] = (f[k][0:1, :] * scale)[:, col]
    return f


def edge_case_inputs(n: int, dof: int, *, seed: int = 0) -> dict:
    """Bench distribution salted with the special cases the reference tests exercise: start states outside the
    limits (brake pre-trajectories), targets at rest, zero displacement, tiny and huge limits, asymmetric minimum
    limits, disabled DoFs."""
    f = bench_inputs(n, dof, seed=seed)
    rng = np.random.Generator(np.random.PCG64(46 + 1000 * seed))
    shape = (dof, n)
    r = rng.random(shape)
    # start beyond the limits -> brake
    m = r < 0.10
    f["current_velocity"][m] = (f["max_velocity"] * rng.uniform(1.0, 1.6, shape) * rng.choice([-1.0, 1.0], shape))[m]
    m = (r >= 0.10) & (r < 0.20)
    f["current_acceleration"][m] = (f["max_acceleration"] * rng.uniform(1.0, 1.6, shape) * rng.choice([-1.0, 1.0], shape))[m]
    # rest-to-rest
    m = (r >= 0.20) & (r < 0.30)
    for k in ("current_velocity", "current_acceleration", "target_velocity", "target_acceleration"):
        f[k][m] = 0.0
    # zero displacement
    m = (r >= 0.30) & (r < 0.33)
    f["target_position"][m] = f["current_position"][m]
    # tiny / huge limits
    m = (r >= 0.33) & (r < 0.38)
    f["max_jerk"][m] *= 1e-3
    m = (r >= 0.38) & (r < 0.43)
    f["max_jerk"][m] *= 1e3
    m = (r >= 0.43) & (r < 0.46)
    f["max_velocity"][m] = (0.01 + np.abs(f["target_velocity"]))[m]
    f["min_velocity"] = np.ascontiguousarray(-f["max_velocity"] * rng.uniform(0.3, 1.0, shape))
    f["min_acceleration"] = np.ascontiguousarray(-f["max_acceleration"] * rng.uniform(0.3, 1.0, shape))
    # keep the target state inside the asymmetric limits most of the time
    f["target_velocity"] = np.clip(f["target_velocity"], f["min_velocity"] * 0.98, None)
    f["target_acceleration"] = np.clip(f["target_acceleration"], f["min_acceleration"] * 0.98, None)
    f["enabled"] = np.ascontiguousarray((rng.random(shape) > 0.05).astype(np.uint8))
    return {k: np.ascontiguousarray(v) for k, v in f.items()}


_REFSTREAM = None


def _refstream():
    """librefstream.so (csrc/refstream.c, host-only C, built by __graft_entry__.build())."""
    global _REFSTREAM
    if _REFSTREAM is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "librefstream.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path)
        vp = C.c_void_p
        lib.refstream_pcg_new_next.argtypes = [C.c_uint64, C.c_uint64, vp, C.c_int32]
        lib.refstream_pcg_new_next.restype = C.c_uint64
        lib.refstream_pcg_from_seed_next.argtypes = [vp, vp, C.c_int32]
        lib.refstream_pcg_seed_from_u64_next.argtypes = [C.c_uint64, vp, C.c_int32]
        lib.refstream_ziggurat_tables.argtypes = [vp, vp]
        lib.refstream_normal.argtypes = [C.c_uint64, C.c_double, C.c_double, vp, C.c_int64]
        lib.refstream_uniform.argtypes = [C.c_uint64, C.c_double, C.c_double, vp, C.c_int64]
        lib.refstream_bench_inputs.argtypes = [C.c_int64, C.c_int32, vp]
        _REFSTREAM = lib
    return _REFSTREAM


def reference_bench_inputs(n: int, dof: int = 7) -> dict:
    """The first n problems of the reference benchmark's own input stream (bench/src/benchmark_target.rs:115-153): three
    `Pcg64Mcg::seed_from_u64` streams (42 positions, 43 velocities / accelerations, 44 limits), `Normal` by the 256-layer
    ziggurat, `Uniform` by the 52-bit mantissa method, drawn per problem in the reference's order. Problems the reference skips
    (validate_input(false, false) fails) are kept: they consumed their draws there too and fail validation on every side."""
    import ctypes as C
    lib = _refstream()
    arrays = [np.zeros((dof, n), dtype=np.float64) for _ in FIELDS]
    ptrs = (C.c_void_p * len(FIELDS))(*[a.ctypes.data for a in arrays])
    lib.refstream_bench_inputs(n, dof, ptrs)
    return dict(zip(FIELDS, arrays))
