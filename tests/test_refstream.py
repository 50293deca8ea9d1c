"""The reference benchmark's own input stream (bench/src/benchmark_target.rs:29-70, :115-153): Pcg64Mcg + rand_distr Normal
(ziggurat) + rand Uniform, restated in rsruckig_b200/csrc/refstream.c from the crates' published algorithms (the crates are
not vendored in the reference). Pinned on the known answers rand_pcg's own test-suite holds for this generator."""
import ctypes as C

import numpy as np

from rsruckig_b200 import workloads


def u64s(call, count):
    out = (C.c_uint64 * count)()
    call(out, count)
    return [int(x) for x in out]


def test_pcg64mcg_known_answers():
    lib = workloads._refstream()
    # rand_pcg tests/mcg128xsl64.rs: "Numbers copied from official test suite (C version)", Mcg128Xsl64::new(42)
    got = u64s(lambda o, c: lib.refstream_pcg_new_next(42, 0, o, c), 6)
    assert got == [0x63b4a3a813ce700a, 0x382954200617ab24, 0xa7fd85ae3fe950ce, 0xd715286aa2887737, 0x60c92fee2e59f32c, 0x84c4e96beff30017]
    # from_seed([1..16]) and seed_from_u64(0) (test_mcg128xsl64_construction)
    seed = (C.c_uint8 * 16)(*range(1, 17))
    assert u64s(lambda o, c: lib.refstream_pcg_from_seed_next(seed, o, c), 1) == [7071994460355047496]
    assert u64s(lambda o, c: lib.refstream_pcg_seed_from_u64_next(0, o, c), 1) == [6198063878555692194]


def test_ziggurat_tables():
    lib = workloads._refstream()
    x, f = np.zeros(257), np.zeros(257)
    lib.refstream_ziggurat_tables(x.ctypes.data, f.ctypes.data)
    # the first entries of rand_distr's ZIG_NORM_X / ZIG_NORM_F, ZIG_NORM_R = x[1]
    assert ["%.18f" % v for v in x[:4]] == ["3.910757959537090045", "3.654152885361008796", "3.449278298560964462", "3.320244733839166074"]
    assert ["%.18f" % v for v in f[:4]] == ["0.000477467764586655", "0.001260285930498598", "0.002609072746106363", "0.004037972593371872"]
    assert x[256] == 0.0 and f[256] == 1.0 and np.all(np.diff(x) < 0) and np.all(np.diff(f) > 0)
    # every layer has the same area V (the defining property of the tables), the base strip included
    area = x[1:256] * (f[2:257] - f[1:256])
    assert np.allclose(area, 0.00492867323399, rtol=2e-9)


def test_distributions_and_draw_order():
    lib = workloads._refstream()
    n = 400_000
    z = np.zeros(n)
    lib.refstream_normal(42, 0.0, 4.0, z.ctypes.data, n)
    assert abs(z.mean()) < 0.03 and abs(z.std() - 4.0) < 0.03 and abs(((z / 4.0) ** 4).mean() - 3.0) < 0.06  # N(0, 16), kurtosis 3
    assert (np.abs(z) > 4.0 * 3.654152885361009).sum() > 0  # the tail branch of the ziggurat is exercised
    u = np.zeros(n)
    lib.refstream_uniform(44, 0.1, 12.0, u.ctypes.data, n)
    assert u.min() >= 0.1 and u.max() < 12.0 and abs(u.mean() - 6.05) < 0.03
    # the problem stream: a prefix of a longer run is the shorter run (streams continue from problem to problem), the first
    # position of the first problem is the first draw of stream 42, max_jerk is the third uniform block of stream 44
    a, b = workloads.reference_bench_inputs(1000, 7), workloads.reference_bench_inputs(250, 7)
    for k in workloads.FIELDS:
        assert np.array_equal(a[k][:, :250], b[k])
    assert a["current_position"][0, 0] == z[0] and a["target_position"][0, 0] == z[7]
    assert a["max_jerk"][0, 0] == u[14] and a["max_velocity"][0, 0] == u[0] + abs(a["target_velocity"][0, 0])
    big = workloads.reference_bench_inputs(200_000, 7)
    for name, p in (("current_velocity", 0.9), ("current_acceleration", 0.8), ("target_velocity", 0.7), ("target_acceleration", 0.6)):
        assert abs((big[name] != 0.0).mean() - p) < 0.005, name
