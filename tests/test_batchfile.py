"""The batch input / results files: round trips, and the comparison at north_star's bar applied to the
oracle's two math flavours (the stand-in, on this machine, for "ours against the real rsruckig")."""
import numpy as np

from rsruckig_b200 import _cabi as cabi
from rsruckig_b200 import workloads
from rsruckig_b200.batchfile import compare_results, read_inputs, read_results, write_inputs, write_results


def test_input_file_round_trip(tmp_path):
    n, dof = 257, 5
    f = workloads.edge_case_inputs(n, dof, seed=2)
    f["minimum_duration"] = np.where(np.arange(n) % 3 == 0, 1.5, np.nan)
    path = str(tmp_path / "in.rkb2")
    write_inputs(path, f, n=n, dof=dof, delta_time=0.004, control_interface=1, synchronization=2, duration_discretization=1)
    g, meta = read_inputs(path)
    assert meta == dict(n=n, dof=dof, delta_time=0.004, control_interface=1, synchronization=2, duration_discretization=1)
    assert set(g) == set(f)
    for k in f:
        assert np.array_equal(g[k], f[k], equal_nan=True), k
        assert g[k].dtype == (np.uint8 if k == "enabled" else np.float64)


def test_results_file_and_comparison(tmp_path):
    from oracle.oracle import OracleBatch
    n, dof = 3000, 7
    f = workloads.bench_inputs(n, dof, seed=3)
    files = {}
    for flavour in ("glibc", "portable"):
        b = OracleBatch(n, dof, flavour)
        b.calculate(f, delta_time=0.005, threads=4)
        path = str(tmp_path / f"{flavour}.rkb2")
        write_results(path, status=b.read(cabi.RK_FIELD_STATUS), duration=b.read(cabi.RK_FIELD_DURATION),
                      codes=b.read(cabi.RK_FIELD_CODES), t=b.read(cabi.RK_FIELD_T))
        files[flavour] = read_results(path)
        assert np.array_equal(files[flavour]["duration"], b.read(cabi.RK_FIELD_DURATION))
        assert np.array_equal(files[flavour]["t"], b.read(cabi.RK_FIELD_T))
    r = compare_results(files["glibc"], files["portable"])
    assert r["pass"] and r["status_mismatches"] == 0 and r["code_mismatches"] == 0 and r["max_rel_duration_diff"] <= 1e-9, r
    # a flipped code or a perturbed duration must be caught
    broken = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in files["portable"].items()}
    ok = np.nonzero(broken["status"] == 0)[0]
    broken["codes"][0, ok[0]] ^= 1
    assert not compare_results(files["glibc"], broken)["pass"]
    broken = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in files["portable"].items()}
    broken["duration"][ok[1]] *= 1.0 + 1e-6
    assert not compare_results(files["glibc"], broken)["pass"]
