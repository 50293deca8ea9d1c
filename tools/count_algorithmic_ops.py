#!/usr/bin/env python3
"""Algorithmic FP64 work per trajectory, counted on the reference algorithm itself.

Runs the op-counting build of the oracle (oracle/rko_count.cpp: the restatement of rsruckig compiled with `double`
replaced by a counting scalar; add/sub/mul/div/sqrt and each libm call = 1 operation, compares / abs / min / max /
negation = 0, no FMA) over a workload, checks that it takes exactly the branches of the plain glibc build (status
and duration bit-equal) and writes the per-trajectory averages to profiles/algorithmic_ops_oracle.json.

CPU only; test / measurement infrastructure (never imported by the product).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from rsruckig_b200 import _cabi as cabi  # noqa: E402
from rsruckig_b200 import workloads  # noqa: E402


def load_count():
    oracle.build()
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_count.so"))
    lib.rko_count_calculate.argtypes = [C.POINTER(cabi.RkBatchDesc), C.POINTER(cabi.RkBatchIn), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rko_count_calculate.restype = None
    lib.rko_count_phases.restype = C.c_int32
    return lib


# oracle/rko_math.hpp, enum RKO_PH_*
PHASES = ["other (validation, brake, block, synchronisation, dispatch)", "step1 quartic families NONE/ACC0/ACC1 (without check and root solver)",
          "step1 ACC0_ACC1", "step1 VEL", "step1 rescues", "step2 ACC0_ACC1_VEL", "step2 VEL (without Newton and check)", "step2 ACC0_VEL", "step2 ACC1_VEL",
          "step2 ACC0_ACC1 / ACC1 / ACC0 / NONE", "Profile::check (all callers)", "solve_quart_monic / solve_cub (all callers)", "shrink_interval (all callers)"]


def count(fields, n, dof, **kw):
    lib = load_count()
    desc, keep = cabi.make_desc(n, dof, delta_time=0.005, memory_space=cabi.RK_MEM_HOST, **kw)
    bin_ = cabi.make_in(fields)
    n_ph = lib.rko_count_phases()
    assert n_ph == len(PHASES)
    ops = np.zeros((n_ph, 5), dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    duration = np.zeros(n)
    lib.rko_count_calculate(C.byref(desc), C.byref(bin_), ops.ctypes.data, status.ctypes.data, duration.ctypes.data)
    # same branches as the plain build: fresh-object semantics on both sides
    ob = oracle.OracleBatch(n, dof, flavour="glibc")
    ob.calculate(fields, threads=os.cpu_count() or 1, **kw)
    s2, d2 = ob.read(cabi.RK_FIELD_STATUS), ob.read(cabi.RK_FIELD_DURATION)
    ok = status >= 0
    assert np.array_equal(status, s2), "counting build took a different branch (status)"
    assert np.array_equal(duration[ok].view(np.uint64), d2[ok].view(np.uint64)), "counting build took a different branch (duration)"
    names = ["add_sub", "mul", "div", "sqrt", "libm"]
    per = {k: float(v) / n for k, v in zip(names, ops.sum(axis=0))}
    per["total"] = float(ops.sum()) / n
    per["by_phase"] = {ph: round(float(ops[i].sum()) / n, 2) for i, ph in enumerate(PHASES)}
    return per, status


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=262144)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "algorithmic_ops_oracle.json"))
    a = ap.parse_args()
    res = {"definition": "FP64 operations executed by the reference algorithm (oracle restatement with a counting scalar): "
                         "add/sub/mul/div/sqrt = 1 each, libm call (cbrt, sin, cos, acos, atan2, fmod) = 1, compares/abs/min/max/negation = 0, no FMA; "
                         "per trajectory, averaged over the workload, Ruckig::calculate only (fresh calculator per problem)",
           "n": a.n, "seed": a.seed, "workloads": {}}
    f7 = workloads.bench_inputs(a.n, 7, seed=a.seed)
    per, st = count(f7, a.n, 7)
    per["working_fraction"] = float((st >= 0).mean())
    res["workloads"]["bench_7dof_position_time"] = per
    f6 = workloads.config2_inputs(a.n, seed=a.seed)
    per, st = count(f6, a.n, 6)
    per["working_fraction"] = float((st >= 0).mean())
    res["workloads"]["config2_6dof_position_time"] = per
    # one DoF has no synchronisation partner: no Step 2, so this is validation + brake + Step 1 + block of one DoF
    f1 = workloads.bench_inputs(a.n, 1, seed=a.seed)
    per, st = count(f1, a.n, 1)
    per["working_fraction"] = float((st >= 0).mean())
    res["workloads"]["bench_1dof_position (Step 1 only)"] = per
    t7 = res["workloads"]["bench_7dof_position_time"]["total"]
    res["breakdown_7dof"] = {"step1_and_setup": 7 * per["total"], "synchronisation_and_step2": t7 - 7 * per["total"],
                             "note": "7 x the single-DoF count (validation, brake, Step 1, block) vs the rest (synchronize, Step 2 of the six non-limiting DoFs, "
                                     "profile checks of the re-timed profiles); per 7-DoF trajectory of the bench distribution"}
    res["flop_per_trajectory_7dof"] = t7
    print(json.dumps(res, indent=1))
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)
        fh.write("\n")


if __name__ == "__main__":
    main()
