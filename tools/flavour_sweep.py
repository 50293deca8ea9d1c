#!/usr/bin/env python3
"""Oracle (portable libm restatement, what the CUDA path reproduces bit for bit) against oracle (glibc libm, what a Linux build
of the Rust reference links) on N random 7-DoF bench inputs: north_star's bar — result codes, profile codes exact, durations
within 1e-9 relative — applied between the two math flavours. CPU only.

    python tools/flavour_sweep.py [millions=10]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleBatch  # noqa: E402
from rsruckig_b200 import _cabi as cabi  # noqa: E402
from rsruckig_b200 import workloads  # noqa: E402

millions = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = 500_000
threads = os.cpu_count() or 1
total = status_mismatch = code_mismatch = 0
worst = 0.0
t0 = time.time()
for seed in range(100, 100 + 2 * millions):
    f = workloads.bench_inputs(n, 7, seed=seed)
    res = []
    for flavour in ("portable", "glibc"):
        ob = OracleBatch(n, 7, flavour)
        ob.calculate(f, threads=threads)
        res.append((ob.read(cabi.RK_FIELD_STATUS), ob.read(cabi.RK_FIELD_DURATION), ob.read(cabi.RK_FIELD_CODES), ob.read(cabi.RK_FIELD_LIMITING_DOF)))
        del ob
    (s0, d0, c0, l0), (s1, d1, c1, l1) = res
    ok = (s0 == 0) & (s1 == 0)
    status_mismatch += int((s0 != s1).sum())
    code_mismatch += int(((c0 != c1) & ok[None, :]).sum()) + int(((l0 != l1) & ok).sum())
    rel = np.abs(d0[ok] - d1[ok]) / np.maximum(np.abs(d1[ok]), 1e-300)
    worst = max(worst, float(rel.max()))
    total += n
    print(f"seed {seed}: {total} trajectories, status mismatches {status_mismatch}, code mismatches {code_mismatch}, "
          f"max relative duration difference {worst:.3e}, {time.time() - t0:.0f} s", flush=True)
assert status_mismatch == 0 and code_mismatch == 0 and worst <= 1e-9
print(f"FLAVOUR SWEEP OK: {total} trajectories, 0 status / code mismatches, max relative duration difference {worst:.3e}")
