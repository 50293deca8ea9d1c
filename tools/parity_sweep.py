"""Wide parity sweep (run under gpurun): CUDA path vs the portable oracle, every stored double compared, over many seeds of
the bench distribution and the other workload generators. Complements tests/test_parity.py (fixed seeds, smaller)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from rsruckig_b200 import BatchRuckig, workloads  # noqa: E402
from test_parity import assert_exact, run_both  # noqa: E402

total = 0
t0 = time.time()
for seed in range(10, 10 + int(sys.argv[1]) if len(sys.argv) > 1 else 18):
    n = 300_000
    f = workloads.bench_inputs(n, 7, seed=seed)
    otg, ob = run_both(BatchRuckig, f, n, 7, threads=16)
    total += assert_exact(otg, ob, label=f"7dof seed {seed}")
    del otg, ob
    n = 200_000
    f = workloads.config2_inputs(n, seed=seed)
    otg, ob = run_both(BatchRuckig, f, n, 6, threads=16)
    total += assert_exact(otg, ob, label=f"config2 seed {seed}")
    del otg, ob
    for dof in (1, 3):
        f = workloads.bench_inputs(n, dof, seed=seed)
        otg, ob = run_both(BatchRuckig, f, n, dof, threads=16)
        total += assert_exact(otg, ob, label=f"{dof}dof seed {seed}")
        del otg, ob
    f = workloads.edge_case_inputs(50_000, 4, seed=seed)
    otg, ob = run_both(BatchRuckig, f, 50_000, 4, threads=16)
    total += assert_exact(otg, ob, label=f"edge seed {seed}")
    del otg, ob
    print(f"seed {seed} ok, {total} Working trajectories compared so far, {time.time() - t0:.0f} s", flush=True)
print(f"PARITY SWEEP OK: {total} trajectories bit-identical (status, duration, limiting DoF, codes, all profile arrays)")
