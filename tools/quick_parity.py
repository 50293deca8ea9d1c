"""Smallest useful GPU check (about 2 s of box time): 60 000 7-DoF bench problems, every stored double against the portable
oracle. RK_B200_LIB selects a library variant (tools/build_variant.sh). Run under gpurun from the repository root."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
t0 = time.time()
from rsruckig_b200 import BatchRuckig, workloads
from test_parity import assert_exact, run_both
n = 60000
f = workloads.bench_inputs(n, 7, seed=5)
otg, ob = run_both(BatchRuckig, f, n, 7, threads=16)
print("compared", assert_exact(otg, ob, label="pool"), "in", round(time.time() - t0, 1), "s", flush=True)
