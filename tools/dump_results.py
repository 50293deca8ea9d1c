"""Write a batch input file and this library's results for it (needs a B200), or with --oracle the CPU restatement's results:
    dump_results.py [--oracle] [--reference-stream] <n> <dof> <inputs.rkb2> <results.rkb2>
--reference-stream: the inputs are the first n problems of the reference benchmark's own Pcg64Mcg / ziggurat stream (seeds
42 / 43 / 44, workloads.reference_bench_inputs) — e.g. `dump_results.py --reference-stream 10000000 7 bench_ref_7dof.rkb2 ours.rkb2`,
then `cargo run --release -- bench_ref_7dof.rkb2 theirs.rkb2` in tools/rsruckig_dump and `tools/compare_results.py ours theirs`."""
import sys

import numpy as np

sys.path.insert(0, ".")
from rsruckig_b200 import _cabi as cabi, workloads  # noqa: E402
from rsruckig_b200.batchfile import write_inputs, write_results  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n, dof = int(args[0]), int(args[1])
f = workloads.reference_bench_inputs(n, dof) if "--reference-stream" in sys.argv else workloads.bench_inputs(n, dof, seed=0)
write_inputs(args[2], f, n=n, dof=dof, delta_time=0.005)
if "--oracle" in sys.argv:
    from oracle.oracle import OracleBatch
    b = OracleBatch(n, dof, "glibc")
    b.calculate(f, delta_time=0.005, threads=8)
else:
    from rsruckig_b200 import BatchRuckig
    b = BatchRuckig(dof, 0.005)
    b.calculate(f, n=n, memory_space=cabi.RK_MEM_HOST)
write_results(args[3], status=b.read(cabi.RK_FIELD_STATUS), duration=b.read(cabi.RK_FIELD_DURATION), codes=b.read(cabi.RK_FIELD_CODES),
              t=b.read(cabi.RK_FIELD_T))
print(f"wrote {args[2]} and {args[3]}")
