"""Write the bench distribution as a batch input file and this library's results for it (needs a B200), or with --oracle the
CPU restatement's results: dump_results.py [--oracle] <n> <dof> <inputs.rkb2> <results.rkb2>"""
import sys

import numpy as np

sys.path.insert(0, ".")
from rsruckig_b200 import _cabi as cabi, workloads  # noqa: E402
from rsruckig_b200.batchfile import write_inputs, write_results  # noqa: E402

args = [a for a in sys.argv[1:] if a != "--oracle"]
n, dof = int(args[0]), int(args[1])
f = workloads.bench_inputs(n, dof, seed=0)
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
