#!/bin/bash
# FP64 operations, thread instructions and DRAM bytes per 7-DoF trajectory of one calculate step (run under gpurun).
# Writes gpurun_out/ops_<tag>.csv (ncu launch list) and gpurun_out/algorithmic_ops.json.
TAG=${1:-r2}
N=262144
python bench.py --n $N --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 || exit 1
ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none --csv --log-file gpurun_out/ops_${TAG}.csv python bench.py --n $N --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
python - $TAG $N <<'PY'
import csv, json, sys
tag, n = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(f"gpurun_out/ops_{tag}.csv")))
hdr = [r for r in rows if "Kernel Name" in r][0]
ki, vi, mi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Metric Unit"))
mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "": 1, "inst": 1, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1}
launches = {}
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit() and any(t in r[ki] for t in ("k1_", "k2_", "k_sync", "k_expand")):
        launches.setdefault(int(r[ii]), {"name": r[ki]})[r[mi]] = float(r[vi].replace(",", "")) * mult.get(r[ui], 1)
ids = sorted(launches)
n_steps = 4                       # bench.py runs at least 3 warm-up steps + the timed step, identical launch sequences
per_step = len(ids) // n_steps    # (a batch of this size is computed as two halves on two streams: both are in the step)
step = [launches[i] for i in ids[-per_step:]]
s = lambda m: sum(l.get(m, 0.0) for l in step)
dadd, dmul, dfma = (s(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum") for o in ("dadd", "dmul", "dfma"))
out = {"flop_per_trajectory_7dof": (dadd + dmul + 2 * dfma) / n, "dadd": dadd / n, "dmul": dmul / n, "dfma": dfma / n,
       "fp64_instructions": (dadd + dmul + dfma) / n, "thread_instructions": s("smsp__thread_inst_executed.sum") / n,
       "dram_bytes_per_trajectory": (s("dram__bytes_read.sum") + s("dram__bytes_write.sum")) / n, "kernel_launches_per_step": per_step,
       "sum_of_kernel_durations_us": s("gpu__time_duration.sum") / 1e3,
       "how": f"FP64 operations executed per 7-DoF trajectory by one calculate step (all {per_step} kernel launches), counted by ncu (smsp__sass_thread_inst_executed_op_{{dadd,dmul,dfma}}_pred_on.sum, flop = dadd + dmul + 2*dfma) on {n} trajectories of the bench distribution. The kernels run the reference's operation sequence without FMA contraction, so dadd + dmul are the algorithm's additions and multiplications (plus what the kernel pipeline recomputes when an item is picked up again by a later kernel); the dfma are the Newton iterations inside divisions and square roots. The algorithm's own count (op-counting build of the oracle) is profiles/algorithmic_ops_oracle.json."}
json.dump(out, open("gpurun_out/algorithmic_ops.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "how"}))
PY
