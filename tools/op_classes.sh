#!/bin/bash
# Executed thread instructions per 7-DoF trajectory by instruction class and kernel (run under gpurun):
# ncu smsp__sass_thread_inst_executed_op_<class>_pred_on over all launches of one 262 144-trajectory step.
TAG=${1:-r2}
N=262144
M=smsp__sass_thread_inst_executed_op
python bench.py --n $N --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 || exit 1
ncu --metrics ${M}_fp64_pred_on.sum,${M}_fp32_pred_on.sum,${M}_integer_pred_on.sum,${M}_bit_pred_on.sum,${M}_control_pred_on.sum,${M}_conversion_pred_on.sum,${M}_memory_pred_on.sum,${M}_misc_pred_on.sum,${M}_uniform_pred_on.sum,${M}_inter_thread_communication_pred_on.sum,smsp__thread_inst_executed.sum,smsp__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum \
    --clock-control none --csv --log-file gpurun_out/opclass_${TAG}.csv python bench.py --n $N --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
python - $TAG $N <<'PY'
import csv, sys, collections
tag, n = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(f"gpurun_out/opclass_{tag}.csv")))
hdr = [r for r in rows if "Kernel Name" in r][0]
ki, vi, mi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "ID"))
launches = {}
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit() and any(t in r[ki] for t in ("k1_", "k2_", "k_sync", "k_expand")):
        launches.setdefault(int(r[ii]), {"name": r[ki].split("(")[0].replace("void ", "").replace("rk::", "")})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(launches)
per_step = len(ids) // 4
step = [launches[i] for i in ids[-per_step:]]
classes = ["fp64", "fp32", "integer", "bit", "control", "conversion", "memory", "misc", "uniform", "inter_thread_communication"]
key = lambda c: f"smsp__sass_thread_inst_executed_op_{c}_pred_on.sum"
by = collections.OrderedDict()
for l in step:
    a = by.setdefault(l["name"], collections.Counter())
    for c in classes:
        a[c] += l.get(key(c), 0.0)
    a["total"] += l.get("smsp__thread_inst_executed.sum", 0.0)
    a["arith"] += l.get("smsp__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum", 0.0)
tot = collections.Counter()
for a in by.values():
    tot.update(a)
print(f"thread instructions per 7-DoF trajectory ({n} trajectories, {per_step} launches): total {tot['total'] / n:.0f}")
print("  " + "  ".join(f"{c} {tot[c] / n:.0f} ({100 * tot[c] / tot['total']:.1f} %)" for c in classes + ["arith"]))
print(f"{'kernel':28s} {'total':>8s} " + " ".join(f"{c[:7]:>8s}" for c in classes + ["arith"]))
for name, a in sorted(by.items(), key=lambda kv: -kv[1]["total"]):
    print(f"{name[:28]:28s} {a['total'] / n:8.0f} " + " ".join(f"{a[c] / n:8.0f}" for c in classes + ["arith"]))
PY
