#!/bin/bash
# Per-kernel durations of one chunk for several library variants (run under gpurun): tools/ncu_times.sh v1 v2 ...
for V in "$@"; do
  export RK_B200_LIB=$PWD/variants/test_$V.so
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_${V}.csv python bench.py --n 599186 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
done
python - "$@" <<'PY'
import csv, sys
from collections import OrderedDict
cols = {}
for v in sys.argv[1:]:
    rows = list(csv.reader(open(f"gpurun_out/t_{v}.csv")))
    hdr = [r for r in rows if "Kernel Name" in r][0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows if len(r) == len(hdr) and r[0].isdigit() and any(t in r[ki] for t in ("k1_", "k2_", "k_sync", "k_expand"))]
    n = len(seq) // 2  # warm-up step + timed step
    cols[v] = seq[-n:]
names = [k for k, _ in cols[sys.argv[1]]]
print("kernel".ljust(46), " ".join(v.rjust(9) for v in cols))
tot = {v: 0.0 for v in cols}
for i, k in enumerate(names):
    line = k[:46].ljust(46)
    for v in cols:
        t = cols[v][i][1] / 1e3 if i < len(cols[v]) else float("nan")
        tot[v] += t
        line += f" {t:9.1f}"
    print(line)
print("total (us)".ljust(46), " ".join(f"{tot[v]:9.1f}" for v in cols))
PY
