"""Print the per-kernel table of one bench step from an `ncu --csv` launch list (time share, active lanes, icache stalls)."""
import csv
import sys

path = sys.argv[1]
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 27
rows = list(csv.reader(open(path)))
hdr = [r for r in rows if "Kernel Name" in r][0]
ki, vi, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
d = {}
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit() and "rk::" in r[ki]:
        d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(d)[-n_last:]
T = "gpu__time_duration.sum"
tot = sum(d[k][T] for k in ids)
for k in ids:
    m = d[k]
    lanes = m.get("smsp__thread_inst_executed_per_inst_executed.ratio", float("nan"))
    noinst = m.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", float("nan"))
    print(f"{m[T] / 1e6:9.3f} ms {100 * m[T] / tot:5.1f}%  lanes {lanes:5.1f}  no_inst {noinst:5.2f}  {k[1][:48]}")
print(f"total {tot / 1e6:.3f} ms")
