"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line: where do the samples go, and
with how many active lanes. Usage: src_hot.py dump.csv [kernel-substring] [top-n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
cur_file, cur_fn, hdr = None, None, None
agg = defaultdict(lambda: [0, 0, 0, ""])  # samples, inst, thread_inst, text
per_file = defaultdict(lambda: [0, 0, 0])
total = [0, 0, 0]
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or want not in (cur_fn or ""):
        continue
    if r[0].isdigit():  # a source line row
        try:
            s, i, t = int(r[iS]), int(r[iI]), int(r[iT])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        a = agg[key]
        a[0] += s; a[1] += i; a[2] += t; a[3] = r[1][:90]
        f = per_file[cur_file]
        f[0] += s; f[1] += i; f[2] += t
        total[0] += s; total[1] += i; total[2] += t
print(f"kernel filter '{want}': samples {total[0]}, warp-inst {total[1]}, lanes {total[2] / max(1, total[1]):.1f}")
for f, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:22s} samples {100 * v[0] / total[0]:5.1f}%  inst {100 * v[1] / total[1]:5.1f}%  lanes {v[2] / max(1, v[1]):5.1f}")
print("top lines by samples:")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"  {100 * v[0] / total[0]:5.1f}%  inst {100 * v[1] / total[1]:5.1f}%  lanes {v[2] / max(1, v[1]):5.1f}  {f}:{ln}  {v[3]}")
