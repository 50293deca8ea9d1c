"""Walk one kernel of an `ncu --page source --print-source cuda,sass --csv` dump in address order, in buckets of B SASS
instructions: executed warp-instructions, share, active lanes, stall samples and the dominant source lines of the bucket.
Usage: sass_map.py dump.csv kernel-substring [bucket=64]"""
import csv
import sys
from collections import Counter

path, want = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ins = {}
fn = fil = None
line = None
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        fil = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        continue
    if want not in (fn or ""):
        continue
    if r[0].isdigit():
        line = f"{fil}:{r[0]}"; continue
    if r[2].startswith("0x"):
        try:
            ins[int(r[2], 16)] = (int(r[7]), int(r[8]), int(r[6]), line, r[3].strip()[:40])
        except ValueError:
            pass
addrs = sorted(ins)
tot = sum(ins[a][0] for a in addrs)
tots = sum(ins[a][2] for a in addrs)
base = addrs[0]
print(f"{len(addrs)} instr, {tot/1e6:.1f}M warp-instr, {tots} samples")
for i in range(0, len(addrs), B):
    chunk = [ins[a] for a in addrs[i:i + B]]
    e = sum(c[0] for c in chunk); t = sum(c[1] for c in chunk); s = sum(c[2] for c in chunk)
    if e == 0:
        continue
    cnt = Counter()
    for c in chunk:
        cnt[c[3]] += c[0]
    top = ", ".join(f"{k}" for k, _ in cnt.most_common(3))
    print(f"+{(addrs[i]-base)//16:5d}  inst {100*e/tot:5.1f}%  smp {100*s/tots:5.1f}%  lanes {t/max(1,e):5.1f}  exec/instr {e/len(chunk)/1e3:8.1f}k  {top}")
