"""Print the per-kernel table of one bench step from an `ncu --csv` launch list (time share, active lanes, issue rate and
the main stall reasons in warp-cycles per issued instruction, DRAM and local-memory traffic)."""
import csv
import sys

path = sys.argv[1]
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
hdr = [r for r in rows if "Kernel Name" in r][0]
ki, vi, mi, ii, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID"), hdr.index("Metric Unit")
d = {}
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit() and any(t in r[ki] for t in ("k1_", "k2_", "k_sync", "k_expand", "k_at_time", "k_validate")):
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if u in ("Gbyte", "GB"): v *= 1e9
        elif u in ("Mbyte", "MB"): v *= 1e6
        elif u in ("Kbyte", "KB"): v *= 1e3
        elif u in ("ms", "msecond"): v *= 1e6
        elif u in ("us", "usecond"): v *= 1e3
        elif u in ("s", "second"): v *= 1e9
        d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = v
ids = sorted(d)[-n_last:]
T = "gpu__time_duration.sum"
S = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
tot = sum(d[k][T] for k in ids)
nan = float("nan")
print("      ms     %  lanes issue% warps%  no_inst  long_sb  wait  short  math  lg_thr  barr  brnch  fp64%   dramGB  localGB  Minst  kernel")
for k in ids:
    m = d[k]
    g = lambda n: m.get(n, nan)
    local = (g("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") + g("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum")) * 32 / 1e9
    dram = (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / 1e9
    print(f"{m[T] / 1e6:8.3f} {100 * m[T] / tot:5.1f}  {g('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} "
          f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} {g('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f}   "
          f"{g(S % 'no_instruction'):6.2f}  {g(S % 'long_scoreboard'):6.2f} {g(S % 'wait'):6.2f} {g(S % 'short_scoreboard'):5.2f} "
          f"{g(S % 'math_pipe_throttle'):5.2f}  {g(S % 'lg_throttle'):5.2f} {g(S % 'barrier'):5.2f}  {g(S % 'branch_resolving'):5.2f}  {g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):5.1f}  "
          f"{dram:7.3f}  {local:7.3f} {g('smsp__inst_executed.sum') / 1e6:6.1f}  {k[1][:44]}")
print(f"total {tot / 1e6:.3f} ms")
