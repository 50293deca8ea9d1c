#!/bin/bash
# One GPU iteration: parity tests, a short bench, and the per-launch ncu metrics of one step (run under gpurun).
set -o pipefail
TAG=${1:-iter}
M=smsp__average_warps_issue_stalled
python -m pytest tests/test_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --n 1048576 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain.log 2>gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,${M}_no_instruction_per_issue_active.ratio,${M}_long_scoreboard_per_issue_active.ratio,${M}_wait_per_issue_active.ratio,${M}_short_scoreboard_per_issue_active.ratio,${M}_math_pipe_throttle_per_issue_active.ratio,${M}_lg_throttle_per_issue_active.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__inst_executed.sum \
    --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv -s 90 python bench.py --n 1048576 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-140
python tools/launch_table.py gpurun_out/launches_${TAG}.csv 44
