#!/bin/bash
# Stall/issue metrics of selected kernels for one library variant (run under gpurun): tools/ncu_kernels.sh <variant> <kernel-regex> <count> [skip]
V=$1; K=$2; C=${3:-4}; SKIP=${4:-0}
M=smsp__average_warps_issue_stalled
export RK_B200_LIB=$PWD/variants/test_$V.so
python bench.py --n 599186 --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain.log 2>gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,${M}_no_instruction_per_issue_active.ratio,${M}_long_scoreboard_per_issue_active.ratio,${M}_wait_per_issue_active.ratio,${M}_short_scoreboard_per_issue_active.ratio,${M}_math_pipe_throttle_per_issue_active.ratio,${M}_lg_throttle_per_issue_active.ratio,${M}_barrier_per_issue_active.ratio,${M}_branch_resolving_per_issue_active.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__inst_executed.sum \
    --clock-control none --csv --log-file gpurun_out/k_${V}.csv -k regex:"$K" -s $SKIP -c $C python bench.py --n 599186 --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
python tools/launch_table.py gpurun_out/k_${V}.csv $C
