#!/bin/bash
# One GPU iteration: parity tests, a short bench, and the per-launch ncu metrics of one step (run under gpurun).
set -o pipefail
TAG=${1:-iter}
python -m pytest tests/test_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --n 1048576 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain.log 2>gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio \
    --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv -s 90 python bench.py --n 1048576 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-140
python tools/launch_table.py gpurun_out/launches_${TAG}.csv 30
