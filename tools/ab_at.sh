#!/bin/bash
# A/B of the at_time rollout for library variants (run under gpurun): tools/ab_at.sh name1 name2 ...
for v in "$@"; do
  echo -n "variant $v: "
  RK_B200_LIB=$PWD/variants/test_$v.so python bench.py --n 262144 --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); a=d['at_time']; print(round(a['ms_per_rollout'],3), 'ms', round(a['roofline']['achieved']), 'GB/s')"
done
