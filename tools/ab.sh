#!/bin/bash
# A/B of library variants built into variants/test_<name>.so (run under gpurun): tools/ab.sh name1 name2 ...
for v in "$@"; do
  echo -n "variant $v: "
  RK_B200_LIB=$PWD/variants/test_$v.so python bench.py --n ${N:-2097152} --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'M traj/s', round(d['ms_per_step'],2), 'ms')"
done
