#!/bin/bash
# The variants prepared at the end of round 1, ready for one A/B call:
#   tools/prepared_variants.sh build                      here (nvcc cross-compiles): variants/test_{cur,pool1,qp6,qp7,qp6k4,ccb,all}.so
#   gpurun --timeout 600 -- 'tools/prepared_variants.sh run'   on a B200: parity of each variant, then throughput and latency
set -e
cd "$(dirname "$0")/.."
VARIANTS="cur pool1 qp6 qp7 qp6k4 ccb all"
case "${1:-}" in
build)
    tools/build_variant.sh cur &
    tools/build_variant.sh pool1 -DRK_UDUD_WARP_POOL=1 &
    tools/build_variant.sh qp6 -DRK_QUARTIC_POOL=6 &
    tools/build_variant.sh qp7 -DRK_QUARTIC_POOL=7 &
    wait
    tools/build_variant.sh qp6k4 -DRK_QUARTIC_POOL=6 -DRK_QP_ITEMS=4 &
    tools/build_variant.sh ccb -DRK_CC_BARRIER=1 &
    tools/build_variant.sh all -DRK_UDUD_WARP_POOL=1 -DRK_QUARTIC_POOL=6 -DRK_CC_BARRIER=1 &
    wait
    ;;
run)
    for v in $VARIANTS; do
        echo -n "parity $v: "
        RK_B200_LIB=$PWD/variants/test_$v.so python tools/quick_parity.py 2>&1 | tail -1
    done
    N=4194304 tools/ab.sh $VARIANTS cur
    for v in $VARIANTS; do
        echo -n "latency $v: "
        RK_B200_LIB=$PWD/variants/test_$v.so python tools/latency.py 65536 2>&1 | tail -1
    done
    echo "a variant that wins: make it the default, then the full parity suite (pytest -m gpu) and tools/parity_sweep.py"
    ;;
*)
    echo "usage: $0 build|run"; exit 1 ;;
esac
