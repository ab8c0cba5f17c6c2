#!/bin/bash
# tools/tensor_tune.sh build | run -- kernel-variant sweep of the tcgen05 screen (csrc/tensor_search.cu).
#   build : compile one libnns_b200_<name>.so per variant into nns-cuda_b200/lib/exp/ (no GPU needed)
#   run   : on a B200, bench every variant on the workloads in $WORKLOADS and print one line each
# Variants are "name:DEF=VAL,DEF=VAL..." (the NNS_T_* knobs at the top of tensor_search.cu).
set -e
cd "$(dirname "$0")/../nns-cuda_b200"
VARIANTS=${VARIANTS:-"ts1:NNS_T_TS=1 ts0:NNS_T_TS=0"}
WORKLOADS=${WORKLOADS:-"c2 c3 c4"}
if [ "$1" = build ]; then
    make >/dev/null
    mkdir -p build/exp lib/exp
    for v in $VARIANTS; do
        name=${v%%:*}; defs=${v#*:}
        dflags=$(echo "$defs" | tr ',' '\n' | sed 's/^/-D/' | tr '\n' ' ')
        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 \
            $dflags $EXTRA_DEFS -c -o build/exp/tensor_search_$name.o csrc/tensor_search.cu &
    done
    wait
    for v in $VARIANTS; do
        name=${v%%:*}
        objs=$(ls build/*.o | grep -v tensor_search.o)
        nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/exp/libnns_b200_$name.so $objs build/exp/tensor_search_$name.o -lpthread
    done
    ls -la lib/exp
else
    cd ..
    for v in $VARIANTS; do
        name=${v%%:*}
        for wl in $WORKLOADS; do
            NNS_B200_LIB=$PWD/nns-cuda_b200/lib/exp/libnns_b200_$name.so python bench.py --workload $wl --flags 8 --steps ${STEPS:-3} \
                --no-cpu-baseline --no-e2e --also none 2>&1 | python -c "
import sys, json
line = sys.stdin.read().strip().splitlines()[-1]
try:
    d = json.loads(line)
    print('$name $wl %.3e pairs/s %.3f ms sm_mhz=%s %s cand=%s' % (d['value'], d['ms_per_step'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['roofline'].get('tensor_stats')))
except Exception as e:
    print('$name $wl FAILED', line[-300:])
"
        done
    done
fi
