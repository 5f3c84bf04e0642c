#!/bin/bash
# ncu --set full captures of the SLIM kernel on config 3 (100 k pairs) and config 5 (1 250 pairs), source pages exported on the box
cd "$(dirname "$0")/.."
TAG=${1:-p1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
K=${2:-slim_kernel}
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_cfg3 python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 100000 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_cfg3.log 2>&1
if [ "$3" != "nocfg5" ]; then
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_cfg5 python bench.py --workload cfg5_100kbp_e15_global_adaptive --pairs 1250 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_cfg5.log 2>&1
fi
ls -la $OUT
