#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-wf1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 2 -c 1 -f -o $OUT/prof_cfg4 python bench.py --workload cfg4_10kbp_in_12kbp_e5_semiglobal --pairs 296 --steps 1 --warmup 1 --only-headline --no-cpu-baseline > $OUT/ncu_full_cfg4.log 2>&1
tail -2 $OUT/ncu_full_cfg4.log | cut -c1-300
