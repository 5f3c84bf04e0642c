#!/bin/bash
# ncu --set full of the WIDE kernel on config 4 + compute-sanitizer memcheck / synccheck of the WIDE worker
cd "$(dirname "$0")/.."
TAG=${1:-wp1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 1 -c 1 -f -o $OUT/prof_cfg4 python bench.py --workload cfg4_10kbp_in_12kbp_e5_semiglobal --pairs 296 --steps 1 --warmup 1 --only-headline --no-cpu-baseline > $OUT/ncu_full_cfg4.log 2>&1
tail -3 $OUT/ncu_full_cfg4.log
for tool in memcheck synccheck; do
timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_wide.py > $OUT/san_$tool.log 2>&1; echo "exit $?" >> $OUT/san_$tool.log; grep -c "ok:" $OUT/san_$tool.log; grep "ERROR SUMMARY\|exit" $OUT/san_$tool.log
done
ls -la $OUT
