#!/bin/bash
# quick GPU pass for one workload: bench summary (+ optional ncu: "inst" = instruction counts only, "full" = --set full with source)
# usage: quick_cfg.sh <tag> <workload> <pairs> [inst|full]
cd "$(dirname "$0")/.."
TAG=$1; WL=$2; N=$3; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python bench.py --workload $WL --pairs $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err
python - <<PY
import json
d=json.load(open("$OUT/bench_$WL.json"))
print("$WL value %.4gM  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.4gM  launches %d  work %s" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["gpu_launches"], d["work"]))
PY
if [ "$4" == "full" ]; then
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 3 -c 1 -f -o $OUT/prof_$WL python bench.py --workload $WL --pairs $N --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_$WL.log 2>&1
fi
if [ "$4" == "inst" ]; then
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:align_kernel -s 3 -c 1 python bench.py --workload $WL --pairs $N --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -A8 "align_kernel" | grep "inst_executed\|duration\|issue_active"
fi
