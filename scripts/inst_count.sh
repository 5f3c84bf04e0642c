#!/bin/bash
# warp-instructions executed by the align kernel for a 200k-pair batch of the given workload
cd "$(dirname "$0")/.."
W=${1:-cfg2_150bp_e5_global}; P=${2:-200000}
WFACUDA_NO_PIPELINE=1 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none -k regex:align_kernel -s 1 -c 1 python bench.py --workload $W --pairs $P --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "align_kernel|inst_executed|time_duration|issue_active|warps_active|grid_size"
