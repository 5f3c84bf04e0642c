#!/bin/bash
# Final captures of the round: launch lists + traffic / instruction counters of the headline kernels (cheap metric set,
# every launch of one short bench run), and one `--set full` report per worker class (LANE stages, SLIM on configs 3 / 5, WIDE).
cd "$(dirname "$0")/.."
TAG=${1:-c2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio
run() { # name workload pairs
  WFACUDA_NO_PIPELINE=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_$1.csv python bench.py --workload $2 --pairs $3 --steps 1 --warmup 3 --only-headline --no-cpu-baseline > $OUT/ncu_$1.log 2>&1
}
run cfg2 cfg2_150bp_e5_global 1000000
run cfg3 cfg3_1kbp_e10_global_adaptive 1000000
run cfg5 cfg5_100kbp_e15_global_adaptive 1250
run cfg4 cfg4_10kbp_in_12kbp_e5_semiglobal 296
if [ "$2" == "full" ]; then
full() { # name regex skip count workload pairs
  WFACUDA_NO_PIPELINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/prof_$1 python bench.py --workload $5 --pairs $6 --steps 1 --warmup 3 --only-headline --no-cpu-baseline > $OUT/ncu_full_$1.log 2>&1
}
full cfg2 lane_ 12 4 cfg2_150bp_e5_global 1000000
full cfg3 slim_kernel 6 1 cfg3_1kbp_e10_global_adaptive 100000
full cfg5 slim_kernel 3 1 cfg5_100kbp_e15_global_adaptive 1250
full cfg4 wide_kernel 3 1 cfg4_10kbp_in_12kbp_e5_semiglobal 296
fi
ls -la $OUT
