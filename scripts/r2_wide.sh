#!/bin/bash
# WIDE worker pass: parity (forced cluster sizes, cta worker test, config 4 sample), config-4 bench, optional ncu
cd "$(dirname "$0")/.."
TAG=${1:-w1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
WFACUDA_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide_worker or (random_small and cta) or (synthetic and cfg4)" > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; grep -v "^\[wfacuda\]" $OUT/pytest.log | tail -15; grep "launch wide" $OUT/pytest.log | tail -3
WFACUDA_DEBUG=1 timeout 600 python bench.py --workload cfg4_10kbp_in_12kbp_e5_semiglobal --pairs ${3:-296} --steps 2 --warmup 2 --only-headline --no-cpu-baseline > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cfg4.json").read().strip().splitlines()[-1])
    print("cfg4 value %.5g  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.5g  launches %d  work %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["work"]))
except Exception as e: print("cfg4 failed", e)
PY
grep "launch wide\|launch cta" $OUT/bench_cfg4.err | tail -3; tail -3 $OUT/bench_cfg4.err
if [ "$2" == "ncu" ]; then
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --metrics $M --clock-control none -k regex:wide_ -s 2 -c 2 python bench.py --workload cfg4_10kbp_in_12kbp_e5_semiglobal --pairs ${3:-296} --steps 1 --warmup 1 --only-headline --no-cpu-baseline > $OUT/ncu_cfg4.log 2>&1
grep -A14 "wide_" $OUT/ncu_cfg4.log | grep "wide_\|inst_executed\|duration\|issue_active\|warps_active\|registers\|dram" | head -30
fi
if [ "$2" == "full" ]; then
WFACUDA_NO_PIPELINE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 1 -c 1 -f -o $OUT/prof_cfg4 python bench.py --workload cfg4_10kbp_in_12kbp_e5_semiglobal --pairs ${3:-296} --steps 1 --warmup 1 --only-headline --no-cpu-baseline > $OUT/ncu_full_cfg4.log 2>&1
fi
ls -la $OUT
