#!/bin/bash
# quick pass: SLIM parity subset + cfg3/cfg5 bench + instruction counts
cd "$(dirname "$0")/.."
TAG=${1:-q1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slim_worker or synthetic or shard_shape" > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log
WFACUDA_DEBUG=1 timeout 600 python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 200000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err
WFACUDA_DEBUG=1 timeout 600 python bench.py --workload cfg5_100kbp_e15_global_adaptive --pairs 1250 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err
if [ "$2" == "full5" ]; then WFACUDA_DEBUG=1 timeout 900 python bench.py --workload cfg5_100kbp_e15_global_adaptive --pairs 10000 --steps 1 --warmup 2 --no-cpu-baseline > $OUT/bench_cfg5full.json 2> $OUT/bench_cfg5full.err; fi
python - <<PY
import json
for c in ("cfg3","cfg5","cfg5full"):
    try:
        d=json.load(open("$OUT/bench_%s.json" % c))
        print(c, "value %.4gM  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.4gM  launches %d  work %s" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["gpu_launches"], d["work"]))
    except Exception as e: print(c, "failed", e)
PY
grep "launch slim" $OUT/bench_cfg3.err | tail -1; grep "launch slim" $OUT/bench_cfg5.err | tail -1
for c in "cfg3_1kbp_e10_global_adaptive 100000" "cfg5_100kbp_e15_global_adaptive 1250"; do set -- $c
WFACUDA_NO_PIPELINE=1 timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:slim_kernel -s 3 -c 1 python bench.py --workload $1 --pairs $2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_inst_$1.log 2>&1
grep -A12 "slim_kernel" $OUT/ncu_inst_$1.log | grep "inst_executed\|duration\|issue_active\|warps_active\|registers"
done
