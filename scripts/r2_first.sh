#!/bin/bash
# round 2, first GPU pass of the REG worker: parity first, then cfg3 / cfg5 bench lines with and without it
cd "$(dirname "$0")/.."
TAG=${1:-r4a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slim_worker or random_small" > $OUT/pytest_reg.log 2>&1; echo "exit $?" >> $OUT/pytest_reg.log; tail -30 $OUT/pytest_reg.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "synthetic or shard_shape" > $OUT/pytest_cfg.log 2>&1; echo "exit $?" >> $OUT/pytest_cfg.log; tail -30 $OUT/pytest_cfg.log
for mode in slim noslim; do
  if [ $mode == noslim ]; then export WFACUDA_NO_SLIM=1; else unset WFACUDA_NO_SLIM; fi
  WFACUDA_DEBUG=1 timeout 600 python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 200000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg3_$mode.json 2> $OUT/bench_cfg3_$mode.err
  WFACUDA_DEBUG=1 timeout 600 python bench.py --workload cfg5_100kbp_e15_global_adaptive --pairs 1250 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_$mode.json 2> $OUT/bench_cfg5_$mode.err
  python - <<PY
import json
for c in ("cfg3","cfg5"):
    try:
        d=json.load(open("$OUT/bench_%s_$mode.json" % c))
        print("$mode", c, "value %.4gM  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.4gM  launches %d  work %s" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["gpu_launches"], d["work"]))
    except Exception as e: print("$mode", c, "failed", e)
PY
done
unset WFACUDA_NO_SLIM
grep "launch slim" $OUT/bench_cfg3_slim.err | tail -3; grep "launch slim" $OUT/bench_cfg5_slim.err | tail -3
WFACUDA_NO_PIPELINE=1 timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:slim_kernel -s 3 -c 1 python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 100000 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_inst_cfg3.log 2>&1
grep -A12 "slim_kernel" $OUT/ncu_inst_cfg3.log | grep "inst_executed\|duration\|issue_active\|warps_active\|registers" 
