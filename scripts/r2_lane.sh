#!/bin/bash
# LANE class pass: parity (lane tests, fixtures, edge cases), staged-run check, config-2 bench
cd "$(dirname "$0")/.."
TAG=${1:-l1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x -k "lane or random_small or fixture or edge or cfg2 or readme or errors or context or multi" > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
timeout 300 python scripts/staged_check.py 200000 > $OUT/staged_check.log 2>&1; tail -2 $OUT/staged_check.log
for v in run; do
timeout 600 python bench.py --steps 5 --warmup 3 --only-headline --no-cpu-baseline > $OUT/bench_cfg2_$v.json 2> $OUT/bench_cfg2_$v.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cfg2_$v.json").read().strip().splitlines()[-1])
    print("$v: cfg2 value %.5g  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.5g (%.2f ms) launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step_mean"], d["gpu_launches"]))
except Exception as e: print("cfg2 $v failed", e)
PY
done
