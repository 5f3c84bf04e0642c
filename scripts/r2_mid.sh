#!/bin/bash
# mid-round pass: full GPU test suite, then the side configs' bench numbers
cd "$(dirname "$0")/.."
TAG=${1:-mid1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
for c in "cfg2_150bp_e5_global 1000000" "cfg3_1kbp_e10_global_adaptive 1000000" "cfg5_100kbp_e15_global_adaptive 1250" "cfg4_10kbp_in_12kbp_e5_semiglobal 296"; do set -- $c
timeout 600 python bench.py --workload $1 --pairs $2 --steps 3 --warmup 3 --only-headline --no-cpu-baseline > $OUT/bench_$1.json 2> $OUT/bench_$1.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$1.json").read().strip().splitlines()[-1])
    print("$1 value %.5g  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.5g  launches %d  work %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["work"]))
except Exception as e: print("$1 failed", e)
PY
done
# the reference's call shape through the C++ mirror, with its own breakdown (flatten / C-ABI call / result objects)
WFACUDA_DEBUG= timeout 300 wfa_b200/host/bench_api 2 150 8 1000000 1 0 10 3 > $OUT/bench_api_cfg2.json 2> $OUT/bench_api_cfg2.err; cat $OUT/bench_api_cfg2.json | cut -c1-400
