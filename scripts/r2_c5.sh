#!/bin/bash
# config 5: parity of the pipelined long-pair path, e2e at the shard size and at the full size
cd "$(dirname "$0")/.."
TAG=${1:-c5a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shard_shape or (synthetic and cfg5)" > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
for np_ in 1250 10000; do
WFACUDA_DEBUG=1 timeout 600 python bench.py --workload cfg5_100kbp_e15_global_adaptive --pairs $np_ --steps 2 --warmup 2 --only-headline --no-cpu-baseline > $OUT/bench_cfg5_$np_.json 2> $OUT/bench_cfg5_$np_.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cfg5_$np_.json").read().strip().splitlines()[-1])
    print("cfg5 $np_ pairs: value %.5g  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.5g (%.2f ms) pageable %.5g" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step_mean"], d["e2e"].get("pageable_value") or 0))
except Exception as e: print("cfg5 $np_ failed", e)
PY
grep "align_batch:" $OUT/bench_cfg5_$np_.err | tail -2
done
