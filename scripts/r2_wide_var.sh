#!/bin/bash
# WIDE worker: parity, then config 4 with different block sizes
cd "$(dirname "$0")/.."
TAG=${1:-wv1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide_worker or (random_small and cta) or (synthetic and cfg4)" > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
for th in ${2:-1024 768 512}; do
WFACUDA_WIDE_THREADS=$th timeout 600 python bench.py --workload cfg4_10kbp_in_12kbp_e5_semiglobal --pairs 296 --steps 2 --warmup 2 --only-headline --no-cpu-baseline > $OUT/bench_cfg4_$th.json 2> $OUT/bench_cfg4_$th.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cfg4_$th.json").read().strip().splitlines()[-1])
    print("threads $th: cfg4 value %.5g  ms/step %.3f  kernel_ms %.3f  frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e: print("cfg4 $th failed", e)
PY
done
