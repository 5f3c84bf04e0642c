#!/bin/bash
# The library's own multi-GPU entry on all GPUs of the box, ONE process: config 5 (10 k pairs cut into N shards) and
# config 2 (N x 1 M pairs), plus the multi-device parity test.
cd "$(dirname "$0")/.."
TAG=${1:-m8b}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi" > $OUT/pytest_multi.log 2>&1; tail -2 $OUT/pytest_multi.log
for wl in cfg5_100kbp_e15_global_adaptive cfg2_150bp_e5_global; do
  PAIRS=""; if [ $wl == cfg2_150bp_e5_global ]; then PAIRS="--pairs $((N * 1000000))"; fi
  timeout 600 python bench.py --gpus $N --multi-entry --workload $wl --steps 3 $PAIRS > $OUT/me_$wl.json 2> $OUT/me_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/me_$wl.json")); m=d["multi_entry"]
    print("$wl N=$N value %.5g  ms mean %.2f min %.2f  per_dev %s ok %s" % (m["value"], m["ms_per_call_mean"], m["ms_per_call_min"], m["pairs_per_device"], m["pairs_ok"]))
except Exception as e:
    print("$wl failed", e); print(open("$OUT/me_$wl.err").read()[-800:])
PY
done
