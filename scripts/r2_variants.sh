#!/bin/bash
# cfg3 bench with builds of the library that differ in one compile-time knob (wfa_b200/_variants/*.so)
cd "$(dirname "$0")/.."
TAG=${1:-v1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for so in wfa_b200/_variants/*.so; do
  name=$(basename $so .so)
  WFACUDA_LIB=$PWD/$so timeout 600 python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 200000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json"))
    print("$name", "value %.4gM  ms/step %.3f  kernel_ms %.3f  frac %.3f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e: print("$name failed", e)
PY
done
