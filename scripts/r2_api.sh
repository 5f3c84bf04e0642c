#!/bin/bash
# cfg2 end to end: C ABI with spinning / sleeping waits, and the reference's call shape through the C++ mirror
cd "$(dirname "$0")/.."
TAG=${1:-a1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for b in 0 1; do
WFACUDA_BLOCKING_SYNC=$b timeout 300 python bench.py --only-headline --no-cpu-baseline > $OUT/bench_b$b.json 2> $OUT/bench_b$b.err
python - <<PY
import json
d=json.load(open("$OUT/bench_b$b.json")); e=d["e2e"]
print("blocking=$b value %.4g e2e %.4g mean %.2f median %.2f min %.2f max %.2f ms  api %.4g (%.2f ms, x%.2f)" % (d["value"], e["value"], e["ms_per_step_mean"], e["ms_per_step_median"], e["ms_per_step_min"], e["ms_per_step_max"], e.get("api_value",0), e.get("api_ms_per_call_mean",0), e.get("api_vs_c_abi",0)))
PY
done
wfa_b200/host/bench_api 2 150 8 1000000 1 0 10 3
timeout 120 python -m pytest tests/test_host_cpp.py -m gpu -q 2>&1 | tail -2
