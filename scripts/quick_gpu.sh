#!/bin/bash
# quick GPU pass: parity tests, cfg2 bench summary, per-launch kernel times (ncu, serialised)
cd "$(dirname "$0")/.."
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$2" != "notest" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err
python - <<PY
import json
d=json.load(open("$OUT/bench_cfg2.json"))
print("cfg2 value %.1fM  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.1fM (%.2f ms, min %.2f)  launches %d" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step_mean"], d["e2e"]["ms_per_step_min"], d["gpu_launches"]))
PY
WFACUDA_NO_PIPELINE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
for r in csv.reader(open("$OUT/launches.csv")):
    if len(r) > 5 and r[0].isdigit(): print(r[4][:60], r[-1])
PY
