#!/bin/bash
# one GPU: the whole GPU test suite, smoke(), the default bench line (all five configs) and the reference arm
cd "$(dirname "$0")/.."
TAG=${1:-f1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
( time timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
def row(name, r):
    if not isinstance(r, dict) or "value" not in r: print(name, r); return
    e=r.get("e2e",{}); rf=r.get("roofline",{}); cb=r.get("cpu_baseline",{})
    print("%-36s value %.4g  ms %.3f  frac %.3f  int32 %.3f  e2e %.4g (%.2f ms)  cpu %.4g  api %s" % (name, r["value"], r["ms_per_step"], rf.get("frac",0), r.get("roofline_int32",{}).get("frac_of_measured") or 0, e.get("value",0), e.get("ms_per_step_mean",0), cb.get("value",0), e.get("api_value")))
row("headline", d)
for k,v in d.get("configs",{}).items(): row(k, v)
print("api_vs_c_abi", d["e2e"].get("api_vs_c_abi"), "launches", d["gpu_launches"], "clocks", d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
