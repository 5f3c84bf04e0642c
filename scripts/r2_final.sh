#!/bin/bash
# Final pass of the round on one GPU: smoke, the whole GPU suite, bench.py exactly as the driver runs it (both arms),
# compute-sanitizer on the LANE class (racecheck: the clamped speculative loads) and on the WIDE worker.
cd "$(dirname "$0")/.."
TAG=${1:-fin1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ) 2>&1 | grep real
( time timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
def show(k,c):
    if not isinstance(c,dict) or "value" not in c: print(k,c); return
    r=c.get("roofline") or {}; e=c.get("e2e") or {}
    print("%-36s value %.5g ms %.3f e2e %.5g api %s frac %s traffic %s kernel %s cpu %s" % (k,c["value"],c.get("ms_per_step",0),e.get("value",0),e.get("api_value"),r.get("frac"),r.get("traffic"),r.get("kernel"),(c.get("cpu_baseline") or {}).get("value")))
show("headline",d)
for k,c in d["configs"].items(): show(k,c)
print("clocks",d.get("clocks"))
PY
if [ "$2" == "san" ]; then
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_lane.py > $OUT/san_racecheck_lane.log 2>&1; echo "exit $?" >> $OUT/san_racecheck_lane.log; grep "RACECHECK SUMMARY\|ERROR SUMMARY\|exit" $OUT/san_racecheck_lane.log | tail -3
for tool in memcheck synccheck; do
timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_wide.py > $OUT/san_${tool}_wide.log 2>&1; echo "exit $?" >> $OUT/san_${tool}_wide.log; grep "ERROR SUMMARY\|exit" $OUT/san_${tool}_wide.log | tail -2
done
fi
