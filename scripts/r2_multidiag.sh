#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-md1}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT; nproc > $OUT/nproc.txt
for wl in cfg2_150bp_e5_global cfg5_100kbp_e15_global_adaptive; do
for env in "X=1" "WFACUDA_PIPE_WORKERS=6 WFACUDA_BLOCKING_SYNC=1"; do
  ( export $env; timeout 600 python bench.py --gpus $N --multi-entry --workload $wl --steps 3 > $OUT/me_$wl.json 2> $OUT/me_$wl.err )
  python - <<PY
import json
d=json.load(open("$OUT/me_$wl.json")); m=d["multi_entry"]
print("$wl [$env] value %.4g  ms mean %.2f min %.2f  per_dev %s" % (m["value"], m["ms_per_call_mean"], m["ms_per_call_min"], m["pairs_per_device"]))
PY
done
done
WFACUDA_DEBUG=1 timeout 600 python bench.py --gpus $N --multi-entry --workload cfg5_100kbp_e15_global_adaptive --steps 1 > /dev/null 2> $OUT/dbg_cfg5.err; grep -E "upload:|run:|download:|launch slim" $OUT/dbg_cfg5.err | tail -12
cat $OUT/nproc.txt
