#!/bin/bash
# N GPUs of one box: PCIe ceiling at 1..N concurrent devices, the multi-device GPU test, bench.py under torchrun
cd "$(dirname "$0")/.."
TAG=${1:-m1}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt; nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
timeout 300 python scripts/pcie_ceiling.py > $OUT/pcie_ceiling.jsonl 2> $OUT/pcie.err; cat $OUT/pcie_ceiling.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi" > $OUT/pytest_multi.log 2>&1; tail -3 $OUT/pytest_multi.log
for n in $3; do
EXTRA=""; if [ "$n" != "$N" ]; then EXTRA="--only-headline --no-cpu-baseline"; fi
( time timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 $EXTRA > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err ) 2>&1 | grep real
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n$n.json"))
    def row(name, r):
        if not isinstance(r, dict) or "value" not in r: return
        e=r.get("e2e",{})
        print("N=$n %-36s value %.4g  ms %.3f  e2e %.4g (mean %.2f min %.2f max %.2f ms)" % (name, r["value"], r["ms_per_step"], e.get("value",0), e.get("ms_per_step_mean",0), e.get("ms_per_step_min",0), e.get("ms_per_step_max",0)))
    row("headline cfg2", d)
    for k,v in d.get("configs",{}).items(): row(k, v)
    print("multi_entry", json.dumps(d.get("multi_entry"))[:1500])
except Exception as e:
    print("N=$n failed", e); import subprocess; print(open("$OUT/bench_n$n.err").read()[-1500:])
PY
done
