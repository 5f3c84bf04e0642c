#!/bin/bash
# One GPU-box pass: parity tests, bench lines for every workload, ncu launch list + full capture.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag> [steps...]   steps: tests bench ncu
cd "$(dirname "$0")/.."
TAG=${1:-rX}; shift
STEPS=${@:-tests bench ncu}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
for s in $STEPS; do
case $s in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
  ;;
bench)
  timeout 900 python bench.py > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err; tail -c 3000 $OUT/bench_cfg2.json
  timeout 900 python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 200000 --steps 3 --warmup 3 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; tail -c 3000 $OUT/bench_cfg3.json
  timeout 900 python bench.py --workload cfg4_10kbp_in_12kbp_e5_semiglobal --pairs 296 --steps 2 --warmup 3 > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; tail -c 3000 $OUT/bench_cfg4.json
  timeout 900 python bench.py --workload cfg5_100kbp_e15_global_adaptive --pairs 296 --steps 2 --warmup 3 > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; tail -c 3000 $OUT/bench_cfg5.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_cfg2.json 2> $OUT/bench_ref_cfg2.err; tail -c 1000 $OUT/bench_ref_cfg2.json
  ;;
ncu)
  WFACUDA_NO_PIPELINE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_cfg2.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_cfg2.log 2>&1
  WFACUDA_NO_PIPELINE=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 3 -c 1 -f -o $OUT/prof_cfg2 \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg2.log 2>&1
  WFACUDA_NO_PIPELINE=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 3 -c 1 -f -o $OUT/prof_cfg3 \
      python bench.py --workload cfg3_1kbp_e10_global_adaptive --pairs 100000 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg3.log 2>&1
  ls -la $OUT
  ;;
esac
done
