#!/bin/bash
# runs the resident-batch bench for each libwfacuda_*.so variant (swapped into place)
cd "$(dirname "$0")/.."
cp wfa_b200/libwfacuda.so /tmp/libwfacuda_orig.so
for v in "$@"; do
  cp wfa_b200/libwfacuda_$v.so wfa_b200/libwfacuda.so
  echo "VARIANT $v"
  WFACUDA_DEBUG=1 WFACUDA_NO_PIPELINE=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | grep "run:" | head -9 | sed "s/.*align \([0-9.]*\) total.*/\1/" | tr "\n" " "; echo
done
cp /tmp/libwfacuda_orig.so wfa_b200/libwfacuda.so
