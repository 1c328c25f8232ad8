#!/bin/bash
# round-2 call 29: the batched passes as a PDL chain (set-up of a launch overlaps its predecessor's tail)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call29
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-300)" | tee -a $OUT/summary.txt
}
step tests_rows   200 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k rows_gemm
step tests_fast   400 python -m pytest tests/test_gpu_fast.py -m gpu -q -x
step forward      120 python profiles/bench_forward.py in1400m 64
step fwd_654m     120 python profiles/bench_forward.py cc3m654m 32
echo "----"; cat $OUT/summary.txt
