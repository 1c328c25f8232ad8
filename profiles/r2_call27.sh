#!/bin/bash
# round-2 call 28: two pairs per cluster sharing the W tile by multicast (RQB200_ROWS_GEMM_CL=2: single pairs)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call28
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-300)" | tee -a $OUT/summary.txt
}
step tests_rows   200 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k rows_gemm
step gemm_pair    120 python profiles/bench_rows_gemm.py
step tests_fwd    300 python -m pytest tests/test_gpu_fast.py -m gpu -q -x -k "forward or prefill"
step forward      120 python profiles/bench_forward.py in1400m 64
echo "----"; cat $OUT/summary.txt
