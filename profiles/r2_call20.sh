#!/bin/bash
# round-2 call 21 (eight epilogue warps, 64-column fp32 steps): the CTA-pair rows GEMM (cta_group::2, 256 x 256 tiles): unit tests, shape-by-shape rate against the single-CTA
# persistent kernel, the forward pass
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call20
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step tests_rows   240 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k rows_gemm
step gemm_pair    180 python profiles/bench_rows_gemm.py
RQB200_ROWS_GEMM_1CTA=1 step gemm_1cta 180 python profiles/bench_rows_gemm.py
step forward      240 python profiles/bench_forward.py in1400m 64
step tests_fwd    600 python -m pytest tests/test_gpu_fast.py -m gpu -q -x -k "forward or prefill"
echo "----"; cat $OUT/summary.txt
