#!/bin/bash
# round-2 call 31: two pairs per cluster sharing the W tile by multicast, grid = co-resident clusters (cudaOccupancyMaxActiveClusters)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call31
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-300)" | tee -a $OUT/summary.txt
}
export RQB200_ROWS_GEMM_VERBOSE=1
step tests_rows   120 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k rows_gemm
step gemm_cl4     100 python profiles/bench_rows_gemm.py
RQB200_ROWS_GEMM_CL=2 step gemm_cl2 100 python profiles/bench_rows_gemm.py
step forward      100 python profiles/bench_forward.py in1400m 64
echo "----"; cat $OUT/summary.txt
