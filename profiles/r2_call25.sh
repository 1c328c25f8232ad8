#!/bin/bash
# round-2 call 25: split-K block GEMMs as thread-block clusters with multicast activation loads (2 / 4 weight tiles per cluster)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call25
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-300)" | tee -a $OUT/summary.txt
}
step tests_mc    120 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "cluster_multicast"
step exp         420 python profiles/exp_env.py "" "RQB200_GEMM_CLUSTER=2" "RQB200_GEMM_CLUSTER=4" "" "RQB200_GEMM_CLUSTER=2" "RQB200_GEMM_CLUSTER=4"
RQB200_TRACE=1 RQB200_GEMM_CLUSTER=4 step trace_c4 150 python profiles/trace_ar.py in1400m 64
echo "----"; cat $OUT/summary.txt
