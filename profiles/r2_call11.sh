#!/bin/bash
# round-2 call 11: four-warp cp.async attention (vs one warp), split weight / activation barriers, where the GEMM time goes
# (weights-landed stamp), floor of a well-formed dependent stage
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call11
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step tests_fast  900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_tc.py -m gpu -x -q
step exp         600 python profiles/exp_env.py "" "RQB200_ATTN_ONE_WARP=1" "" "RQB200_ATTN_ONE_WARP=1"
RQB200_TRACE=1 step trace1 300 python profiles/trace_ar.py in1400m 64
RQB200_TRACE=2 step trace2 300 python profiles/trace_ar.py in1400m 64
step chain       300 python profiles/bench_chain.py 0 1 4
echo "----"; cat $OUT/summary.txt
