#!/bin/bash
# round-2 call 15: LayerNorm folded into qkv / fc1 (elementwise x-reduce on 148 CTAs instead of the 64-CTA row kernel): A/B + tests + trace
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call15
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step exp         900 python profiles/exp_env.py "" "RQB200_LN_FOLD=0" "" "RQB200_LN_FOLD=0"
step tests_fast  900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_tc.py -m gpu -q
RQB200_TRACE=1 step trace1 300 python profiles/trace_ar.py in1400m 64
echo "----"; cat $OUT/summary.txt
