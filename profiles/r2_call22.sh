#!/bin/bash
# round-2 call 22: small-vector L2 prefetch A/B; mma.sync causal attention of the batched passes (T <= 64): forward / prefill tests, forward timing
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call22
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step tests_fwd    300 python -m pytest tests/test_gpu_fast.py -m gpu -q -x -k "forward or prefill or consistency"
step forward      120 python profiles/bench_forward.py in1400m 64
step exp          420 python profiles/exp_env.py "" "RQB200_NO_PARAM_PREFETCH=1" "" "RQB200_NO_PARAM_PREFETCH=1"
echo "----"; cat $OUT/summary.txt
