#!/bin/bash
# round-2 call 16: are the per-layer small vectors (biases, LayerNorm parameters: DRAM misses behind the next GEMM's weight prefetch)
# what makes the reduce stages slow?  timing-only experiment with every block reading block 0's vectors; scattered-writer chain
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call16
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step exp         900 python profiles/exp_env.py "RQB200_LN_FOLD=0" "RQB200_LN_FOLD=0,RQB200_DBG_SHARED_PARAMS=1" "" "RQB200_DBG_SHARED_PARAMS=1" "RQB200_LN_FOLD=0"
RQB200_TRACE=1 RQB200_LN_FOLD=0 RQB200_DBG_SHARED_PARAMS=1 step trace_shared 300 python profiles/trace_ar.py in1400m 64
step chain4      300 python profiles/bench_chain4.py
echo "----"; cat $OUT/summary.txt
