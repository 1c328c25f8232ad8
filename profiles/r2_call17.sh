#!/bin/bash
# round-2 call 17: keep the block scheduler from stacking the small kernels' CTAs on few SMs (unused dynamic shared memory as a spacer)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call17
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step exp         900 python profiles/exp_env.py "RQB200_LN_FOLD=0" "RQB200_LN_FOLD=0,RQB200_PAD_SMEM_KB=28" "RQB200_LN_FOLD=0,RQB200_PAD_SMEM_KB=14" "RQB200_LN_FOLD=0,RQB200_PAD_SMEM_KB=28,RQB200_PAD_SMEM_ACT_KB=9" "RQB200_PAD_SMEM_KB=28" "RQB200_PAD_SMEM_KB=28,RQB200_PAD_SMEM_ACT_KB=9" "RQB200_LN_FOLD=0"
RQB200_TRACE=1 RQB200_LN_FOLD=0 RQB200_PAD_SMEM_KB=28 step trace_pad 300 python profiles/trace_ar.py in1400m 64
RQB200_TRACE=1 RQB200_PAD_SMEM_KB=28 RQB200_PAD_SMEM_ACT_KB=9 step trace_pad_fold 300 python profiles/trace_ar.py in1400m 64
echo "----"; cat $OUT/summary.txt
