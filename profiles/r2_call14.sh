#!/bin/bash
# round-2 call 14: ln_reduce grid padding / 192-thread CTAs (chain3: 64- and 128-CTA grids of >= 256 threads hand over 1.1 us slower
# than 74 / 148), GEMM warps with one polling lane
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call14
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step exp         900 python profiles/exp_env.py "" "RQB200_LN_GRID=74" "RQB200_LN_GRID=148" "RQB200_LN_THREADS=192" "RQB200_LN_THREADS=192,RQB200_LN_GRID=74" "RQB200_LN_THREADS=192,RQB200_LN_GRID=148" ""
RQB200_TRACE=1 step trace1 300 python profiles/trace_ar.py in1400m 64
RQB200_TRACE=1 RQB200_LN_GRID=148 step trace_g148 300 python profiles/trace_ar.py in1400m 64
step tests_fast  900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_tc.py -m gpu -q -x
echo "----"; cat $OUT/summary.txt
