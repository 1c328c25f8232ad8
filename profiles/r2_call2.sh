#!/bin/bash
# round-2 call 2: the refactored fast tier (fp16 default, spans, batched prefill, trace) -- tests, then the stage trace, then timings
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call2
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
step tests_tc    300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q
step tests_fast  900 python -m pytest tests/test_gpu_fast.py -m gpu -q -s
step tests_par   900 python -m pytest tests/test_gpu_parity.py -m gpu -q
step trace       300 python profiles/trace_ar.py in1400m 64
step exp_env     300 python profiles/exp_env.py "" "RQB200_GEMM_SHALLOW=1,RQB200_GEMM_L2PF=1" "RQB200_FAST_DTYPE=bf16" "RQB200_NO_PDL=1"
echo "----"; cat $OUT/summary.txt
