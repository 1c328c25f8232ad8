#!/bin/bash
# round-2 call 12: fast-tier suite on the four-warp attention (alignment fix) + warp-per-row LayerNorm; forward timing;
# read / write / clean-read split of one dependent stage
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call12
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step tests_fast  900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_tc.py -m gpu -q
step chain2      300 python profiles/bench_chain2.py
step forward     300 python profiles/bench_forward.py in1400m 64
echo "----"; cat $OUT/summary.txt
