#!/bin/bash
# round-2 call 5: encoder on tcgen05 (stride-2 convs), attention phase stamps, TMA fill-rate micro-benchmark, config 1 on the GPU
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call5
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
step tests_tc    300 python -m pytest tests/test_gpu_tc.py -m gpu -q
step tests_par   1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "vae or embed or rq_"
step tma         300 python profiles/bench_tma.py
step trace       300 python profiles/trace_ar.py in1400m 64
RQB200_NO_KV_PF=1 step trace_nopf 300 python profiles/trace_ar.py in1400m 64
step codes       600 python profiles/prof_codes.py
echo "----"; cat $OUT/summary.txt
