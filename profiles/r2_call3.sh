#!/bin/bash
# round-2 call 3: epilogue / attention / prefetch changes -- fast-tier tests, stage trace, switch experiments, first full bench line
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call3
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
step tests_tc    300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q
step tests_fast  1500 python -m pytest tests/test_gpu_fast.py -m gpu -q -s
step trace       300 python profiles/trace_ar.py in1400m 64
step exp_env     600 python profiles/exp_env.py "" "RQB200_NO_NEXT_PF=1" "RQB200_LN_CLUSTER=1" "RQB200_SPLIT_PROJ=6,RQB200_SPLIT_FC2=6" \
                     "RQB200_SPLIT_PROJ=8,RQB200_SPLIT_FC2=8" "RQB200_SPLIT_QKV=2" "RQB200_SPLIT_FC1=2" "RQB200_GEMM_SHALLOW=1,RQB200_GEMM_L2PF=1"
step bench       900 python bench.py --steps 5 --warmup 3
step tests_par   1500 python -m pytest tests/test_gpu_parity.py -m gpu -q
echo "----"; cat $OUT/summary.txt
