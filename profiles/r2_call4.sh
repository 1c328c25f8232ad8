#!/bin/bash
# round-2 call 4: attention v3 (coalesced rows, next-layer KV prefetch), batched forward, smoke with the fast tier first
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call4
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
step smoke       300 python __graft_entry__.py smoke
step tests_fast  1500 python -m pytest tests/test_gpu_fast.py -m gpu -q -s -x
step trace       300 python profiles/trace_ar.py in1400m 64
step exp_env     600 python profiles/exp_env.py "" "RQB200_NO_KV_PF=1" "RQB200_LN_CLUSTER=1" "RQB200_NO_PDL=1"
step fwd16       300 python profiles/bench_forward.py in1400m 16
step fwd64       300 python profiles/bench_forward.py in1400m 64
step bench       900 python bench.py --steps 5 --warmup 3
echo "----"; cat $OUT/summary.txt
