#!/bin/bash
# First GPU call of the next round: exercise every prepared default-off path once, with per-step timeouts, logs under gpurun_out/.
#   gpurun --timeout 1500 -- 'bash profiles/r2_bringup.sh'
# Each step is independent: a hang or a trap in one (bounded by `timeout`) does not stop the others.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_bringup
mkdir -p $OUT
step() {  # name, seconds, command...
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
step default_tests_tc      300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q
step gemm_gr               300 python -m pytest tests/pending_r2_gr.py -m gpu -q -k "gemm_gr"
step rq_v2                 300 python -m pytest tests/pending_r2_gr.py -m gpu -q -k "rq_search_v2"
step sampler_v2            300 python -m pytest tests/pending_r2_gr.py -m gpu -q -k "sampler_v2"
step enc_fast              300 python -m pytest tests/pending_r2_gr.py -m gpu -q -k "fast_tier_encode"
step gr_chain_tiny         300 python -m pytest tests/pending_r2_gr.py -m gpu -q -k "gr_chain and (tiny or consistency)"
step gr_chain_big          600 python -m pytest tests/pending_r2_gr.py -m gpu -q -k "gr_chain and (ffhq355m or in1400m)"
step exp_env               300 python profiles/exp_env.py "" "RQB200_GEMM_STAGES=4,RQB200_GEMM_L2PF=1" "RQB200_GEMM_STAGES=4,RQB200_GEMM_RELINQ=1" \
                               "RQB200_GEMM_STAGES=4,RQB200_GEMM_L2PF=1,RQB200_GEMM_RELINQ=1" "RQB200_GEMM_STAGES=3,RQB200_GEMM_L2PF=1,RQB200_GEMM_RELINQ=1" "RQB200_GR=1" "RQB200_GR=1,RQB200_LNFOLD=1" \
                               "RQB200_GR=1,RQB200_LNFOLD=1,RQB200_GEMM_STAGES=4,RQB200_GEMM_L2PF=1" "RQB200_SAMPLER_V2=1" \
                               "RQB200_GR=1,RQB200_LNFOLD=1,RQB200_SAMPLER_V2=1"
step rq_v1                 120 python profiles/prof_rq.py 64 16384
RQB200_RQ_V2=1 step rq_v2_time 120 python profiles/prof_rq.py 64 16384
step sampler_v1            120 python profiles/bench_sampler.py
RQB200_SAMPLER_V2=1 step sampler_v2_time 120 python profiles/bench_sampler.py
step gemm_iso              120 python profiles/bench_gemm.py
RQB200_GEMM_STAGES=4 RQB200_GEMM_L2PF=1 RQB200_GEMM_RELINQ=1 step gemm_iso_s4 120 python profiles/bench_gemm.py
step chain                 120 python profiles/bench_chain.py
echo "----"; cat $OUT/summary.txt
