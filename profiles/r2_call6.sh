#!/bin/bash
# round-2 call 6: attention with K rows fetched ahead of the dependency; large-M ring depth; ncu launch lists + one full capture
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call6
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
step tests_fast  1500 python -m pytest tests/test_gpu_fast.py -m gpu -q -x
step exp_env     600 python profiles/exp_env.py "" "RQB200_NO_KV_PF=1"
step trace       300 python profiles/trace_ar.py in1400m 64
step fwd64       300 python profiles/bench_forward.py in1400m 64
RQB200_BATCHED_DEEP=1 step fwd64_deep 300 python profiles/bench_forward.py in1400m 64
# ncu: launch lists (shares only; cold-cache, serialised) and one --set full capture of the block GEMMs + attention
OURS='regex:rqb|gemm_tc|conv_tc|attn_fast|ln_reduce|act_reduce|sample_kernel|code_sum|cond_tok|advance|gn_|cast_f16|vae_attn|rq_|prefill'
step ncu_ar      900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k "$OURS" -c 1300 --csv --log-file gpurun_out/launches_ar_r2.csv python profiles/prof_ar.py 64 1 2
step ncu_dec     900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv --log-file gpurun_out/launches_decode_r2.csv python profiles/prof_decode.py 64 1
step ncu_full    900 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:'gemm_tc|attn_fast|ln_reduce' -s 60 -c 9 -o gpurun_out/ncu_ar_chain_r2 python profiles/prof_ar.py 64 1 1
echo "----"; cat $OUT/summary.txt
