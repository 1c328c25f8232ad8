#!/bin/bash
# round-2 call 24: ncu evidence of the final build -- launch lists (AR, decode, forward) and full captures (AR chain, pair GEMM)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call24
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-300)" | tee -a $OUT/summary.txt
}
OURS='regex:rqb|gemm_tc|conv_tc|rows_gemm2|attn|ln_reduce|ln_rows|act_reduce|sample_kernel|code_sum|cond_tok|advance|gn_|cast_f16|vae_attn|rq_|prefill|init_state'
step ncu_ar      600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k "$OURS" -c 1300 --csv --log-file gpurun_out/launches_ar_r2.csv python profiles/prof_ar.py 64 1 2
step ncu_dec     400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv --log-file gpurun_out/launches_decode_r2.csv python profiles/prof_decode.py 64 1
step ncu_fwd     600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1500 --csv --log-file gpurun_out/launches_forward_r2.csv python profiles/bench_forward.py in1400m 64
step ncu_chain   600 ncu --set full --clock-control none --import-source on --graph-profiling node -k "regex:gemm_tc|attn_fast2|ln_reduce|act_reduce" -s 60 -c 9 -o gpurun_out/ncu_ar_chain_r2 -f python profiles/prof_ar.py 64 1 1
step ncu_pair    600 ncu --set full --clock-control none --import-source on -k "regex:rows_gemm2" -s 8 -c 4 -o gpurun_out/ncu_rows_gemm2_r2 -f python profiles/bench_forward.py in1400m 64
echo "----"; cat $OUT/summary.txt
