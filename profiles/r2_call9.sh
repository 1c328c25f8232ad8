#!/bin/bash
# round-2 call 9: fused GroupNorm statistics with the cheaper reduction (vs stand-alone pass), soft codes, forward launch list, trace
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call9
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-400)" | tee -a $OUT/summary.txt
}
step tests_par   900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "vae or soft"
step decode      300 python profiles/prof_decode.py 64 3
RQB200_GN_FUSE=0 step decode_nofuse 300 python profiles/prof_decode.py 64 3
step codes       600 python profiles/prof_codes.py
step trace       300 python profiles/trace_ar.py in1400m 64
OURS='regex:rqb|gemm_tc|conv_tc|attn|ln_reduce|act_reduce|sample_kernel|code_sum|cond_tok|advance|gn_|cast_f16|vae_attn|rq_|prefill|init_state'
step ncu_fwd     900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1500 --csv --log-file gpurun_out/launches_forward_r2.csv python profiles/bench_forward.py in1400m 64
echo "----"; cat $OUT/summary.txt
