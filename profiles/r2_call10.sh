#!/bin/bash
# round-2 call 10: full GPU suite on the current tree, smoke, default bench, decoder with / without the fused GroupNorm
# statistics, config 1, forward launch list, decode launch list
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call10
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-1500)" | tee -a $OUT/summary.txt
}
step tests       1200 python -m pytest tests -m gpu -x -q
step smoke       300 python -c "import __graft_entry__ as g; g.smoke()"
step bench       900 python bench.py --steps 8 --warmup 3
step decode      300 python profiles/prof_decode.py 64 3
RQB200_GN_FUSE=0 step decode_nofuse 300 python profiles/prof_decode.py 64 3
step codes       600 python profiles/prof_codes.py
OURS='regex:rqb|gemm_tc|conv_tc|attn|ln_reduce|act_reduce|sample_kernel|code_sum|cond_tok|advance|gn_|cast_f16|vae_attn|rq_|prefill|init_state'
step ncu_fwd     900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1500 --csv --log-file gpurun_out/launches_forward_r2.csv python profiles/bench_forward.py in1400m 64
step ncu_dec     900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1500 --csv --log-file gpurun_out/launches_decode_r2b.csv python profiles/prof_decode.py 64 1
echo "----"; cat $OUT/summary.txt
