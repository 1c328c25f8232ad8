#!/bin/bash
# round-2 call 7: full GPU test suite, the bench line of every BASELINE config, forward through the rows GEMM
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call7
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-600)" | tee -a $OUT/summary.txt
}
step tests       2400 python -m pytest tests -m gpu -q -x
step smoke       300 python __graft_entry__.py smoke
step fwd64       300 python profiles/bench_forward.py in1400m 64
RQB200_BATCHED_STREAMER=1 step fwd64_streamer 300 python profiles/bench_forward.py in1400m 64
step fwd_654m    300 python profiles/bench_forward.py cc3m654m 32
step bench       900 python bench.py --steps 5 --warmup 3
step bench_cfg2  600 python bench.py --model ffhq355m --steps 5 --warmup 3 --no-cpu-baseline
step bench_cfg4  600 python bench.py --model cc3m654m --steps 5 --warmup 3 --no-cpu-baseline
step bench_cfg4b 900 python bench.py --model cc3m654m_16 --steps 3 --warmup 3 --no-cpu-baseline --no-extras
step bench_cfg5  900 python bench.py --model t2i3900m --steps 3 --warmup 3 --no-cpu-baseline --no-extras
step bench_ref   600 python bench.py --impl reference --steps 3 --warmup 1
echo "----"; cat $OUT/summary.txt
