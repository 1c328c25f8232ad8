#!/bin/bash
# round-2 call 23: final state -- full GPU suite, smoke, the bench line of every BASELINE config + the CPU arm, forward, config 1
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call23
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-300)" | tee -a $OUT/summary.txt
}
step tests       1500 python -m pytest tests -m gpu -q -x
step smoke       300 python -c "import __graft_entry__ as g; g.smoke()"
step bench       600 python bench.py --steps 10 --warmup 3
step bench_cfg2  300 python bench.py --model ffhq355m --steps 5 --warmup 3 --no-cpu-baseline
step bench_cfg4  300 python bench.py --model cc3m654m --steps 5 --warmup 3 --no-cpu-baseline
step bench_cfg4b 400 python bench.py --model cc3m654m_16 --steps 3 --warmup 3 --no-cpu-baseline --no-extras
step bench_cfg5  400 python bench.py --model t2i3900m --steps 3 --warmup 3 --no-cpu-baseline --no-extras
step bench_ref   400 python bench.py --impl reference --steps 3 --warmup 1
step fwd64       200 python profiles/bench_forward.py in1400m 64
step fwd_654m    200 python profiles/bench_forward.py cc3m654m 32
step codes       300 python profiles/prof_codes.py
step trace       200 python profiles/trace_ar.py in1400m 64
echo "----"; cat $OUT/summary.txt
