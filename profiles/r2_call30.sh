#!/bin/bash
# round-2 call 30: the final build once more -- full GPU suite, smoke, default bench line, forward
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call30
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-300)" | tee -a $OUT/summary.txt
}
step tests       1200 python -m pytest tests -m gpu -q -x
step smoke       200 python -c "import __graft_entry__ as g; g.smoke()"
step bench       500 python bench.py --steps 10 --warmup 3
step fwd64       120 python profiles/bench_forward.py in1400m 64
echo "----"; cat $OUT/summary.txt
