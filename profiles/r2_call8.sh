#!/bin/bash
# round-2 call 8 (2 GPUs): the driver's multi-GPU launch of both bench arms + the gloo-free NCCL path; decoder with fused GN stats
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call8
mkdir -p $OUT
step() {
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $secs "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 $OUT/$name.log | cut -c1-900)" | tee -a $OUT/summary.txt
}
step tests_vae   900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "vae"
step decode      300 python profiles/prof_decode.py 64 3
step codes       600 python profiles/prof_codes.py
step bench2      900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3
step bench2_ref  900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1
echo "----"; cat $OUT/summary.txt
