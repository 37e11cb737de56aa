#!/usr/bin/env bash
# 2-GPU re-check after the CTA-pair GEMM / fused conv: default frame-sharded step (graph replay + frame-sharded decoder) at NS.
set -u
N=2
out=gpurun_out/r2_multi2b
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -1) $(grep -o '"value": [0-9.]*' "$out/$name.log" | head -1) $(grep -o '"execution": "[^"]*"' "$out/$name.log" | head -1) $(tail -n 2 "$out/$name.log" | tr '\n' ' ' | cut -c1-160))" | tee -a "$out/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run ns_default 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline
run ns_ref_arm_rank_contract 400 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 0
cat "$out/summary.txt"
