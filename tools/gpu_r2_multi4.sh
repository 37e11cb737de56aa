#!/usr/bin/env bash
# 4-GPU check: ragged frame split (T=5 over 4 ranks: 2,1,1,1) on NCCL with graph replay + frame-sharded decoder
set -u
N=4
out=gpurun_out/r2_multi$N
mkdir -p "$out"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline ) > "$out/ns_default.log" 2>&1
echo "exit $? $(grep -o '"ms_per_step": [0-9.]*' "$out/ns_default.log" | head -1) $(grep -o '"value": [0-9.]*' "$out/ns_default.log" | head -1) $(grep -o '"execution": "[^"]*"' "$out/ns_default.log" | head -1)" | tee "$out/summary.txt"
tail -3 "$out/ns_default.log" | cut -c1-300
