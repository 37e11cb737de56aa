#!/usr/bin/env bash
# Multi-GPU check (gpurun --gpus N): frame-sharded step as a CUDA-graph replay, feature all-gather vs frame-sharded decoder.
set -u
N=${1:-2}
out=gpurun_out/r2_multi$N
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -1) $(grep -o '"execution": "[^"]*"' "$out/$name.log" | head -1) $(tail -n 2 "$out/$name.log" | tr '\n' ' ' | cut -c1-200))" | tee -a "$out/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run graph_gather 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline
run graph_sharddec 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --shard-decoder
UNIVS_GRAPH_MULTI=0 run eager_gather 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline
cat "$out/summary.txt"
