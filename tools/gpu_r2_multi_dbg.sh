#!/usr/bin/env bash
set -u
N=${1:-2}
out=gpurun_out/r2_multi_dbg
mkdir -p "$out"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > "$out/gather.log" 2>&1
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --shard-decoder > "$out/sharddec.log" 2>&1
grep -h -A16 "capture failed on rank 0" "$out/gather.log" "$out/sharddec.log" | head -60
grep -ho '"ms_per_step": [0-9.]*, "higher\|"execution": "[^"]*"' "$out/gather.log" "$out/sharddec.log"
