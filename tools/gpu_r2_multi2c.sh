#!/usr/bin/env bash
# 2-GPU check of the final defaults (gpurun --gpus 2): the driver's own launch line, default switches (frame-sharded decoder)
set -u
out=gpurun_out/r2_multi2c
mkdir -p "$out"
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline ) > "$out/bench_n2.log" 2>&1
echo "exit $?"; grep -o '"ms_per_step": [0-9.]*' "$out/bench_n2.log" | head -2; grep -o '"value": [0-9.]*' "$out/bench_n2.log" | head -1; tail -n 3 "$out/bench_n2.log" | cut -c1-300
