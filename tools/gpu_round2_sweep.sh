#!/usr/bin/env bash
# Tuning sweeps (optional third GPU call of round 2, after the parity steps of tools/gpu_round2_open.sh are green).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2_sweep.sh'
set -u
out=gpurun_out/r2_sweep
mkdir -p "$out"
timeout 600 python tools/msda_tile_sweep.py > "$out/msda_tiles.log" 2>&1
for s in 1 2 3 4 7; do
  UNIVS_MHA_TC_SPLITS=$s timeout 300 python tests/tools/mha_tc_check.py --time > "$out/mha_tc_splits_$s.log" 2>&1
done
for g in 2 3 5; do
  UNIVS_FRAME_STREAMS=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$out/bench_streams_$g.log" 2>&1
done
# L2-resident MLP schedule (validated kernels, another order): working-set sizes around the 126 MB L2
for mb in 32 64 96 160; do
  UNIVS_MLP_CHUNK_MB=$mb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$out/bench_mlp_chunk_$mb.log" 2>&1
done
tail -n 12 "$out"/*.log
