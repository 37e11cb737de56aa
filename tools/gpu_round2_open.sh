#!/usr/bin/env bash
# First GPU call of round 2: validates everything that was written without GPU time at the end of round 1 and measures
# it, in ONE box session.  Every step runs under its own `timeout` (a hung kernel must not hang the box) and logs to
# gpurun_out/r2_open/; a failing step does not stop the next one.
#   /usr/local/graft/bin/gpurun --timeout 3300 -- 'bash tools/gpu_round2_open.sh'      (≈ 40 min of box time)
# STEPS=a,b,c restricts the run to the named steps (names as in the `run` lines below).
set -u
out=gpurun_out/r2_open
mkdir -p "$out"
run() {  # name, seconds, command...
  local name=$1 secs=$2; shift 2
  if [ -n "${STEPS:-}" ] && [[ ",$STEPS," != *",$name,"* ]]; then return 0; fi
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300))" | tee -a "$out/summary.txt"
}
# 0. the validated default path still green on this box
run default_gpu_tests 1500 python -m pytest tests -m gpu -q -x
# 1. opt-in kernels, one family per step so that a trapped launch (sticky CUDA error) only spoils its own process
UNIVS_GPU_WINTC=1 run wintc_tests 600 python -m pytest tests/test_window_attn_tc.py -m gpu -q
run wintc_check 600 python tests/tools/win_tc_check.py --time
UNIVS_GPU_MHATC=1 run mhatc_tests 600 python -m pytest tests/test_mha_tc.py -m gpu -q
run mhatc_check 600 python tests/tools/mha_tc_check.py --time
UNIVS_GPU_EINSUM_MC=1 run einsum_mc_tests 600 python -m pytest tests/test_einsum_mc.py -m gpu -q
EINSUM_MC=1 run einsum_mc_check 600 python tools/einsum_tc_check.py
UNIVS_GPU_ROWWISE_V2=1 run rowwise_v2_tests 600 python -m pytest tests/test_rowwise_v2.py -m gpu -q
UNIVS_GPU_GLUE=1 run glue_tests 900 python -m pytest tests/test_fused_glue.py -m gpu -q
UNIVS_GPU_COMPAT=1 run compat_tests 300 python -m pytest tests/test_compat_msda_module.py -m gpu -q
UNIVS_GPU_HEADS=1 run heads_tests 900 python -m pytest tests/test_heads_golden.py -m gpu -q
# 2. measurements: default, then each opt-in on top of it (a path that failed above still runs: its number is void)
run bench_default 900 python bench.py --steps 10 --warmup 3
UNIVS_FUSED_GLUE=1 UNIVS_MSDA_TILE=8 run bench_glue 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_WIN_TC=1 run bench_wintc 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_ROWWISE_V2=3 run bench_rowwise_v2 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_FRAME_STREAMS=2 run bench_streams2 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_FRAME_STREAMS=5 run bench_streams5 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_MLP_CHUNK_MB=96 run bench_mlp_chunk 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_POOLED_MASKS=1 run bench_pooled_masks 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_EINSUM_MC=1 run bench_einsum_mc 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_MHA_TC=1 run bench_mhatc 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_FUSED_GLUE=1 UNIVS_MSDA_TILE=8 UNIVS_WIN_TC=1 UNIVS_MHA_TC=1 UNIVS_ROWWISE_V2=3 UNIVS_EINSUM_MC=1 UNIVS_POOLED_MASKS=1 run bench_all 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_video 900 python bench.py --steps 2 --video-frames 12
# 3. end-to-end parity of the opt-in paths at the north-star geometry (T=2): same tool and thresholds as round 1
PARITY_MODES=fp16x3 UNIVS_FUSED_GLUE=1 UNIVS_MSDA_TILE=8 UNIVS_WIN_TC=1 UNIVS_MHA_TC=1 UNIVS_ROWWISE_V2=3 UNIVS_EINSUM_MC=1 UNIVS_POOLED_MASKS=1 run parity_at_scale 1200 python tests/tools/parity_at_scale.py
# 4. the prompt configurations of BASELINE.json at full geometry (default path): C3 sot memory over 3 clips, C4 grounding, C5 1080p
PARITY_CONFIG=c3 run parity_c3 900 python tests/tools/parity_configs.py
PARITY_CONFIG=c4 run parity_c4 900 python tests/tools/parity_configs.py
PARITY_CONFIG=c5 run parity_c5 900 python tests/tools/parity_configs.py
cat "$out/summary.txt"
