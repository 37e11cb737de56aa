#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call12
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run kernel_tests 900 python -m pytest tests/test_ops_gpu.py tests/test_fused_glue.py tests/test_window_attn_tc.py tests/test_model_gpu.py tests/test_golden.py -m gpu -q
run msda_sweep 300 python tools/msda_tile_sweep.py
run bench_ns 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run torch_eager 600 python bench.py --impl torch-eager --steps 3
run parity 1500 python -m pytest tests/test_parity_full_geometry.py -m gpu -q
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:msda_encoder_staged -c 1 -o "$out/msda" python bench.py --ncu-step --no-cpu-baseline > "$out/msda.log" 2>&1
cat "$out/summary.txt"
