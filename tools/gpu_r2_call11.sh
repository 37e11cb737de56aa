#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call11
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run kernel_tests 600 python -m pytest tests/test_window_attn_tc.py tests/test_gemm_tc_gpu.py tests/test_model_gpu.py -m gpu -q
run wintc_check 300 python tests/tools/win_tc_check.py --time
run shapes 300 python tools/gemm_tc_shapes.py
run bench_ns 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run torch_eager 600 python bench.py --impl torch-eager --steps 3
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:swin_window_attn_tc12 -c 1 -o "$out/wintc" python tools/prof_targets.py win > "$out/wintc.log" 2>&1
cat "$out/summary.txt"
