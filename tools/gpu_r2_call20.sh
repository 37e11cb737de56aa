#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call20
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run gemm_tests 600 python -m pytest tests/test_gemm_tc_gpu.py -q -x
run bench_conv_off 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run bench_conv_on 600 env UNIVS_CONV_FUSED=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run parity_ns_conv 900 env UNIVS_CONV_FUSED=1 python -m pytest tests/test_parity_full_geometry.py -q -x -k ns
cat "$out/summary.txt"
