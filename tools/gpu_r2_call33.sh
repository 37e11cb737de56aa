#!/usr/bin/env bash
# GEMM producer without the local Args copy: GEMM tests + bench
set -u
out=gpurun_out/r2_call33
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? $(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300)" | tee -a "$out/summary.txt"; }
run gemm_tests 300 python -m pytest tests/test_gemm_tc_gpu.py tests/test_model_gpu.py -m gpu -q -x
run bench 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
grep -ho '"ms_per_step": [0-9.]*' "$out/bench.log"
