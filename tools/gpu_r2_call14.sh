#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call14
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run gemm_mc 240 python tools/gemm_tc_mc.py
run bench_mc_auto 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
UNIVS_GEMM_MC=0 run bench_mc_off 600 env UNIVS_GEMM_MC=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
UNIVS_GEMM_MC=1 run bench_mc_on 600 env UNIVS_GEMM_MC=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
cat "$out/gemm_mc.log" | tail -20
cat "$out/summary.txt"
