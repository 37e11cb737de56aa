#!/usr/bin/env bash
# Round 2, call 5: GEMM epilogue v2 (parity, timing, bench, ncu), compat / boundary tests after the in-kernel level tables.
set -u
out=gpurun_out/r2_call5
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300))" | tee -a "$out/summary.txt"; }
run gemm_tests 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_compat_msda_module.py tests/test_ops_gpu.py -m gpu -q
run gemm_check 600 python tools/gemm_tc_check.py
UNIVS_GEMM_TC=1 UNIVS_WIN_TC=1 run bench_both 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:gemm_f16x3_tc -s 2 -c 4 -o "$out/gemmtc" python tools/gemm_tc_check.py > "$out/gemmtc_ncu.log" 2>&1
cat "$out/summary.txt"
