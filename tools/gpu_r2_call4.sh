#!/usr/bin/env bash
# Round 2, call 4: the tcgen05 GEMM on hardware (parity, timing next to the library path, bench) + the re-worked window loader.
set -u
out=gpurun_out/r2_call4
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300))" | tee -a "$out/summary.txt"; }
run gemm_tests 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x
run wintc_tests 600 python -m pytest tests/test_window_attn_tc.py -m gpu -q
run wintc_check 600 python tests/tools/win_tc_check.py --time
run gemm_check 600 python tools/gemm_tc_check.py
run bench_base 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_GEMM_TC=1 run bench_gemmtc 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_WIN_TC=1 run bench_wintc 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_GEMM_TC=1 UNIVS_WIN_TC=1 run bench_both 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:gemm_f16x3_tc -s 2 -c 3 -o "$out/gemmtc" python tools/gemm_tc_check.py > "$out/gemmtc_ncu.log" 2>&1
cat "$out/summary.txt"
