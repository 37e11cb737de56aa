#!/usr/bin/env bash
# window attention v2 (two tail warps): parity + timing against v1
set -u
out=gpurun_out/r2_call30
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? $(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300)" | tee -a "$out/summary.txt"; }
run wintc_tests 300 python -m pytest tests/test_window_attn_tc.py -m gpu -q -x
run check_v1 200 python tests/tools/win_tc_check.py --time --ver 0
run check_v2 200 python tests/tools/win_tc_check.py --time --ver 1
run bench_v1 300 env UNIVS_WIN_TC=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_v2 300 env UNIVS_WIN_TC=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -h "stage" "$out/check_v1.log" "$out/check_v2.log"
grep -ho '"ms_per_step": [0-9.]*' "$out/bench_v1.log" "$out/bench_v2.log"
