#!/usr/bin/env bash
# mha_tc with 144 registers (no spills) + WIN_TC=2 promoted: targeted tests, timings, bench
set -u
out=gpurun_out/r2_call32
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? $(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300)" | tee -a "$out/summary.txt"; }
run tc_tests 300 python -m pytest tests/test_mha_tc.py tests/test_window_attn_tc.py tests/test_model_gpu.py -m gpu -q -x
run mha_check 200 python tests/tools/mha_tc_check.py --time
run bench 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -h "ms" "$out/mha_check.log" | tail -12
grep -ho '"ms_per_step": [0-9.]*' "$out/bench.log"
