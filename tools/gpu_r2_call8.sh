#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call8
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-200))" | tee -a "$out/summary.txt"; }
run shapes 600 python tools/gemm_tc_shapes.py
run tests 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_full_geometry.py
run bench_ns 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_c3 600 python bench.py --workload c3 --steps 5 --warmup 3
run parity_ns 900 python -m pytest tests/test_parity_full_geometry.py -m gpu -q -k "north_star or c3"
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:gemm_f16x3_tc -s 13 -c 1 -o "$out/gemm_add" python tools/gemm_tc_shapes.py > "$out/gemm_add.log" 2>&1
timeout 600 $NCU -k regex:gemm_f16x3_tc -s 39 -c 1 -o "$out/gemm_gelu" python tools/gemm_tc_shapes.py > "$out/gemm_gelu.log" 2>&1
cat "$out/summary.txt"
