#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call7
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-200))" | tee -a "$out/summary.txt"; }
run shapes 600 python tools/gemm_tc_shapes.py
run tests 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_mha_tc.py -m gpu -q
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k "regex:gemm_f16x3_tc_kernel<0, 1, 0, 1>" -c 1 -o "$out/gemm_add" python tools/gemm_tc_shapes.py > "$out/gemm_add.log" 2>&1
timeout 600 $NCU -k "regex:gemm_f16x3_tc_kernel<1, 0, 1, 0>" -c 1 -o "$out/gemm_gelu" python tools/gemm_tc_shapes.py > "$out/gemm_gelu.log" 2>&1
run bench_c3 600 python bench.py --workload c3 --steps 5 --warmup 3
run bench_c4 600 python bench.py --workload c4 --steps 5 --warmup 3
cat "$out/summary.txt"
