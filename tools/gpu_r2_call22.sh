#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call22
mkdir -p "$out"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3_tc -c 4 -o "$out/gemm_pair" python tools/prof_gemm_pair.py > "$out/ncu.log" 2>&1
tail -3 "$out/ncu.log"
ncu -i "$out/gemm_pair.ncu-rep" --page raw --csv > "$out/gemm_pair_raw.csv" 2>/dev/null
ls -la "$out"
