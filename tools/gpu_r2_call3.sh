#!/usr/bin/env bash
# Round 2, call 3: full-geometry parity tests + ncu --set full of the kernels to optimise (large launches only).
set -u
out=gpurun_out/r2_call3
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300))" | tee -a "$out/summary.txt"; }
run parity_full 1500 python -m pytest tests/test_parity_full_geometry.py -m gpu -q -x
run glue_tests 300 python -m pytest tests/test_fused_glue.py -m gpu -q
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:swin_window_attn_tc12 -c 2 -o "$out/wintc" python tools/prof_targets.py win > "$out/wintc.log" 2>&1
UNIVS_WIN_TC=0 timeout 600 $NCU -k regex:swin_window_attn_f16x3 -c 2 -o "$out/winmma" python tools/prof_targets.py winmma > "$out/winmma.log" 2>&1
timeout 600 $NCU -k regex:mha_tc_kernel -c 2 -o "$out/mhatc" python tools/prof_targets.py mha > "$out/mhatc.log" 2>&1
timeout 600 $NCU -k regex:gelu_split -c 1 -o "$out/gelu" python tools/prof_targets.py gelu > "$out/gelu.log" 2>&1
ls -la "$out"; cat "$out/summary.txt"
