#!/usr/bin/env bash
# ncu source-level capture of window attention v2
set -u
out=gpurun_out/r2_call34
mkdir -p "$out"
export UNIVS_WIN_TC=2
timeout 300 ncu --clock-control none --set full --import-source on -k "regex:swin_window_attn_tc12v2" -s 1 -c 1 -o "$out/wintc2" python tools/prof_targets.py win > "$out/wintc2.log" 2>&1
ncu -i "$out/wintc2.ncu-rep" --page raw --csv > "$out/wintc2_raw.csv" 2>/dev/null
ncu -i "$out/wintc2.ncu-rep" --page source --csv --print-source sass > "$out/wintc2_sass.csv" 2>/dev/null
ls -la "$out"
