#!/usr/bin/env bash
# ncu source-level captures (stall samples per SASS line) of the window-attention and decoder-attention tcgen05 kernels
set -u
out=gpurun_out/r2_call24
mkdir -p "$out"
NCU="ncu --clock-control none"
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 $NCU --set full --import-source on -k "regex:$rx" -s "$skip" -c 1 -o "$out/$name" "$@" > "$out/$name.log" 2>&1
  ncu -i "$out/$name.ncu-rep" --page raw --csv > "$out/${name}_raw.csv" 2>/dev/null
  ncu -i "$out/$name.ncu-rep" --page source --csv --print-source sass > "$out/${name}_sass.csv" 2>/dev/null
  ncu -i "$out/$name.ncu-rep" --page details > "$out/${name}_details.txt" 2>/dev/null
}
cap wintc 'swin_window_attn_tc12' 1 python tools/prof_targets.py win
cap mhatc 'mha_tc_kernel' 1 python tools/prof_targets.py mha
ls -la "$out"
