#!/usr/bin/env bash
# Round 2, call 2: promoted defaults (tuned.json) -- full GPU suite without gates, per-switch parity at the north-star
# geometry, launch list and ncu --set full captures of the kernels to optimise next.
set -u
out=gpurun_out/r2_call2
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(tail -n 4 "$out/$name.log" | tr '\n' ' ' | cut -c1-300))" | tee -a "$out/summary.txt"; }
run gpu_tests 1500 python -m pytest tests -m gpu -q
run bench 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
PARITY_MODES=fp16x3 run parity_promoted 600 python tests/tools/parity_at_scale.py
cp gpurun_out/parity_at_scale_large_T2.json "$out/parity_promoted.json" 2>/dev/null
PARITY_MODES=fp16x3 UNIVS_POOLED_MASKS=1 run parity_pooled 600 python tests/tools/parity_at_scale.py
cp gpurun_out/parity_at_scale_large_T2.json "$out/parity_pooled.json" 2>/dev/null
PARITY_MODES=fp16x3 UNIVS_WIN_TC=1 run parity_wintc 600 python tests/tools/parity_at_scale.py
cp gpurun_out/parity_at_scale_large_T2.json "$out/parity_wintc.json" 2>/dev/null
NCU="ncu --clock-control none"
timeout 1200 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file "$out/launches.csv" \
    python bench.py --ncu-step --no-cpu-baseline > "$out/launches.log" 2>&1
python tools/summarize_launches.py "$out/launches.csv" 60 > "$out/launches_summary.txt" 2>&1
cap() { local name=$1 rx=$2; shift 2
  timeout 900 $NCU --set full --import-source on -k "regex:$rx" -c 2 -o "$out/$name" "$@" > "$out/$name.log" 2>&1; }
cap wintc 'swin_window_attn_tc12' python tests/tools/win_tc_check.py --time
cap mhatc 'mha_tc_kernel' python tests/tools/mha_tc_check.py --time
cap msda 'msda_encoder' python bench.py --ncu-step --no-cpu-baseline
cap gelu 'gelu_split' python bench.py --ncu-step --no-cpu-baseline
cap layernorm 'layernorm' python bench.py --ncu-step --no-cpu-baseline
ls -la "$out"
cat "$out/summary.txt"
