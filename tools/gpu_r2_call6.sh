#!/usr/bin/env bash
# Round 2, call 6: graph capture restored for the fused-glue step; GEMM_TC with compact operands + fused residuals; candidates
# for promotion validated at full geometry; MSDeformAttn tile sweep; smoke().
set -u
out=gpurun_out/r2_call6
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(grep -o '"execution": "[^"]*"' "$out/$name.log" | head -1) $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-200))" | tee -a "$out/summary.txt"; }
run smoke 300 python __graft_entry__.py smoke
run gpu_tests 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_full_geometry.py
run bench_base 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_GEMM_TC=1 run bench_gemmtc 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_GEMM_TC=1 UNIVS_WIN_TC=1 run bench_both 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
UNIVS_GEMM_TC=1 UNIVS_WIN_TC=1 run parity_full_candidates 1500 python -m pytest tests/test_parity_full_geometry.py -m gpu -q
run msda_sweep 300 python tools/msda_tile_sweep.py
NCU="ncu --clock-control none"
UNIVS_GEMM_TC=1 UNIVS_WIN_TC=1 timeout 900 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file "$out/launches.csv" \
    python bench.py --ncu-step --no-cpu-baseline > "$out/launches.log" 2>&1
python tools/summarize_launches.py "$out/launches.csv" 60 > "$out/launches_summary.txt" 2>&1
timeout 600 $NCU --set full --import-source on -k regex:swin_window_attn_tc12 -c 1 -o "$out/wintc" python tools/prof_targets.py win > "$out/wintc.log" 2>&1
cat "$out/summary.txt"
