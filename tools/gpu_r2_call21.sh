#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call21
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run smoke 300 python __graft_entry__.py smoke
run gpu_tests 1800 python -m pytest tests -m gpu -q
run bench_ns 900 python bench.py --steps 20 --warmup 5
run bench_ref 900 python bench.py --impl reference --steps 2 --warmup 1
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file "$out/launches.csv" \
    python bench.py --ncu-step --no-cpu-baseline > "$out/launches.log" 2>&1
python tools/summarize_launches.py "$out/launches.csv" 40 > "$out/launches_summary.txt" 2>&1
cat "$out/summary.txt"
