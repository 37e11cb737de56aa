#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call15
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run bench_t1 300 python bench.py --frames 1 --steps 20 --warmup 5 --no-cpu-baseline
run bench_t2 300 python bench.py --frames 2 --steps 20 --warmup 5 --no-cpu-baseline
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file "$out/launches_t1.csv" \
    python bench.py --frames 1 --ncu-step --no-cpu-baseline > "$out/launches_t1.log" 2>&1
python tools/summarize_launches.py "$out/launches_t1.csv" 40 > "$out/launches_t1_summary.txt" 2>&1
cat "$out/launches_t1_summary.txt"
cat "$out/summary.txt"
