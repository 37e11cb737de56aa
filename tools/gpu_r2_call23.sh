#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call23
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-200))" | tee -a "$out/summary.txt"; }
run bench_t1 300 python bench.py --frames 1 --steps 20 --warmup 5 --no-cpu-baseline
run bench_c2 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline
run bench_c5 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline
run bench_c3 300 python bench.py --workload c3 --steps 5 --warmup 3
run bench_c4 300 python bench.py --workload c4 --steps 5 --warmup 3
cat "$out/summary.txt"
