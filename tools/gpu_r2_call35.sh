#!/usr/bin/env bash
# validation of the session's defaults (window attention v2 + compact operand, 12-warp mha_tc, GEMM without local Args):
# smoke, the whole -m gpu suite, the default bench line, the ncu launch list of one step, the other workloads
set -u
out=gpurun_out/r2_call35
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run smoke 300 python __graft_entry__.py smoke
run gpu_tests 1500 python -m pytest tests -m gpu -q
run bench_ns 600 python bench.py --steps 20 --warmup 5
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file "$out/launches.csv" \
    python bench.py --ncu-step --no-cpu-baseline > "$out/launches.log" 2>&1
python tools/summarize_launches.py "$out/launches.csv" 40 > "$out/launches_summary.txt" 2>&1
run bench_c2 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline
run bench_c5 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline
run bench_c3 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline
run bench_c4 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline
run bench_t1 300 python bench.py --frames 1 --steps 20 --warmup 5 --no-cpu-baseline
cat "$out/summary.txt"
