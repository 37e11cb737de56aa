#!/usr/bin/env bash
set -u
out=gpurun_out/r2_call10
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run smoke 300 python __graft_entry__.py smoke
run gpu_tests 1800 python -m pytest tests -m gpu -q
run bench_ns 900 python bench.py --steps 20 --warmup 5
run torch_eager 600 python bench.py --impl torch-eager --steps 3
UNIVS_POOLED_MASKS=1 run parity_pooled 900 python -m pytest tests/test_parity_full_geometry.py -m gpu -q -k "north_star or c2 or c5"
UNIVS_POOLED_MASKS=1 run bench_pooled 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_c2 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline
run bench_c5 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline
run bench_video 600 python bench.py --steps 2 --video-frames 12
cat "$out/summary.txt"
