#!/usr/bin/env bash
# streaming LayerNorm (UNIVS_ROWWISE_V2 bit 2): bit identity + bench against the default
set -u
out=gpurun_out/r2_call36
mkdir -p "$out"
run() { local name=$1 secs=$2; shift 2
  echo "=== $name: $*" | tee -a "$out/summary.txt"
  ( time timeout "$secs" "$@" ) > "$out/$name.log" 2>&1
  echo "    exit $? ($(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -2 | tr '\n' ' ') $(tail -n 3 "$out/$name.log" | tr '\n' ' ' | cut -c1-220))" | tee -a "$out/summary.txt"; }
run rowwise_tests 300 python -m pytest tests/test_rowwise_v2.py -m gpu -q -x
run bench_v3 300 env UNIVS_ROWWISE_V2=3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run bench_v7 300 env UNIVS_ROWWISE_V2=7 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
python - <<'PY'
import json,re
for n in ("bench_v3","bench_v7"):
    t=open(f"gpurun_out/r2_call36/{n}.log").read()
    m=re.findall(r'^\{.*\}$', t, re.M)
    if m:
        d=json.loads(m[-1]); k=d.get("kernels",{})
        print(n, d["ms_per_step"], {x:(round(k[x]["ms_per_launch"]*k[x]["launches_per_step"],3)) for x in k if "layernorm" in x})
PY
