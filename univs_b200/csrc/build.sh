#!/usr/bin/env bash
# Builds univs_b200/lib/libunivs_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v)
objs=()
for f in common msda swin_window_attn swin_window_attn_tc swin_window_attn_tc2 mha mha_tc mask_einsum mask_einsum_tc mask_einsum_mc gemm_tc elementwise groupnorm swin_glue decoder_glue ${UNIVS_EXTRA_SRCS:-}; do
  [ -f "$here/$f.cu" ] || continue
  "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$out/$f.o" 2> "$out/$f.ptxas.log" || { cat "$out/$f.ptxas.log"; exit 1; }
  sed -i '/Compile time = /d' "$out/$f.ptxas.log"      # keep the tracked resource-usage logs stable from build to build
  objs+=("$out/$f.o")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libunivs_b200.so" "${objs[@]}" -lcudart -lcuda
echo "built $out/libunivs_b200.so"
