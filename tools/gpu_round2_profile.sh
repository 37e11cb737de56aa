#!/usr/bin/env bash
# Second GPU call of round 2 (after tools/gpu_round2_open.sh is green): ncu evidence for the kernels that became default.
# One GPU, every capture under `timeout`; reports land in gpurun_out/r2_prof/ (copy the summaries into profiles/).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round2_profile.sh'
# Environment switches of the paths to profile are taken from the caller, e.g.
#   UNIVS_FUSED_GLUE=1 UNIVS_MSDA_TILE=8 UNIVS_WIN_TC=1 UNIVS_MHA_TC=1 UNIVS_ROWWISE_V2=1 bash tools/gpu_round2_profile.sh
set -u
out=gpurun_out/r2_prof
mkdir -p "$out"
NCU="ncu --clock-control none"
# 1. launch list of ONE step (cold-cache, serialised: compare shares, not absolutes)
timeout 1200 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file "$out/launches.csv" \
    python bench.py --ncu-step --no-cpu-baseline > "$out/launches.log" 2>&1
python tools/summarize_launches.py "$out/launches.csv" 40 > "$out/launches_summary.txt" 2>&1
# 2. full captures of the named kernels (3 launches each; -lineinfo is in the build, so the source page maps to csrc/)
cap() {  # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  timeout 900 $NCU --set full --import-source on -k "regex:$rx" -c 3 -o "$out/$name" "$@" > "$out/$name.log" 2>&1
  ncu -i "$out/$name.ncu-rep" --page raw --csv 2>/dev/null | \
    grep -E 'Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput|sm__pipe_tensor.*cycles_active|sm__throughput|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|launch__registers_per_thread|sm__warps_active' \
    > "$out/$name.summary.csv"
}
cap wintc 'swin_window_attn_tc12' python tests/tools/win_tc_check.py --time
cap winmma 'swin_window_attn_f16x3' python tests/tools/win_tc_check.py --time
cap einsum 'mask_einsum_tc' python tools/einsum_tc_check.py
cap mhatc 'mha_tc_kernel' python bench.py --ncu-step --no-cpu-baseline
cap msda 'msda_encoder' python bench.py --ncu-step --no-cpu-baseline
cap gelu 'gelu_split' python bench.py --ncu-step --no-cpu-baseline
ls -la "$out"
