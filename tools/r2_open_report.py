#!/usr/bin/env python
"""Reads what tools/gpu_round2_open.sh left in gpurun_out/r2_open/ and prints, per opt-in switch, whether its gated GPU tests
passed and what its bench line says against the default path -- and, with --write, puts the switches that are green AND
faster into univs_b200/tuned.json (the file univs_b200/switches.py reads after the environment).

    python tools/r2_open_report.py [--dir gpurun_out/r2_open] [--min-gain 0.3] [--write]

A switch is promoted only if (a) every test step that covers it exited 0, (b) its bench run printed a line, saw no thermal /
hardware slowdown, and beats the default path's ms_per_step by at least --min-gain percent.  Parity of everything together is
the parity_at_scale step: if that step failed, nothing is written."""
import argparse
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# switch -> (value to promote, test steps that must be green, bench step)
SWITCHES = {
    "FUSED_GLUE": (1, ["glue_tests"], "bench_glue"),
    "MSDA_TILE": (8, ["glue_tests"], "bench_glue"),
    "WIN_TC": (1, ["wintc_tests", "wintc_check"], "bench_wintc"),
    "MHA_TC": (1, ["mhatc_tests", "mhatc_check"], "bench_mhatc"),
    "ROWWISE_V2": (3, ["rowwise_v2_tests"], "bench_rowwise_v2"),
    "EINSUM_MC": (1, ["einsum_mc_tests", "einsum_mc_check"], "bench_einsum_mc"),
    "POOLED_MASKS": (1, ["glue_tests"], "bench_pooled_masks"),
    "MLP_CHUNK_MB": (96, [], "bench_mlp_chunk"),
    "FRAME_STREAMS": (2, [], "bench_streams2"),
}


def exit_codes(summary_path):
    codes, name = {}, None
    for line in open(summary_path):
        m = re.match(r"=== (\w+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+exit (\d+)", line)
        if m and name:
            codes[name] = int(m.group(1))
    return codes


def bench_line(log_path):
    if not os.path.exists(log_path):
        return None
    for line in reversed(open(log_path).read().splitlines()):
        line = line.strip()
        if line.startswith("{") and '"ms_per_step"' in line:
            try:
                return json.loads(line)
            except ValueError:
                continue
    return None


def throttled(line):
    reasons = ((line or {}).get("clocks") or {}).get("reasons") or []
    return any(r in reasons for r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"))


def report(directory, min_gain):
    codes = exit_codes(os.path.join(directory, "summary.txt"))
    base = bench_line(os.path.join(directory, "bench_default.log"))
    rows, promote = [], {}
    for name, (value, tests, bench) in SWITCHES.items():
        green = all(codes.get(t) == 0 for t in tests)
        line = bench_line(os.path.join(directory, bench + ".log"))
        ms = line["ms_per_step"] if line else None
        gain = (1 - ms / base["ms_per_step"]) * 100 if (line and base) else None
        ok = green and line is not None and not throttled(line) and gain is not None and gain >= min_gain
        rows.append((name, value, {t: codes.get(t) for t in tests}, ms, gain, ok))
        if ok:
            promote[name] = value
    return codes, base, rows, promote


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=os.path.join(ROOT, "gpurun_out", "r2_open"))
    ap.add_argument("--min-gain", type=float, default=0.3, help="percent of ms_per_step a switch must save to be promoted")
    ap.add_argument("--write", action="store_true", help="write the promoted switches to univs_b200/tuned.json")
    args = ap.parse_args()
    codes, base, rows, promote = report(args.dir, args.min_gain)
    print(f"default path: {base['ms_per_step']:.2f} ms/step ({base['value']:.1f} {base['unit']})" if base else "default path: NO BENCH LINE")
    print(f"default GPU suite: exit {codes.get('default_gpu_tests')}")
    for name, value, tests, ms, gain, ok in rows:
        ms_s = f"{ms:.2f} ms" if ms is not None else "no line"
        gain_s = f"{gain:+.1f} %" if gain is not None else "-"
        print(f"  {name}={value:<3} tests {tests}  bench {ms_s} ({gain_s})  -> {'PROMOTE' if ok else 'keep opt-in'}")
    both = bench_line(os.path.join(args.dir, "bench_all.log"))
    if both and base:
        print(f"all switches together: {both['ms_per_step']:.2f} ms/step ({(1 - both['ms_per_step'] / base['ms_per_step']) * 100:+.1f} %)")
    parity_ok = codes.get("parity_at_scale") == 0
    print(f"parity at scale with everything on: exit {codes.get('parity_at_scale')}")
    if args.write:
        if not parity_ok or codes.get("default_gpu_tests") != 0:
            sys.exit("not writing tuned.json: the default GPU suite or the combined parity step is not green")
        path = os.path.join(ROOT, "univs_b200", "tuned.json")
        json.dump(promote, open(path, "w"), indent=1, sort_keys=True)
        print(f"wrote {path}: {promote}")


if __name__ == "__main__":
    main()
