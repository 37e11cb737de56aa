"""bench.py arithmetic that can be checked without a GPU: the algorithmic-byte formulas of the four named kernels
(SURVEY.md 8d figures at the north-star shape) and the JSON contract of the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_named_kernel_algorithmic_bytes_match_survey_figures():
    ms = {"mask_einsum": 1.0, "ms_deform_attn_encoder": 1.0, "swin_window_attention": 1.0, "mha": 1.0}
    calls = {k: 1 for k in ms}
    r = bench.named_kernel_rooflines("large", 1, 720, 1280, 200, ms, calls, 1000.0, dec_layers=0, enc_layers=1)
    assert abs(r["mask_einsum"]["bytes_per_step"] / 1e6 - 107.6) < 0.1          # per frame per call
    assert abs(r["ms_deform_attn"]["bytes_per_step"] / 1e6 - 61.8) < 0.1        # per frame per layer
    r = bench.named_kernel_rooflines("large", 5, 720, 1280, 200, ms, dict(calls, mask_einsum=10), 1000.0)
    assert abs(r["mask_einsum"]["bytes_per_step"] / 1e9 - 5.38) < 0.01          # x10 calls x T=5
    # pooled intermediate heads: one full-resolution call + nine at the memory resolutions (1/32, 1/16, 1/8 in turn)
    pooled = bench.named_kernel_rooflines("large", 5, 720, 1280, 200, dict(ms, mask_einsum_pooled=1.0),
                                          dict(calls, mask_einsum=1, mask_einsum_pooled=9), 1000.0)
    assert abs(pooled["mask_einsum"]["bytes_per_step"] / 1e9 - 0.538) < 0.001
    S = [23 * 40, 46 * 80, 92 * 160]
    want = sum(4 * 5 * (200 * 256 + 256 * S[i % 3] + 200 * S[i % 3]) for i in range(9))
    assert pooled["mask_einsum_pooled"]["bytes_per_step"] == want
    # stage 1 of Swin-L at 736x1280: 184x320 tokens padded to 192x324, C = 192 -> 191 MB per block per frame
    stage1 = 4 * 4 * 192 * 324 * 192
    assert abs(stage1 / 1e6 - 191) < 0.5
    assert r["swin_window_attention"]["bytes_per_step"] > 2 * 5 * stage1
    # achieved = bytes / time, frac = achieved / peak
    one = r["ms_deform_attn"]
    assert abs(one["achieved"] - one["bytes_per_step"] / 1e-3 / 1e9) < 1e-6 and abs(one["frac"] - one["achieved"] / 1000.0) < 1e-12
    # a kernel that was not bracketed is simply absent
    assert "decoder_attention" not in bench.named_kernel_rooflines("tiny", 5, 480, 864, 100, {"mask_einsum": 1.0}, {}, 1.0)


def test_committed_round1_line_reproduces_the_documented_fractions():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_ns_fp16x3_1gpu.json")))
    km = {k: v["ms_per_launch"] for k, v in d["kernels"].items()}
    kc = {k: v["launches_per_step"] for k, v in d["kernels"].items()}
    r = bench.named_kernel_rooflines("large", 5, 720, 1280, 200, km, kc, d["roofline"]["peak"])
    assert abs(r["mask_einsum"]["frac"] - d["roofline"]["frac"]) < 1e-6
    assert 0.10 < r["ms_deform_attn"]["frac"] < 0.12 and 0.17 < r["swin_window_attention"]["frac"] < 0.21


def test_reference_arm_contract_on_non_zero_ranks():
    """Under torchrun only rank 0 runs the CPU reference; the other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_round2_report_promotes_only_green_and_faster_switches(tmp_path):
    """tools/r2_open_report.py on a synthetic result directory"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("r2_open_report", os.path.join(ROOT, "tools", "r2_open_report.py"))
    rep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rep)
    d = tmp_path
    (d / "summary.txt").write_text(
        "=== default_gpu_tests: python -m pytest\n    exit 0 (98 passed)\n=== wintc_tests: x\n    exit 0 (ok)\n=== wintc_check: x\n    exit 0 (ok)\n"
        "=== mhatc_tests: x\n    exit 1 (boom)\n=== mhatc_check: x\n    exit 0 (ok)\n=== glue_tests: x\n    exit 0 (ok)\n"
        "=== rowwise_v2_tests: x\n    exit 0 (ok)\n=== parity_at_scale: x\n    exit 0 (ok)\n")
    line = lambda ms, reasons=(): json.dumps({"ms_per_step": ms, "value": 5e3 / ms, "unit": "frames/s", "clocks": {"reasons": list(reasons)}})
    (d / "bench_default.log").write_text("noise\n" + line(52.9) + "\n")
    (d / "bench_wintc.log").write_text(line(49.5) + "\n")                 # green + faster
    (d / "bench_mhatc.log").write_text(line(51.0) + "\n")                 # faster but its tests failed
    (d / "bench_glue.log").write_text(line(47.0, ["hw_thermal_slowdown"]) + "\n")     # throttled run: void
    (d / "bench_rowwise_v2.log").write_text(line(52.85) + "\n")           # green but no real gain
    codes, base, rows, promote = rep.report(str(d), 0.3)
    assert base["ms_per_step"] == 52.9 and codes["mhatc_tests"] == 1
    assert promote == {"WIN_TC": 1}


def test_dense_layer_roofline_is_algorithmic_flops_over_event_time():
    """bench.py `roofline` (dominant kernel): achieved = sum of 2*M*N*K of a step's dense-layer launches / their CUDA-event time;
    `executed` counts the three fp16 products the strict policy issues per algorithmic product."""
    r = bench.dense_layer_roofline({"gemm_f16x3_tc": 0.08}, {"gemm_f16x3_tc": 250}, 7.0e12)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["kernel"] == "gemm_f16x3_tc"
    assert abs(r["ms_per_step"] - 20.0) < 1e-9 and abs(r["achieved"] - 350.0) < 1e-6
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and abs(r["executed"] - 3 * r["achieved"]) < 1e-9
    assert abs(r["frac_executed"] - 3 * r["frac"]) < 1e-12 and r["traffic"] is None
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        assert r["peak"] == json.load(open(peaks_file))["bf16_tflops_sustained"] and r["peak_kind"].startswith("measured")
    # no dense-layer brackets (library-GEMM policies, CPU dry run): no object
    assert bench.dense_layer_roofline({"mask_einsum": 0.1}, {"mask_einsum": 1}, 0) is None
