#!/usr/bin/env python
"""Benchmark of the UniVS per-clip forward hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ns|c2|c5] [--precision fp32|tf32]

One step = one per-clip forward (normalise+pad -> Swin backbone -> MSDeformAttn pixel decoder -> UniVS decoder ->
mask logits) over one synthetic clip.  Default workload = the configuration BASELINE.json's metric is quoted on:
Swin-L, T=5, 720x1280 (padded to 736x1280), Q=200, detection task without prompts.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (swin variant, T, H, W, Q)
    "ns": ("large", 5, 720, 1280, 200),      # north-star / metric configuration
    "c2": ("tiny", 5, 480, 864, 100),        # BASELINE.json configs[1]
    "c5": ("large", 8, 1080, 1920, 200),     # BASELINE.json configs[4]: meant for --gpus 8 (one frame per GPU); fits one GPU too
}
# prompt configurations (BASELINE.json configs[2], configs[3]): eager execution (data-dependent host logic), N=1
PROMPT_WORKLOADS = {
    # name: (swin variant, T, H, W, Q, task, clips that grow the prompt memory before the timed clip)
    "c3": ("base", 5, 720, 1280, 200, "sot", 2),          # P=10 visual prompts, R=128 points, memory of 2 earlier clips
    "c4": ("large", 10, 720, 1280, 200, "grounding", 0),  # P=32 text prompts + lang->vision
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def tensor_peak():
    """Dense tensor denominator for a kernel timed inside a long step: the SUSTAINED cuBLAS bf16 rate (B200_PROFILING.md)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured (cuBLAS bf16, sustained under the power cap)"
    return 1400.0, "fallback (sustained)"


def dense_layer_roofline(kernel_ms, kernel_calls, flops_per_step):
    """Tensor roofline of the dense-layer kernel (csrc/gemm_tc.cu), the kernel most of the step's time is spent in.
    `achieved` = ALGORITHMIC flops (2*M*N*K of every launch of a step: the fp32 products the reference computes) over the
    launches' CUDA-event time; `executed` = the fp16 MMA flops the strict hi|lo policy issues for them (3 products per
    algorithmic one) -- the number that says how busy the tensor pipe is."""
    if "gemm_f16x3_tc" not in kernel_ms or not flops_per_step:
        return None
    peak, kind = tensor_peak()
    ms = kernel_ms["gemm_f16x3_tc"] * kernel_calls["gemm_f16x3_tc"]
    ach = flops_per_step / (ms * 1e-3) / 1e12
    return {"kernel": "gemm_f16x3_tc", "bound": "tensor", "achieved": ach, "peak": peak, "peak_kind": kind, "unit": "TFLOP/s",
            "frac": ach / peak, "executed": 3 * ach, "frac_executed": 3 * ach / peak,
            "note": "strict fp16 hi|lo policy: 3 tensor-core products per fp32-equivalent product (parity bound 1e-3 rules out 1 pass)",
            "launches_per_step": kernel_calls["gemm_f16x3_tc"], "ms_per_step": ms,
            "algorithmic_flops_per_step": float(flops_per_step), "traffic": None}


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        clocks, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                clocks.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(clocks) if clocks else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(clocks)}


def named_kernel_rooflines(variant, T, H, W, Q, kernel_ms, kernel_calls, peak, dec_layers=9, enc_layers=6, sm_mhz=1965.0,
                           T_local=None, T_dec=None):
    """Algorithmic HBM bytes (SURVEY.md 8(d) formulas, fp32 element size) of the four kernels BASELINE.json names that THIS
    RANK's launches move, divided by their CUDA-event time on this rank.  `T_local` = frames this rank runs through
    backbone + pixel decoder (frame sharding: ceil(T / world) on rank 0), `T_dec` = frames its decoder kernels see (T when
    the decoder is replicated, T_local when it is frame-sharded).  Pure arithmetic on the per-kernel brackets; returns
    {name: {"bytes_per_step", "ms_per_step", "achieved" (GB/s), "frac"}}."""
    T_local = T if T_local is None else T_local
    T_dec = T if T_dec is None else T_dec
    from univs_b200.config import SWIN_VARIANTS
    sw = SWIN_VARIANTS[variant]
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    e, C = 4, 256
    out = {}

    def put(name, key, nbytes):
        if key in kernel_ms and kernel_ms[key] > 0:
            ms = kernel_ms[key] * kernel_calls.get(key, 1)
            ach = nbytes / (ms * 1e-3) / 1e9
            out[name] = {"bytes_per_step": nbytes, "ms_per_step": ms, "achieved": ach, "frac": ach / peak}

    # mask einsum: e*(Q*C + C*HW + Q*HW) per frame per call, dec_layers + 1 calls
    HW = (Hp // 4) * (Wp // 4)
    # full-resolution launches only: with pooled intermediate heads (SURVEY 8a note a11') nine of the ten head calls run on
    # the pooled features at the memory resolutions (reported separately); the last one writes the full-resolution logits
    n_full = kernel_calls.get("mask_einsum", dec_layers + 1)
    put("mask_einsum", "mask_einsum", float(e * T_dec * (Q * C + C * HW + Q * HW) * n_full))
    if "mask_einsum_pooled" in kernel_ms:
        Sp = [(Hp // s) * (Wp // s) for s in (32, 16, 8)]
        calls = kernel_calls.get("mask_einsum_pooled", dec_layers)
        # head i (0-based, i < dec_layers) feeds layer i, whose memory level is i % 3
        put("mask_einsum_pooled", "mask_einsum_pooled",
            float(sum(e * T_dec * (Q * C + C * Sp[i % 3] + Q * Sp[i % 3]) for i in range(calls))))
    # MSDeformAttn core: e*(2*Len*256 + 3*Len*M*L*P) per frame per layer, M*L*P = 96
    Len = sum((Hp // s) * (Wp // s) for s in (8, 16, 32))
    put("ms_deform_attn", "ms_deform_attn_encoder", float(e * T_local * (2 * Len * C + 3 * Len * 96) * enc_layers))
    if "ms_deform_attn" in out:
        # the gather is bound by L1 request throughput long before HBM: every (query, head) fetches L*P*4 = 48 corners of
        # 32 fp32 channels = 48 requests of one 128-byte line, and an SM's L1 serves one 128-byte wavefront per clock
        wavefronts = float(T_local * Len * 8 * 48 * enc_layers)
        floor_ms = wavefronts / (148 * sm_mhz * 1e6) * 1e3
        out["ms_deform_attn"].update({"l1_wavefronts_per_step": wavefronts, "l1_floor_ms_per_step": floor_ms,
                                      "frac_of_l1_floor": floor_ms / out["ms_deform_attn"]["ms_per_step"]})
    # Swin window attention: e*4*nW*N*C per block per frame (q, k, v in, out), windows padded to the window size
    ws, total = sw["WINDOW_SIZE"], 0
    for s, depth in enumerate(sw["DEPTHS"]):
        hs, wsz = Hp // (4 << s), Wp // (4 << s)
        tokens = (-(-hs // ws) * ws) * (-(-wsz // ws) * ws)
        total += depth * e * 4 * tokens * (sw["EMBED_DIM"] << s)
    put("swin_window_attention", "swin_window_attention", float(total * T_local))
    # decoder attention: cross-attention e*(2*S_l*256 + 2*Q*256) per frame per layer (level l = layer % 3, 1/32 first)
    # + the Q*T self-attention e*4*Q*T*256 per layer
    S = [(Hp // s) * (Wp // s) for s in (32, 16, 8)]
    mha = sum(e * T_dec * (2 * S[i % 3] * C + 2 * Q * C) + e * 4 * Q * T * C for i in range(dec_layers))
    put("decoder_attention", "mha", float(mha))
    return out


def make_targets(T, device):
    return [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual",
             "frame_indices": torch.arange(T, device=device)}]


def run_cpu_reference(args, workload, steps, warmup, as_line):
    """The reference algorithm on the host cores: product host logic with every operator replaced by its CPU oracle
    (oracle/ops_ref.py; the reference tree itself does not travel to the GPU box), fp32, all host threads.
    Sample: ONE full clip of the workload (all T frames, the same configuration as the GPU arm) per step; the number of
    timed steps is bounded by a wall-clock budget (UNIVS_CPU_BUDGET_S, default 240 s: the CPU needs ~18 s per north-star
    clip), and at least one timed step always runs."""
    from oracle.cpu_backend import oracle_ops
    from univs_b200.build import build_model, make_cfg
    variant, T, H, W, Q = WORKLOADS[workload]
    # the torch CPU path scales to ~32 threads on this workload and slows down beyond (measured on the 128-thread
    # GPU-box host: 16 thr 8.7 s/frame, 32 thr 7.7 s, 64 thr 13.2 s, 128 thr 115 s) -- use what it can use
    cores = min(os.cpu_count() or 1, int(os.environ.get("UNIVS_CPU_THREADS", "32")))
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    Ts = int(os.environ.get("UNIVS_CPU_SAMPLE_FRAMES", T))              # frames per sampled clip (default: the whole clip)
    cfg = make_cfg(variant, Q, Ts, clip_emb=torch.randn(3938, 640, generator=g), TEXT_PROMPT_TO_IMAGE_ENABLE=False)
    model = build_model(cfg)
    frames = torch.rand(Ts, 3, H, W, generator=g) * 255
    times, budget_s, t_begin = [], float(os.environ.get("UNIVS_CPU_BUDGET_S", "240")), time.perf_counter()
    with oracle_ops():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = model.clip_forward(frames, make_targets(Ts, "cpu"))
            float(out["pred_masks"].sum())
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            # the whole run has to end within a few minutes on any host: stop early once one timed step exists
            if times and time.perf_counter() - t_begin + dt > budget_s:
                break
            if not times and time.perf_counter() - t_begin + 2 * dt > budget_s:
                warmup = i + 1                       # slow host: the next step is the (first) timed one
    mean = sum(times) / len(times)
    steps = len(times)
    sample = (f"{steps} timed step(s) of 1 clip of T={Ts} frames ({variant} {H}x{W}, Q={Q}); the workload's clip has T={T} frames"
              + ("" if Ts == T else " (reduced sample)"))
    base = {"value": Ts / mean, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
            "same_config": Ts == T}
    if not as_line:
        return base
    return {
        "impl": "reference",
        "metric": "frames/sec (Swin-L 720p T=5 Q=200)" if workload == "ns" else "frames/sec (per-clip forward)",
        "value": Ts / mean, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload}: Swin-{variant} T={T} {H}x{W}->pad32 Q={Q} detection, no prompts, random init",
                   "precision": "fp32 (CPU)", "sample": sample},
        "cpu_baseline": base,
        "e2e": {"value": Ts / mean, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_prompt_workload(args, dev):
    """BASELINE configs[2] / configs[3]: one step = the per-clip forward of a clip WITH prompts (ProCA path).  sot: the
    visual-prompt memory has been grown by the preceding stride-1 clips (kv length 1 + R * (1 + frames in the pool)); every
    step starts from a copy of that state.  Runs eagerly (the prompt sampler has data-dependent host control flow)."""
    from univs_b200 import ops, switches
    from univs_b200.build import build_model, make_cfg
    from univs_b200.synthetic import ClipSource, clone_targets, univs_overrides
    variant, T, H, W, Q, task, grow = PROMPT_WORKLOADS[args.workload]
    src = ClipSource(task, T, T + grow, H, W, annotate_all=True)    # later frames: masks fed back by the task head
    cfg = make_cfg(variant, Q, T, clip_emb=src.clip_emb, **univs_overrides(task))
    model = build_model(cfg).to(dev)
    tg = src.targets(dev)
    for c in range(grow):                                       # earlier clips of the video: fill the prompt memory pool
        torch.manual_seed(100 + c)
        model.clip_forward(src.clip_inputs(tg, c, dev), tg)
    frames_dev = src.clip_inputs(tg, grow, dev).to(torch.uint8)
    frames_host = frames_dev.cpu().pin_memory()
    state = clone_targets(tg)
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    def step(frames):
        torch.manual_seed(7)
        return model.clip_forward(frames, clone_targets(state))

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(warmup):
        out = step(frames_dev)
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(lambda: step(frames_dev), steps)
    clocks = sampler.stop()
    sink = ops.profile_events(True)
    l0 = ops.launch_count
    timed(lambda: step(frames_dev), steps)
    launches = (ops.launch_count - l0) // steps
    torch.cuda.synchronize()
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in sink.items()}
    kernel_calls = {k: len(v) // steps for k, v in sink.items()}
    ops.profile_events(False)
    pinned = {}

    def step_e2e():
        o = step(frames_host)
        for k, v in (("pred_logits", o["pred_logits"]), ("pred_embds", o["pred_embds"]), ("pred_masks_bin", o["pred_masks"] > 0)):
            if k not in pinned:
                pinned[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
            pinned[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    ms_e2e = timed(step_e2e, steps)
    peak, peak_kind = peaks()
    n_lp = out["pred_masks"].shape[1]
    P = n_lp - Q
    roof = None
    if "proca" in kernel_ms:
        # ProCA: per layer e*2*Q_p*(1+L)*256 (T-invariant memory, read once) + e*2*Q_p*T*256; L from the pool
        pool = state[0].get("prompt_feats")
        L = int(pool.shape[1] * min(pool.shape[2], 1 + cfg.MODEL.UniVS.TEST.NUM_PREV_FRAMES_MEMORY)) if pool is not None else 78
        nbytes = 4.0 * (2 * P * (1 + L) * 256 + 2 * P * T * 256)
        ach = nbytes / (kernel_ms["proca"] * 1e-3) / 1e9
        roof = {"kernel": "proca", "bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "launch_ms": kernel_ms["proca"], "algorithmic_bytes_per_launch": nbytes,
                "prompts": P, "memory_tokens_per_prompt": L,
                "note": "latency-bound by size: a few MB per launch; the K/V in-projection of the memory tokens (library "
                        "/ own GEMM) dominates the ProCA layer"}
    print(json.dumps({
        "metric": "frames/sec (per-clip forward with prompts)", "value": T / (ms / 1e3), "unit": "frames/s", "n_gpus": 1,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{args.workload}: Swin-{variant} T={T} {H}x{W}->pad32 Q={Q} task={task} P={P} prompts"
                               + (f", prompt memory grown over {grow} earlier clips" if grow else ", lang->vision on"),
                   "precision": args.precision, "execution": "eager", "switches": switches.active(),
                   "l2": "per-step working set (multi-GB activations) exceeds the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": T / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": frames_host.numel(),
                "d2h_bytes_per_step": sum(v.numel() * v.element_size() for v in pinned.values()), "ms_per_step": ms_e2e},
        "gpu_launches": launches, "roofline": roof,
        "kernels": {k: {"ms_per_launch": kernel_ms[k], "launches_per_step": kernel_calls[k]} for k in sorted(kernel_ms)},
        "cpu_baseline": None}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ns", choices=list(WORKLOADS) + list(PROMPT_WORKLOADS))
    ap.add_argument("--video-frames", type=int, default=0,
                    help="> 0: secondary benchmark (SURVEY 8f rank 1) -- the VIS sliding-window head over a synthetic video of this "
                         "many frames of the workload's geometry (stride-1 clips of T frames), with per-frame feature reuse "
                         "(ClipStream) and with the reference's schedule; one step = one video; N=1 only")
    ap.add_argument("--precision", default=os.environ.get("UNIVS_PRECISION", "fp16x3"), choices=["fp16x3", "tf32x3", "fp32", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the clip eagerly instead of replaying a CUDA graph")
    ap.add_argument("--shard-decoder", action="store_true",
                    help="N>1: keep the decoder frame-sharded and exchange query tokens per layer instead of all-gathering "
                         "the pixel features (also UNIVS_SHARD_DECODER=1)")
    ap.add_argument("--frames", type=int, default=0,
                    help="diagnostic: override the workload's T (e.g. --frames 1 on one GPU = the compute share of one rank of "
                         "an 8-GPU run, without the exchange); the line's config.workload says so -- not a headline number")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: after warm-up run ONE step between cudaProfilerStart/Stop and exit "
                         "(use with `ncu --profile-from-start off`); prints no bench line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_cpu_reference(args, args.workload, max(1, args.steps), max(0, args.warmup), True)))
        return

    import torch.distributed as dist
    from univs_b200 import ops, switches
    from univs_b200.build import build_model, make_cfg
    from univs_b200.precision import set_precision

    # UNIVS_BENCH_DEVICE exists for tests/test_bench_dryrun.py (control-flow check of this script on a CPU with the oracle
    # operator backend patched in by the test); the operators themselves have no CPU path, so a plain run on "cpu" fails
    dev_type = os.environ.get("UNIVS_BENCH_DEVICE", "cuda")
    if dev_type == "cuda":
        torch.cuda.set_device(local_rank)
    dev = torch.device(dev_type, local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if dev_type == "cuda":
            dist.init_process_group("nccl", device_id=dev)
        else:                                   # tests/test_bench_dryrun.py: the same control flow over gloo on the CPU
            dist.init_process_group("gloo")
        group = dist.group.WORLD
    set_precision(args.precision)
    if args.workload in PROMPT_WORKLOADS:
        if world > 1:
            raise SystemExit("the prompt workloads are single-GPU benchmarks")
        return run_prompt_workload(args, dev)

    variant, T, H, W, Q = WORKLOADS[args.workload]
    if args.frames > 0:
        T = args.frames
    g = torch.Generator().manual_seed(0)
    cfg = make_cfg(variant, Q, T, clip_emb=torch.randn(3938, 640, generator=g), TEXT_PROMPT_TO_IMAGE_ENABLE=False)
    model = build_model(cfg, process_group=group).to(dev)
    if args.shard_decoder:
        model.shard_decoder = True
    frames_host = (torch.rand(T, 3, H, W, generator=g) * 255).to(torch.uint8).pin_memory()
    frames_dev = frames_host.to(dev)
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.video_frames > 0:
        if world > 1:
            raise SystemExit("--video-frames is a single-GPU benchmark")
        from univs_b200.inference import InferenceVideoVISFast
        V = max(args.video_frames, T)
        video = [f for f in (torch.rand(V, 3, H, W, generator=g) * 255).to(torch.uint8).pin_memory()]
        res = {}
        for reuse in (True, False, "rle"):     # "rle": feature reuse + COCO RLE results (run scan on the device)
            head = InferenceVideoVISFast(num_queries=Q, num_frames=T, num_frames_window_test=T, reuse_features=bool(reuse),
                                         rle_output=reuse == "rle")
            run = lambda: head.eval(model, [{"image": video, "height": H, "width": W, "dataset_name": "ytvis21"}])
            run()                                                       # warm-up (weight caches, allocator)
            torch.cuda.synchronize()
            t0 = time.perf_counter()                                    # host wall clock: the head has host control flow
            for _ in range(steps):                                      # (Hungarian matching, result lists); sync'd both sides
                out_v = run()
            torch.cuda.synchronize()
            res[reuse] = (time.perf_counter() - t0) / steps
        print(json.dumps({
            "metric": "video frames/sec (VIS fast head: stride-1 clips, MinVIS tracker, masks at output size)",
            "value": V / res[True], "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": 1,
            "ms_per_step": res[True] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.workload}_video: Swin-{variant} V={V} frames {H}x{W}, T={T}, Q={Q}, "
                                   f"{V - T + 1} clips, frames from pinned host memory, results (bool masks) to the host",
                       "precision": args.precision, "execution": "eager"},
            "reference_schedule": {"value": V / res[False], "unit": "frames/s", "ms_per_step": res[False] * 1e3,
                                   "what": "same head with reuse_features=False: pixel decoder re-run for every clip, as "
                                           "inference_video_vis_fast.py:223-236 does"},
            "rle_results": {"value": V / res["rle"], "unit": "frames/s", "ms_per_step": res["rle"] * 1e3,
                            "what": "feature reuse + results as COCO RLE built from device-side run boundaries instead of "
                                    "dense bool masks copied to the host"},
            "instances": len(out_v["pred_scores"])}))
        return

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident throughput ------------------------------------------------------------------------
    def step_eager():
        return model.clip_forward(frames_dev, make_targets(T, dev))

    for _ in range(warmup):
        out = step_eager()
    n_lp = out["pred_masks"].shape[1]
    if args.ncu_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_eager()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # frame-sharded runs replay a graph too (the NCCL all-gather is captured with the rest of the step; round 1 ran them
    # eagerly and was host-launch-bound: ~900 launches of 16-20 us each per 20 ms step at N=8).  UNIVS_GRAPH_MULTI=0 opts
    # out; a failed capture on any rank sends every rank down the eager path.
    use_graph = not args.no_graph and (world == 1 or os.environ.get("UNIVS_GRAPH_MULTI", "1") == "1")
    graphed = None
    if use_graph:
        from univs_b200.runtime import GraphedClip
        try:
            graphed = GraphedClip(model, frames_dev, lambda: make_targets(T, dev))
            ok = torch.ones(1, device=dev)
        except Exception as e:  # noqa: BLE001  (e.g. a collective that refuses stream capture)
            import traceback
            sys.stderr.write(f"[bench] CUDA-graph capture failed on rank {rank}: {e}\n"
                             + "".join(traceback.format_exc().splitlines(True)[-14:]))
            graphed, ok = None, torch.zeros(1, device=dev)
        if world > 1:                                   # all ranks must take the same path
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            graphed, use_graph = None, False
    if graphed is not None:
        step_resident = lambda: graphed(frames_dev)
    else:
        step_resident = step_eager
    for _ in range(2):
        out = step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / steps
    fps = T / (ms_per_step / 1e3)

    # ---- per-kernel CUDA-event brackets (eager pass: events cannot be read back from inside a graph replay) ---
    sink = ops.profile_events(True)
    launches0 = ops.launch_count
    timed(step_eager, steps)
    launches = (ops.launch_count - launches0) // steps
    torch.cuda.synchronize()
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in sink.items()}
    kernel_calls = {k: len(v) // steps for k, v in sink.items()}
    dense = dense_layer_roofline(kernel_ms, kernel_calls, ops.flop_count.get("gemm_f16x3_tc", 0) // steps)
    ops.profile_events(False)

    # ---- end-to-end through the public API with host buffers ------------------------------------------------
    d2h_pinned = {}

    def step_e2e():
        if graphed is not None:
            o = graphed(frames_host)                                    # H2D of the pinned uint8 frames, then replay
        else:
            o = model.clip_forward(frames_host, make_targets(T, dev))
        res = {"pred_logits": o["pred_logits"], "pred_embds": o["pred_embds"], "pred_masks_bin": o["pred_masks"] > 0}
        for k, v in res.items():
            if k not in d2h_pinned:
                d2h_pinned[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
            d2h_pinned[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()                       # the caller reads the result every step

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, steps) / steps
    h2d = frames_host.numel() * frames_host.element_size()
    d2h = sum(v.numel() * v.element_size() for v in d2h_pinned.values())

    # ---- roofline of the mask einsum (the kernel BASELINE.json's metric names) -----------------------------
    HW = out["pred_masks"].shape[-1] * out["pred_masks"].shape[-2]
    C = 256
    t_einsum = out["pred_masks"].shape[2]
    alg_bytes = 4.0 * t_einsum * (n_lp * C + C * HW + n_lp * HW)        # per launch (all T frames), SURVEY 8(d)
    peak, peak_kind = peaks()
    roof = None
    if "mask_einsum" in kernel_ms:
        ach = alg_bytes / (kernel_ms["mask_einsum"] * 1e-3) / 1e9
        traffic, traffic_source = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_einsum_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("workload") == args.workload and tj.get("precision") == args.precision:
                traffic = tj.get("dram_bytes_per_launch")     # dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full
                traffic_source = "static: ncu --set full capture of this kernel at this shape, profiles/r1_einsum_traffic.json (not re-measured by this run)"
        roof = {"kernel": "mask_einsum", "bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_source, "launch_ms": kernel_ms["mask_einsum"],
                "algorithmic_bytes_per_launch": alg_bytes}

    named = None
    try:    # per-kernel roofline fractions of the four named kernels (metric (iii) of SURVEY.md 8d); never blocks the line
        t_local = len(range(rank, T, world))                  # frames of this rank (frame_plan: round robin)
        named = named_kernel_rooflines(variant, T, H, W, n_lp, kernel_ms, kernel_calls, peak,
                                       dec_layers=cfg.MODEL.MASK_FORMER.DEC_LAYERS - 1,
                                       enc_layers=cfg.MODEL.SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS,
                                       sm_mhz=float((clocks or {}).get("sm_mhz") or 1965.0), T_local=t_local,
                                       T_dec=t_local if (world > 1 and model.shard_decoder) else T)
    except Exception as exc:  # noqa: BLE001
        sys.stderr.write(f"[bench] named-kernel rooflines unavailable: {exc}\n")

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = run_cpu_reference(args, args.workload, 1, 1, False)
            except BaseException as e:  # noqa: BLE001
                cpu = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                       "sample": f"failed: {str(e)[:200]}"}
        line = {
            "metric": ("frames/sec (Swin-L 720p T=5 Q=200)" if args.workload == "ns" and not args.frames
                       else "frames/sec (per-clip forward)"),
            "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp16x3": "fp16x3 (2-term fp16 split operands, fp32-equivalent products, fp32 accumulate)", "tf32x3": "tf32x3 (3-pass TF32 split, fp32-equivalent products, fp32 accumulate)", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": f"{args.workload}: Swin-{variant} T={T} {H}x{W}->pad32 Q={Q} detection, no prompts, random init"
                                   + (" [DIAGNOSTIC --frames override of the workload's T]" if args.frames else ""),
                       "precision": args.precision, "parallelism": (f"frame-shard x{world}" + (", frame-sharded decoder (token all-gather per layer)"
                                                                  if model.shard_decoder else ", feature all-gather"))
                       if world > 1 else "single",
                       "execution": "CUDA graph replay" if use_graph else "eager",
                       "switches": switches.active(),      # opt-in paths that were on ({} = the round-1 default path)
                       "e2e_result": "pred_logits + pred_embds + BINARISED pred_masks (pred_masks > 0, 1 byte/pixel) copied to pinned host memory "
                                     "every step -- what the task heads move to the host; the fp32 mask logits stay on the device",
                       "l2": "per-step working set (multi-GB activations) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": T / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            # the DOMINANT kernel of the step (dense layers: ~57 % of the NS step) against the tensor roofline; the mask
            # einsum -- SURVEY 8(d) metric (ii), the kernel BASELINE.json's north star names -- against the HBM roofline
            "roofline": dense if dense is not None else roof,
            "roofline_mask_einsum": roof,
            "named_kernel_rooflines": named,        # algorithmic bytes per clip / event time per clip, vs the same HBM peak
            "kernels": {k: {"ms_per_launch": kernel_ms[k], "launches_per_step": kernel_calls[k]} for k in sorted(kernel_ms)},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        # Tear-down: a CUDA graph that captured NCCL kernels keeps the communicator busy and destroy_process_group() then
        # blocks (seen on the B200 box: the line was out, the ranks never exited).  Drop the graph, drain the device, agree
        # that every rank is done, and leave without the communicator destructors -- the results are already printed.
        if dev_type != "cuda":                   # gloo dry run (tests/test_bench_dryrun.py): ordinary tear-down
            dist.destroy_process_group()
            return
        sys.stdout.flush()
        sys.stderr.flush()
        graphed = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        try:
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
        os._exit(0)


if __name__ == "__main__":
    main()
