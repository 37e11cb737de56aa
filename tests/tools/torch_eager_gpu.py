"""Informative GPU-side baseline (SURVEY.md 8d "GPU-side reference"): the reference algorithm on the SAME B200 through torch's
LIBRARY kernels -- the product host model with every hot-path operator replaced by its torch restatement (oracle/ops_ref.py:
F.grid_sample MSDeformAttn, matmul / softmax attention, einsum), eager, fp32 with torch's default TF32 settings.  What "beat
torch library dispatch on the same GPU" is measured against; a diagnostic that uses the oracle, hence under tests/tools/
(bench.py itself may only execute the oracle in its CPU arm).
  python tests/tools/torch_eager_gpu.py [--workload ns|c2|c5] [--steps 3]        -> one JSON line"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import WORKLOADS, make_targets          # noqa: E402


def run_torch_eager_gpu(args):
    """Informative arm (SURVEY.md 8d "GPU-side reference"): the reference algorithm on the SAME B200 through torch's
    library kernels -- the product host model with every hot-path operator replaced by its torch restatement
    (oracle/ops_ref.py: F.grid_sample MSDeformAttn, matmul / softmax attention, einsum), eager, fp32 with torch's default
    TF32 settings.  What "beat torch library dispatch on the same GPU" is measured against; not a parity-checked path."""
    try:
        from oracle.cpu_backend import oracle_ops
        from univs_b200.build import build_model, make_cfg
        variant, T, H, W, Q = WORKLOADS[args.workload]
        dev = torch.device("cuda", 0)
        g = torch.Generator().manual_seed(0)
        cfg = make_cfg(variant, Q, T, clip_emb=torch.randn(3938, 640, generator=g), TEXT_PROMPT_TO_IMAGE_ENABLE=False)
        model = build_model(cfg).to(dev)
        frames = (torch.rand(T, 3, H, W, generator=g) * 255).to(dev)
        torch.backends.cudnn.allow_tf32 = True           # torch's shipped defaults
        torch.backends.cuda.matmul.allow_tf32 = False
        steps, warmup = max(1, min(args.steps, 5)), 2
        with oracle_ops():
            for _ in range(warmup):
                model.clip_forward(frames, make_targets(T, dev))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model.clip_forward(frames, make_targets(T, dev))
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"impl": "torch-eager", "metric": "frames/sec (Swin-L 720p T=5 Q=200)" if args.workload == "ns" else "frames/sec (per-clip forward)",
                "value": T / (ms / 1e3), "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms,
                "higher_is_better": True, "dtype": "f32 (torch defaults: IEEE matmul, TF32 cuDNN)", "data": "synthetic",
                "config": {"workload": f"{args.workload}: Swin-{variant} T={T} {H}x{W}->pad32 Q={Q} detection, no prompts, random init",
                           "execution": "eager, torch library kernels (port of the reference algorithm)"}}
    except Exception as e:  # noqa: BLE001
        import traceback
        where = " <- ".join(f"{fr.filename.split('/')[-1]}:{fr.lineno}" for fr in traceback.extract_tb(e.__traceback__)[-4:])
        return {"impl": "torch-eager", "unavailable": f"{type(e).__name__}: {str(e)[:300]} [{where}]"}



if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ns", choices=list(WORKLOADS))
    ap.add_argument("--steps", type=int, default=3)
    print(json.dumps(run_torch_eager_gpu(ap.parse_args())))
