import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
sys.argv = ["x"]
import bench
from oracle.cpu_backend import oracle_ops
from univs_b200.build import build_model, make_cfg
g = torch.Generator().manual_seed(0)
cfg = make_cfg("large", 200, 1, clip_emb=torch.randn(3938, 640, generator=g), TEXT_PROMPT_TO_IMAGE_ENABLE=False)
model = build_model(cfg)
frames = torch.rand(1, 3, 720, 1280, generator=g) * 255
for n in (16, 32, 64):
    torch.set_num_threads(n)
    with oracle_ops():
        t0 = time.perf_counter(); model.clip_forward(frames, bench.make_targets(1, "cpu")); dt = time.perf_counter() - t0
    print(f"threads={n}: {dt:.1f} s/frame", flush=True)
