"""Parity of the CUDA path against the CPU oracle at the north-star geometry (Swin-L, 736x1280, Q=200), bench
initialisation, for each arithmetic policy.  T frames (default 2) to bound the CPU time.  Writes a JSON summary."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.cpu_backend import oracle_ops, unpack_bits
from univs_b200 import ops
from univs_b200.build import build_model, make_cfg
from univs_b200.precision import set_precision

T = int(os.environ.get("PARITY_T", "2"))
variant = os.environ.get("PARITY_VARIANT", "large")
H, W, Q = 720, 1280, 200
g = torch.Generator().manual_seed(0)
clip = torch.randn(3938, 640, generator=g)
frames = torch.rand(T, 3, H, W, generator=g) * 255
tg = lambda dev: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "frame_indices": torch.arange(T, device=dev)}]


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


cfg = make_cfg(variant, Q, T, clip_emb=clip, TEXT_PROMPT_TO_IMAGE_ENABLE=False)
cpu_model = build_model(cfg)
torch.set_num_threads(min(32, os.cpu_count() or 1))
t0 = time.time()
with oracle_ops():
    x, _ = cpu_model.preprocess(frames)
    wf = cpu_model.backbone(x)
    wmf, _, _, wms = cpu_model.sem_seg_head.pixel_decoder.forward_features(wf)
    wout = cpu_model.sem_seg_head.predictor(wms, wmf, wmf, None, tg("cpu"))
cpu_s = time.time() - t0
res = {"geometry": f"Swin-{variant} T={T} {H}x{W} Q={Q}", "cpu_oracle_seconds": cpu_s, "modes": {}}
gpu_model = build_model(cfg).cuda()
gpu_model.load_state_dict(cpu_model.state_dict())
for mode in ("tf32x3", "fp32", "tf32"):
    set_precision(mode)
    x, _ = gpu_model.preprocess(frames.cuda())
    gf = gpu_model.backbone(x)
    gmf, _, _, gms = gpu_model.sem_seg_head.pixel_decoder.forward_features(gf)
    gout = gpu_model.sem_seg_head.predictor(gms, gmf, gmf, None, tg("cuda"))
    torch.cuda.synchronize()
    sign = ((gout["pred_masks"].cpu() < 0) != (wout["pred_masks"] < 0)).float().mean().item()
    res["modes"][mode] = {
        "res2": rel(gf["res2"], wf["res2"]), "res5": rel(gf["res5"], wf["res5"]),
        "mask_features": rel(gmf, wmf), "ms_1_8": rel(gms[2], wms[2]),
        "pred_masks": rel(gout["pred_masks"], wout["pred_masks"]),
        "pred_logits": rel(gout["pred_logits"], wout["pred_logits"]),
        "pred_embds": rel(gout["pred_embds"], wout["pred_embds"]),
        "mask_sign_disagreement": sign,
    }
    print(mode, json.dumps(res["modes"][mode]), flush=True)
set_precision("fp32")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/parity_at_scale_{variant}_T{T}.json", "w"), indent=1)
print(json.dumps(res))
