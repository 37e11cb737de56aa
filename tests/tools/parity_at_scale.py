"""Parity of the CUDA path against the CPU oracle at the north-star geometry (Swin-L, 736x1280, Q=200), bench
initialisation, per arithmetic policy; T frames (default 2) to bound the CPU time.

The decoder contains hard decisions (attention-mask bits = sign of a mask logit; the "fully blocked row attends
everywhere" rule, ..._univs.py:390): a logit within rounding distance of zero flips a decision and moves that query by
far more than any arithmetic error.  So two numbers are reported per policy:
  free-running  : the CUDA path on its own decisions (max-norm error, relative L2, #queries beyond 1e-3, #bits flipped)
  same-decisions: the CUDA path replaying the oracle's attention-mask bits (isolates the arithmetic error)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle.cpu_backend import oracle_ops
from univs_b200.build import build_model, make_cfg
from univs_b200.precision import set_precision

T = int(os.environ.get("PARITY_T", "2"))
variant = os.environ.get("PARITY_VARIANT", "large")
modes = os.environ.get("PARITY_MODES", "fp16x3,tf32x3,fp32,tf32").split(",")
H, W, Q = 720, 1280, 200
g = torch.Generator().manual_seed(0)
clip = torch.randn(3938, 640, generator=g)
frames = torch.rand(T, 3, H, W, generator=g) * 255
tg = lambda dev: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "frame_indices": torch.arange(T, device=dev)}]


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def popcount_diff(a, b):
    x = (a.cpu() ^ b.cpu()).to(torch.int64) & 0xFFFFFFFF
    n = 0
    for s in range(0, 32, 8):
        byte = (x >> s) & 0xFF
        n += int(sum(((byte >> k) & 1).sum().item() for k in range(8)))
    return n


cfg = make_cfg(variant, Q, T, clip_emb=clip, TEXT_PROMPT_TO_IMAGE_ENABLE=False)
cpu_model = build_model(cfg)
torch.set_num_threads(min(32, os.cpu_count() or 1))
recorded = {}
cpu_model.sem_seg_head.predictor.attn_mask_hook = lambda i, b, r: (recorded.__setitem__(i, (b.clone(), r.clone())) or (b, r))
t0 = time.time()
with oracle_ops():
    x, _ = cpu_model.preprocess(frames)
    wf = cpu_model.backbone(x)
    wmf, _, _, wms = cpu_model.sem_seg_head.pixel_decoder.forward_features(wf)
    wout = cpu_model.sem_seg_head.predictor(wms, wmf, wmf, None, tg("cpu"))
cpu_s = time.time() - t0
res = {"geometry": f"Swin-{variant} T={T} {H}x{W} Q={Q}", "cpu_oracle_seconds": cpu_s, "tolerance_metric": "max|a-b|/max|b| per tensor",
       "modes": {}}
gpu_model = build_model(cfg).cuda()
gpu_model.load_state_dict(cpu_model.state_dict())
dec = gpu_model.sem_seg_head.predictor
for mode in modes:
    set_precision(mode)
    x, _ = gpu_model.preprocess(frames.cuda())
    gf = gpu_model.backbone(x)
    gmf, _, _, gms = gpu_model.sem_seg_head.pixel_decoder.forward_features(gf)
    flips = {}

    def count(i, b, r):
        flips[i] = popcount_diff(b, recorded[i][0])
        return b, r
    dec.attn_mask_hook = count
    gout = dec(gms, gmf, gmf, None, tg("cuda"))
    dec.attn_mask_hook = lambda i, b, r: (recorded[i][0].cuda(), recorded[i][1].cuda())
    fout = dec(gms, gmf, gmf, None, tg("cuda"))
    dec.attn_mask_hook = None
    torch.cuda.synchronize()
    pm, wm = gout["pred_masks"].cpu(), wout["pred_masks"]
    perq = (pm - wm).abs().flatten(2).amax(2)[0] / wm.abs().max()
    total_bits = sum(int(b.numel()) * 32 for b, _ in recorded.values())
    res["modes"][mode] = {
        "features": {"res2": rel(gf["res2"], wf["res2"]), "res5": rel(gf["res5"], wf["res5"]),
                     "mask_features": rel(gmf, wmf), "ms_1_8": rel(gms[2], wms[2])},
        "free_running": {"pred_masks": rel(pm, wm), "pred_logits": rel(gout["pred_logits"], wout["pred_logits"]),
                         "pred_embds": rel(gout["pred_embds"], wout["pred_embds"]),
                         "pred_masks_rel_l2": ((pm - wm).norm() / wm.norm()).item(),
                         "queries_beyond_1e-3": int((perq > 1e-3).sum()), "queries": int(perq.numel()),
                         "attn_mask_bits_flipped": int(sum(flips.values())), "attn_mask_bits_total": total_bits,
                         "final_mask_sign_disagreement": ((pm < 0) != (wm < 0)).float().mean().item()},
        "same_decisions": {"pred_masks": rel(fout["pred_masks"], wm), "pred_logits": rel(fout["pred_logits"], wout["pred_logits"]),
                           "pred_embds": rel(fout["pred_embds"], wout["pred_embds"])},
    }
    print(mode, json.dumps(res["modes"][mode]), flush=True)
set_precision("fp32")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/parity_at_scale_{variant}_T{T}.json", "w"), indent=1)
