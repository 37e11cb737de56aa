"""Diagnostic: same GPU features -> decoder under the fp32 and tf32x3 policies, per-layer differences."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from univs_b200.build import build_model, make_cfg
from univs_b200.precision import set_precision
T, H, W, Q = 2, 720, 1280, 200
g = torch.Generator().manual_seed(0)
cfg = make_cfg("large", Q, T, clip_emb=torch.randn(3938, 640, generator=g), TEXT_PROMPT_TO_IMAGE_ENABLE=False)
m = build_model(cfg).cuda()
frames = (torch.rand(T, 3, H, W, generator=g) * 255).cuda()
tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "frame_indices": torch.arange(T, device="cuda")}]
set_precision("fp32")
x, _ = m.preprocess(frames)
f = m.backbone(x)
mf, _, _, ms = m.sem_seg_head.pixel_decoder.forward_features(f)
dec = m.sem_seg_head.predictor
dec.return_aux_outputs = True
outs = {}
for pol in ("fp32", "tf32x3"):
    set_precision(pol)
    outs[pol] = dec(ms, mf, mf, None, tg())
a, b = outs["fp32"], outs["tf32x3"]
rel = lambda u, v: ((u - v).abs().max() / v.abs().max()).item()
for i, (x1, x2) in enumerate(zip(a["aux_outputs"], b["aux_outputs"])):
    d = (x1["pred_masks"] - x2["pred_masks"]).abs()
    perq = d.flatten(2).amax(2)[0] / x1["pred_masks"].abs().max()
    print(f"layer {i}: masks {rel(x2['pred_masks'], x1['pred_masks']):.2e} embds {rel(x2['pred_embds'], x1['pred_embds']):.2e} "
          f"queries>1e-4: {(perq > 1e-4).sum().item()} top {perq.topk(3).values.tolist()}", flush=True)
print("final", rel(b["pred_masks"], a["pred_masks"]), rel(b["pred_embds"], a["pred_embds"]), rel(b["pred_logits"], a["pred_logits"]))
d = (a["pred_masks"] - b["pred_masks"]).abs().flatten(3).amax(3)[0] / a["pred_masks"].abs().max()   # [Q,T]
print("per-query max err (sorted desc):", d.amax(1).sort(descending=True).values[:8].tolist())
print("rel L2:", ((a["pred_masks"] - b["pred_masks"]).norm() / a["pred_masks"].norm()).item())
set_precision("fp32")
