"""Parity of the CUDA path against the CPU oracle on the prompt configurations of BASELINE.json at full geometry
(SURVEY.md 8d "Synthetic inputs"), with the frame count reduced to bound the CPU time:

  c3  Swin-B, 720x1280, Q=200, task "sot": P=10 objects given as first-frame rectangle masks (area 2-20 % of the frame,
      seed 1), R=32 points per prompt, visual-prompt memory grown over PARITY_CLIPS (default 3) consecutive stride-1
      clips of T frames (default 2; BASELINE: 5)
  c4  Swin-L, 720x1280, Q=200 + P=32 text prompts, task "grounding", self-attention mask "sep-blocked",
      lang->vision on, T frames (default 2; BASELINE: 10)
  c5  Swin-L, 1080x1920, Q=200, detection, T frames (default 1; BASELINE: 8 frames sharded over 8 GPUs -- the sharded
      run itself is `torchrun ... bench.py`; here the single-GPU result at that geometry is checked)

  PARITY_CONFIG=c3 python tests/tools/parity_configs.py        -> gpurun_out/parity_config_c3.json

Per clip: max|a-b|/max|b| of pred_masks / pred_logits / pred_embds (tolerance 1e-3, BASELINE.json north_star), and for
c3 the prompt memory written into `targets` (prompt_feats, prompt_attn_masks).  Same caveat as tests/tools/parity_at_scale.py:
an attention-mask bit is the sign of a logit, so the count of queries beyond 1e-3 is reported next to the maximum."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle.cpu_backend import oracle_ops
from univs_b200.build import build_model, make_cfg
from univs_b200.precision import set_precision

CONFIGS = {
    # name: (variant, H, W, Q, default T, task)
    "c3": ("base", 720, 1280, 200, 2, "sot"),
    "c4": ("large", 720, 1280, 200, 2, "grounding"),
    "c5": ("large", 1080, 1920, 200, 1, "detection"),
}


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def rectangle_masks(P, frames, H, W, seed=1):
    """P axis-aligned rectangles of 2-20 % of the frame, drifting a few pixels per frame: masks [P, frames, H, W] float,
    boxes [P, frames, 4] XYXY normalised (what PrepareTargets hands to the sampler, prepare_targets.py:327)."""
    g = torch.Generator().manual_seed(seed)
    masks, boxes = torch.zeros(P, frames, H, W), torch.zeros(P, frames, 4)
    for p in range(P):
        area = (0.02 + 0.18 * torch.rand(1, generator=g).item()) * H * W
        aspect = 0.5 + 1.5 * torch.rand(1, generator=g).item()
        h = int(min(H - 8, max(8, (area / aspect) ** 0.5)))
        w = int(min(W - 8, max(8, area / h)))
        y0 = int(torch.randint(0, H - h - 4, (1,), generator=g))
        x0 = int(torch.randint(0, W - w - 4, (1,), generator=g))
        for f in range(frames):
            y, x = min(y0 + f, H - h), min(x0 + 2 * f, W - w)
            masks[p, f, y:y + h, x:x + w] = 1
            boxes[p, f] = torch.tensor([x / W, y / H, (x + w) / W, (y + h) / H])
    return masks, boxes


def main():
    name = os.environ.get("PARITY_CONFIG", "c3")
    variant, H, W, Q, T, task = CONFIGS[name]
    T = int(os.environ.get("PARITY_T", T))
    variant = os.environ.get("PARITY_VARIANT", variant)                     # smaller geometry for a quick plumbing check
    if "PARITY_HW" in os.environ:
        H, W = (int(v) for v in os.environ["PARITY_HW"].split("x"))
    dry = not torch.cuda.is_available()        # no GPU: run the oracle side only (checks the target construction)
    clips = int(os.environ.get("PARITY_CLIPS", "3")) if task == "sot" else 1
    precision = os.environ.get("PARITY_PRECISION", "fp16x3")
    g = torch.Generator().manual_seed(0)
    clip_emb = torch.randn(3938, 640, generator=g)
    over = {}
    if name == "c3":
        over = dict(VISUAL_PROMPT_PIXELS_PER_IMAGE=32)
    if name == "c4":
        over = dict(MASKDEC_SELF_ATTN_MASK_TYPE="sep-blocked", TEXT_PROMPT_TO_IMAGE_ENABLE=True)
    if name == "c5":
        over = dict(TEXT_PROMPT_TO_IMAGE_ENABLE=False)
    cfg = make_cfg(variant, Q, T, clip_emb=clip_emb, **over)
    cpu_model = build_model(cfg)
    gpu_model = None
    if not dry:
        gpu_model = build_model(cfg).cuda()
        gpu_model.load_state_dict(cpu_model.state_dict())
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    V = T + clips - 1
    frames = torch.rand(V, 3, H, W, generator=g) * 255
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    P = {"sot": 10, "grounding": 32, "detection": 0}[task]

    def targets(dev):
        tg = {"task": task, "dataset_name": {"sot": "davis", "grounding": "refytvos", "detection": "ytvis21"}[task],
              "prompt_type": "text" if task == "grounding" else "visual"}
        if task == "sot":
            tg["ids"] = torch.arange(P, device=dev)
            tg["first_appear_frame_idxs"] = torch.zeros(P, dtype=torch.long, device=dev)
        if task == "grounding":
            gg = torch.Generator().manual_seed(2)
            tg["exp_word_feats"] = torch.randn(P, 77, T, 640, generator=gg).to(dev)
            tg["exp_sentence_feats"] = torch.randn(P, T, 640, generator=gg).to(dev)
            tg["exp_word_len"] = torch.full((P,), 12, dtype=torch.long, device=dev)
        return [tg]

    masks = boxes = None
    if task == "sot":
        masks, boxes = rectangle_masks(P, V, Hp, Wp)
        masks[:, 1:] = 0                 # only the first frame is annotated; later frames are filled by the caller's
        boxes[:, 1:] = 0                 # pseudo annotations -- here: left blank, the memory carries the objects

    def clip_inputs(tg, c, dev):
        tg[0]["first_frame_idx"] = c
        tg[0]["frame_indices"] = torch.arange(c, c + T, device=dev)
        if task == "sot":
            tg[0]["masks"] = masks[:, : c + T].clone().to(dev)
            tg[0]["boxes"] = boxes[:, : c + T].clone().to(dev)
        return frames[c: c + T].to(dev)

    res = {"config": name, "geometry": f"Swin-{variant} T={T} {H}x{W} Q={Q} P={P} task={task} clips={clips}",
           "precision": precision, "tolerance": 1e-3, "clips": []}
    ctg, gtg = targets("cpu"), (None if dry else targets("cuda"))
    set_precision(precision)
    for c in range(clips):
        t0 = time.time()
        torch.manual_seed(100 + c)
        with oracle_ops():
            wout = cpu_model.clip_forward(clip_inputs(ctg, c, "cpu"), ctg)
        cpu_s = time.time() - t0
        if dry:
            print(json.dumps({"clip": c, "cpu_oracle_seconds": cpu_s, "shape": list(wout["pred_masks"].shape),
                              "prompt_feats": list(ctg[0]["prompt_feats"].shape) if "prompt_feats" in ctg[0] else None}))
            continue
        torch.manual_seed(100 + c)
        gout = gpu_model.clip_forward(clip_inputs(gtg, c, "cuda"), gtg)
        torch.cuda.synchronize()
        pm, wm = gout["pred_masks"].cpu(), wout["pred_masks"]
        perq = (pm - wm).abs().flatten(2).amax(2)[0] / wm.abs().max()
        row = {"clip": c, "cpu_oracle_seconds": cpu_s, "shape": list(pm.shape),
               "pred_masks": rel(pm, wm), "pred_logits": rel(gout["pred_logits"], wout["pred_logits"]),
               "pred_embds": rel(gout["pred_embds"], wout["pred_embds"]),
               "queries_beyond_1e-3": int((perq > 1e-3).sum()), "queries": int(perq.numel())}
        if task == "sot" and "prompt_feats" in ctg[0]:
            row["prompt_feats"] = rel(gtg[0]["prompt_feats"], ctg[0]["prompt_feats"])
            row["prompt_attn_masks_equal"] = bool(torch.equal(gtg[0]["prompt_attn_masks"].cpu(), ctg[0]["prompt_attn_masks"]))
        if task == "grounding" and torch.is_tensor(gout.get("pred_reid_logits")):
            row["pred_reid_logits"] = rel(gout["pred_reid_logits"], wout["pred_reid_logits"])
        res["clips"].append(row)
        print(json.dumps(row), flush=True)
    set_precision("fp32")
    if dry:
        print("no GPU: oracle side only")
        return
    res["pass"] = all(r["pred_masks"] <= 1e-3 for r in res["clips"])
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/parity_config_{name}.json", "w"), indent=1)
    print("PASS" if res["pass"] else "FAIL (see queries_beyond_1e-3: decision flips vs arithmetic)")


if __name__ == "__main__":
    main()
