"""Synthetic inputs of the BASELINE.json configurations (SURVEY.md 8d "Synthetic inputs"): shared by bench.py and the
full-geometry parity tests so that both drive the path with the same clips, prompts and targets.

  detection   no prompts (learnable queries only)
  sot         P objects given as first-frame rectangle masks (area 2-20 % of the frame, seed 1); the visual-prompt memory
              grows over consecutive stride-1 clips (prompt_encoder.py:781-960 keeps it in `targets[0]`)
  grounding   P referring expressions as CLIP text features: exp_word_feats [P,77,T,640], exp_sentence_feats [P,T,640]
              (the CLIP text tower itself runs once per video and is outside the path)"""
from __future__ import annotations

import torch

PROMPTS = {"sot": 10, "grounding": 32, "detection": 0}
DATASET = {"sot": "davis", "grounding": "refytvos", "detection": "ytvis21"}


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


class ClipSource:
    """Frames, targets and per-clip annotations of one synthetic video of `task` at padded size (Hp, Wp)."""

    def __init__(self, task, T, V, H, W, seed_frames=0, annotate_all=False):
        """annotate_all (sot): every frame carries the (drifting) object masks, as when a task head feeds the previous
        clip's predictions back as pseudo annotations; default: only frame 0 is annotated and the memory carries the objects"""
        self.task, self.T, self.V, self.H, self.W = task, T, V, H, W
        self.Hp, self.Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
        self.P = PROMPTS[task]
        g = torch.Generator().manual_seed(seed_frames)
        self.clip_emb = torch.randn(3938, 640, generator=g)
        self.frames = torch.rand(V, 3, H, W, generator=g) * 255
        self.masks = self.boxes = None
        if task == "sot":
            self.masks, self.boxes = rectangle_masks(self.P, V, self.Hp, self.Wp)
            if not annotate_all:
                self.masks[:, 1:] = 0        # only the first frame is annotated; the memory carries the objects afterwards
                self.boxes[:, 1:] = 0

    def targets(self, dev):
        task, P, T = self.task, self.P, self.T
        tg = {"task": task, "dataset_name": DATASET[task], "prompt_type": "text" if task == "grounding" else "visual"}
        if task == "sot":
            tg["ids"] = torch.arange(P, device=dev)
            tg["first_appear_frame_idxs"] = torch.zeros(P, dtype=torch.long, device=dev)
        if task == "grounding":
            gg = torch.Generator().manual_seed(2)
            tg["exp_word_feats"] = torch.randn(P, 77, T, 640, generator=gg).to(dev)
            tg["exp_sentence_feats"] = torch.randn(P, T, 640, generator=gg).to(dev)
            tg["exp_word_len"] = torch.full((P,), 12, dtype=torch.long, device=dev)
        return [tg]

    def clip_inputs(self, tg, c, dev):
        """prepares targets `tg` for the clip starting at frame c (as the task heads do per clip) and returns its frames"""
        T = self.T
        tg[0]["first_frame_idx"] = c
        tg[0]["frame_indices"] = torch.arange(c, c + T, device=dev)
        if self.task == "sot":
            tg[0]["masks"] = self.masks[:, : c + T].clone().to(dev)
            tg[0]["boxes"] = self.boxes[:, : c + T].clone().to(dev)
        return self.frames[c: c + T].to(dev)


def univs_overrides(task, points=128):
    """MODEL.UniVS overrides of the configuration (make_cfg(**overrides))"""
    return {"sot": dict(VISUAL_PROMPT_PIXELS_PER_IMAGE=points),
            "grounding": dict(MASKDEC_SELF_ATTN_MASK_TYPE="sep-blocked", TEXT_PROMPT_TO_IMAGE_ENABLE=True),
            "detection": dict(TEXT_PROMPT_TO_IMAGE_ENABLE=False)}[task]


def clone_targets(tg):
    import copy
    return [{k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in t.items()} for t in tg]
