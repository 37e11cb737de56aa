"""Prompt-specified video segmentation head: semi-supervised VOS (task "sot", visual mask prompts) and referring VOS
(task "grounding", text prompts) -- univs/inference/inference_video_vos.py.

The head owns the per-video annotation state that the decoder's prompt sampler reads from `targets[0]`
(`masks`, `boxes`, `ids`, `first_appear_frame_idxs`, `first_frame_idx`, ...; prompt_encoder.py:844-960): before each
clip it opens slots for the new frames and writes the given first-appearance masks
(`write_targets_into_annotations_per_clip`, reference :532-618), after each clip it writes the predictions back as
pseudo annotations (`write_predictions_into_annotations_per_clip`, :286-530) so that they prompt the following frames.
Same state layout and update rules as the reference; what differs:

* clips come from `ClipStream` (each frame through backbone + pixel decoder once instead of once per clip);
* results are returned in memory (`eval` -> {"frames": {frame_idx: uint8 id map}} for sot,
  {"objects": {obj_id: {frame_idx: bool mask}}} for grounding); PNG files in the reference's layout are written only
  when `output_dir` is set (:620-705);
* everything stays on the model's device; the only host syncs are the small index computations the update rules need.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch import nn

from ..modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO
from ..modeling.visual_prompts import mask_to_box
from ..registry import is_cfg
from ..streaming import ClipStream
from .comm import (calculate_mask_quality_scores, check_consistency_with_prev_frames, match_from_learnable_embds,
                   pair_mask_iou, video_box_iou)
from .video_vis_fast import _Images

_PROMPT_MODES = ("prompt", "prompt+learn", "learn+prompt")
_LEARN_MODES = ("learn", "prompt+learn", "learn+prompt")


class FrameAnnotations:
    """Per-frame annotations of a VOS video for callers without detectron2: the fields of `Instances` the head reads
    (:587-606) -- ori_ids (list[int]), gt_masks [n,h,w], gt_boxes [n,4] absolute XYXY, gt_classes [n], image_size."""

    def __init__(self, image_size, ori_ids=(), gt_masks=None, gt_boxes=None, gt_classes=None):
        self.image_size = tuple(image_size)
        self.ori_ids = list(ori_ids)
        n = len(self.ori_ids)
        self.gt_masks = gt_masks if gt_masks is not None else torch.zeros((n, *self.image_size))
        self.gt_boxes = gt_boxes if gt_boxes is not None else mask_to_box(self.gt_masks > 0).float()
        self.gt_classes = gt_classes if gt_classes is not None else torch.zeros(n, dtype=torch.long)

    def __len__(self):
        return len(self.ori_ids)

    def to(self, device):
        return FrameAnnotations(self.image_size, self.ori_ids, self.gt_masks.to(device), self.gt_boxes.to(device),
                                self.gt_classes.to(device))


def _tensor_of(x):
    """detectron2 Boxes / BitMasks wrap a `.tensor`; plain tensors pass through."""
    return getattr(x, "tensor", x)


def _exclusive_assignment(logits, weight):
    """Per pixel, the object with the largest weighted probability owns it; pixels where no object is positive are
    background.  logits [n,T,H,W], weight [n] -> one-hot float [n,T,H,W] (:392-401, :489-495)."""
    background = (logits <= 0).all(0)
    owner = (logits.sigmoid() * weight.view(-1, 1, 1, 1)).argmax(0)
    owner[background] = -1
    return (owner.unsqueeze(0) == torch.arange(logits.shape[0], device=logits.device).view(-1, 1, 1, 1)).float()


class InferenceVideoVOS(nn.Module):
    def __init__(self, cfg=None, *, hidden_dim=256, num_queries=200, num_frames=5, size_divisibility=32,
                 prompt_as_queries=True, num_frames_window_test=5, clip_stride=1, output_dir=None,
                 video_unified_inference_queries="prompt", num_prev_frames_memory=5, lsj_aug_enable_test=False,
                 lsj_aug_image_size=1024, metadata=None, reuse_features=True):
        super().__init__()
        if cfg is not None and is_cfg(cfg):
            mf, bv, uv = cfg.MODEL.MASK_FORMER, cfg.MODEL.BoxVIS.TEST, cfg.MODEL.UniVS
            hidden_dim = mf.HIDDEN_DIM
            num_queries = mf.NUM_OBJECT_QUERIES
            num_frames = cfg.INPUT.SAMPLING_FRAME_NUM
            size_divisibility = mf.SIZE_DIVISIBILITY
            prompt_as_queries = uv.PROMPT_AS_QUERIES
            num_frames_window_test = bv.NUM_FRAMES_WINDOW
            clip_stride = bv.CLIP_STRIDE
            output_dir = cfg.get("OUTPUT_DIR", None)
            video_unified_inference_queries = uv.TEST.get("VIDEO_UNIFIED_INFERENCE_QUERIES", "prompt")
            num_prev_frames_memory = uv.TEST.NUM_PREV_FRAMES_MEMORY
            lsj_aug_enable_test = cfg.INPUT.LSJ_AUG.SQUARE_ENABLED
            lsj_aug_image_size = cfg.INPUT.LSJ_AUG.IMAGE_SIZE
        if video_unified_inference_queries not in ("prompt", "learn", "prompt+learn", "learn+prompt"):
            raise ValueError(video_unified_inference_queries)
        self.hidden_dim = hidden_dim
        self.num_queries = num_queries
        self.num_frames = num_frames
        self.size_divisibility = size_divisibility
        self.prompt_as_queries = prompt_as_queries
        self.num_frames_window_test = max(num_frames_window_test, num_frames)
        self.clip_stride = clip_stride
        self.output_dir = output_dir
        self.video_unified_inference_queries = video_unified_inference_queries
        self.num_prev_frames_memory = max(num_prev_frames_memory, num_frames)
        self.LSJ_aug_enable_test = lsj_aug_enable_test
        self.LSJ_aug_image_size = lsj_aug_image_size
        self.metadata = metadata
        self.use_semseg_pvos = True
        self.reuse_features = reuse_features
        self.results = None
        self._last_targets = None

    # ------------------------------------------------------------------ entry point (reference :207-245)
    @torch.no_grad()
    def eval(self, model, batched_inputs):
        if len(batched_inputs) != 1:
            raise ValueError("one video per call")
        video = batched_inputs[0]
        x, image_size = model.preprocess(video["image"])
        if self.LSJ_aug_enable_test:
            d, S = self.size_divisibility, self.LSJ_aug_image_size
            S = (max(S, *x.shape[-2:]) + d - 1) // d * d
            x = F.pad(x, (0, S - x.shape[-1], 0, S - x.shape[-2]), value=0.0)
        out_size = (video.get("height", image_size[0]), video.get("width", image_size[1]))
        task = video["task"]
        if task not in ("sot", "grounding"):
            raise ValueError(f"the VOS head serves the prompt-specified tasks, got {task!r}")
        V = x.shape[0]
        tg = {"task": task, "dataset_name": video["dataset_name"], "video_len": V, "num_frames": self.num_frames,
              "prompt_type": "text" if task == "grounding" else "visual",     # prepare_targets.py:58-64
              "inter_image_size": tuple(x.shape[-2:]), "image_size": image_size,
              "file_names": video.get("file_names", [f"video/{i:05d}.jpg" for i in range(V)])}
        passthrough = ("instances", "mask_palette", "expressions", "exp_obj_ids", "exp_word_feats",
                       "exp_sentence_feats", "exp_word_len", "prompt_obj_ids")
        tg.update({k: video[k] for k in passthrough if k in video})
        images = _Images(x, [image_size] * V)
        return self.inference_video_vos(model, batched_inputs, images, [tg], image_size, out_size)

    # ------------------------------------------------------------------ clip loop (reference :247-284)
    @torch.no_grad()
    def inference_video_vos(self, model, batched_inputs, images, targets, image_size, out_size):
        x = images.tensor
        image_size = images.image_sizes[0]
        V, T = x.shape[0], self.num_frames
        self._device = x.device
        stride = min(self.clip_stride, T)
        task = targets[0]["task"]
        self.results = {"frames": {}} if (task == "sot" or "davis" in targets[0]["dataset_name"]) else {"objects": {}}

        stream = ClipStream(model, T, max_cached_frames=2 * max(T, self.num_frames_window_test)) \
            if self.reuse_features else None
        pushed, window, is_last = 0, (0, 0, None), False
        for i in range(0, V, stride):
            if is_last and i + T > V:
                break
            is_last = i + T >= V
            n = min(T, V - i)
            targets[0]["frame_indices"] = torch.arange(i, i + n)
            # step 1: open annotation slots for the new frames, write the given first-appearance masks
            self.write_targets_into_annotations_per_clip(targets, i, stride)
            # step 2: the hot path
            if stream is not None:
                while pushed < i + n:
                    k = min(self.num_frames_window_test, V - pushed)
                    stream.push_preprocessed(pushed, x[pushed:pushed + k])
                    pushed += k
                out = stream.clip(i, targets, length=n)
            else:
                if i + T > window[1]:
                    end = min(i + self.num_frames_window_test, V)
                    window = (i, end, model.backbone(x[i:end]))
                feats = {k: v[i - window[0]:i - window[0] + T] for k, v in window[2].items()}
                out = model.sem_seg_head(feats, targets=targets)
            out = {k: v for k, v in out.items() if torch.is_tensor(v)}
            # step 3: predictions become the pseudo annotations that prompt the following frames
            self.write_predictions_into_annotations_per_clip(out, image_size, targets, i, stride)
            if task == "sot" or "davis" in targets[0]["dataset_name"]:
                self.save_vos_results(i, targets, image_size, out_size, is_last, stride)
            elif task == "grounding":
                self.save_rvos_results(i, targets, image_size, out_size, is_last, stride)
        self._last_targets = targets
        return self.results

    # ------------------------------------------------------------------ annotation slots (reference :532-618)
    def write_targets_into_annotations_per_clip(self, targets, first_frame_idx, stride):
        dev = self._device
        for tg in targets:
            V = tg["video_len"]
            Hp, Wp = tg["inter_image_size"]
            if "ids" not in tg:      # first clip of the video: enumerate the objects
                if tg["task"] == "grounding":
                    ids = [int(o) for o in tg["exp_obj_ids"]]
                    first_appear = torch.zeros(len(ids), dtype=torch.long, device=dev)
                else:
                    ids = list(set(sum([list(f.ori_ids) for f in tg["instances"]], [])))
                    ids = [o for o in ids if o != -1]
                    first_appear = torch.full((len(ids),), -1, dtype=torch.long, device=dev)
                tg["ids"] = ids
                tg["first_appear_frame_idxs"] = first_appear
                tg["labels"] = torch.full((len(ids),), -1, dtype=torch.long, device=dev)
            tg["first_frame_idx"] = first_frame_idx

            n_clip = min(self.num_frames, V - first_frame_idx)          # the last clip may be short
            n_new = n_clip if first_frame_idx == 0 else min(stride, V - first_frame_idx)
            ids, labels, first_appear = tg["ids"], tg["labels"], tg["first_appear_frame_idxs"]
            N = len(ids)
            masks = torch.zeros((N, n_new, Hp, Wp), dtype=torch.float32, device=dev)
            logits = masks.clone()
            boxes = torch.zeros((N, n_new, 4), dtype=torch.float32, device=dev)
            if first_frame_idx == 0:
                embds = torch.zeros((N, n_new, self.hidden_dim), dtype=torch.float32, device=dev)
            else:
                # masks keep the last `num_prev_frames_memory` frames, boxes / embeddings keep every frame; new
                # embedding slots start from the mean of the newest ones
                keep = self.num_prev_frames_memory
                embds = tg["embds"][:, -n_new:].mean(1, keepdim=True).repeat(1, n_new, 1)
                masks = torch.cat([tg["masks"][:, -keep:], masks], 1)
                logits = torch.cat([tg["mask_logits"][:, -keep:], logits], 1)
                boxes = torch.cat([tg["boxes"], boxes], 1)
                embds = torch.cat([tg["embds"], embds], 1)

            if tg["task"] == "sot":
                scale = torch.tensor([Wp, Hp, Wp, Hp], dtype=torch.float32, device=dev)
                for f, ann in enumerate(tg["instances"]):
                    if not (first_frame_idx <= f < first_frame_idx + n_clip) or len(ann) == 0:
                        continue
                    ann = ann.to(dev)
                    h, w = ann.image_size
                    rows = [ids.index(o) for o in ann.ori_ids]
                    boxes[rows, f] = _tensor_of(ann.gt_boxes) / scale           # XYXY, absolute frame index
                    rel = f - (first_frame_idx + n_clip)                         # masks are indexed from the end
                    given = _tensor_of(ann.gt_masks).float()
                    masks[rows, rel, :h, :w] = given
                    logits[rows, rel, :h, :w] = given
                    labels[rows] = ann.gt_classes.to(labels.dtype)
                    first_appear[rows] = f       # VOS objects may enter in the middle of the video
            tg.update({"labels": labels, "masks": masks, "mask_logits": logits, "boxes": boxes, "embds": embds,
                       "first_appear_frame_idxs": first_appear})

    # ------------------------------------------------------------------ prediction write-back (reference :286-530)
    def _is_stuff(self, label):
        table = getattr(self.metadata, "stuff_dataset_id_to_contiguous_id", None)
        return table is not None and (label + 1) in table

    def write_predictions_into_annotations_per_clip(self, out, image_size, targets, first_frame_idx, stride):
        tg = targets[0]
        Q, mode, task = self.num_queries, self.video_unified_inference_queries, tg["task"]
        if task == "grounding" and not self.prompt_as_queries:
            raise ValueError("only support prompts as queries for referring segmentation task")
        viposeg = "viposeg" in tg["dataset_name"]
        probs = out["pred_logits"][0].float().sigmoid()      # [Q+P, K]
        pmasks = out["pred_masks"][0].float()                # [Q+P, T, h, w]
        pembds = out["pred_embds"][0].float()                # [Q+P, T, C]
        h, w = pmasks.shape[-2:]
        pboxes = mask_to_box(pmasks > 0) / torch.tensor([w, h, w, h], device=pmasks.device)   # normalised XYXY
        T = pmasks.shape[1]

        gmasks, glogits, gboxes, gembds, glabels = tg["masks"], tg["mask_logits"], tg["boxes"], tg["embds"], tg["labels"]
        pmasks = F.interpolate(pmasks, gmasks.shape[-2:], mode="bilinear", align_corners=False)
        quality = calculate_mask_quality_scores(pmasks[..., :image_size[0], :image_size[1]])
        sem = None
        if viposeg and self.use_semseg_pvos:     # per-pixel semantic map from the learnable queries (:318-323)
            k, s = COMBINED_DATASETS_CATEGORY_INFO["vipseg"]
            probs = probs[..., s:s + k]
            sem = torch.einsum("qc,qthw->cthw", (probs * quality.view(-1, 1))[:Q], pmasks[:Q].sigmoid()).argmax(0)

        first_appear = tg["first_appear_frame_idxs"]
        prompt_on = self.prompt_as_queries and mode in _PROMPT_MODES
        learn_on = mode in _LEARN_MODES

        # ---- objects whose first annotated frame lies in this clip
        enters = (first_appear >= first_frame_idx) & (first_appear < first_frame_idx + T)
        if enters.any():
            obj = enters.nonzero().reshape(-1)
            rel = first_appear[enters] - (first_frame_idx + T)           # negative: frames counted from the clip end
            n_in = obj.numel()
            rng = torch.arange(n_in, device=obj.device)
            prompt_only = task == "sot"        # given masks: only the prompt queries are trusted in the first clip
            given_masks, given_boxes = gmasks[obj, rel], gboxes[obj, rel]
            idx_p = obj + Q if (prompt_only or prompt_on) else None
            idx_l = None
            if not prompt_only and learn_on:
                # re-identify among all queries: top-5 by box IoU in the first frame, then best mask IoU (:347-357)
                biou = video_box_iou(given_boxes[:, None].repeat(1, T, 1), pboxes)[rng, :, rel]
                top = biou.topk(5, dim=-1)[1]
                cand = pmasks[top.flatten(), rel[:, None].repeat(1, 5).flatten()].reshape(n_in, 5, *pmasks.shape[-2:]) > 0
                miou = pair_mask_iou(given_masks.unsqueeze(1).repeat(1, 5, 1, 1), cand)
                idx_l = top[rng, miou.argmax(-1)]
            if prompt_only or (self.prompt_as_queries and mode == "prompt"):
                m_masks, m_quality, m_embds, m_boxes = pmasks[idx_p], quality[idx_p], pembds[idx_p], pboxes[idx_p]
            elif mode == "learn":
                m_masks, m_quality, m_embds, m_boxes = pmasks[idx_l], quality[idx_l], pembds[idx_l], pboxes[idx_l]
            else:                               # quality-weighted blend of the prompt query and its re-identified twin
                tot = (quality[idx_p] + quality[idx_l]).clamp(min=1e-5)
                wp, wl = quality[idx_p] / tot, quality[idx_l] / tot
                m_masks = wp.view(-1, 1, 1, 1) * pmasks[idx_p] + wl.view(-1, 1, 1, 1) * pmasks[idx_l]
                m_quality = calculate_mask_quality_scores(m_masks)
                m_embds = wp.view(-1, 1, 1) * pembds[idx_p] + wl.view(-1, 1, 1) * pembds[idx_l]
                m_boxes = wp.view(-1, 1, 1) * pboxes[idx_p] + wl.view(-1, 1, 1) * pboxes[idx_l]
            gembds[enters, -T:] = m_embds

            if task == "sot":
                # agreement with the given mask in its own frame weights the exclusive pixel assignment; objects whose
                # assigned region no longer overlaps the given mask are not propagated (:388-410)
                agree = pair_mask_iou(given_masks, m_masks[rng, rel] > 0)
                onehot = _exclusive_assignment(m_masks, agree ** 2 * m_quality)
                m_masks = m_masks * onehot
                agree = pair_mask_iou(given_masks, onehot[rng, rel])
                area = given_masks.flatten(1).sum(1) / (96 * 96)
                accept = agree > 0.15 * area.clamp(max=1)
            else:
                accept = torch.ones(n_in, dtype=torch.bool, device=obj.device)

            for j, (ok, o, r) in enumerate(zip(accept.tolist(), obj.tolist(), rel.tolist())):
                r = r + 1 if task == "sot" else r        # sot: the given frame itself keeps its annotation
                label = int(glabels[o])
                stuff = viposeg and self._is_stuff(label)
                if (not ok and not stuff) or r == 0:
                    continue
                cur = m_masks[j, r:]
                if stuff and sem is not None:            # stuff regions follow the semantic map
                    cur[sem[r:] == label] = 10.0
                gmasks[o, r:] = (cur > 0).float()
                glogits[o, r:] = cur
                gboxes[o, r:] = m_boxes[j, r:]

        # ---- objects already being tracked
        tracked = (first_appear < first_frame_idx) & (first_appear != -1)
        if tracked.any():
            memory = gembds[tracked, -self.num_prev_frames_memory:]       # [n, V_mem, C]
            if not (prompt_on or learn_on):
                raise ValueError("Must use at least one of prompt or learn queries")
            if prompt_on:
                idx_p = tracked.nonzero().reshape(-1) + Q
                ok, sim_p = check_consistency_with_prev_frames(memory, pembds[idx_p], sim_threshold=0.5,
                                                               return_similarity=True)
                keep = ok.float()
                masks_p = pmasks[idx_p] * keep.view(-1, 1, 1, 1)
                quality_p, embds_p = quality[idx_p] * keep, pembds[idx_p] * keep.view(-1, 1, 1)
                boxes_p, sim_p = pboxes[idx_p] * keep.view(-1, 1, 1), sim_p * keep
            if learn_on:
                use_norm = not viposeg
                idx_l, sim_l = match_from_learnable_embds(memory, pembds[:Q], return_similarity=True,
                                                          use_norm=use_norm)
                idx_l = torch.as_tensor(idx_l, device=pmasks.device)
                keep = (sim_l >= (0.65 if use_norm else 0.5)).float()
                masks_l = pmasks[idx_l] * keep.view(-1, 1, 1, 1)
                quality_l, embds_l = quality[idx_l] * keep, pembds[idx_l] * keep.view(-1, 1, 1)
                boxes_l, sim_l = pboxes[idx_l] * keep.view(-1, 1, 1), sim_l * keep
            if prompt_on and learn_on:
                sim = (sim_p + sim_l) / ((sim_p > 0).float() + (sim_l > 0).float()).clamp(min=1)
                tot = (sim_p + sim_l).clamp(min=1e-5)
                wp, wl = sim_p / tot, sim_l / tot
                inter = ((masks_p > 0) & (masks_l > 0)).flatten(1).sum(1)
                union = ((masks_p > 0) | (masks_l > 0)).flatten(1).sum(1)
                disagree = inter / union.clamp(min=1) < 0.5       # the two candidates are different regions: prompt wins
                wp = torch.where(disagree, torch.ones_like(wp), wp)
                wl = torch.where(disagree, torch.zeros_like(wl), wl)
                m_masks = wp.view(-1, 1, 1, 1) * masks_p + wl.view(-1, 1, 1, 1) * masks_l
                m_quality = calculate_mask_quality_scores(m_masks)
                m_embds = wp.view(-1, 1, 1) * embds_p + wl.view(-1, 1, 1) * embds_l
                m_boxes = wp.view(-1, 1, 1) * boxes_p + wl.view(-1, 1, 1) * boxes_l
            elif prompt_on:
                sim, m_masks, m_quality, m_embds, m_boxes = sim_p, masks_p, quality_p, embds_p, boxes_p
            else:
                sim, m_masks, m_quality, m_embds, m_boxes = sim_l, masks_l, quality_l, embds_l, boxes_l

            if task == "sot":
                area_before = (m_masks > 0).flatten(1).sum(1).clamp(min=1)
                weight = sim ** 2 * m_quality
                if viposeg and self.use_semseg_pvos:
                    forced = torch.zeros_like(m_masks, dtype=torch.bool)
                    for j, label in enumerate(glabels[tracked].tolist()):
                        if self._is_stuff(int(label)):
                            forced[j] = sem == label
                    m_masks = torch.where(forced, torch.full_like(m_masks, 10.0), m_masks)
                    background = (m_masks <= 0).all(0)
                    prob = torch.where(forced, torch.ones_like(m_masks), m_masks.sigmoid())
                    owner = (prob * weight.view(-1, 1, 1, 1)).argmax(0)
                    owner[background] = -1
                    onehot = (owner.unsqueeze(0) == torch.arange(m_masks.shape[0], device=owner.device)
                              .view(-1, 1, 1, 1)).float()
                else:
                    onehot = _exclusive_assignment(m_masks, weight)
                area_after = onehot.flatten(1).sum(1)
                # an object that lost more than 3/4 of its pixels to others is dropped for this clip
                alive = ((area_after / area_before) > 0.25) & (area_before > 0) & (area_after > 0)
                onehot = onehot * alive.view(-1, 1, 1, 1).float()
                m_masks = m_masks * onehot

            glogits[tracked, -T:] += m_masks
            gboxes[tracked, -T:] = m_boxes
            seen = (gembds[tracked, -T:] != 0).any(-1)
            gembds[tracked, -T:] = (gembds[tracked, -T:] + m_embds) / (seen.unsqueeze(-1) + 1.0)

        tg["masks"] = (glogits > 0).float()
        tg["mask_logits"] = glogits
        tg["boxes"] = gboxes
        tg["embds"] = gembds

    # ------------------------------------------------------------------ results (reference :620-705)
    def _finished_frames(self, tg, first_frame_idx, image_size, out_size, is_last, stride):
        """Logits [N, n, H_out, W_out] > 0 of the frames this clip finalises (the reference's slice, :634-637)."""
        V = len(tg["file_names"])
        n = min(self.num_frames, V - first_frame_idx)
        logits = tg["mask_logits"]
        logits = logits[:, -n:] if is_last else logits[:, -n:min(-n + stride, -1)]
        logits = logits[:, :, :image_size[0], :image_size[1]]
        if tuple(image_size) != tuple(out_size):
            logits = F.interpolate(logits.float(), out_size, mode="bilinear", align_corners=False)
        return logits > 0

    def save_vos_results(self, first_frame_idx, targets, image_size, out_size, is_last, stride):
        tg = targets[0]
        ids = torch.as_tensor(tg["ids"], device=self._device)
        if ids.min() == 0:
            ids = ids + 1          # zero-based expression ids (RefDAVIS): 0 is background in the id map
        fg = self._finished_frames(tg, first_frame_idx, image_size, out_size, is_last, stride)
        for t, m in enumerate(fg.transpose(0, 1)):
            id_map = ids[m.float().argmax(0)]
            id_map[~m.any(0)] = 0
            id_map = id_map.to(torch.uint8)
            self.results["frames"][first_frame_idx + t] = id_map.cpu()
            if self.output_dir is not None:
                self._write_png(tg, tg["file_names"][first_frame_idx + t], id_map, palette=tg.get("mask_palette"))

    def save_rvos_results(self, first_frame_idx, targets, image_size, out_size, is_last, stride):
        tg = targets[0]
        fg = self._finished_frames(tg, first_frame_idx, image_size, out_size, is_last, stride)
        for obj_id, per_obj in zip(tg["ids"], fg):
            store = self.results["objects"].setdefault(obj_id, {})
            for t, m in enumerate(per_obj):
                store[first_frame_idx + t] = m.cpu()
                if self.output_dir is not None:
                    self._write_png(tg, tg["file_names"][first_frame_idx + t], m.to(torch.uint8) * 255,
                                    subdir=str(obj_id))

    def _write_png(self, tg, file_name, image_u8, palette=None, subdir=None):
        from PIL import Image
        video = tg["file_names"][0].split("/")[-2]
        save_dir = os.path.join(self.output_dir, "inference/Annotations", video, *([subdir] if subdir else []))
        os.makedirs(save_dir, exist_ok=True)
        img = Image.fromarray(image_u8.cpu().numpy())
        if palette is not None:
            img.putpalette(palette)
        img.save(os.path.join(save_dir, file_name.split("/")[-1].replace(".jpg", ".png")))
