"""Unified video-entity head (univs/inference/inference_video_entity.py): VIS / VPS / VSS of one video where entities found
by the learnable queries are written into a memory pool and come back as *visual prompts* for the following clips.

Same contract as the reference class -- `eval(model, batched_inputs)` for one video returns
  * VIS  : list of COCO-video json dicts (inference/comm.py:96-195, `results_to_coco_video`),
  * VPS  : {"image_size", "pred_masks" int32 [V, H, W], "segments_infos", "task": "vps"} (:1061-1094),
  * VSS  : {"image_size", "pred_masks" int64 [V, H, W], "task": "vss"} (:1126-1132)
-- and the same protocol with the decoder: the pool lives in the caller's `targets[0]` under the reference's keys
("logits", "masks", "mask_logits", "boxes", "embds", "ids", "first_appear_frame_idxs", "mask_quality_scores", "occurrence";
the visual-prompt sampler reads "masks" / "boxes" / "ids" / "first_frame_idx" / "frame_indices" from there and appends
"prompt_pe" / "prompt_feats" / "prompt_attn_masks", prompt_encoder.py:844-960).  What differs from the reference:

  * frames go through backbone + pixel decoder once (`ClipStream`), not once per clip (:301-309);
  * result masks are run-length encoded from device-side run boundaries (`inference/rle.py`) instead of moving every dense
    mask to the host for pycocotools (:943-947);
  * visualisation / plotting (:1134-1358) is not part of the path.

Per clip (reference :283-431): (1) prompt-query predictions that are consistent with their entity's history are added
to the pool, (2) learnable-query predictions are de-duplicated, matched to the pool (quasi-dense bidirectional softmax
+ Hungarian assignment) and the unmatched confident ones open new entities, (3) every `num_frames_window_output` frames
the finished part of the pool is turned into results and dropped."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO
from ..modeling.visual_prompts import mask_to_box
from ..registry import is_cfg
from ..streaming import ClipStream
from . import rle
from .comm import calculate_mask_quality_scores, check_consistency_with_prev_frames, process_inference, video_box_iou

_ENTITY_DATASETS = {"entity_vss_entityseg": "entityseg_panoptic", "entity_vps_entityseg": "entityseg_panoptic",
                    "entity_vss_vipseg": "vipseg", "entity_vps_vipseg": "vipseg",
                    "entity_vis_entityseg": "entityseg_instance", "entity_vis_coco": "coco"}
_POOL_KEYS = ("logits", "masks", "mask_logits", "boxes", "embds", "ids", "first_appear_frame_idxs",
              "mask_quality_scores", "occurrence")


def group_mask_iou(masks1, masks2):
    """[B, N, H, W] x [B, M, H, W] binary masks -> IoU [B, N, M] (univs/utils/comm.py:196-210)."""
    a, b = masks1.flatten(-2).float(), masks2.flatten(-2).float()
    both = a[:, :, None] + b[:, None]
    return (both == 2).sum(-1) / (both >= 1).sum(-1).clamp(min=1)


def _upsample(x, size):
    return F.interpolate(x, size, mode="bilinear", align_corners=False)


def _suppress_later_duplicates(overlap, thresh):
    """overlap [n, n] between candidates sorted by score: keep candidate j unless an earlier one overlaps it by >= thresh
    (:566-569, :683-693: strict upper triangle, column-wise maximum)."""
    return torch.triu(overlap, diagonal=1).max(0)[0] < thresh


def temporal_consistency_weights(scores):
    """inference/comm.py:197-207: the class scores of window t are weighted by how many of the windows {t-1, t} carry the
    object at all (scores [W, K], modified in place as the reference does)."""
    nonblank = scores.sum(-1) > 0
    for t in range(len(nonblank)):
        lo, hi = max(0, t - 1), min(len(nonblank), t + 1)
        scores[t] *= nonblank[t] * nonblank[lo:hi].sum() / max(hi - lo, 1)
    return scores


def results_to_coco_video(batched_inputs, results_list, apply_cls_thresh=0.05, test_topk_per_video=25):
    """vis_clip_instances_to_coco_json_video (inference/comm.py:96-195): per-window, per-object results ->
    one entry per (object, class) over the whole video, filtered to the top scores."""
    video = batched_inputs[0]
    try:
        video_id = int(video["video_id"])
    except (TypeError, ValueError):
        video_id = video["video_id"]
    V, height, width = int(video["video_len"]), int(video["height"]), int(video["width"])
    blank = rle.encode(torch.zeros((1, height, width), dtype=torch.bool))[0]
    entries, entry_scores, confident = [], [], 0
    for obj_id in {res["obj_id"] for window in results_list for res in window}:
        segm, cls_scores, quality = [blank] * V, [], []
        for window in results_list:
            for res in window:
                if res["obj_id"] != obj_id:
                    continue
                if "mask_quality_score" in res:
                    quality.append(res["mask_quality_score"])
                cls_scores.append(res["score"])
                s = res["frame_id_start"]
                segm[s:s + len(res["segmentations"])] = res["segmentations"]
        assert len(segm) == V, f"The video has {V} frames, but the prediction has {len(segm)} frames!"
        scores = torch.stack(cls_scores, 0)
        if quality:
            quality = sum(quality) / len(quality)
        else:
            quality = ((scores.sum(-1) > 0).sum(0) / V).clamp(min=0.1)
        scores = temporal_consistency_weights(scores)
        scores = scores.sum(0) / (scores.sum(-1) > 0).sum(0).clamp(min=1)
        for c in range(len(scores)):
            if float(scores[c]) < 0.1 * apply_cls_thresh:
                continue
            s = float(scores[c]) * float(quality)
            entries.append({"video_id": video_id, "score": s, "category_id": c, "segmentations": segm,
                            "height": height, "width": width})
            entry_scores.append(s)
            confident += int(scores[c] > apply_cls_thresh)
    if entry_scores:
        entry_scores.sort(reverse=True)
        cut = entry_scores[min(max(int(confident * 1.5), test_topk_per_video), len(entry_scores) - 1)]
        entries = [e for e in entries if e["score"] >= cut]
    return entries


class InferenceVideoEntity(nn.Module):
    def __init__(self, cfg=None, *, hidden_dim=256, num_queries=200, num_frames=5, size_divisibility=32,
                 overlap_threshold=0.8, overlap_threshold_entity=0.5, stability_score_thresh=0.0, test_topk_per_image=100,
                 apply_cls_thres=0.05, box_nms_thresh=0.75, num_frames_window_test=5, clip_stride=1,
                 num_prev_frames_memory=5, video_unified_inference_entities="", temporal_consistency_threshold=0.05,
                 detect_newly_object_threshold=0.05, detect_newly_interval_frames=1, custom_videos_enable=False,
                 thing_ids=(), lsj_aug_enable_test=False, lsj_aug_image_size=1024, reuse_features=True, semantic_on=False):
        super().__init__()
        if cfg is not None and is_cfg(cfg):
            mf, bv, uv = cfg.MODEL.MASK_FORMER, cfg.MODEL.BoxVIS.TEST, cfg.MODEL.UniVS.TEST
            semantic_on = mf.TEST.get("SEMANTIC_ON", False)
            hidden_dim = mf.HIDDEN_DIM
            num_queries = mf.NUM_OBJECT_QUERIES
            num_frames = cfg.INPUT.SAMPLING_FRAME_NUM
            size_divisibility = mf.SIZE_DIVISIBILITY
            overlap_threshold = mf.TEST.OVERLAP_THRESHOLD
            overlap_threshold_entity = mf.TEST.get("OVERLAP_THRESHOLD_ENTITY", 0.5)
            stability_score_thresh = mf.TEST.get("STABILITY_SCORE_THRESH", 0.0)
            test_topk_per_image = cfg.get("TEST", {}).get("DETECTIONS_PER_IMAGE", 100)
            apply_cls_thres = bv.get("APPLY_CLS_THRES", 0.05)
            box_nms_thresh = uv.get("BOX_NMS_THRESH", 0.75)
            num_frames_window_test = bv.NUM_FRAMES_WINDOW
            clip_stride = bv.CLIP_STRIDE
            num_prev_frames_memory = uv.NUM_PREV_FRAMES_MEMORY
            video_unified_inference_entities = uv.get("VIDEO_UNIFIED_INFERENCE_ENTITIES", "")
            temporal_consistency_threshold = uv.get("TEMPORAL_CONSISTENCY_THRESHOLD", 0.05)
            detect_newly_object_threshold = uv.get("DETECT_NEWLY_OBJECT_THRESHOLD", 0.05)
            detect_newly_interval_frames = uv.get("DETECT_NEWLY_INTERVAL_FRAMES", 1)
            custom_videos_enable = uv.get("CUSTOM_VIDEOS_ENABLE", False)
            lsj_aug_enable_test = cfg.INPUT.LSJ_AUG.SQUARE_ENABLED
            lsj_aug_image_size = cfg.INPUT.LSJ_AUG.IMAGE_SIZE
        if custom_videos_enable and not video_unified_inference_entities:
            video_unified_inference_entities = "entity_vps_vipseg"                      # :172-173
        if video_unified_inference_entities and video_unified_inference_entities not in _ENTITY_DATASETS:
            raise ValueError(f"Unsupported inference manner: {video_unified_inference_entities}")
        self.hidden_dim = hidden_dim
        self.num_queries = num_queries
        self.num_frames = num_frames
        self.size_divisibility = size_divisibility
        self.overlap_threshold = overlap_threshold
        self.overlap_threshold_entity = overlap_threshold_entity
        self.stability_score_thresh = stability_score_thresh
        self.test_topk_per_image = test_topk_per_image
        self.apply_cls_thres = apply_cls_thres
        self.box_nms_thresh = box_nms_thresh
        self.num_frames_window_test = max(num_frames_window_test, num_frames)
        self.num_frames_window_output = (math.ceil(self.num_frames_window_test / 5) + 1) * 5      # :155
        self.clip_stride = clip_stride
        self.num_prev_frames_memory = num_prev_frames_memory
        self.video_unified_inference_entities = video_unified_inference_entities
        self.temporal_consistency_threshold = temporal_consistency_threshold
        self.detect_newly_object_threshold = detect_newly_object_threshold
        self.detect_newly_interval_frames = detect_newly_interval_frames
        self.custom_videos_enable = custom_videos_enable
        self.thing_ids = set(int(i) for i in thing_ids)       # 1-based dataset ids of the thing classes (metadata, :674)
        self.LSJ_aug_enable_test, self.LSJ_aug_image_size = lsj_aug_enable_test, lsj_aug_image_size
        self.reuse_features = reuse_features
        self.semantic_on = semantic_on           # PrepareTargets: category prompts ("text") for detection clips
        self._device = torch.device("cpu")
        self._last_targets = None

    # ------------------------------------------------------------------ entry point (reference :237-281)
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
        V = x.shape[0]
        targets = video.get("targets")
        if targets is None:
            targets = process_inference(video, tuple(x.shape[-2:]), image_size, self.num_frames, semantic_on=self.semantic_on)
        if self.video_unified_inference_entities:
            targets[0]["sub_task"] = self.video_unified_inference_entities
        else:
            name = targets[0]["dataset_name"]
            if name.startswith("ytvis") or name.startswith("ovis"):
                targets[0]["sub_task"] = "vis"
            elif name.startswith("vipseg"):
                targets[0]["sub_task"] = "vps"
            elif name.startswith("vspw"):
                targets[0]["sub_task"] = "vss"
            else:
                raise ValueError(f"Not support to eval the dataset {name} yet")
        return self.inference_video(model, batched_inputs, x, image_size, targets)

    # ------------------------------------------------------------------ clip loop (reference :283-431)
    @torch.no_grad()
    def inference_video(self, model, batched_inputs, x, image_size, targets):
        tg = targets[0]
        sub_task = tg["sub_task"]
        V, T = x.shape[0], self.num_frames
        self._device = x.device
        interim_size = tuple(x.shape[-2:])
        out_size = (batched_inputs[0].get("height", image_size[0]), batched_inputs[0].get("width", image_size[1]))
        video_len = int(batched_inputs[0].get("video_len", V))
        name = _ENTITY_DATASETS[sub_task] if sub_task.startswith("entity") else tg["dataset_name"]
        class_slice = COMBINED_DATASETS_CATEGORY_INFO.get(name)
        if sub_task.startswith("entity") and class_slice is None:
            raise ValueError(sub_task)
        stride = min(T if "vss" in sub_task else self.clip_stride, T)
        stream = ClipStream(model, T, max_cached_frames=2 * max(T, self.num_frames_window_test)) \
            if self.reuse_features else None
        pushed, window, is_last, results = 0, (0, 0, None), False, []
        for i in range(0, V, stride):
            if is_last and i + T > V:
                break
            is_last = i + T >= V
            n = min(T, V - i)
            tg["first_frame_idx"] = i
            tg["frame_indices"] = torch.arange(i, i + n)
            if stream is not None:
                while pushed < i + n:
                    k = min(self.num_frames_window_test, V - pushed)
                    stream.push_preprocessed(pushed, x[pushed:pushed + k])
                    pushed += k
                out = stream.clip(i, targets, length=n)
            else:
                if i + T > window[1]:
                    window = (i, i + self.num_frames_window_test, model.backbone(x[i:i + self.num_frames_window_test]))
                feats = {k: v[i - window[0]:i - window[0] + T] for k, v in window[2].items()}
                out = model.sem_seg_head(feats, targets=targets)
            out = {k: v[0] for k, v in out.items() if torch.is_tensor(v)}               # batch size 1
            cls = out["pred_logits"].sigmoid()
            if class_slice is not None:
                assert class_slice[1] + class_slice[0] <= cls.shape[-1]
                cls = cls[..., class_slice[1]:class_slice[1] + class_slice[0]]
            out["pred_logits"] = cls
            learn = {k: v[:self.num_queries] for k, v in out.items()}
            prompt = {k: v[self.num_queries:] for k, v in out.items()}

            if "vss" in sub_task:
                results.append(self.save_results_vss(learn, interim_size, image_size, out_size, is_last, stride))
            elif "vis" in sub_task or "vps" in sub_task:
                self.absorb_prompt_predictions(i, prompt, tg, interim_size, image_size, stride)
                if i % self.detect_newly_interval_frames == 0 or tg["masks"].nelement() == 0:
                    if "vis" in sub_task:
                        fresh = self.detect_new_entities_instance(learn, tg, interim_size)
                    else:
                        fresh = self.detect_new_entities_pixel(learn, tg, interim_size)
                    self.open_new_entities(i, fresh, tg, interim_size)
                is_out = i > self.num_prev_frames_memory and \
                    i % self.num_frames_window_output == self.num_prev_frames_memory
                if is_out or is_last:
                    if "vis" in sub_task:
                        results.append(self.save_results_vis(i, tg, image_size, out_size, is_last))
                    else:
                        results.append(self.save_results_vps(tg, image_size, out_size, is_last))
                    w = self.num_frames_window_output     # the finished frames leave the pool (boxes / embds stay whole)
                    for key in ("mask_logits", "masks", "occurrence"):
                        tg[key] = tg[key][:, w:]
            else:
                raise ValueError(f"Not support to eval the sub-task {sub_task} yet")
            if not is_last and "masks" in tg:
                self.open_slots_for_next_clip(tg, min(stride, video_len - i - T))
        self._last_targets = targets
        if "vis" in sub_task:
            return results_to_coco_video(batched_inputs, results, test_topk_per_video=self.test_topk_per_image)
        if "vps" in sub_task:
            return self.vps_output_results(tg, results, out_size)
        return {"image_size": out_size, "pred_masks": torch.cat(results, 0).cpu(), "task": "vss"}

    # ------------------------------------------------------------------ step 1: prompt queries -> pool (reference :433-515)
    def absorb_prompt_predictions(self, first_frame_idx, out, tg, interim_size, image_size, stride):
        if out["pred_masks"].nelement() == 0:
            return                                                       # no prompt queries in this clip
        masks = _upsample(out["pred_masks"], interim_size)               # [P, T, Hp, Wp]
        embds = out["pred_embds"]                                        # [P, T, C]
        n = masks.shape[1]
        pool_logits, pool_masks, pool_boxes = tg["logits"], tg["mask_logits"], tg["boxes"]
        pool_embds, pool_occ, pool_quality = tg["embds"], tg["occurrence"], tg["mask_quality_scores"]

        thresh = self.temporal_consistency_threshold * (0.5 if first_frame_idx < self.num_frames else 1.0)
        history = max(int(self.num_prev_frames_memory / stride), 3)
        consistent, sim = check_consistency_with_prev_frames(pool_embds[:, -history:], embds, sim_threshold=thresh,
                                                             return_similarity=True)
        cropped = masks[:, :, :image_size[0], :image_size[1]]
        quality = calculate_mask_quality_scores(cropped)
        if "vis" in tg["sub_task"]:
            # instances may not overlap: a pixel belongs to the best-scoring entity; an entity that loses most of its
            # area this way is not written back
            score = pool_logits.mean(1).max(-1)[0] * sim * quality
            prob = cropped.sigmoid().flatten(1)
            owner = (score.view(-1, 1) * prob).argmax(0)
            owner[(prob < 0.5).sum(0) == len(prob)] = -1
            owned = owner[None] == torch.arange(prob.shape[0], device=prob.device).view(-1, 1)
            inside = prob > 0.5
            keeps_area = owned.sum(1) / inside.sum(1).clamp(min=1) > self.overlap_threshold_entity
            consistent = consistent & keeps_area & ((owned & inside).sum(1) > 0)

        if consistent.sum():
            good = masks[consistent]
            norm = torch.as_tensor([interim_size[1], interim_size[0], interim_size[1], interim_size[0]], device=self._device)
            pool_occ[consistent, -n:] += good.flatten(-2).gt(0.).any(-1).float()
            pool_masks[consistent, -n:] += good.clone()
            pool_boxes[consistent, -n:] = mask_to_box(pool_masks[consistent, -n:] > 0) / norm.view(1, 1, -1)
            seen = (pool_embds[consistent, -1] != 0).any(-1)
            pool_embds[consistent, -1] = (pool_embds[consistent, -1] + embds[consistent].mean(1)) / (seen[..., None] + 1.)
            pool_quality[consistent] += quality[consistent]
        tg["masks"] = pool_masks.gt(0.).float()

    # ------------------------------------------------------------------ step 2: learnable queries (reference :517-765)
    def _match_to_pool(self, tg, embds):
        """Quasi-dense association of the pool's last three embeddings with the clip's query embeddings (:608-613):
        bidirectional softmax over all (slot, frame) pairs, thresholded, Hungarian assignment.
        Returns (pool indices, prediction indices, similarity) of the assignment."""
        from scipy.optimize import linear_sum_assignment
        sim = torch.einsum("ntc,mfc->nmtf", tg["embds"][:, -3:], embds).flatten(2)
        sim = (sim.softmax(1) + sim.softmax(0)).mean(-1) / 2.
        sim[sim < self.detect_newly_object_threshold] = 0
        rows, cols = linear_sum_assignment((1 - sim).cpu())
        matched = sim[rows, cols]
        return torch.as_tensor(rows, device=sim.device), torch.as_tensor(cols, device=sim.device), matched

    def _refresh_matched(self, tg, rows, cols, logits, embds):
        tg["logits"][rows, -1] = 0.5 * (tg["logits"][rows, -1] + logits[cols])
        seen = (tg["embds"][rows, -1] != 0).any(-1)
        tg["embds"][rows, -1] = (tg["embds"][rows, -1] + embds[cols].mean(1)) / (seen[..., None] + 1.)

    def _add_matched_masks(self, tg, rows, cols, masks, quality, interim_size):
        n = masks.shape[1]
        up = _upsample(masks[cols], interim_size)
        tg["occurrence"][rows, -n:] += up.flatten(-2).gt(0.).any(-1).float()
        tg["mask_logits"][rows, -n:] += up.clone()
        tg["mask_quality_scores"][rows] += quality[cols]
        tg["masks"] = tg["mask_logits"].gt(0.).float()

    def _unseen(self, tg, matched_cols, logits, masks, min_score):
        """Confident predictions that were not matched and overlap no pooled entity in any frame of the clip (:640-646)."""
        n = masks.shape[1]
        pooled = _upsample(tg["mask_logits"][:, -n:], masks.shape[-2:]).transpose(0, 1).gt(0.)
        matched_cols = set(matched_cols.tolist())
        fresh = []
        for idx in range(masks.shape[0]):
            if idx in matched_cols or not logits[idx].max() > min_score:
                continue
            iou = group_mask_iou(masks[idx][:, None].gt(0.), pooled)
            if iou.nelement() and iou.max() < 0.5:
                fresh.append(idx)
        return fresh

    def detect_new_entities_instance(self, out, tg, interim_size):
        logits, masks, embds = out["pred_logits"].float(), out["pred_masks"].float(), out["pred_embds"].float()
        quality = calculate_mask_quality_scores(masks)
        logits = logits * quality.view(-1, 1)
        if self.stability_score_thresh > 0.:
            keep = quality > self.stability_score_thresh
            logits, masks, embds, quality = logits[keep], masks[keep], embds[keep], quality[keep]
        keep = logits.max(-1)[0].sort(descending=True)[1][:self.test_topk_per_image]
        logits, masks, embds, quality = logits[keep], masks[keep], embds[keep], quality[keep]
        h, w = masks.shape[-2:]
        boxes = mask_to_box(masks > 0) / torch.as_tensor([w, h, w, h], device=self._device)
        if masks.shape[0] > 1:           # box NMS over the clip: duplicates of a better-scoring query go
            order = logits.max(-1)[0].sort(descending=True)[1]
            overlap = video_box_iou(boxes[order], boxes[order]).max(-1)[0]
            keep = order[_suppress_later_duplicates(overlap, self.box_nms_thresh)]
            logits, masks, embds, boxes, quality = logits[keep], masks[keep], embds[keep], boxes[keep], quality[keep]
        if "masks" not in tg:            # first clip: every confident entity opens a slot
            fresh = logits.max(-1)[0] > max(self.apply_cls_thres, 0.1)
        else:
            rows, cols, sim = self._match_to_pool(tg, embds)
            ok = sim > self.detect_newly_object_threshold
            self._refresh_matched(tg, rows[ok], cols[ok], logits, embds)
            sure = sim > 2 * self.detect_newly_object_threshold
            self._add_matched_masks(tg, rows[sure], cols[sure], masks, quality, interim_size)
            # (the reference's membership test reads the indices of the stricter threshold, :625-627, :642)
            fresh = self._unseen(tg, cols[sure], logits, masks, self.apply_cls_thres)
        return {"pred_logits": logits[fresh], "pred_masks": masks[fresh], "pred_embds": embds[fresh],
                "pred_boxes": boxes[fresh], "mask_quality_scores": quality[fresh]}

    def detect_new_entities_pixel(self, out, tg, interim_size):
        logits, masks, embds = out["pred_logits"].float(), out["pred_masks"].float(), out["pred_embds"].float()
        h, w = masks.shape[-2:]
        boxes = mask_to_box(masks > 0) / torch.as_tensor([w, h, w, h], device=self._device)
        quality = calculate_mask_quality_scores(masks)
        logits = logits * quality.view(-1, 1)
        scores, labels = logits.max(-1)
        if "masks" not in tg:
            # first clip: things are de-duplicated by box IoU over the clip, stuff by mask IoU in the first frame
            order = scores.sort(descending=True)[1][:100]
            is_thing = torch.as_tensor([int(l) + 1 in self.thing_ids for l in labels[order]], dtype=torch.bool)
            things, stuff = order[is_thing], order[~is_thing]
            if len(things):
                things = things[:70]
                overlap = video_box_iou(boxes[things], boxes[things]).max(-1)[0]
                things = things[_suppress_later_duplicates(overlap, self.box_nms_thresh)]
            if len(stuff):
                stuff = stuff[:30]
                first = masks[stuff][:, 0].gt(0.).float().unsqueeze(0)
                stuff = stuff[_suppress_later_duplicates(group_mask_iou(first, first).max(0)[0], 0.6)]
            fresh = torch.cat([things, stuff])
            fresh = fresh[scores[fresh] > self.apply_cls_thres]
        else:
            rows, cols, sim = self._match_to_pool(tg, embds)
            ok = sim > self.detect_newly_object_threshold
            rows, cols = rows[ok], cols[ok]
            n = masks.shape[1]
            up = _upsample(masks[cols], interim_size)
            tg["mask_logits"][rows, -n:] += up.clone()
            tg["occurrence"][rows, -n:] += up.flatten(-2).gt(0.).any(-1).float()
            self._refresh_matched(tg, rows, cols, logits, embds)
            tg["mask_quality_scores"][rows] += quality[cols]
            tg["masks"] = tg["mask_logits"].gt(0.).float()
            fresh = self._unseen(tg, cols, logits, masks, 2 * self.apply_cls_thres)
        return {"pred_logits": logits[fresh], "pred_masks": masks[fresh], "pred_embds": embds[fresh],
                "pred_boxes": boxes[fresh], "mask_quality_scores": quality[fresh]}

    # ------------------------------------------------------------------ new entities -> pool (reference :767-876)
    def open_new_entities(self, first_frame_idx, new, tg, interim_size):
        dev = self._device
        logits = new["pred_logits"].unsqueeze(1)                          # [n, 1, K]
        embds = new["pred_embds"].mean(1, keepdim=True)                   # [n, 1, C]
        boxes, quality = new["pred_boxes"], new["mask_quality_scores"]
        n_new, n = new["pred_masks"].shape[:2]
        first_appear = torch.full((n_new,), first_frame_idx, dtype=torch.long, device=dev)
        if n_new == 0:
            masks = torch.zeros((0, self.num_frames, *interim_size), device=new["pred_masks"].device)
        else:
            masks = _upsample(new["pred_masks"], interim_size)
        occurrence = torch.ones(masks.shape[:2], device=masks.device)
        if "masks" not in tg:
            tg.update({"logits": logits, "masks": masks.gt(0.), "mask_logits": masks, "boxes": boxes, "embds": embds,
                       "ids": torch.arange(n_new, device=dev), "first_appear_frame_idxs": first_appear,
                       "mask_quality_scores": quality, "occurrence": occurrence})
            return
        if n_new == 0:
            return

        def left_pad(new_part, like, dim_size=None):
            """zeros for the frames / clips before the entity appeared, then its values"""
            shape = list(new_part.shape)
            shape[1] = like.shape[1] - (new_part.shape[1] if dim_size is None else dim_size)
            return torch.cat([torch.zeros(shape, dtype=torch.float32, device=dev), new_part], 1)

        masks_full = left_pad(masks, tg["masks"], n)
        tg.update({
            "logits": torch.cat([tg["logits"], left_pad(logits, tg["logits"])]),
            "masks": torch.cat([tg["masks"], masks_full.gt(0.)]),
            "mask_logits": torch.cat([tg["mask_logits"], masks_full]),
            "boxes": torch.cat([tg["boxes"], left_pad(boxes, tg["boxes"], n)]),
            "embds": torch.cat([tg["embds"], left_pad(embds, tg["embds"])]),
            "ids": torch.cat([tg["ids"], torch.arange(n_new, device=dev) + len(tg["ids"])]),
            "first_appear_frame_idxs": torch.cat([tg["first_appear_frame_idxs"], first_appear]),
            "mask_quality_scores": torch.cat([tg["mask_quality_scores"], quality]),
            "occurrence": torch.cat([tg["occurrence"], left_pad(occurrence, tg["occurrence"])]),
        })
        if "prompt_pe" in tg:             # the sampler's memory keeps one row per entity: blank rows for the new ones
            pe, feats, am = tg["prompt_pe"], tg["prompt_feats"], tg["prompt_attn_masks"]
            tg["prompt_pe"] = torch.cat([pe, torch.zeros((n_new, *pe.shape[1:]), device=dev)])
            tg["prompt_feats"] = torch.cat([feats, torch.zeros((n_new, *feats.shape[1:]), device=dev)])
            tg["prompt_attn_masks"] = torch.cat(
                [am, torch.zeros((am.shape[0], am.shape[1], n_new, am.shape[-1]), dtype=torch.bool, device=dev)], dim=-2)

    def open_slots_for_next_clip(self, tg, stride):
        """One more clip slot for logits / embds, `stride` more frame slots for masks / boxes / occurrence (:878-912)."""
        dev = self._device
        n = tg["embds"].shape[0]
        frames = torch.zeros((n, stride, *tg["masks"].shape[-2:]), dtype=torch.float, device=dev)
        tg.update({
            "logits": torch.cat([tg["logits"], tg["logits"][:, -1:].clone()], 1),
            "masks": torch.cat([tg["masks"], frames], 1),
            "mask_logits": torch.cat([tg["mask_logits"], frames], 1),
            "boxes": torch.cat([tg["boxes"], torch.zeros((n, stride, 4), dtype=torch.float32, device=dev)], 1),
            "embds": torch.cat([tg["embds"], tg["embds"][:, -3:].mean(1, keepdim=True).clone()], 1),
            "occurrence": torch.cat([tg["occurrence"], torch.zeros((n, stride), device=dev)], 1),
        })

    # ------------------------------------------------------------------ step 3: results (reference :914-1132)
    def save_results_vis(self, first_frame_idx, tg, image_size, out_size, is_last):
        if "masks" not in tg:
            return []                                                    # no entity detected so far
        start = min(first_frame_idx + self.num_frames, tg["video_len"]) - tg["mask_logits"].shape[1]
        quality = tg["mask_quality_scores"]
        scores = tg["logits"].mean(1).cpu()                              # [N, K]
        masks, occ = tg["mask_logits"], tg["occurrence"]
        if not is_last:
            masks, occ = masks[:, :self.num_frames_window_output], occ[:, :self.num_frames_window_output]
        masks = masks / occ[..., None, None].clamp(min=1)
        masks = _upsample(masks[:, :, :image_size[0], :image_size[1]].float(), out_size) > 0.
        N, W = masks.shape[:2]
        segms = rle.encode(masks.flatten(0, 1)) if masks.numel() else []
        results = []
        for i, obj_id in enumerate(tg["ids"]):
            res = {"obj_id": int(obj_id), "score": scores[i], "segmentations": segms[i * W:(i + 1) * W],
                   "frame_id_start": start}
            if is_last:
                res["mask_quality_score"] = quality[i] / (int(quality.max()) + 1)
            results.append(res)
        return results

    def save_results_vps(self, tg, image_size, out_size, is_last):
        masks = tg["mask_logits"]
        if not is_last:
            masks = masks[:, :self.num_frames_window_output]
        masks = _upsample(masks[:, :, :image_size[0], :image_size[1]].float(), out_size)
        things = tg.setdefault("thing_memory_list", {})                  # entity id -> segment id
        stuff = tg.setdefault("stuff_memory_list", {})                   # class id  -> segment id
        used = list(things.values()) + list(stuff.values())
        scores, classes = tg["logits"].mean(1).max(-1)
        classes = classes + 1                                            # dataset ids start from 1
        scores = scores * calculate_mask_quality_scores(masks)
        for k, c in enumerate(classes):
            if k not in things and int(c) not in self.thing_ids:        # things win ties against stuff
                scores[k] *= 0.75
        panoptic = torch.zeros((masks.size(1), out_size[0], out_size[1]), dtype=torch.int32, device=masks.device)
        if masks.shape[0] == 0:
            return panoptic.cpu()
        ids = tg["ids"]
        assert ids.min() == 0 and ids.max() == len(ids) - 1
        owner = (scores.view(-1, 1, 1, 1).to(masks.device) * masks).argmax(0)            # weighted *logits* (:1014)
        prob = masks.sigmoid()
        owner[(prob < 0.5).sum(0) == len(prob)] = -1
        next_id = max(used) + 1 if used else 0
        for k in range(classes.shape[0]):
            c, obj = int(classes[k]), int(ids[k])
            is_thing = c in self.thing_ids
            own, inside = owner == k, prob[k] >= 0.5
            area, full = int(own.sum()), int(inside.sum())
            region = own & inside
            if not (area > 0 and full > 0 and int(region.sum()) > 0):
                continue
            limit = 0.5 * self.overlap_threshold if obj in things else self.overlap_threshold
            if is_thing and area / full < limit:
                continue
            book, key = (things, obj) if is_thing else (stuff, c)
            if key not in book:
                book[key] = next_id + 1
                next_id += 1
            panoptic[region] = book[key]
        return panoptic.cpu()

    def vps_output_results(self, tg, panoptic_list, out_size):
        classes = tg["logits"].mean(1).max(-1)[1] + 1
        infos = [{"id": seg, "isthing": int(classes[obj]) in self.thing_ids, "category_id": int(classes[obj])}
                 for obj, seg in tg.get("thing_memory_list", {}).items()]
        infos += [{"id": seg, "isthing": False, "category_id": int(c)} for c, seg in tg.get("stuff_memory_list", {}).items()]
        return {"image_size": out_size, "pred_masks": torch.cat(panoptic_list, 0).cpu(), "segments_infos": infos,
                "task": "vps"}

    def save_results_vss(self, out, interim_size, image_size, out_size, is_last, stride):
        logits, masks = out["pred_logits"], out["pred_masks"]
        if not is_last:
            masks = masks[:, :stride]
        masks = _upsample(masks, interim_size)[:, :, :image_size[0], :image_size[1]]
        masks = F.interpolate(masks.float(), size=out_size, mode="nearest")
        logits = logits * calculate_mask_quality_scores(masks).view(-1, 1)
        return torch.einsum("qc,qthw->cthw", logits, masks.sigmoid()).argmax(0).cpu()
