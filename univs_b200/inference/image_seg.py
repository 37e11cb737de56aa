"""Image-level generic segmentation head for the COCO / ADE20K vocabularies (univs/inference/inference_image_generic_seg.py):
one image = a clip of one frame through the same hot path, then Mask2Former-style semantic / panoptic / instance
post-processing with class-wise box NMS.

Same contract as the reference class: `eval(model, batched_inputs)` returns a list with one dict per image holding
"sem_seg" [K, H, W], "panoptic_seg" (int32 [H, W], segments_info) and / or "instances" -- here a plain dict
{"image_size", "pred_masks" float [n, H, W], "pred_boxes" [n, 4] XYXY pixels, "scores" [n], "pred_classes" [n]} with the
fields the reference puts into a detectron2 `Instances` (:421-430).  `thing_contiguous_ids` replaces
`metadata.thing_dataset_id_to_contiguous_id.values()` (0-based class indices of the thing classes).

Everything stays on the model's device; the NMS works on the [n, n] IoU matrix (one device pass) and a host loop over at
most a few hundred boxes (torchvision's `batched_nms` semantics: per class, highest score first, suppress IoU > thr)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ..modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO
from ..modeling.visual_prompts import mask_to_box
from ..registry import is_cfg
from .comm import calculate_mask_quality_scores, process_inference


def classwise_box_nms(boxes, scores, labels, iou_threshold):
    """torchvision.ops.batched_nms (coordinate-offset formulation, boxes.py:35-80): indices of the kept boxes, highest
    score first.  boxes [n, 4] XYXY, scores [n], labels [n]."""
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    b = boxes.float()
    b = b + (labels.to(b) * (b.max() + 1))[:, None]                     # classes never overlap
    order = scores.sort(descending=True, stable=True)[1]
    b = b[order]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt, rb = torch.maximum(b[:, None, :2], b[None, :, :2]), torch.minimum(b[:, None, 2:], b[None, :, 2:])
    inter = (rb - lt).clamp(min=0).prod(-1)
    over = (inter / (area[:, None] + area[None] - inter) > iou_threshold).cpu().numpy()       # NaN (0 / 0) compares false
    alive, keep = [True] * n, []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(i)
        row = over[i]
        for j in range(i + 1, n):
            if row[j]:
                alive[j] = False
    return order[torch.as_tensor(keep, dtype=torch.int64, device=boxes.device)]


def resize_to_output(x, image_size, height, width):
    """detectron2 sem_seg_postprocess: crop the padding, bilinear resize to the output resolution.  x [C, Hp, Wp]"""
    x = x[:, :image_size[0], :image_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(x, size=(height, width), mode="bilinear", align_corners=False)[0]


class InferenceImageGenericSeg(nn.Module):
    def __init__(self, cfg=None, *, num_queries=200, size_divisibility=32, object_mask_threshold=0.05, overlap_threshold=0.8,
                 stability_score_thresh=0.0, prompt_as_queries=True, semantic_on=False, instance_on=True, panoptic_on=False,
                 disable_semantic_queries=False, test_topk_per_image=100, sem_seg_postprocess_before_inference=False,
                 thing_contiguous_ids=(), lsj_aug_enable_test=False, lsj_aug_image_size=1024):
        super().__init__()
        if cfg is not None and is_cfg(cfg):
            mf = cfg.MODEL.MASK_FORMER
            num_queries = mf.NUM_OBJECT_QUERIES
            size_divisibility = mf.SIZE_DIVISIBILITY
            object_mask_threshold = mf.TEST.OBJECT_MASK_THRESHOLD
            overlap_threshold = mf.TEST.OVERLAP_THRESHOLD
            stability_score_thresh = mf.TEST.get("STABILITY_SCORE_THRESH", 0.0)
            prompt_as_queries = cfg.MODEL.UniVS.PROMPT_AS_QUERIES
            semantic_on, instance_on, panoptic_on = mf.TEST.SEMANTIC_ON, mf.TEST.INSTANCE_ON, mf.TEST.PANOPTIC_ON
            disable_semantic_queries = cfg.MODEL.UniVS.TEST.get("DISABLE_SEMANTIC_QUERIES", False)
            test_topk_per_image = cfg.get("TEST", {}).get("DETECTIONS_PER_IMAGE", 100)
            sem_seg_postprocess_before_inference = mf.TEST.get("SEM_SEG_POSTPROCESSING_BEFORE_INFERENCE", False)
            lsj_aug_enable_test = cfg.INPUT.LSJ_AUG.SQUARE_ENABLED
            lsj_aug_image_size = cfg.INPUT.LSJ_AUG.IMAGE_SIZE
        self.num_queries = num_queries
        self.size_divisibility = size_divisibility
        self.object_mask_threshold = object_mask_threshold
        self.overlap_threshold = overlap_threshold
        self.stability_score_thresh = stability_score_thresh
        self.prompt_as_queries = prompt_as_queries
        self.semantic_on, self.instance_on, self.panoptic_on = semantic_on, instance_on, panoptic_on
        self.disable_semantic_queries = disable_semantic_queries
        self.test_topk_per_image = test_topk_per_image
        self.sem_seg_postprocess_before_inference = sem_seg_postprocess_before_inference
        self.thing_contiguous_ids = [int(i) for i in thing_contiguous_ids]
        self.LSJ_aug_enable_test, self.LSJ_aug_image_size = lsj_aug_enable_test, lsj_aug_image_size

    # ------------------------------------------------------------------ entry point (reference :175-209)
    @torch.no_grad()
    def eval(self, model, batched_inputs):
        if len(batched_inputs) != 1 or len(batched_inputs[0]["image"]) != 1:
            raise ValueError("one image per call (the reference reads frame 0 of a batch-size-1 clip, :218-219)")
        inp = batched_inputs[0]
        name = inp["dataset_name"]
        if not (name.startswith("coco") or name.startswith("ade20k")):
            raise ValueError(f"do not support the model inference on {name}.")
        x, image_size = model.preprocess(inp["image"])
        if self.LSJ_aug_enable_test:
            d, S = self.size_divisibility, self.LSJ_aug_image_size
            S = (max(S, *x.shape[-2:]) + d - 1) // d * d
            x = F.pad(x, (0, S - x.shape[-1], 0, S - x.shape[-2]), value=0.0)
        targets = inp.get("targets")
        if targets is None:
            targets = process_inference(inp, tuple(x.shape[-2:]), image_size, 1, semantic_on=self.semantic_on)
            targets[0]["frame_indices"] = torch.arange(1)
        return self.inference_image(model, batched_inputs, x, image_size, targets)

    # ------------------------------------------------------------------ reference :211-289
    @torch.no_grad()
    def inference_image(self, model, batched_inputs, x, image_size, targets):
        out = model.sem_seg_head(model.backbone(x), targets=targets)
        name = batched_inputs[0]["dataset_name"]
        n_cls, start = COMBINED_DATASETS_CATEGORY_INFO[name]
        cls = out["pred_logits"][0, :, start:start + n_cls].sigmoid()                       # [Q, K]
        masks = F.interpolate(out["pred_masks"][..., 0, :, :], size=tuple(x.shape[-2:]), mode="bilinear",
                              align_corners=False)[0]                                        # [Q, Hp, Wp]
        height, width = batched_inputs[0].get("height", image_size[0]), batched_inputs[0].get("width", image_size[1])
        quality = calculate_mask_quality_scores(masks)
        cls = cls * quality.unsqueeze(-1)
        if self.stability_score_thresh > 0:
            keep = quality > self.stability_score_thresh
            cls, masks = cls[keep], masks[keep]
        if self.sem_seg_postprocess_before_inference:
            masks = resize_to_output(masks, image_size, height, width)
        else:
            masks = masks[:, :image_size[0], :image_size[1]]
        result = {}
        if self.semantic_on:
            r = self.semantic_inference(cls, masks)
            if not self.sem_seg_postprocess_before_inference:
                r = resize_to_output(r, image_size, height, width)
            result["sem_seg"] = r
        if self.panoptic_on:
            seg, infos = self.panoptic_inference(cls, masks)
            if not self.sem_seg_postprocess_before_inference:
                seg = F.interpolate(seg[None, None].float(), size=(height, width), mode="nearest")[0, 0].to(torch.int32)
                present = set(seg.unique().tolist())
                infos = [i for i in infos if i["id"] in present]
            result["panoptic_seg"] = (seg, infos)
        if self.instance_on:
            result["instances"] = self.instance_inference(cls, masks, (height, width))
        return [result]

    def _without_thing_prompts(self, cls, masks):
        """category-prompt queries (rows >= num_queries, one per class) of thing classes are dropped (:307-313)"""
        keep = [i for i in range(cls.shape[0])
                if i < self.num_queries or i - self.num_queries not in self.thing_contiguous_ids]
        return cls[keep], masks[keep]

    def semantic_inference(self, cls, masks):
        if self.prompt_as_queries and self.disable_semantic_queries:
            cls, masks = cls[self.num_queries:], masks[self.num_queries:]
        top = torch.topk(cls.max(-1)[0], k=min(200, cls.shape[0]))[1]       # (the reference asks for exactly 200)
        cls, masks = cls[top], masks[top]
        return torch.einsum("qc,qhw->chw", (cls / 0.06).softmax(-1), masks.sigmoid())      # 0.06: open-vocabulary temperature

    def panoptic_inference(self, cls, masks):
        if self.prompt_as_queries:
            cls, masks = self._without_thing_prompts(cls, masks)
        cls, masks, _ = self.postprocess_nms(cls, masks, biou_threshold=0.9)
        keep = cls.max(-1)[0] > self.object_mask_threshold
        scores, labels = (cls / 0.06).softmax(-1).max(-1)
        scores, labels, prob = scores[keep], labels[keep], masks.sigmoid()[keep]
        h, w = prob.shape[-2:]
        seg = torch.zeros((h, w), dtype=torch.int32, device=prob.device)
        infos = []
        if prob.shape[0] == 0:
            return seg, infos
        owner = (scores.view(-1, 1, 1) * prob).argmax(0)
        stuff, next_id = {}, 0
        for k in range(prob.shape[0]):
            c = int(labels[k])
            is_thing = c in self.thing_contiguous_ids
            own, inside = owner == k, prob[k] >= 0.5
            area, full = int(own.sum()), int(inside.sum())
            region = own & inside
            if not (area > 0 and full > 0 and int(region.sum()) > 0) or area / full < self.overlap_threshold:
                continue
            if not is_thing:
                if c in stuff:                                            # stuff regions of one class share a segment
                    seg[region] = stuff[c]
                    continue
                stuff[c] = next_id + 1
            next_id += 1
            seg[region] = next_id
            infos.append({"id": next_id, "isthing": bool(is_thing), "category_id": c})
        return seg, infos

    def instance_inference(self, cls, masks, out_size):
        image_size = tuple(masks.shape[-2:])
        boxes = mask_to_box(masks.gt(0))
        if self.prompt_as_queries:
            cls, masks, boxes = cls[:self.num_queries], masks[:self.num_queries], boxes[:self.num_queries]
        things = self.thing_contiguous_ids
        if len(things) != cls.shape[-1]:          # panoptic vocabulary: instances are the thing classes only (:369-382)
            labels = cls.max(-1)[1]
            cls = cls[..., things]
            keep = torch.as_tensor([int(l) in things for l in labels], dtype=torch.bool, device=cls.device)
            if keep.sum() == 0:
                s = cls.max(-1)[0]
                keep = s >= min(0.1, s.max())
            cls, masks, boxes = cls[keep], masks[keep], boxes[keep]
        cls, masks, boxes = self.postprocess_nms(cls, masks, boxes)
        K = cls.shape[-1]
        scores, top = cls.flatten(0, 1).topk(min(self.test_topk_per_image, cls.nelement()), sorted=False)
        labels = top % K
        top = torch.div(top, K, rounding_mode="floor").long()
        masks, boxes = masks[top], boxes[top]
        if image_size != tuple(out_size):
            masks = F.interpolate(masks[None], size=out_size, mode="bilinear", align_corners=False)[0]
            boxes = mask_to_box(masks.gt(0))
        return {"image_size": tuple(out_size), "pred_masks": (masks > 0).float(), "pred_boxes": boxes, "scores": scores,
                "pred_classes": labels}

    def postprocess_nms(self, cls, masks, boxes=None, biou_threshold=0.85):
        if boxes is None:
            boxes = mask_to_box(masks.gt(0.))
        s, l = cls.max(-1)
        keep = classwise_box_nms(boxes.float(), s, l, biou_threshold)
        return cls[keep], masks[keep], boxes[keep]
