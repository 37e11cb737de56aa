"""Sliding-window task heads over the hot path (callers of `model.backbone` / `model.sem_seg_head`,
univs/inference/*).  Built on `ClipStream`: every frame is encoded once."""
from .comm import (TemporalMaskMean, calculate_mask_quality_scores, check_consistency_with_prev_frames,
                   generate_temporal_weights, is_semseg_dataset, match_from_learnable_embds, pair_mask_iou,
                   process_inference, video_box_iou)
from .image_seg import InferenceImageGenericSeg, classwise_box_nms
from .video_entity import InferenceVideoEntity, results_to_coco_video
from .video_semantic_extraction import InferenceVideoSemanticExtraction
from .video_vis_fast import InferenceVideoVISFast
from .video_vos import FrameAnnotations, InferenceVideoVOS
from .video_vps import InferenceVideoVPS

__all__ = ["InferenceVideoVISFast", "InferenceVideoVOS", "InferenceVideoVPS", "process_inference", "is_semseg_dataset", "InferenceVideoEntity", "InferenceVideoSemanticExtraction", "InferenceImageGenericSeg", "classwise_box_nms", "results_to_coco_video", "FrameAnnotations", "match_from_learnable_embds",
           "check_consistency_with_prev_frames", "generate_temporal_weights", "calculate_mask_quality_scores",
           "video_box_iou", "pair_mask_iou", "TemporalMaskMean"]
