"""Config node with the keys of the reference's stacked `add_*_config` functions that the hot path reads
(mask2former/config.py:6-129 `MODEL.MASK_FORMER.*`, `MODEL.SWIN.*`; univs/config.py:4-160 `MODEL.UniVS.*`,
`MODEL.BoxVIS.TEST.*`, `INPUT.*`).  When detectron2 is importable use its CfgNode + the reference's add_*_config
instead; the modules only need attribute access.  YAML files with `_BASE_` inheritance (configs/**) load through
`merge_from_file`; unknown keys are accepted (the reference YAMLs carry many training-only keys)."""
from __future__ import annotations

import copy
import os

import yaml


class CfgNode(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_from_dict(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def merge_from_file(self, path):
        with open(path) as f:
            d = yaml.safe_load(f) or {}
        base = d.pop("_BASE_", None)
        if base:
            self.merge_from_file(os.path.join(os.path.dirname(path), base))
        self.merge_from_dict(d)

    def merge_from_list(self, kv):
        assert len(kv) % 2 == 0
        for k, v in zip(kv[0::2], kv[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = yaml.safe_load(v) if isinstance(v, str) else v


def get_cfg() -> CfgNode:
    """Defaults = detectron2 defaults + add_maskformer2_config + add_univs_config, restricted to what is read."""
    return CfgNode({
        "MODEL": {
            "META_ARCHITECTURE": "UniVS_Prompt",
            "DEVICE": "cuda",
            "PIXEL_MEAN": [123.675, 116.280, 103.530],
            "PIXEL_STD": [58.395, 57.120, 57.375],
            "BACKBONE": {"NAME": "D2SwinTransformer", "FREEZE_AT": 0},
            "SWIN": {"PRETRAIN_IMG_SIZE": 224, "PATCH_SIZE": 4, "EMBED_DIM": 96, "DEPTHS": [2, 2, 6, 2],
                     "NUM_HEADS": [3, 6, 12, 24], "WINDOW_SIZE": 7, "MLP_RATIO": 4.0, "QKV_BIAS": True,
                     "QK_SCALE": None, "DROP_RATE": 0.0, "ATTN_DROP_RATE": 0.0, "DROP_PATH_RATE": 0.3, "APE": False,
                     "PATCH_NORM": True, "OUT_FEATURES": ["res2", "res3", "res4", "res5"], "USE_CHECKPOINT": False},
            "SEM_SEG_HEAD": {"NAME": "MaskFormerHead", "IGNORE_VALUE": 255, "NUM_CLASSES": 133, "LOSS_WEIGHT": 1.0,
                             "CONVS_DIM": 256, "MASK_DIM": 256, "NORM": "GN",
                             "PIXEL_DECODER_NAME": "MSDeformAttnPixelDecoder",
                             "IN_FEATURES": ["res2", "res3", "res4", "res5"],
                             "DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES": ["res3", "res4", "res5"],
                             "COMMON_STRIDE": 4, "TRANSFORMER_ENC_LAYERS": 6},
            "MASK_FORMER": {"TRANSFORMER_DECODER_NAME": "VideoMultiScaleMaskedTransformerDecoderUniVS",
                            "TRANSFORMER_IN_FEATURE": "multi_scale_pixel_decoder", "HIDDEN_DIM": 256,
                            "NUM_OBJECT_QUERIES": 200, "NHEADS": 8, "DROPOUT": 0.0, "DIM_FEEDFORWARD": 2048,
                            "ENC_LAYERS": 0, "PRE_NORM": False, "ENFORCE_INPUT_PROJ": False, "SIZE_DIVISIBILITY": 32,
                            "DEC_LAYERS": 10,
                            "TEST": {"SEMANTIC_ON": False, "INSTANCE_ON": True, "PANOPTIC_ON": False,
                                     "OVERLAP_THRESHOLD": 0.8, "OBJECT_MASK_THRESHOLD": 0.05,
                                     "STABILITY_SCORE_THRESH": 0.0}},
            "BoxVIS": {"TEST": {"NUM_FRAMES_WINDOW": 5, "CLIP_STRIDE": 1, "NUM_FRAMES": 3, "TRACKER_TYPE": "minvis",
                                "ZERO_SHOT_INFERENCE": False, "MERGE_ON_CPU": False, "NUM_MAX_INST": 50,
                                "APPLY_CLS_THRES": 0.05, "MULTI_CLS_ON": True, "WINDOW_INFERENCE": False}},
            "UniVS": {"CLIP_CLASS_EMBED_PATH": "datasets/concept_emb/combined_datasets_cls_emb_rn50x4.pth",
                      "VISUAL_PROMPT_ENCODER": True, "TEXT_PROMPT_ENCODER": True, "PROMPT_AS_QUERIES": True,
                      "TEXT_PROMPT_TO_IMAGE_ENABLE": True, "MASKDEC_SELF_ATTN_MASK_TYPE": "sep",
                      "DISABLE_LEARNABLE_QUERIES_SA1B": False, "VISUAL_PROMPT_PIXELS_PER_IMAGE": 32,
                      "PROMPT_SELF_ATTN_LAYERS": -1, "POSITION_EMBEDDING_SINE3D": "ArbitraryT",
                      "TEST": {"NUM_PREV_FRAMES_MEMORY": 5, "ENABLED_PREV_FRAMES_MEMORY": True,
                               "ENABLED_PREV_VISUAL_PROMPTS_FOR_GROUNDING": False, "CUSTOM_VIDEOS_TEXT": [],
                               "SEMANTIC_EXTRACTION": {"ENABLE": False}}},
        },
        "INPUT": {"SAMPLING_FRAME_NUM": 5, "FORMAT": "RGB", "LSJ_AUG": {"IMAGE_SIZE": 1024, "SQUARE_ENABLED": False}},
        "TEST": {"DETECTIONS_PER_IMAGE": 100},
    })


SWIN_VARIANTS = {
    # configs/univs/univs_swin{t,b,l}_stage1.yaml:5-9
    "tiny": dict(EMBED_DIM=96, DEPTHS=[2, 2, 6, 2], NUM_HEADS=[3, 6, 12, 24], WINDOW_SIZE=7),
    "base": dict(EMBED_DIM=128, DEPTHS=[2, 2, 18, 2], NUM_HEADS=[4, 8, 16, 32], WINDOW_SIZE=12),
    "large": dict(EMBED_DIM=192, DEPTHS=[2, 2, 18, 2], NUM_HEADS=[6, 12, 24, 48], WINDOW_SIZE=12),
}
