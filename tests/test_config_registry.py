"""Host-side boundary checks that need no GPU: registry names, cfg-driven construction (the `@configurable`
`(cfg, ...)` calling convention), YAML loading with _BASE_, preprocessing (normalise + pad to /32)."""
import os

import pytest
import torch

from univs_b200 import modeling  # noqa: F401
from univs_b200.build import build_model, make_cfg
from univs_b200.config import get_cfg
from univs_b200.registry import (BACKBONE_REGISTRY, META_ARCH_REGISTRY, SEM_SEG_HEADS_REGISTRY,
                                 TRANSFORMER_DECODER_REGISTRY)


def test_reference_registry_names_resolve():
    import univs_b200.meta_arch  # noqa: F401
    assert BACKBONE_REGISTRY.get("D2SwinTransformer").__name__ == "D2SwinTransformer"
    assert SEM_SEG_HEADS_REGISTRY.get("MaskFormerHead").__name__ == "MaskFormerHead"
    assert SEM_SEG_HEADS_REGISTRY.get("MSDeformAttnPixelDecoder").__name__ == "MSDeformAttnPixelDecoder"
    assert TRANSFORMER_DECODER_REGISTRY.get("VideoMultiScaleMaskedTransformerDecoderUniVS") is not None
    assert META_ARCH_REGISTRY.get("UniVS_Prompt").__name__ == "UniVS_Prompt"


def test_build_from_cfg_and_state_dict_layout():
    cfg = make_cfg("tiny", num_queries=7, num_frames=2, clip_emb=torch.randn(3938, 640))
    cfg.MODEL.SWIN.DEPTHS = [1, 1, 1, 1]
    cfg.MODEL.SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS = 1
    cfg.MODEL.MASK_FORMER.DEC_LAYERS = 2
    model = build_model(cfg)
    sd = model.state_dict()
    for k in ("backbone.patch_embed.proj.weight", "backbone.layers.0.blocks.0.attn.relative_position_bias_table",
              "backbone.layers.0.blocks.0.attn.relative_position_index", "backbone.layers.0.downsample.reduction.weight",
              "backbone.norm3.weight", "sem_seg_head.pixel_decoder.input_proj.0.0.weight",
              "sem_seg_head.pixel_decoder.transformer.level_embed",
              "sem_seg_head.pixel_decoder.transformer.encoder.layers.0.self_attn.sampling_offsets.weight",
              "sem_seg_head.pixel_decoder.adapter_1.norm.weight", "sem_seg_head.pixel_decoder.layer_1.weight",
              "sem_seg_head.pixel_decoder.mask_features.bias",
              "sem_seg_head.predictor.transformer_cross_attention_layers.0.multihead_attn.in_proj_weight",
              "sem_seg_head.predictor.transformer_prompt_self_attention_layers.0.multihead_attn.out_proj.bias",
              "sem_seg_head.predictor.query_feat.weight", "sem_seg_head.predictor.mask_embed.layers.2.weight",
              "sem_seg_head.predictor.lang2vision_cross_attention_layer.norm.weight",
              "sem_seg_head.predictor.cls_temp.weight", "sem_seg_head.predictor.prompt_sot.weight"):
        assert k in sd, k
    assert "pixel_mean" not in sd and "pixel_std" not in sd          # non-persistent (univs_prompt.py:169-170)
    assert sd["sem_seg_head.predictor.query_feat.weight"].shape == (7, 256)
    assert model.backbone.size_divisibility == 32
    assert model.backbone.output_shape()["res5"].channels == 768
    # static_query -> query_feat upgrade hook (..._univs.py:32-53)
    old = {k.replace("query_feat", "static_query"): v for k, v in model.sem_seg_head.predictor.state_dict().items()}
    model.sem_seg_head.predictor.load_state_dict(old)


def test_preprocess_normalise_and_pad():
    cfg = make_cfg("tiny", 5, 2, clip_emb=torch.randn(3938, 640))
    cfg.MODEL.SWIN.DEPTHS = [1, 1, 1, 1]
    cfg.MODEL.SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS = 1
    cfg.MODEL.MASK_FORMER.DEC_LAYERS = 2
    m = build_model(cfg)
    frames = (torch.rand(2, 3, 50, 70) * 255).to(torch.uint8)
    x, size = m.preprocess(frames)
    assert x.shape == (2, 3, 64, 96) and size == (50, 70)
    want = (frames.float() - m.pixel_mean) / m.pixel_std
    assert torch.allclose(x[:, :, :50, :70], want)
    assert x[:, :, 50:].abs().sum() == 0 and x[:, :, :, 70:].abs().sum() == 0   # zero padding AFTER normalisation
    x2, _ = m.preprocess([f for f in frames])
    assert torch.equal(x, x2)


@pytest.mark.reference
def test_reference_yaml_loads_with_base():
    cfg = get_cfg()
    cfg.merge_from_file("/root/reference/configs/univs/univs_swinl_stage1.yaml")
    assert cfg.MODEL.SWIN.EMBED_DIM == 192 and cfg.MODEL.SWIN.WINDOW_SIZE == 12
    assert cfg.MODEL.MASK_FORMER.NUM_OBJECT_QUERIES == 200 and cfg.MODEL.MASK_FORMER.DEC_LAYERS == 10
    assert cfg.MODEL.META_ARCHITECTURE == "UniVS_Prompt"
    cfg.merge_from_list(["MODEL.UniVS.TEST.NUM_PREV_FRAMES_MEMORY", "10", "INPUT.SAMPLING_FRAME_NUM", "5"])
    assert cfg.MODEL.UniVS.TEST.NUM_PREV_FRAMES_MEMORY == 10


def test_training_mode_is_rejected():
    cfg = make_cfg("tiny", 5, 2, clip_emb=torch.randn(3938, 640))
    cfg.MODEL.SWIN.DEPTHS = [1, 1, 1, 1]
    cfg.MODEL.SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS = 1
    cfg.MODEL.MASK_FORMER.DEC_LAYERS = 2
    m = build_model(cfg)
    m.train()
    with pytest.raises(NotImplementedError):
        m([{"image": [torch.zeros(3, 32, 32)] * 2}])


def test_switch_defaults_are_the_round1_path(monkeypatch):
    """every opt-in path is off unless asked for (environment > univs_b200/tuned.json > built-in default)"""
    import importlib
    from univs_b200 import switches
    for k in switches.DEFAULTS:
        monkeypatch.delenv("UNIVS_" + k, raising=False)
    importlib.reload(switches)
    assert switches.active() == {k: v for k, v in switches._TUNED.items() if v != switches.DEFAULTS[k]}
    monkeypatch.setenv("UNIVS_WIN_TC", "1")
    monkeypatch.setenv("UNIVS_FRAME_STREAMS", "2")
    assert switches.get("win_tc") == 1 and switches.active()["FRAME_STREAMS"] == 2


def test_task_heads_build_from_the_model_cfg():
    """attach_task_heads() reads the keys the reference heads' from_config read (defaults where a config omits them)"""
    cfg = make_cfg("tiny", 10, 2, clip_emb=torch.randn(3938, 640))
    model = build_model(cfg)
    assert model.task_heads is None
    model.attach_task_heads(thing_ids={1, 2}, thing_contiguous_ids=[0, 1])
    h = model.task_heads
    assert {"vis_fast", "vos", "vps", "entity", "image", "semantic_extraction"} <= set(h)
    assert h["entity"].num_frames == 2 and h["entity"].num_queries == 10 and h["entity"].num_frames_window_output == 10
    assert h["vps"].thing_ids == {1, 2} and h["image"].thing_contiguous_ids == [0, 1]
    assert h["unified"] is False and h["tracker_type"] == "minvis"
    cfg.MODEL.UniVS.TEST["VIDEO_UNIFIED_INFERENCE_ENABLE"] = True
    assert build_model(cfg).attach_task_heads().task_heads["unified"] is True
