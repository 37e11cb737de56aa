"""MaskFormerHead glue (mask2former/modeling/meta_arch/mask_former_head.py:19-191): builds the pixel decoder and the
transformer predictor and routes `features -> pixel_decoder.forward_features -> predictor(...)` (:148-154)."""
from __future__ import annotations

import torch.nn as nn

from ..registry import SEM_SEG_HEADS_REGISTRY, build_pixel_decoder, build_transformer_decoder, is_cfg


@SEM_SEG_HEADS_REGISTRY.register()
class MaskFormerHead(nn.Module):
    def __init__(self, input_shape, *args, **kwargs):
        super().__init__()
        if is_cfg(input_shape):
            kwargs = self.from_config(input_shape, *args)
            input_shape = kwargs.pop("input_shape")
        self._build(input_shape, **kwargs)

    @classmethod
    def from_config(cls, cfg, input_shape):
        """mask_former_head.py:100-143 (multi_scale_pixel_decoder route only)"""
        h, mf = cfg.MODEL.SEM_SEG_HEAD, cfg.MODEL.MASK_FORMER
        if mf.TRANSFORMER_IN_FEATURE != "multi_scale_pixel_decoder":
            raise NotImplementedError("only TRANSFORMER_IN_FEATURE='multi_scale_pixel_decoder' (Base.yaml:36) is built")
        return dict(input_shape={k: v for k, v in input_shape.items() if k in h.IN_FEATURES},
                    ignore_value=h.IGNORE_VALUE, num_classes=h.NUM_CLASSES,
                    pixel_decoder=build_pixel_decoder(cfg, input_shape), pixel_decoder_name=h.PIXEL_DECODER_NAME,
                    loss_weight=h.LOSS_WEIGHT, transformer_in_feature=mf.TRANSFORMER_IN_FEATURE,
                    transformer_predictor=build_transformer_decoder(cfg, h.CONVS_DIM, mask_classification=True))

    def _build(self, input_shape, *, num_classes, pixel_decoder, pixel_decoder_name="MSDeformAttnPixelDecoder",
               loss_weight=1.0, ignore_value=-1, transformer_predictor, transformer_in_feature="multi_scale_pixel_decoder",
               **_unused):
        shapes = sorted(input_shape.items(), key=lambda kv: kv[1].stride)
        self.in_features = [k for k, _ in shapes]
        self.ignore_value, self.loss_weight, self.num_classes = ignore_value, loss_weight, num_classes
        self.common_stride = 4
        self.pixel_decoder = pixel_decoder
        self.pixel_decoder_name = pixel_decoder_name
        self.predictor = transformer_predictor
        self.transformer_in_feature = transformer_in_feature
        if pixel_decoder_name != "MSDeformAttnPixelDecoder":
            raise NotImplementedError("only MSDeformAttnPixelDecoder is selected by the shipped UniVS configs")

    def forward(self, features, mask=None, targets=None):
        return self.layers(features, mask, targets)

    def layers(self, features, mask=None, targets=None):
        mask_features, mask_features_bfe_conv, _enc, multi_scale = self.pixel_decoder.forward_features(features)
        return self.predictor(multi_scale, mask_features, mask_features_bfe_conv, mask, targets)
