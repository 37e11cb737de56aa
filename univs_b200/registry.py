"""Registries the drop-in classes register under.  When detectron2 is importable its own registries are used, so
`MODEL.BACKBONE.NAME: D2SwinTransformer` etc. resolve to these classes (reference: swin.py:686, msdeformattn.py:166,
mask_former_head.py:19, ..._univs.py:27, univs_prompt.py:66); otherwise a minimal stand-in with the same
`.register()` / `.get()` surface is used."""
from __future__ import annotations

from collections import namedtuple


class _Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        def deco(o):
            self[o.__name__] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


try:  # pragma: no cover - detectron2 is not installed in the build container
    from detectron2.modeling import BACKBONE_REGISTRY, META_ARCH_REGISTRY, SEM_SEG_HEADS_REGISTRY
    from detectron2.layers import ShapeSpec
    from detectron2.utils.registry import Registry
    TRANSFORMER_DECODER_REGISTRY = Registry("TRANSFORMER_MODULE")
    HAVE_DETECTRON2 = True
except Exception:  # ImportError and friends
    BACKBONE_REGISTRY = _Registry("BACKBONE")
    SEM_SEG_HEADS_REGISTRY = _Registry("SEM_SEG_HEADS")
    META_ARCH_REGISTRY = _Registry("META_ARCH")
    TRANSFORMER_DECODER_REGISTRY = _Registry("TRANSFORMER_MODULE")
    ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=[None] * 4)
    HAVE_DETECTRON2 = False


def build_backbone(cfg, input_shape=None):
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)


def build_pixel_decoder(cfg, input_shape):
    """mask2former/modeling/pixel_decoder/fpn.py:21-33"""
    model = SEM_SEG_HEADS_REGISTRY.get(cfg.MODEL.SEM_SEG_HEAD.PIXEL_DECODER_NAME)(cfg, input_shape)
    if not callable(getattr(model, "forward_features", None)):
        raise ValueError("Pixel decoders must expose forward_features()")
    return model


def build_transformer_decoder(cfg, in_channels, mask_classification=True):
    """mask2former/modeling/transformer_decoder/maskformer_transformer_decoder.py:22-27"""
    name = cfg.MODEL.MASK_FORMER.TRANSFORMER_DECODER_NAME
    return TRANSFORMER_DECODER_REGISTRY.get(name)(cfg, in_channels, mask_classification)


def build_sem_seg_head(cfg, input_shape):
    return SEM_SEG_HEADS_REGISTRY.get(cfg.MODEL.SEM_SEG_HEAD.NAME)(cfg, input_shape)


def is_cfg(obj) -> bool:
    return hasattr(obj, "MODEL") and hasattr(obj.MODEL, "META_ARCHITECTURE")
