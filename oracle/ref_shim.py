"""TEST INFRASTRUCTURE ONLY -- loader that executes the reference's own hot-path
source files (read-only at /root/reference) on CPU, without detectron2 / timm /
fvcore / the compiled MultiScaleDeformableAttention extension.

Nothing in the product package (univs_b200/) imports this module.  It is used by
 * tests/golden/make_golden.py  -- to generate the committed golden vectors,
 * tests/test_oracle_*.py        -- to pin oracle/ops_ref.py against the reference
                                    (skipped when /root/reference is absent, i.e.
                                    on the GPU box),
 * bench.py --impl reference     -- never (the reference tree does not travel).

What is stubbed (third-party symbols the reference imports):
  timm.models.layers.{DropPath,to_2tuple,trunc_normal_}
  fvcore.nn.weight_init.c2_xavier_fill
  detectron2.config.configurable            (returns the plain __init__)
  detectron2.layers.{Conv2d,ShapeSpec,get_norm,DeformConv}
  detectron2.modeling.{BACKBONE,SEM_SEG_HEADS,META_ARCH}_REGISTRY, Backbone
  detectron2.utils.registry.Registry
  detectron2.projects.point_rend.point_features.point_sample
  MultiScaleDeformableAttention             (empty; MSDeformAttnFunction.apply is
                                             routed to the reference's own
                                             ms_deform_attn_core_pytorch,
                                             ops/functions/ms_deform_attn_func.py:52-72)
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("UNIVS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "univs"))


class _Registry(dict):
    def __init__(self, name="registry"):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        return self[name]


class _DropPath(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        return x


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class _Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d semantics: conv -> norm -> activation."""

    def __init__(self, *a, **kw):
        norm = kw.pop("norm", None)
        act = kw.pop("activation", None)
        super().__init__(*a, **kw)
        self.norm = norm
        self.activation = act

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def _get_norm(norm, ch):
    if norm is None or norm == "":
        return None
    if norm == "GN":
        return nn.GroupNorm(32, ch)
    raise ValueError(norm)


_ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=[None] * 4)


def _configurable(init_func=None, *, from_config=None):
    if init_func is not None:
        return init_func

    def wrap(f):
        return f
    return wrap


def _point_sample(inp, coords, **kw):
    add_dim = False
    if coords.dim() == 3:
        add_dim = True
        coords = coords.unsqueeze(2)
    out = F.grid_sample(inp, 2.0 * coords - 1.0, **kw)
    if add_dim:
        out = out.squeeze(3)
    return out


def _c2_xavier_fill(m):
    nn.init.kaiming_uniform_(m.weight, a=1)
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


_LOADED = None


def load():
    """Install the stubs and import the reference hot-path files by path.

    Returns a namespace with the reference classes / functions."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    # ---- third-party stubs -------------------------------------------------
    _mod("timm"); _mod("timm.models")
    _mod("timm.models.layers", DropPath=_DropPath, to_2tuple=_to_2tuple,
         trunc_normal_=nn.init.trunc_normal_)
    _mod("fvcore"); _mod("fvcore.nn")
    wi = _mod("fvcore.nn.weight_init", c2_xavier_fill=_c2_xavier_fill)
    sys.modules["fvcore.nn"].weight_init = wi
    _mod("detectron2")
    _mod("detectron2.config", configurable=_configurable)
    _mod("detectron2.layers", Conv2d=_Conv2d, ShapeSpec=_ShapeSpec, get_norm=_get_norm,
         DeformConv=object, ModulatedDeformConv=object)
    _mod("detectron2.modeling", BACKBONE_REGISTRY=_Registry("BACKBONE"),
         SEM_SEG_HEADS_REGISTRY=_Registry("SEM_SEG_HEADS"),
         META_ARCH_REGISTRY=_Registry("META_ARCH"), Backbone=nn.Module, ShapeSpec=_ShapeSpec)
    _mod("detectron2.utils"); _mod("detectron2.utils.registry", Registry=_Registry)
    _mod("detectron2.projects"); _mod("detectron2.projects.point_rend")
    _mod("detectron2.projects.point_rend.point_features", point_sample=_point_sample)
    _mod("MultiScaleDeformableAttention")

    # ---- path-only namespace packages (their __init__.py are NOT executed) --
    R = REF_ROOT
    _pkg("mask2former", f"{R}/mask2former")
    _pkg("mask2former.modeling", f"{R}/mask2former/modeling")
    _pkg("mask2former.modeling.backbone", f"{R}/mask2former/modeling/backbone")
    _pkg("mask2former.modeling.pixel_decoder", f"{R}/mask2former/modeling/pixel_decoder")
    _pkg("mask2former.modeling.pixel_decoder.ops", f"{R}/mask2former/modeling/pixel_decoder/ops")
    _pkg("mask2former.modeling.pixel_decoder.ops.functions",
         f"{R}/mask2former/modeling/pixel_decoder/ops/functions")
    _pkg("mask2former.modeling.pixel_decoder.ops.modules",
         f"{R}/mask2former/modeling/pixel_decoder/ops/modules")
    _pkg("mask2former.modeling.transformer_decoder", f"{R}/mask2former/modeling/transformer_decoder")
    _pkg("mask2former.modeling.meta_arch", f"{R}/mask2former/modeling/meta_arch")
    _pkg("univs", f"{R}/univs")
    _pkg("univs.modeling", f"{R}/univs/modeling")
    _pkg("univs.modeling.transformer_decoder", f"{R}/univs/modeling/transformer_decoder")
    pe_pkg = _pkg("univs.modeling.prompt_encoder", f"{R}/univs/modeling/prompt_encoder")
    lang = _pkg("univs.modeling.language", f"{R}/univs/modeling/language")
    lang.pre_tokenize_expression = lambda *a, **k: None
    _pkg("univs.utils", f"{R}/univs/utils")
    _pkg("datasets", f"{R}/datasets")
    _pkg("datasets.concept_emb", f"{R}/datasets/concept_emb")

    imp = importlib.import_module
    func = imp("mask2former.modeling.pixel_decoder.ops.functions.ms_deform_attn_func")
    fpk = sys.modules["mask2former.modeling.pixel_decoder.ops.functions"]
    fpk.ms_deform_attn_core_pytorch = func.ms_deform_attn_core_pytorch

    class _MSDAFunctionViaCore:
        """Routes the CUDA autograd Function to the reference's own pure-PyTorch core
        (ops/modules/ms_deform_attn.py:118-119 suggests exactly this swap)."""

        @staticmethod
        def apply(value, shapes, lsi, loc, w, step):
            return func.ms_deform_attn_core_pytorch(value, shapes.tolist(), loc, w)

    fpk.MSDeformAttnFunction = _MSDAFunctionViaCore
    msda_mod = imp("mask2former.modeling.pixel_decoder.ops.modules.ms_deform_attn")
    msda_mod.MSDeformAttnFunction = _MSDAFunctionViaCore
    sys.modules["mask2former.modeling.pixel_decoder.ops.modules"].MSDeformAttn = msda_mod.MSDeformAttn

    swin = imp("mask2former.modeling.backbone.swin")
    pe2d = imp("mask2former.modeling.transformer_decoder.position_encoding")
    pix = imp("mask2former.modeling.pixel_decoder.msdeformattn")
    pe3d = imp("univs.modeling.transformer_decoder.position_encoding")
    tl = imp("univs.modeling.transformer_decoder.transformer_layers")
    penc = imp("univs.modeling.prompt_encoder.prompt_encoder")
    pe_pkg.VisualPromptEncoder = penc.VisualPromptEncoder
    pe_pkg.VisualPromptSampler = penc.VisualPromptSampler
    pe_pkg.TextPromptEncoder = penc.TextPromptEncoder
    dec = imp("univs.modeling.transformer_decoder.video_mask2former_transformer_decoder_univs")
    cat = imp("datasets.concept_emb.combined_datasets_category_info")

    ns = types.SimpleNamespace(
        swin=swin, pix=pix, msda_mod=msda_mod, msda_func=func, pe2d=pe2d, pe3d=pe3d, layers=tl,
        prompt_encoder=penc, decoder=dec, category_info=cat.combined_datasets_category_info,
        ShapeSpec=_ShapeSpec,
        SwinTransformer=swin.SwinTransformer,
        WindowAttention=swin.WindowAttention,
        MSDeformAttnPixelDecoder=pix.MSDeformAttnPixelDecoder,
        MSDeformAttn=msda_mod.MSDeformAttn,
        ms_deform_attn_core_pytorch=func.ms_deform_attn_core_pytorch,
        VisualPromptSampler=penc.VisualPromptSampler,
        Decoder=dec.VideoMultiScaleMaskedTransformerDecoderUniVS,
        CrossAttentionLayer=tl.CrossAttentionLayer,
        SelfAttentionLayer=tl.SelfAttentionLayer,
    )
    _LOADED = ns
    return ns


# --------------------------------------------------------------------------
# Builders for the reference modules with explicit kwargs (bypassing from_config)
# --------------------------------------------------------------------------
SWIN_VARIANTS = {
    # configs/univs/univs_swin{t,b,l}_stage1.yaml:5-9
    "tiny": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7),
    "base": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=12),
    "large": dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48], window_size=12),
}


def build_reference_model(swin_kwargs, *, num_queries, num_frames, clip_emb, conv_dim=256,
                          enc_layers=6, dec_layers=9, dim_feedforward=2048, num_dense_points=32,
                          num_prev_frames_memory=5, text_prompt_to_image_enable=False,
                          self_attn_mask_type="sep", tmp_dir="/tmp"):
    """Returns (backbone, pixel_decoder, predictor) built from the reference classes."""
    ref = load()
    bb = ref.SwinTransformer(drop_path_rate=0.3, **swin_kwargs)
    bb.eval()  # NB: returns None (swin.py:680-683)
    E = swin_kwargs["embed_dim"]
    shapes = {f"res{i + 2}": ref.ShapeSpec(channels=E * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    pix = ref.MSDeformAttnPixelDecoder(
        shapes, transformer_dropout=0.0, transformer_nheads=conv_dim // 32,
        transformer_dim_feedforward=1024, transformer_enc_layers=enc_layers, conv_dim=conv_dim,
        mask_dim=conv_dim, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4)
    pix.eval()
    sampler = ref.VisualPromptSampler(
        pretrain_img_size=1024, hidden_dim=conv_dim, num_heads=conv_dim // 32, num_frames=num_frames,
        num_prev_frames_memory=num_prev_frames_memory, num_dense_points=num_dense_points,
        position_embedding_sin3d_type="ArbitraryT", clip_stride=1)
    path = os.path.join(tmp_dir, f"_clip_emb_{os.getpid()}.pth")
    torch.save(clip_emb, path)
    dec = ref.Decoder(
        conv_dim, True, num_classes=133, hidden_dim=conv_dim, num_queries=num_queries,
        nheads=conv_dim // 32, dim_feedforward=dim_feedforward, dec_layers=dec_layers, pre_norm=False,
        mask_dim=conv_dim, enforce_input_project=False, num_frames=num_frames,
        clip_class_embed_path=path, visual_prompt_sampler=sampler, num_dense_points=num_dense_points,
        text_prompt_enable=True, prompt_as_queries=True,
        text_prompt_to_image_enable=text_prompt_to_image_enable,
        maskdec_self_attn_mask_type=self_attn_mask_type, position_embedding_sin3d_type="ArbitraryT",
        num_prev_frames_memory=num_prev_frames_memory)
    os.remove(path)
    dec.eval()
    return bb, pix, dec


@torch.no_grad()
def reference_clip_forward(bb, pix, dec, frames, targets):
    """MaskFormerHead.layers glue (mask_former_head.py:148-154), restated."""
    feats = bb(frames)
    mf, mf_bfe, _enc0, ms = pix.forward_features(feats)
    out = dec(ms, mf, mf_bfe, None, targets)
    return feats, (mf, ms), out


# --------------------------------------------------------------------------
# Sliding-window task heads (univs/inference/*), loaded by path for the head parity tests
# --------------------------------------------------------------------------
_HEADS = None


class _RefModel:
    """What a reference head touches on `model`: .backbone(x) and .sem_seg_head(features, targets=...)
    (MaskFormerHead.layers glue, mask_former_head.py:148-154)."""

    def __init__(self, bb, pix, dec):
        self.backbone, self._pix, self._dec = bb, pix, dec

    @torch.no_grad()
    def sem_seg_head(self, features, targets=None):
        mf, mf_bfe, _enc0, ms = self._pix.forward_features(features)
        return self._dec(ms, mf, mf_bfe, None, targets)


class _ImageList:
    def __init__(self, tensor, image_sizes):
        self.tensor, self.image_sizes = tensor, image_sizes


class _TensorBox:
    """detectron2 Boxes / BitMasks: a tensor behind `.tensor`."""

    def __init__(self, tensor):
        self.tensor = tensor

    def to(self, device):
        return _TensorBox(self.tensor.to(device))


class _Instances:
    """detectron2 Instances as the VOS head reads it (inference_video_vos.py:587-606): image_size, free fields,
    len() = number of objects, .to(device)."""

    def __init__(self, image_size, **fields):
        self.image_size = image_size
        self._fields = dict(fields)

    def __getattr__(self, k):
        try:
            return self.__dict__["_fields"][k]
        except KeyError:
            raise AttributeError(k)

    def __len__(self):
        return len(self._fields.get("ori_ids", []))

    def to(self, device):
        return _Instances(self.image_size, **{k: (v.to(device) if hasattr(v, "to") else v)
                                              for k, v in self._fields.items()})


def load_inference_heads():
    """Imports the reference's inference heads with their heavy imports (kornia, pycocotools, detectron2 structures,
    the training-side `univs` exports) stubbed; only the inference code paths are exercised."""
    global _HEADS
    if _HEADS is not None:
        return _HEADS
    load()
    R = REF_ROOT
    _mod("kornia", color=types.SimpleNamespace())
    def _encode(arr):
        """pycocotools.mask.encode for an [h, w, n] uint8 array (restated codec, oracle/rle_ref.py); counts as bytes"""
        from . import rle_ref
        out = []
        for k in range(arr.shape[2]):
            r = rle_ref.encode(arr[:, :, k])
            out.append({"size": r["size"], "counts": r["counts"].encode("ascii")})
        return out
    mu = _mod("pycocotools.mask", encode=_encode)
    _mod("pycocotools", mask=mu)
    _mod("detectron2.data", MetadataCatalog=types.SimpleNamespace(get=lambda name: types.SimpleNamespace()))
    def _sem_seg_postprocess(result, img_size, output_height, output_width):
        """detectron2/modeling/postprocessing.py: crop to the unpadded size, bilinear resize to the output size"""
        result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
        return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]
    _mod("detectron2.modeling.postprocessing", sem_seg_postprocess=_sem_seg_postprocess)
    _mod("detectron2.structures", Boxes=_TensorBox, ImageList=_ImageList, Instances=_Instances, BitMasks=_TensorBox)
    _mod("detectron2.utils.memory", retry_if_cuda_oom=lambda f: f)
    _pkg("mask2former.utils", f"{R}/mask2former/utils")
    univs = sys.modules["univs"]
    for name in ("VideoSetCriterionUni", "VideoHungarianMatcherUni", "BoxVISTeacherSetPseudoMask",
                 "TextPromptEncoder", "build_clip_language_encoder", "Clips", "FastOverTracker_DET",
                 "MDQE_OverTrackerEfficient"):
        setattr(univs, name, object)
    _mod("univs.prepare_targets", PrepareTargets=object)
    _pkg("univs.inference", f"{R}/univs/inference")
    imp = importlib.import_module
    comm = imp("univs.inference.comm")
    ucomm = imp("univs.utils.comm")
    vis_fast = imp("univs.inference.inference_video_vis_fast")
    _mod("matplotlib"); _mod("matplotlib.pyplot")
    _mod("univs.inference.visualization", visualization_query_embds=lambda **kw: None, display_instance_masks=None)
    _mod("univs.utils.visualizer", VisualizerFrame=object)
    vos = imp("univs.inference.inference_video_vos")
    vps = imp("univs.inference.inference_video_vps")
    _mod("univs.data", datasets=None)
    _mod("univs.data.datasets", _get_vspw_vss_metadata=None, _get_vipseg_panoptic_metadata_val=None)
    entity = imp("univs.inference.inference_video_entity")
    image = imp("univs.inference.inference_image_generic_seg")
    semx = imp("univs.inference.inference_video_semantic_extraction")
    _HEADS = types.SimpleNamespace(comm=comm, utils_comm=ucomm, vis_fast=vis_fast, vos=vos, vps=vps,
                                   entity=entity, InferenceVideoEntity=entity.InferenceVideoEntity,
                                   image=image, InferenceImageGenericSeg=image.InferenceImageGenericSegmentation,
                                   semx=semx,
                                   InferenceVideoVPS=vps.InferenceVideoVPS,
                                   InferenceVideoVISFast=vis_fast.InferenceVideoVISFast,
                                   InferenceVideoVOS=vos.InferenceVideoVOS,
                                   RefModel=_RefModel, ImageList=_ImageList, Instances=_Instances,
                                   BitMasks=_TensorBox, Boxes=_TensorBox)
    return _HEADS
