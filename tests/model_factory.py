"""TEST INFRASTRUCTURE: deterministic weights + tiny configs shared by the parity tests and the golden generator.

Weights are generated per state_dict key from a seed derived from the key name (crc32), so the reference modules
(built here, where /root/reference exists) and the product modules (built anywhere) receive bit-identical parameters
without shipping checkpoints."""
import zlib

import torch

TINY_SWIN = dict(embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], window_size=4)
SMALL_SWIN7 = dict(embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], window_size=7)


def keyed_state_dict(template_sd, seed=0):
    """template_sd: {key: tensor} giving shapes/dtypes.  Returns {key: tensor} with deterministic values."""
    out = {}
    for k, v in template_sd.items():
        if not v.is_floating_point():
            out[k] = v.clone()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) % (2 ** 31))
        base = torch.randn(v.shape, generator=g, dtype=torch.float32)
        name = k.rsplit(".", 1)[-1]
        if "norm" in k and name == "weight":
            t = 1.0 + 0.1 * base
        elif name == "bias" or "bias" in name:
            t = 0.1 * base if "sampling_offsets" not in k else base * 0.5
        elif "relative_position_bias_table" in k:
            t = 0.3 * base
        elif "sampling_offsets" in k:
            t = 0.05 * base
        elif "attention_weights" in k:
            t = 0.1 * base
        elif k.endswith("cls_temp.weight") or k.endswith("reid_temp.weight"):
            t = torch.full(v.shape, 2.659)
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            t = base * (1.0 / fan_in) ** 0.5
        else:
            t = 0.5 * base
        out[k] = t.to(v.dtype)
    return out


def make_clip_emb(seed=0):
    g = torch.Generator().manual_seed(1234 + seed)
    return torch.randn(3938, 640, generator=g)


def build_product_model(swin_kwargs, *, num_queries, num_frames, clip_emb, enc_layers=6, dec_layers=9,
                        dim_feedforward=2048, num_dense_points=32, num_prev_frames_memory=5,
                        text_prompt_to_image_enable=False, self_attn_mask_type="sep"):
    from univs_b200.modeling import MSDeformAttnPixelDecoder, SwinTransformer, VideoMultiScaleMaskedTransformerDecoderUniVS
    from univs_b200.modeling.prompt_sampler import VisualPromptSampler
    from univs_b200.registry import ShapeSpec
    bb = SwinTransformer(drop_path_rate=0.3, **swin_kwargs)
    E = swin_kwargs["embed_dim"]
    shapes = {f"res{i + 2}": ShapeSpec(channels=E * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    pix = MSDeformAttnPixelDecoder(
        shapes, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
        transformer_enc_layers=enc_layers, conv_dim=256, mask_dim=256, norm="GN",
        transformer_in_features=["res3", "res4", "res5"], common_stride=4)
    sampler = VisualPromptSampler(pretrain_img_size=1024, hidden_dim=256, num_heads=8, num_frames=num_frames,
                                  num_prev_frames_memory=num_prev_frames_memory, num_dense_points=num_dense_points,
                                  position_embedding_sin3d_type="ArbitraryT", clip_stride=1)
    dec = VideoMultiScaleMaskedTransformerDecoderUniVS(
        256, True, num_classes=133, hidden_dim=256, num_queries=num_queries, nheads=8,
        dim_feedforward=dim_feedforward, dec_layers=dec_layers, pre_norm=False, mask_dim=256,
        enforce_input_project=False, num_frames=num_frames, clip_class_embed_path=clip_emb,
        visual_prompt_sampler=sampler, num_dense_points=num_dense_points, text_prompt_enable=True,
        prompt_as_queries=True, text_prompt_to_image_enable=text_prompt_to_image_enable,
        maskdec_self_attn_mask_type=self_attn_mask_type, position_embedding_sin3d_type="ArbitraryT",
        num_prev_frames_memory=num_prev_frames_memory)
    return bb, pix, dec


def load_keyed(modules, seed=0):
    for m in modules:
        sd = keyed_state_dict(m.state_dict(), seed)
        missing, unexpected = m.load_state_dict(sd, strict=True)
        assert not missing and not unexpected


@torch.no_grad()
def product_clip_forward(bb, pix, dec, frames, targets):
    feats = bb(frames)
    mf, mf_bfe, _enc0, ms = pix.forward_features(feats)
    out = dec(ms, mf, mf_bfe, None, targets)
    return feats, (mf, ms), out
