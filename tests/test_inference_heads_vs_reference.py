"""Sliding-window task heads against the reference heads executed by path: same tiny model,
same key-seeded weights, same synthetic video.  Product operators = CPU oracles; the product head runs every frame
through backbone + pixel decoder once (ClipStream), the reference re-runs the pixel decoder per clip."""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle.cpu_backend import oracle_ops
from tests import model_factory as mf
from univs_b200.inference import (InferenceVideoVISFast, TemporalMaskMean, calculate_mask_quality_scores,
                                  generate_temporal_weights, match_from_learnable_embds)
from univs_b200.meta_arch import UniVS_Prompt
from univs_b200.modeling.head import MaskFormerHead
from univs_b200.registry import ShapeSpec

pytestmark = pytest.mark.reference
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _pair(T, Q, **kw):
    clip = mf.make_clip_emb()
    ref = ref_shim.build_reference_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=clip, **kw)
    prod = mf.build_product_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=clip, **kw)
    for r, p in zip(ref, prod):
        sd = mf.keyed_state_dict(r.state_dict())
        r.load_state_dict(sd)
        p.load_state_dict(sd)
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=prod[1], transformer_predictor=prod[2])
    model = UniVS_Prompt(backbone=prod[0], sem_seg_head=head, pixel_mean=MEAN, pixel_std=STD)
    return ref, model


def _ref_head(heads, T, Q, **over):
    kw = dict(hidden_dim=256, num_queries=Q, object_mask_threshold=0.05, overlap_threshold=0.8,
              stability_score_thresh=0.0, metadata=None, size_divisibility=32, LSJ_aug_image_size=1024,
              LSJ_aug_enable_test=False, sem_seg_postprocess_before_inference=False, pixel_mean=MEAN, pixel_std=STD,
              num_frames=T, num_classes=133, data_name="ytvis_2021_val", prompt_as_queries=True,
              zero_shot_inference=False, semantic_on=False, instance_on=True, panoptic_on=False,
              test_topk_per_image=10, tracker_type="minvis", mdqe_tracker=None, window_inference=False,
              is_multi_cls=True, apply_cls_thres=0.05, merge_on_cpu=False, num_max_inst_test=50,
              num_frames_window_test=T, clip_stride=1)
    kw.update(over)
    return heads.InferenceVideoVISFast(**kw)


@pytest.mark.parametrize("reuse,stability", [(True, 0.0), (False, 0.0), (True, 0.5)])
def test_vis_fast_minvis_head(reuse, stability):
    heads = ref_shim.load_inference_heads()
    T, Q, V, H, W = 2, 12, 5, 60, 90
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=3)
    g = torch.Generator().manual_seed(5)
    frames = [(torch.rand(3, H, W, generator=g) * 255).round() for _ in range(V)]
    inputs = [{"image": frames, "height": 75, "width": 120, "dataset_name": "ytvis21", "task": "detection",
               "video_len": V, "file_names": [f"{i}.jpg" for i in range(V)]}]

    # reference: normalise + pad exactly as eval() does (:198-207), then its own clip loop / tracker / post-processing
    rhead = _ref_head(heads, T, Q, stability_score_thresh=stability)
    xs = torch.stack([(f - rhead.pixel_mean) / rhead.pixel_std for f in frames])
    xs = torch.nn.functional.pad(xs, (0, 96 - W, 0, 64 - H))
    rimages = heads.ImageList(xs, [(H, W)] * V)
    tg = [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}]
    with torch.no_grad():
        want = rhead.inference_video_vis_minvis(heads.RefModel(*ref), inputs, rimages, tg)

    phead = InferenceVideoVISFast(num_queries=Q, num_frames=T, stability_score_thresh=stability, test_topk_per_image=10,
                                  num_frames_window_test=T, reuse_features=reuse)
    with oracle_ops():
        got = phead.eval(model, inputs)

    assert got["image_size"] == want["image_size"] == (75, 120)
    # (query, class) detections: same set, same scores
    order_w = np.lexsort((want["pred_labels"], want["pred_scores"]))
    order_g = np.lexsort((got["pred_labels"], got["pred_scores"]))
    assert len(order_w) == len(order_g) >= 5
    assert [want["pred_labels"][i] for i in order_w] == [got["pred_labels"][i] for i in order_g]
    np.testing.assert_allclose([got["pred_scores"][i] for i in order_g], [want["pred_scores"][i] for i in order_w],
                               rtol=2e-4, atol=1e-6)
    flips = total = 0
    for iw, ig in zip(order_w, order_g):
        mw, mg = want["pred_masks"][iw], got["pred_masks"][ig]
        assert mw.shape == mg.shape == (V, 75, 120) and mg.dtype == torch.bool
        flips += (mw != mg).sum().item()
        total += mw.numel()
    assert flips <= 1e-4 * total, (flips, total)


def test_tracking_helpers_match_reference():
    heads = ref_shim.load_inference_heads()
    g = torch.Generator().manual_seed(0)
    for V in (1, 2, 3):
        tgt = torch.randn(7, V, 16, generator=g)
        tgt[2, 0] = 0            # blank memory slot
        cur = torch.randn(9, 4, 16, generator=g)
        for use_norm in (True, False):
            want, ws = heads.comm.match_from_learnable_embds(tgt, cur, return_similarity=True, use_norm=use_norm)
            got, gs = match_from_learnable_embds(tgt, cur, return_similarity=True, use_norm=use_norm)
            assert list(want) == list(got)
            torch.testing.assert_close(gs, ws, rtol=1e-5, atol=1e-6)
        w = torch.rand(5, V, generator=g)
        for sm in (False, True):
            torch.testing.assert_close(generate_temporal_weights(V, w, enable_softmax=sm),
                                       heads.comm.generate_temporal_weights(V, w, enable_softmax=sm))
    m = torch.randn(6, 3, 8, 8, generator=g) * 2
    torch.testing.assert_close(calculate_mask_quality_scores(m), heads.utils_comm.calculate_mask_quality_scores(m))


def test_temporal_mask_mean_equals_list_average():
    g = torch.Generator().manual_seed(1)
    T, V, Q = 3, 6, 4
    clips = [torch.randn(Q, T, 5, 7, generator=g) for _ in range(V - T + 1)]
    acc = TemporalMaskMean(Q, V, (5, 7), "cpu")
    for s, c in enumerate(clips):
        acc.add(s, c)
    got = acc.mean()
    for v in range(V):       # reference rule: frame v = mean of clips[v - t][:, t] over valid (clip, t)
        parts = [clips[v - t][:, t] for t in range(T) if 0 <= v - t < len(clips)]
        torch.testing.assert_close(got[:, v], torch.stack(parts).mean(0), rtol=1e-6, atol=1e-6)
