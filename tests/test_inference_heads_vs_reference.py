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
    from univs_b200.inference import check_consistency_with_prev_frames, pair_mask_iou, video_box_iou
    prev, cur = torch.randn(6, 4, 16, generator=g), torch.randn(6, 3, 16, generator=g)
    prev[1, :2] = 0
    for use_norm in (True, False):
        wk, ws = heads.comm.check_consistency_with_prev_frames(prev, cur, 0.1, True, use_norm)
        gk, gs = check_consistency_with_prev_frames(prev, cur, 0.1, True, use_norm)
        assert torch.equal(wk, gk)
        torch.testing.assert_close(gs, ws, rtol=1e-5, atol=1e-6)
    b1, b2 = torch.rand(3, 2, 4, generator=g), torch.rand(5, 2, 4, generator=g)
    b1[..., 2:] += b1[..., :2]
    b2[..., 2:] += b2[..., :2]
    torch.testing.assert_close(video_box_iou(b1, b2), heads.utils_comm.video_box_iou(b1, b2)[0])
    m1, m2 = (torch.rand(4, 2, 9, 9, generator=g) > 0.5).float(), torch.rand(4, 2, 9, 9, generator=g) > 0.5
    torch.testing.assert_close(pair_mask_iou(m1, m2), heads.utils_comm.batched_pair_mask_iou(m1, m2))


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


# ---------------------------------------------------------------------------------------------------------------
# VOS head: task "sot" (mask prompts + memory) and "grounding" (text prompts)
# ---------------------------------------------------------------------------------------------------------------
def _ref_vos_head(heads, T, Q, out_dir, **over):
    kw = dict(hidden_dim=256, num_queries=Q, object_mask_threshold=0.05, overlap_threshold=0.8,
              overlap_threshold_entity=0.5, stability_score_thresh=0.0, metadata=None, size_divisibility=32,
              LSJ_aug_image_size=1024, LSJ_aug_enable_test=False, sem_seg_postprocess_before_inference=False,
              pixel_mean=MEAN, pixel_std=STD, num_frames=T, num_classes=133, data_name="davis_val",
              prompt_as_queries=True, zero_shot_inference=False, semantic_on=False, instance_on=True,
              panoptic_on=False, test_topk_per_image=10, tracker_type="minvis", window_inference=False,
              num_frames_window_test=T, clip_stride=1, output_dir=out_dir, video_unified_inference_queries="prompt",
              num_prev_frames_memory=4)
    kw.update(over)
    return heads.InferenceVideoVOS(**kw)


def _rects(ids, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    masks = torch.zeros(len(ids), H, W)
    boxes = torch.zeros(len(ids), 4)
    for j in range(len(ids)):
        bw, bh = int(W * (0.25 + 0.3 * torch.rand(1, generator=g))), int(H * (0.25 + 0.3 * torch.rand(1, generator=g)))
        x0, y0 = int((W - bw) * torch.rand(1, generator=g)), int((H - bh) * torch.rand(1, generator=g))
        masks[j, y0:y0 + bh, x0:x0 + bw] = 1.0
        boxes[j] = torch.tensor([x0, y0, x0 + bw, y0 + bh], dtype=torch.float32)
    return masks, boxes


def _read_png(path):
    from PIL import Image
    return torch.from_numpy(np.array(Image.open(path)))


@pytest.mark.parametrize("mode,dataset,reuse", [("prompt", "davis", True), ("prompt", "davis", False),
                                                ("prompt+learn", "davis", True), ("learn", "davis", True),
                                                ("prompt", "viposeg", True), ("prompt+learn", "viposeg", True)])
def test_vos_head_sot(tmp_path, mode, dataset, reuse):
    import types
    stride = 1           # the only stride the reference's prompt sampler supports for sot (prompt_encoder.py:944)
    # panoptic VOS: classes 10 and 30 are "stuff" (dataset ids are label + 1)
    metadata = types.SimpleNamespace(stuff_dataset_id_to_contiguous_id={11: 0, 31: 1}) if dataset == "viposeg" else None
    classes = {4: 10, 7: 3, 5: 30}
    from univs_b200.inference import FrameAnnotations, InferenceVideoVOS
    heads = ref_shim.load_inference_heads()
    T, Q, V, H, W = 3, 8, 6, 60, 90
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=3, num_dense_points=8, num_prev_frames_memory=4)
    g = torch.Generator().manual_seed(9)
    frames = [(torch.rand(3, H, W, generator=g) * 255).round() for _ in range(V)]
    names = [f"davis/vid0/{i:05d}.jpg" for i in range(V)]
    # objects 4 and 7 are given in frame 0, object 5 enters (and is given) in frame 2
    given = {0: [4, 7], 2: [5]}
    palette = [(i * 37) % 256 for i in range(768)]

    def annotations(make):
        out = []
        for f in range(V):
            ids = given.get(f, [])
            m, b = _rects(ids, H, W, 50 + f)
            out.append(make(f, ids, m, b))
        return out

    # ---- reference
    rhead = _ref_vos_head(heads, T, Q, str(tmp_path / "ref"), video_unified_inference_queries=mode, clip_stride=stride,
                          metadata=metadata)
    xs = torch.stack([(f - rhead.pixel_mean) / rhead.pixel_std for f in frames])
    xs = torch.nn.functional.pad(xs, (0, 96 - W, 0, 64 - H))
    rinst = annotations(lambda f, ids, m, b: heads.Instances(
        (H, W), ori_ids=ids, gt_masks=heads.BitMasks(m), gt_boxes=heads.Boxes(b),
        gt_classes=torch.tensor([classes[i] for i in ids], dtype=torch.long)))
    rtg = [{"task": "sot", "dataset_name": dataset, "prompt_type": "visual", "video_len": V, "num_frames": T,
            "inter_image_size": (64, 96), "image_size": (H, W), "file_names": names, "instances": rinst,
            "mask_palette": palette}]
    torch.manual_seed(21)
    with torch.no_grad():
        rhead.inference_video_vos(heads.RefModel(*ref), None, heads.ImageList(xs, [(H, W)] * V), rtg, (H, W), (H, W))

    # ---- product
    phead = InferenceVideoVOS(num_queries=Q, num_frames=T, num_frames_window_test=T, clip_stride=stride,
                              video_unified_inference_queries=mode, num_prev_frames_memory=4, reuse_features=reuse,
                              output_dir=str(tmp_path / "prod"), metadata=metadata)
    pinst = annotations(lambda f, ids, m, b: FrameAnnotations(
        (H, W), ids, m, b, torch.tensor([classes[i] for i in ids], dtype=torch.long)))
    inputs = [{"image": frames, "task": "sot", "dataset_name": dataset, "file_names": names, "instances": pinst,
               "mask_palette": palette}]
    torch.manual_seed(21)
    with oracle_ops():
        got = phead.eval(model, inputs)

    # written id maps: identical frames, (almost) identical pixels; in-memory results == written files
    rdir, pdir = tmp_path / "ref/inference/Annotations/vid0", tmp_path / "prod/inference/Annotations/vid0"
    rfiles = sorted(p.name for p in rdir.iterdir())
    assert rfiles == sorted(p.name for p in pdir.iterdir()) and len(rfiles) >= V - 1
    flips = total = labelled = 0
    for name in rfiles:
        want, mine = _read_png(rdir / name), _read_png(pdir / name)
        assert torch.equal(mine, got["frames"][int(name[:5])])
        assert want.shape == mine.shape == (H, W)
        flips += (want != mine).sum().item()
        total += want.numel()
        labelled += (want > 0).sum().item()
    assert flips <= 2e-4 * total, (flips, total)
    assert labelled > 0                       # the comparison is not vacuous: objects are being propagated
    # the annotation state that prompts the next clip
    ptg = phead._last_targets[0]
    assert ptg["ids"] == rtg[0]["ids"] == [4, 5, 7]
    assert torch.equal(ptg["first_appear_frame_idxs"], rtg[0]["first_appear_frame_idxs"])
    for k in ("mask_logits", "boxes", "embds"):
        assert ptg[k].shape == rtg[0][k].shape, k
        scale = rtg[0][k].abs().max().item()
        assert (ptg[k] - rtg[0][k]).abs().max().item() <= 1e-3 * max(scale, 1e-6), k
    assert (ptg["masks"] != rtg[0]["masks"]).float().mean().item() < 2e-4


@pytest.mark.parametrize("mode", ["prompt", "prompt+learn", "learn"])
def test_vos_head_grounding(tmp_path, mode):
    from univs_b200.inference import InferenceVideoVOS
    heads = ref_shim.load_inference_heads()
    T, Q, V, H, W, P = 2, 8, 4, 60, 90, 3
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=3, text_prompt_to_image_enable=True,
                       self_attn_mask_type="sep-blocked")
    g = torch.Generator().manual_seed(13)
    frames = [(torch.rand(3, H, W, generator=g) * 255).round() for _ in range(V)]
    names = [f"refytvos/vid7/{i:05d}.jpg" for i in range(V)]
    text = {"exp_obj_ids": [0, 1, 2], "exp_word_feats": torch.randn(P, 77, T, 640, generator=g),
            "exp_sentence_feats": torch.randn(P, T, 640, generator=g), "exp_word_len": [10] * P,
            "expressions": ["a", "b", "c"]}
    clone = lambda d: {k: (v.clone() if torch.is_tensor(v) else list(v)) for k, v in d.items()}

    rhead = _ref_vos_head(heads, T, Q, str(tmp_path / "ref"), video_unified_inference_queries=mode,
                          data_name="rvos-refytb-val")
    xs = torch.stack([(f - rhead.pixel_mean) / rhead.pixel_std for f in frames])
    xs = torch.nn.functional.pad(xs, (0, 96 - W, 0, 64 - H))
    rtg = [{"task": "grounding", "dataset_name": "refytvos", "prompt_type": "text", "video_len": V, "num_frames": T,
            "inter_image_size": (64, 96), "image_size": (H, W), "file_names": names, **clone(text)}]
    rtg[0]["prompt_obj_ids"] = rtg[0]["exp_obj_ids"]
    with torch.no_grad():
        rhead.inference_video_vos(heads.RefModel(*ref), None, heads.ImageList(xs, [(H, W)] * V), rtg, (H, W), (75, 120))

    phead = InferenceVideoVOS(num_queries=Q, num_frames=T, num_frames_window_test=T, num_prev_frames_memory=4,
                              video_unified_inference_queries=mode, output_dir=str(tmp_path / "prod"))
    inputs = [{"image": frames, "task": "grounding", "dataset_name": "refytvos", "file_names": names,
               "height": 75, "width": 120, **clone(text)}]
    with oracle_ops():
        got = phead.eval(model, inputs)

    flips = total = 0
    for obj in (0, 1, 2):
        rdir, pdir = tmp_path / f"ref/inference/Annotations/vid7/{obj}", tmp_path / f"prod/inference/Annotations/vid7/{obj}"
        rfiles = sorted(p.name for p in rdir.iterdir())
        assert rfiles == sorted(p.name for p in pdir.iterdir()) and len(rfiles) >= V - 1
        for name in rfiles:
            want, mine = _read_png(rdir / name), _read_png(pdir / name)
            assert want.shape == mine.shape == (75, 120)
            assert torch.equal(mine > 0, got["objects"][obj][int(name[:5])])
            flips += (want != mine).sum().item()
            total += want.numel()
    assert flips <= 2e-4 * total, (flips, total)
    ptg = phead._last_targets[0]
    for k in ("mask_logits", "boxes", "embds"):
        scale = rtg[0][k].abs().max().item()
        assert (ptg[k] - rtg[0][k]).abs().max().item() <= 1e-3 * max(scale, 1e-6), k


def test_vos_head_refdavis_id_maps(tmp_path):
    """Referring VOS on a DAVIS-style dataset is written as palette id maps with zero-based expression ids shifted by
    one (inference_video_vos.py:276-277, :630-632)."""
    from univs_b200.inference import InferenceVideoVOS
    heads = ref_shim.load_inference_heads()
    T, Q, V, H, W, P = 2, 8, 3, 64, 96, 2
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=2, text_prompt_to_image_enable=True)
    g = torch.Generator().manual_seed(2)
    frames = [(torch.rand(3, H, W, generator=g) * 255).round() for _ in range(V)]
    names = [f"refdavis/dog/{i:05d}.jpg" for i in range(V)]
    palette = list(range(256)) * 3
    text = {"exp_obj_ids": [0, 1], "exp_word_feats": torch.randn(P, 77, T, 640, generator=g),
            "exp_sentence_feats": torch.randn(P, T, 640, generator=g), "exp_word_len": [7] * P}
    clone = lambda d: {k: (v.clone() if torch.is_tensor(v) else list(v)) for k, v in d.items()}
    rhead = _ref_vos_head(heads, T, Q, str(tmp_path / "ref"))
    xs = torch.stack([(f - rhead.pixel_mean) / rhead.pixel_std for f in frames])
    rtg = [{"task": "grounding", "dataset_name": "refdavis", "prompt_type": "text", "video_len": V, "num_frames": T,
            "inter_image_size": (H, W), "image_size": (H, W), "file_names": names, "mask_palette": palette,
            **clone(text)}]
    with torch.no_grad():
        rhead.inference_video_vos(heads.RefModel(*ref), None, heads.ImageList(xs, [(H, W)] * V), rtg, (H, W), (H, W))
    phead = InferenceVideoVOS(num_queries=Q, num_frames=T, num_frames_window_test=T, num_prev_frames_memory=4,
                              output_dir=str(tmp_path / "prod"))
    with oracle_ops():
        got = phead.eval(model, [{"image": frames, "task": "grounding", "dataset_name": "refdavis",
                                  "file_names": names, "mask_palette": palette, **clone(text)}])
    rdir, pdir = tmp_path / "ref/inference/Annotations/dog", tmp_path / "prod/inference/Annotations/dog"
    assert sorted(p.name for p in rdir.iterdir()) == sorted(p.name for p in pdir.iterdir())
    flips = total = 0
    for p in rdir.iterdir():
        want, mine = _read_png(p), _read_png(pdir / p.name)
        assert set(mine.unique().tolist()) <= {0, 1, 2}
        flips += (want != mine).sum().item()
        total += want.numel()
    assert flips <= 2e-4 * total and sorted(got["frames"]) == list(range(V))


def test_heads_build_from_cfg_and_pad_square():
    """Constructors read the reference's config keys (from_config, inference_video_vis_fast.py:141-179,
    inference_video_vos.py:158-201); LSJ square padding follows ImageList.from_tensors(square_size=...)."""
    from univs_b200.config import get_cfg
    from univs_b200.inference import InferenceVideoVOS
    cfg = get_cfg()
    cfg.merge_from_list(["MODEL.BoxVIS.TEST.NUM_FRAMES_WINDOW", 7, "INPUT.SAMPLING_FRAME_NUM", 3,
                         "MODEL.MASK_FORMER.NUM_OBJECT_QUERIES", 50, "TEST.DETECTIONS_PER_IMAGE", 35,
                         "INPUT.LSJ_AUG.SQUARE_ENABLED", True, "INPUT.LSJ_AUG.IMAGE_SIZE", 128])
    vis, vos = InferenceVideoVISFast(cfg), InferenceVideoVOS(cfg)
    assert (vis.num_queries, vis.num_frames, vis.num_frames_window_test, vis.test_topk_per_image) == (50, 3, 7, 35)
    assert (vos.num_queries, vos.num_frames, vos.num_prev_frames_memory, vos.clip_stride) == (50, 3, 5, 1)
    assert vos.video_unified_inference_queries == "prompt" and vis.LSJ_aug_enable_test and vos.LSJ_aug_image_size == 128

    T, Q, V = 2, 6, 3
    _ref, model = _pair(T, Q, enc_layers=1, dec_layers=2)
    head = InferenceVideoVISFast(num_queries=Q, num_frames=T, num_frames_window_test=T, lsj_aug_enable_test=True,
                                 lsj_aug_image_size=128, test_topk_per_image=5)
    g = torch.Generator().manual_seed(0)
    frames = [(torch.rand(3, 60, 90, generator=g) * 255).round() for _ in range(V)]
    with oracle_ops():
        out = head.eval(model, [{"image": frames, "dataset_name": "ovis", "height": 60, "width": 90}])
    assert out["image_size"] == (60, 90) and out["pred_masks"][0].shape == (V, 60, 90)
    assert model.sem_seg_head.predictor is not None and len(out["pred_scores"]) == len(out["pred_labels"]) >= 5


def test_vps_online_head():
    """InferenceVideoVPS (inference_video_vps.py): clip loop + tracker + panoptic assembly against the reference head."""
    import types
    from univs_b200.inference import InferenceVideoVPS
    heads = ref_shim.load_inference_heads()
    T, Q, V, H, W = 2, 12, 4, 60, 90
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=3)
    g = torch.Generator().manual_seed(11)
    frames = [(torch.rand(3, H, W, generator=g) * 255).round() for _ in range(V)]
    inputs = [{"image": frames, "height": 75, "width": 120, "dataset_name": "vipseg", "task": "detection",
               "video_len": V, "file_names": [f"{i}.jpg" for i in range(V)]}]
    things = {c for c in range(1, 125) if c % 3 == 0}            # synthetic thing / stuff split of the 124 VIPSeg classes
    meta = types.SimpleNamespace(thing_dataset_id_to_contiguous_id={c: i for i, c in enumerate(sorted(things))})
    kw = dict(hidden_dim=256, num_queries=Q, object_mask_threshold=0.0, overlap_threshold=0.3, overlap_threshold_entity=0.5,
              stability_score_thresh=0.0, metadata=meta, size_divisibility=32, LSJ_aug_image_size=1024,
              LSJ_aug_enable_test=False, sem_seg_postprocess_before_inference=False, pixel_mean=MEAN, pixel_std=STD,
              num_frames=T, data_name="vipseg_val", prompt_as_queries=True, zero_shot_inference=False, semantic_on=False,
              instance_on=False, panoptic_on=True, test_topk_per_image=8, tracker_type="minvis", window_inference=False,
              is_multi_cls=True, apply_cls_thres=0.05, merge_on_cpu=False, num_max_inst_test=50, num_frames_window_test=T,
              clip_stride=1)
    import inspect
    accepted = set(inspect.signature(heads.InferenceVideoVPS.__init__).parameters)
    rhead = heads.InferenceVideoVPS(**{k: v for k, v in kw.items() if k in accepted})
    xs = torch.stack([(f - rhead.pixel_mean) / rhead.pixel_std for f in frames])
    xs = torch.nn.functional.pad(xs, (0, 96 - W, 0, 64 - H))
    rimages = heads.ImageList(xs, [(H, W)] * V)
    tg = [{"task": "detection", "dataset_name": "vipseg", "prompt_type": "visual"}]
    with torch.no_grad():
        want = rhead.inference_video_vps_online(heads.RefModel(*ref), inputs, rimages, tg)

    for reuse in (True, False):
        phead = InferenceVideoVPS(num_queries=Q, num_frames=T, object_mask_threshold=0.0, overlap_threshold=0.3,
                                  test_topk_per_image=8, num_frames_window_test=T, thing_ids=things, reuse_features=reuse)
        with oracle_ops():
            got = phead.eval(model, inputs)
        assert got["image_size"] == want["image_size"] == (720, 1152) and got["task"] == "vps"
        assert got["segments_infos"] == want["segments_infos"] and len(got["segments_infos"]) > 0
        assert [int(i) for i in got["pred_ids"]] == [int(i) for i in want["pred_ids"]]
        assert got["pred_masks"].shape == want["pred_masks"].shape == (V, 720, 1152)
        diff = (got["pred_masks"] != want["pred_masks"]).float().mean().item()
        assert diff <= 1e-4, diff


def _entity_pair(T, Q, V, H, W, seed, dataset, **over):
    """Reference entity head + its inputs, and the keyword arguments of the product head for the same setting."""
    import inspect
    import types
    heads = ref_shim.load_inference_heads()
    g = torch.Generator().manual_seed(seed)
    frames = [(torch.rand(3, H, W, generator=g) * 255).round() for _ in range(V)]
    inputs = [{"image": frames, "height": 75, "width": 120, "dataset_name": dataset, "task": "detection",
               "video_len": V, "video_id": 7, "file_names": [f"{i}.jpg" for i in range(V)]}]
    things = {c for c in range(1, 125) if c % 3 == 0}
    meta = types.SimpleNamespace(thing_dataset_id_to_contiguous_id={c: i for i, c in enumerate(sorted(things))})
    kw = dict(hidden_dim=256, num_queries=Q, overlap_threshold=0.3, overlap_threshold_entity=0.2,
              stability_score_thresh=0.0, metadata=meta, size_divisibility=32, LSJ_aug_image_size=1024,
              LSJ_aug_enable_test=False, sem_seg_postprocess_before_inference=False, pixel_mean=MEAN, pixel_std=STD,
              num_frames=T, num_classes=133, data_name="ytvis_2021_val", prompt_as_queries=True, zero_shot_inference=False,
              semantic_on=False, instance_on=True, panoptic_on=False, test_topk_per_image=8, tracker_type="minvis",
              window_inference=False, is_multi_cls=True, apply_cls_thres=0.02, merge_on_cpu=False, box_nms_thresh=0.75,
              num_max_inst_test=50, num_frames_window_test=T, clip_stride=1, output_dir="/tmp/univs_entity_test",
              num_prev_frames_memory=1, video_unified_inference_entities="", temporal_consistency_threshold=0.05,
              detect_newly_object_threshold=0.02, detect_newly_interval_frames=1, custom_videos_enable=False)
    kw.update(over)
    accepted = set(inspect.signature(heads.InferenceVideoEntity.__init__).parameters)
    rhead = heads.InferenceVideoEntity(**{k: v for k, v in kw.items() if k in accepted})
    xs = torch.stack([(f - rhead.pixel_mean) / rhead.pixel_std for f in frames])
    xs = torch.nn.functional.pad(xs, (0, 96 - W, 0, 64 - H))
    rimages = heads.ImageList(xs, [(H, W)] * V)
    pkw = dict(num_queries=Q, num_frames=T, overlap_threshold=kw["overlap_threshold"],
               overlap_threshold_entity=kw["overlap_threshold_entity"], stability_score_thresh=kw["stability_score_thresh"],
               test_topk_per_image=kw["test_topk_per_image"], apply_cls_thres=kw["apply_cls_thres"],
               box_nms_thresh=kw["box_nms_thresh"], num_frames_window_test=T, clip_stride=kw["clip_stride"],
               num_prev_frames_memory=kw["num_prev_frames_memory"],
               video_unified_inference_entities=kw["video_unified_inference_entities"],
               temporal_consistency_threshold=kw["temporal_consistency_threshold"],
               detect_newly_object_threshold=kw["detect_newly_object_threshold"],
               detect_newly_interval_frames=kw["detect_newly_interval_frames"], thing_ids=things)
    return heads, rhead, rimages, inputs, pkw


def _ref_entity_run(heads, rhead, ref, inputs, rimages, dataset, sub_task):
    import contextlib
    import io
    tg = [{"task": "detection", "dataset_name": dataset, "prompt_type": "visual", "video_len": inputs[0]["video_len"],
           "sub_task": sub_task}]
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):     # the reference prints pool shapes per clip
        rmodel = ref[0] if len(ref) == 1 else heads.RefModel(*ref)      # a scripted stand-in, or (backbone, pixdec, decoder)
        want = rhead.inference_video(rmodel, inputs, rimages, tg)
    return want, tg


class _ScriptedScene:
    """Stand-in for `model` with a scripted decoder: three objects in separate image regions (the third one enters at
    frame 5), learnable queries that find them (with duplicates), prompt queries that follow the pooled entities.  Both
    heads see identical clip outputs, so every pool update / result must agree exactly."""

    def __init__(self, Q, class_start, classes=(3, 7, 11), enter=(0, 0, 5), C=256):
        self.Q, self.class_start, self.classes, self.enter = Q, class_start, classes, enter
        g = torch.Generator().manual_seed(1)
        self.embds = torch.nn.functional.normalize(torch.randn(len(classes), C, generator=g), dim=-1) * 4
        self.noise = torch.randn(64, C, generator=g) * 0.05
        self.sem_seg_head = self._head
        self.pixel_mean = torch.tensor(MEAN).view(-1, 1, 1)
        self.pixel_std = torch.tensor(STD).view(-1, 1, 1)
        self.device = torch.device("cpu")

    def preprocess(self, frames):
        x = (torch.stack(list(frames)).float() - self.pixel_mean) / self.pixel_std
        H, W = x.shape[-2:]
        return torch.nn.functional.pad(x, (0, (W + 31) // 32 * 32 - W, 0, (H + 31) // 32 * 32 - H)), (H, W)

    def backbone(self, x):
        return {"res2": x[:, :1, ::4, ::4]}

    def _rect(self, k, f, h, w):
        m = torch.full((h, w), -5.0)
        if f >= self.enter[k]:
            x0 = 1 + k * (w // 3) + (f % 3)
            m[2 + (f % 2):h - 3, x0:x0 + w // 3 - 3] = 5.0
        return m

    def _head(self, features, targets=None):
        tg = targets[0]
        fi = [int(f) for f in tg["frame_indices"]]
        n, (h, w) = len(fi), features["res2"].shape[-2:]
        assert features["res2"].shape[0] == n
        rows = []
        for j in range(self.Q):                       # learnable queries: query j looks at object j % 3
            rows.append((j % 3, 1.0 - 0.1 * (j // 3), j))
        if "masks" in tg and tg["masks"].nelement():  # prompt queries: one per pooled entity, following "its" object
            seen = tg["mask_logits"].gt(0).any(1).float()[:, ::4, ::4]                         # [N, h, w]
            regions = torch.stack([torch.stack([self._rect(k, f, h, w) for f in range(12)]).gt(0).any(0).float()
                                   for k in range(3)])                                          # [3, h, w]
            owner = torch.einsum("nhw,khw->nk", seen[:, :h, :w], regions).argmax(1)
            for e, k in enumerate(owner.tolist()):
                rows.append((k, 1.0, 32 + e))
        logits = torch.full((1, len(rows), 3938), -4.0)
        masks = torch.empty((1, len(rows), n, h, w))
        embds = torch.empty((1, len(rows), n, self.embds.shape[1]))
        for r, (k, conf, tag) in enumerate(rows):
            visible = any(f >= self.enter[k] for f in fi)
            logits[0, r, self.class_start + self.classes[k]] = (3.0 * conf) if visible else -4.0
            for t, f in enumerate(fi):
                masks[0, r, t] = self._rect(k, f, h, w) * conf
                embds[0, r, t] = self.embds[k] + self.noise[(tag + f) % 64]
        return {"pred_logits": logits, "pred_masks": masks, "pred_embds": embds, "pred_reid_logits": [None],
                "aux_outputs": []}


@pytest.mark.parametrize("sub_task,dataset,stride", [("vis", "ytvis21", 1), ("entity_vis_coco", "ytvis21", 2)])
def test_entity_head_vis_scripted(sub_task, dataset, stride):
    """InferenceVideoEntity, instance sub-tasks, on a scripted scene: NMS of duplicate queries, pool updates from prompt
    and learnable queries, an entity entering mid-video, windowed result output -> COCO-video json, exactly."""
    from univs_b200.inference import InferenceVideoEntity
    from univs_b200.modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO as INFO
    T, Q, V, H, W = 2, 6, 14, 60, 90
    heads, rhead, rimages, inputs, pkw = _entity_pair(
        T, Q, V, H, W, 21, dataset, clip_stride=stride, apply_cls_thres=0.05, detect_newly_object_threshold=0.05,
        video_unified_inference_entities=sub_task if sub_task.startswith("entity") else "")
    scene = _ScriptedScene(Q, INFO["coco" if sub_task == "entity_vis_coco" else dataset][1])
    want, rtg = _ref_entity_run(heads, rhead, (scene,), inputs, rimages, dataset, sub_task)
    assert rtg[0]["ids"].tolist() == [0, 1, 2] and rtg[0]["first_appear_frame_idxs"].tolist()[:2] == [0, 0]
    assert rtg[0]["first_appear_frame_idxs"][2] >= 4               # the third object enters later
    got = InferenceVideoEntity(reuse_features=False, **pkw).eval(scene, inputs)
    key = lambda r: (r["category_id"], round(r["score"], 5))
    want_s, got_s = sorted(want, key=key), sorted(got, key=key)
    assert len(got_s) == len(want_s) >= 3
    for rw, rg in zip(want_s, got_s):
        assert rg["category_id"] == rw["category_id"] and rg["video_id"] == rw["video_id"] == 7
        assert abs(rg["score"] - rw["score"]) <= 1e-6
        assert rg["segmentations"] == rw["segmentations"]          # same RLE strings, frame by frame


def test_entity_head_vps_vss_scripted():
    """Panoptic (things tracked per entity, stuff merged per class) and semantic sub-tasks on the scripted scene."""
    from univs_b200.inference import InferenceVideoEntity
    from univs_b200.modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO as INFO
    T, Q, V, H, W = 2, 6, 14, 60, 90
    heads, rhead, rimages, inputs, pkw = _entity_pair(T, Q, V, H, W, 22, "vipseg", apply_cls_thres=0.05,
                                                      detect_newly_object_threshold=0.05)
    scene = _ScriptedScene(Q, INFO["vipseg"][1], classes=(2, 7, 10))     # dataset ids 3 (thing), 8 (stuff), 11 (stuff)
    want, rtg = _ref_entity_run(heads, rhead, (scene,), inputs, rimages, "vipseg", "vps")
    assert len(want["segments_infos"]) == 3 and sorted(i["isthing"] for i in want["segments_infos"]) == [False, False, True]
    got = InferenceVideoEntity(reuse_features=False, **pkw).eval(scene, inputs)
    assert got["task"] == "vps" and got["image_size"] == want["image_size"] == (75, 120)
    assert got["segments_infos"] == want["segments_infos"]
    assert got["pred_masks"].dtype == torch.int32 and torch.equal(got["pred_masks"], want["pred_masks"])
    assert got["pred_masks"].shape == (V, 75, 120) and got["pred_masks"].unique().numel() == 4

    heads, rhead, rimages, inputs, pkw = _entity_pair(T, Q, 5, H, W, 23, "vspw")
    scene = _ScriptedScene(Q, INFO["vspw"][1])
    want, _ = _ref_entity_run(heads, rhead, (scene,), inputs, rimages, "vspw", "vss")
    got = InferenceVideoEntity(reuse_features=False, **pkw).eval(scene, inputs)
    assert got["task"] == "vss" and torch.equal(got["pred_masks"], want["pred_masks"])
    assert got["pred_masks"].shape == (5, 75, 120) and got["pred_masks"].unique().tolist() == [3, 7]


def test_entity_head_vis_full_model():
    """The entity head over the real decoder + visual-prompt sampler (tiny model): entities of the first clip come back as
    prompt queries in every later clip; product (feature reuse on / off) against the reference head."""
    from univs_b200.inference import InferenceVideoEntity
    from oracle import rle_ref
    T, Q, V, H, W = 2, 12, 7, 60, 90
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=3)
    heads, rhead, rimages, inputs, pkw = _entity_pair(T, Q, V, H, W, 21, "ytvis21", box_nms_thresh=1.01,
                                                      apply_cls_thres=0.45, detect_newly_object_threshold=0.05)
    torch.manual_seed(3)              # the visual-prompt sampler draws its points from the global generator
    want, rtg = _ref_entity_run(heads, rhead, ref, inputs, rimages, "ytvis21", "vis")
    assert len(want) >= 3 and rtg[0]["ids"].numel() >= 3 and "prompt_feats" in rtg[0]
    for reuse in (True, False):
        phead = InferenceVideoEntity(reuse_features=reuse, **pkw)
        torch.manual_seed(3)
        with oracle_ops():
            got = phead.eval(model, inputs)
        ptg = phead._last_targets[0]
        assert ptg["ids"].tolist() == rtg[0]["ids"].tolist()
        torch.testing.assert_close(ptg["logits"], rtg[0]["logits"], rtol=1e-3, atol=1e-5)
        torch.testing.assert_close(ptg["occurrence"], rtg[0]["occurrence"])
        key = lambda r: (r["category_id"], round(r["score"], 4))
        want_s, got_s = sorted(want, key=key), sorted(got, key=key)
        assert [r["category_id"] for r in got_s] == [r["category_id"] for r in want_s]
        np.testing.assert_allclose([r["score"] for r in got_s], [r["score"] for r in want_s], rtol=1e-3, atol=1e-6)
        flips = total = 0
        for rw, rg in zip(want_s, got_s):
            assert len(rg["segmentations"]) == V
            for sw, sg in zip(rw["segmentations"], rg["segmentations"]):
                flips += int((rle_ref.decode(sw) != rle_ref.decode(sg)).sum())
                total += 75 * 120
        assert flips <= 1e-4 * total, (flips, total)


def test_forward_inference_dispatch():
    """UniVS_Prompt.forward with task heads attached routes whole videos like univs_prompt.py:416-452."""
    from univs_b200.modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO as INFO
    T, Q = 2, 6
    _, model = _pair(T, Q, enc_layers=1, dec_layers=2)
    kw = dict(num_queries=Q, num_frames=T, num_frames_window_test=T, reuse_features=False)
    g = torch.Generator().manual_seed(2)
    frames = [(torch.rand(3, 60, 90, generator=g) * 255).round() for _ in range(3)]
    video = {"image": frames, "height": 60, "width": 90, "task": "detection", "video_len": 3, "video_id": 1}
    calls = []

    class Spy:
        def __init__(self, name):
            self.name = name

        def eval(self, m, inputs):
            calls.append(self.name)
            return self.name

    with pytest.raises(RuntimeError):
        model.forward_inference([dict(video, dataset_name="ytvis21")])
    model.attach_task_heads(thing_ids={3}, **kw)
    assert model.task_heads["entity"].thing_ids == {3} and model.task_heads["vps"].num_frames == T
    for k in ("vis_fast", "vos", "vps", "entity", "image", "semantic_extraction"):
        model.task_heads[k] = Spy(k)
    model.sem_seg_head.predictor.semantic_extraction_enable = True
    assert model([dict(video, dataset_name="ytvis21")]) == "semantic_extraction"
    model.sem_seg_head.predictor.semantic_extraction_enable = False
    assert model([dict(video, dataset_name="ytvis21")]) == "vis_fast"
    assert model([dict(video, dataset_name="vipseg_val")]) == "vps"
    assert model([dict(video, dataset_name="davis17", task="sot")]) == "vos"
    assert model([dict(video, dataset_name="refytvos", task="grounding")]) == "vos"
    with pytest.raises(ValueError):
        model([dict(video, dataset_name="bdd_track")])
    assert model([dict(video, dataset_name="coco_panoptic")]) == "image"
    assert model([dict(video, dataset_name="ade20k_sem_seg")]) == "image"
    model.task_heads["unified"] = True
    assert model([dict(video, dataset_name="ovis")]) == "entity"
    assert model([dict(video, dataset_name="vspw")]) == "entity"
    model.task_heads["unified"], model.task_heads["tracker_type"] = False, "mdqe"
    with pytest.raises(NotImplementedError):
        model([dict(video, dataset_name="ytvis21")])
    # and one real head end to end through forward()
    model.attach_task_heads(**kw)
    with oracle_ops():
        out = model([dict(video, dataset_name="ytvis21")])
    assert out["image_size"] == (60, 90) and len(out["pred_scores"]) == len(out["pred_labels"])
    # demo form (demo/predictor.py:117): no dataset_name / task -> the keys run_on_video reads (:49-52)
    with oracle_ops():
        demo = model([{"image": frames, "height": 60, "width": 90}])
    assert set(demo) >= {"image_size", "pred_scores", "pred_labels", "pred_masks"}
    assert len(demo["pred_masks"]) == len(out["pred_masks"]) and demo["pred_scores"] == out["pred_scores"]


class _ScriptedImage(_ScriptedScene):
    """One image: 220 learnable + K category-prompt queries with blob masks of varied extent and noisy class scores."""

    def __init__(self, Q, class_start, K):
        super().__init__(Q, class_start)
        self.K = K

    def _head(self, features, targets=None):
        h, w = features["res2"].shape[-2:]
        g = torch.Generator().manual_seed(4)
        n = self.Q + self.K
        logits = torch.full((1, n, 3938), -6.0)
        logits[0, :, self.class_start:self.class_start + self.K] = torch.randn(n, self.K, generator=g) * 1.5 - 3.0
        masks = torch.full((1, n, 1, h, w), -4.0)
        for r in range(n):
            y0, x0 = int(torch.randint(0, h - 4, (1,), generator=g)), int(torch.randint(0, w - 6, (1,), generator=g))
            hh, ww = int(torch.randint(2, 7, (1,), generator=g)), int(torch.randint(3, 10, (1,), generator=g))
            masks[0, r, 0, y0:y0 + hh, x0:x0 + ww] = 3.0 + torch.rand(1, generator=g).item()
            logits[0, r, self.class_start + r % self.K] += 4.0 * torch.rand(1, generator=g).item()
        masks += torch.randn(masks.shape, generator=g) * 0.3
        for r in range(0, n, 7):                     # duplicates of the previous query: class-wise NMS must drop them
            if r:
                masks[0, r] = masks[0, r - 1] + 0.01
                logits[0, r] = logits[0, r - 1] - 0.05
        return {"pred_logits": logits, "pred_masks": masks, "pred_embds": torch.zeros(1, n, 1, 256),
                "pred_reid_logits": [None], "aux_outputs": []}


@pytest.mark.parametrize("before", [False, True])
def test_image_generic_seg_head_scripted(before):
    """InferenceImageGenericSeg (inference_image_generic_seg.py): semantic / panoptic / instance post-processing and the
    class-wise box NMS against the reference head (torchvision batched_nms) on scripted decoder outputs."""
    import inspect
    import types
    from univs_b200.inference import InferenceImageGenericSeg
    from univs_b200.modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO as INFO
    heads = ref_shim.load_inference_heads()
    Q, H, W = 220, 60, 90
    K, start = INFO["coco_panoptic"]
    things = list(range(80))                                               # contiguous ids of the thing classes
    meta = types.SimpleNamespace(thing_dataset_id_to_contiguous_id={i + 1: i for i in things})
    kw = dict(hidden_dim=256, num_queries=Q, object_mask_threshold=0.02, overlap_threshold=0.5, stability_score_thresh=0.0,
              metadata=meta, size_divisibility=32, LSJ_aug_image_size=1024, LSJ_aug_enable_test=False,
              sem_seg_postprocess_before_inference=before, pixel_mean=MEAN, pixel_std=STD, num_frames=1,
              data_name="coco_2017_val_panoptic", prompt_as_queries=True, zero_shot_inference=False, semantic_on=True,
              instance_on=True, panoptic_on=True, disable_semantic_queries=False, test_topk_per_image=20,
              tracker_type="minvis", window_inference=False, is_multi_cls=True, apply_cls_thres=0.05, merge_on_cpu=False,
              num_max_inst_test=50, num_frames_window_test=1, clip_stride=1)
    accepted = set(inspect.signature(heads.InferenceImageGenericSeg.__init__).parameters)
    rhead = heads.InferenceImageGenericSeg(**{k: v for k, v in kw.items() if k in accepted})
    scene = _ScriptedImage(Q, start, K)
    g = torch.Generator().manual_seed(8)
    image = (torch.rand(3, H, W, generator=g) * 255).round()
    inputs = [{"image": [image], "height": 75, "width": 120, "dataset_name": "coco_panoptic", "task": "detection"}]
    x, image_size = scene.preprocess([image])
    with torch.no_grad():
        want = rhead.inference_image(scene, inputs, heads.ImageList(x, [image_size]),
                                     [{"task": "detection", "dataset_name": "coco_panoptic", "frame_indices": torch.arange(1)}])[0]
    phead = InferenceImageGenericSeg(num_queries=Q, object_mask_threshold=0.02, overlap_threshold=0.5, semantic_on=True,
                                     instance_on=True, panoptic_on=True, test_topk_per_image=20,
                                     sem_seg_postprocess_before_inference=before, thing_contiguous_ids=things)
    got = phead.eval(scene, inputs)[0]
    torch.testing.assert_close(got["sem_seg"], want["sem_seg"], rtol=1e-5, atol=1e-6)
    assert got["sem_seg"].shape == (K, 75, 120)
    assert torch.equal(got["panoptic_seg"][0], want["panoptic_seg"][0]) and got["panoptic_seg"][1] == want["panoptic_seg"][1]
    assert got["panoptic_seg"][0].shape == (75, 120) and len(got["panoptic_seg"][1]) >= 3
    wi, gi = want["instances"], got["instances"]
    order_w, order_g = wi.scores.sort(descending=True)[1], gi["scores"].sort(descending=True)[1]
    assert len(order_g) == len(order_w) == 20 and gi["image_size"] == wi.image_size == (75, 120)
    torch.testing.assert_close(gi["scores"][order_g], wi.scores[order_w])
    assert torch.equal(gi["pred_classes"][order_g], wi.pred_classes[order_w])
    assert torch.equal(gi["pred_masks"][order_g], wi.pred_masks[order_w])
    assert torch.equal(gi["pred_boxes"][order_g].float(), wi.pred_boxes.tensor[order_w].float())


def test_classwise_box_nms_matches_torchvision():
    from torchvision.ops.boxes import batched_nms
    from univs_b200.inference import classwise_box_nms
    g = torch.Generator().manual_seed(0)
    for n in (0, 1, 7, 150):
        xy = torch.randint(0, 40, (n, 2), generator=g).float()
        wh = torch.randint(0, 25, (n, 2), generator=g).float()               # zero-area boxes included
        boxes = torch.cat([xy, xy + wh], 1)
        if n > 4:
            boxes[3] = boxes[2]                                               # exact duplicate
        scores, labels = torch.rand(n, generator=g), torch.randint(0, 4, (n,), generator=g)
        for thr in (0.3, 0.85):
            assert classwise_box_nms(boxes, scores, labels, thr).tolist() == batched_nms(boxes, scores, labels, thr).tolist()


def test_semantic_extraction_head(tmp_path):
    """InferenceVideoSemanticExtraction: object tokens + compressed mask features of non-overlapping clips (last clip
    shorter), against the files the reference head writes."""
    import inspect
    from univs_b200.inference import InferenceVideoSemanticExtraction
    heads = ref_shim.load_inference_heads()
    T, Q, V, H, W = 2, 8, 5, 60, 90
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=2)
    ref[2].semantic_extraction_enable = True
    model.sem_seg_head.predictor.semantic_extraction_enable = True
    g = torch.Generator().manual_seed(9)
    frames = [(torch.rand(3, H, W, generator=g) * 255).round() for _ in range(V)]
    inputs = [{"image": frames, "height": 64, "width": 96, "dataset_name": "ytvis21", "task": "detection", "video_len": V,
               "video_id": "vid0", "file_names": [f"raw/vid0/{i}.jpg" for i in range(V)]}]
    kw = dict(hidden_dim=256, num_queries=Q, overlap_threshold=0.8, overlap_threshold_entity=0.5, stability_score_thresh=0.0,
              metadata=None, size_divisibility=32, LSJ_aug_image_size=1024, LSJ_aug_enable_test=False,
              sem_seg_postprocess_before_inference=False, pixel_mean=MEAN, pixel_std=STD, num_frames=T, num_classes=133,
              semantic_extraction_enable=True, semantic_extraction_compression_ratio=8,
              semantic_extraction_compression_ratio_temporal=2, semantic_extraction_output_dir=str(tmp_path / "ref"))
    accepted = set(inspect.signature(heads.semx.InferenceVideoSemanticExtraction.__init__).parameters)
    rhead = heads.semx.InferenceVideoSemanticExtraction(**{k: v for k, v in kw.items() if k in accepted})
    xs = torch.stack([(f - rhead.pixel_mean) / rhead.pixel_std for f in frames])
    xs = torch.nn.functional.pad(xs, (0, 96 - W, 0, 64 - H))
    tg = [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "file_names": inputs[0]["file_names"]}]
    with torch.no_grad():
        rhead.inference_video(heads.RefModel(*ref), inputs, heads.ImageList(xs, [(H, W)] * V), tg)
    want_tok = torch.load(tmp_path / "ref" / "vid0._obj_tokens_8_2.pt")
    want_mf = torch.load(tmp_path / "ref" / "vid0._compression_mask_features_8_2.pt")

    phead = InferenceVideoSemanticExtraction(num_frames=T, compression_ratio=8, compression_ratio_temporal=2,
                                             output_dir=str(tmp_path / "prod"))
    with oracle_ops():
        got = phead.eval(model, inputs)
    assert got["obj_tokens"].shape == want_tok.shape == (3, 256, Q)
    assert got["compression_mask_features"].shape == want_mf.shape == (3, 256, 8, 12)
    torch.testing.assert_close(got["obj_tokens"], want_tok, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got["compression_mask_features"], want_mf, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(torch.load(tmp_path / "prod" / "vid0._obj_tokens_8_2.pt"), got["obj_tokens"])
    model.sem_seg_head.predictor.semantic_extraction_enable = False
    with pytest.raises(RuntimeError):
        phead.eval(model, inputs)


def test_entity_head_vss_full_model_uses_category_prompts():
    """VSPW (a semantic dataset): PrepareTargets gives detection clips prompt_type "text" (prepare_targets.py:58-64), so the
    124 category prompts join the learnable queries in the decoder; semantic result of the entity head vs the reference."""
    from univs_b200.inference import InferenceVideoEntity, process_inference
    T, Q, V, H, W = 2, 12, 4, 60, 90
    ref, model = _pair(T, Q, enc_layers=1, dec_layers=3)
    heads, rhead, rimages, inputs, pkw = _entity_pair(T, Q, V, H, W, 31, "vspw")
    tg = [{"task": "detection", "dataset_name": "vspw", "prompt_type": "text", "video_len": V, "sub_task": "vss"}]
    seen = []
    rmodel = heads.RefModel(*ref)
    orig = rmodel.sem_seg_head
    rmodel.sem_seg_head = lambda feats, targets=None: (lambda o: (seen.append(o["pred_masks"].shape[1]), o)[1])(orig(feats, targets=targets))
    with torch.no_grad():
        want = rhead.inference_video(rmodel, inputs, rimages, tg)
    assert seen and all(n == Q + 124 for n in seen)             # the category prompts were in the decoder
    assert process_inference(inputs[0], (64, 96), (H, W), T)[0]["prompt_type"] == "text"
    assert process_inference(dict(inputs[0], dataset_name="ytvis21"), (64, 96), (H, W), T)[0]["prompt_type"] == "visual"
    assert process_inference(dict(inputs[0], dataset_name="ytvis21"), (64, 96), (H, W), T, semantic_on=True)[0]["prompt_type"] == "text"
    for reuse in (True, False):
        with oracle_ops():
            got = InferenceVideoEntity(reuse_features=reuse, **pkw).eval(model, inputs)
        assert got["task"] == "vss" and got["pred_masks"].shape == want["pred_masks"].shape == (V, 75, 120)
        assert (got["pred_masks"] != want["pred_masks"]).float().mean().item() <= 1e-3
