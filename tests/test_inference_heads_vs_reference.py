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
