"""Sliding-window task heads against the committed golden results of the REFERENCE heads
(tests/golden/make_golden_heads.py); needs no reference tree, so it also runs on the GPU box.
 * not-gpu: product heads + host model with CPU oracle operators;
 * gpu:     product heads + host model with the CUDA kernels.  New in this round and not yet run on a B200, therefore
            validated on the B200 in round 2 (gpurun_out/r2_open)."""
import os

import numpy as np
import pytest
import torch

from oracle.cpu_backend import oracle_ops
from tests import model_factory as mf
from tests.golden.make_golden_heads import ENTITY, MEAN, SOT, STD, VIS, VPS, sot_annotations, video
from univs_b200.inference import (FrameAnnotations, InferenceVideoEntity, InferenceVideoVISFast, InferenceVideoVOS,
                                  InferenceVideoVPS)
from univs_b200.meta_arch import UniVS_Prompt
from univs_b200.modeling.head import MaskFormerHead
from univs_b200.registry import ShapeSpec

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(spec, device):
    parts = mf.build_product_model(mf.TINY_SWIN, num_queries=spec["Q"], num_frames=spec["T"],
                                   clip_emb=mf.make_clip_emb(), **spec["model"])
    mf.load_keyed(parts)
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=parts[1], transformer_predictor=parts[2])
    return UniVS_Prompt(backbone=parts[0], sem_seg_head=head, pixel_mean=MEAN, pixel_std=STD).to(device)


def _vis(device):
    s = VIS
    gold = torch.load(os.path.join(HERE, "head_vis_fast.pt"))
    head = InferenceVideoVISFast(num_queries=s["Q"], num_frames=s["T"], test_topk_per_image=s["topk"],
                                 num_frames_window_test=s["T"])
    got = head.eval(_model(s, device), [{"image": video(s), "height": s["out"][0], "width": s["out"][1],
                                         "dataset_name": "ytvis21"}])
    order = np.lexsort((got["pred_labels"], got["pred_scores"]))
    assert [got["pred_labels"][i] for i in order] == gold["labels"].tolist()
    np.testing.assert_allclose([got["pred_scores"][i] for i in order], gold["scores"].numpy(), rtol=1e-3, atol=1e-6)
    want = np.unpackbits(gold["masks_packed"].numpy())[:int(np.prod(gold["masks_shape"]))].reshape(gold["masks_shape"])
    mine = np.stack([got["pred_masks"][i].numpy() for i in order])
    assert mine.shape == want.shape
    assert (mine != want.astype(bool)).mean() <= 1e-3


def _sot(device):
    s = SOT
    gold = torch.load(os.path.join(HERE, "head_vos_sot.pt"))
    head = InferenceVideoVOS(num_queries=s["Q"], num_frames=s["T"], num_frames_window_test=s["T"],
                             num_prev_frames_memory=4)
    inst = sot_annotations(s, lambda f, ids, m, b: FrameAnnotations((s["H"], s["W"]), ids, m, b))
    model = _model(s, device)
    torch.manual_seed(s["rng"])              # the prompt sampler draws its points from the global CPU generator
    got = head.eval(model, [{"image": video(s), "task": "sot", "dataset_name": "davis", "instances": inst}])
    assert sorted(got["frames"]) == sorted(gold["id_maps"])
    flips = sum((got["frames"][f] != gold["id_maps"][f]).sum().item() for f in gold["id_maps"])
    total = sum(v.numel() for v in gold["id_maps"].values())
    assert flips <= 1e-3 * total, (flips, total)
    tg = head._last_targets[0]
    assert tg["ids"] == gold["ids"]
    for k in ("boxes", "embds"):
        err = (tg[k].cpu() - gold[k]).abs().max().item() / max(gold[k].abs().max().item(), 1e-6)
        assert err < (2e-2 if k == "boxes" else 1e-3), (k, err)      # boxes are integer pixel edges / size


def _entity(device):
    from oracle import rle_ref
    s = ENTITY
    gold = torch.load(os.path.join(HERE, "head_entity_vis.pt"))
    head = InferenceVideoEntity(num_queries=s["Q"], num_frames=s["T"], num_frames_window_test=s["T"], **s["head"])
    model = _model(s, device)
    torch.manual_seed(s["rng"])              # the prompt sampler draws its points from the global CPU generator
    got = head.eval(model, [{"image": video(s), "height": s["out"][0], "width": s["out"][1],
                             "dataset_name": "ytvis21", "task": "detection", "video_len": s["V"], "video_id": 7}])
    tg = head._last_targets[0]
    assert tg["ids"].tolist() == gold["ids"].tolist() and tg["first_appear_frame_idxs"].tolist() == gold["first_appear"].tolist()
    torch.testing.assert_close(tg["logits"].cpu(), gold["logits"], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(tg["occurrence"].cpu(), gold["occurrence"])
    got = sorted(got, key=lambda r: (r["category_id"], round(r["score"], 4)))
    assert [r["category_id"] for r in got] == gold["category_ids"]
    np.testing.assert_allclose([r["score"] for r in got], gold["scores"].numpy(), rtol=1e-3, atol=1e-6)
    flips = total = 0
    size = list(s["out"])
    for r, counts in zip(got, gold["segmentations"]):
        assert len(r["segmentations"]) == s["V"]
        for seg, c in zip(r["segmentations"], counts):
            flips += int((rle_ref.decode(seg) != rle_ref.decode({"size": size, "counts": c})).sum())
            total += size[0] * size[1]
    assert flips <= 1e-3 * total, (flips, total)


def _vps(device):
    s = VPS
    gold = torch.load(os.path.join(HERE, "head_vps.pt"))
    head = InferenceVideoVPS(num_queries=s["Q"], num_frames=s["T"], num_frames_window_test=s["T"], thing_ids=s["things"],
                             change_to_720p=False, **s["head"])
    got = head.eval(_model(s, device), [{"image": video(s), "height": s["out"][0], "width": s["out"][1],
                                         "dataset_name": "vipseg", "task": "detection", "video_len": s["V"]}])
    assert got["image_size"] == tuple(gold["image_size"]) and got["segments_infos"] == gold["segments_infos"]
    assert [int(i) for i in got["pred_ids"]] == gold["pred_ids"]
    assert (got["pred_masks"] != gold["pred_masks"].to(torch.int32)).float().mean().item() <= 1e-3


def test_entity_head_with_oracle_ops_matches_golden():
    with oracle_ops():
        _entity("cpu")


def test_vps_head_with_oracle_ops_matches_golden():
    with oracle_ops():
        _vps("cpu")


def test_vis_fast_head_with_oracle_ops_matches_golden():
    with oracle_ops():
        _vis("cpu")


def test_vos_sot_head_with_oracle_ops_matches_golden():
    with oracle_ops():
        _sot("cpu")


_gpu_heads = pytest.mark.filterwarnings("default")      # validated on a B200 (round 2): no gate


@pytest.mark.gpu
@_gpu_heads
@pytest.mark.parametrize("policy", ["fp16x3", "fp32"])
def test_vis_fast_head_cuda_matches_golden(policy):
    from univs_b200 import precision
    precision.set_precision(policy)
    try:
        _vis("cuda")
    finally:
        precision.set_precision("fp32")


@pytest.mark.gpu
@_gpu_heads
@pytest.mark.parametrize("policy", ["fp16x3", "fp32"])
def test_vos_sot_head_cuda_matches_golden(policy):
    from univs_b200 import precision
    precision.set_precision(policy)
    try:
        _sot("cuda")
    finally:
        precision.set_precision("fp32")


@pytest.mark.gpu
@_gpu_heads
@pytest.mark.parametrize("which", ["entity", "vps"])
def test_entity_and_vps_heads_cuda_match_golden(which):
    from univs_b200 import precision
    precision.set_precision("fp16x3")
    try:
        (_entity if which == "entity" else _vps)("cuda")
    finally:
        precision.set_precision("fp32")
