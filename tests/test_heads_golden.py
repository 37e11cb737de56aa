"""Sliding-window task heads against the committed golden results of the REFERENCE heads
(tests/golden/make_golden_heads.py); needs no reference tree, so it also runs on the GPU box.
 * not-gpu: product heads + host model with CPU oracle operators;
 * gpu:     product heads + host model with the CUDA kernels.  New in this round and not yet run on a B200, therefore
            opt-in (UNIVS_GPU_HEADS=1) until validated -- see DESIGN.md section 8."""
import os

import numpy as np
import pytest
import torch

from oracle.cpu_backend import oracle_ops
from tests import model_factory as mf
from tests.golden.make_golden_heads import MEAN, SOT, STD, VIS, sot_annotations, video
from univs_b200.inference import FrameAnnotations, InferenceVideoVISFast, InferenceVideoVOS
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


def test_vis_fast_head_with_oracle_ops_matches_golden():
    with oracle_ops():
        _vis("cpu")


def test_vos_sot_head_with_oracle_ops_matches_golden():
    with oracle_ops():
        _sot("cpu")


_gpu_heads = pytest.mark.skipif(os.environ.get("UNIVS_GPU_HEADS") != "1",
                                reason="task heads on CUDA: opt-in until validated on a B200 (UNIVS_GPU_HEADS=1)")


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
