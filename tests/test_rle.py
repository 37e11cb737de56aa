"""COCO RLE result format (univs_b200/inference/rle.py) against the CPU restatement of pycocotools' maskApi.c
(oracle/rle_ref.py).  The compressed-string form is unpinned against pycocotools itself (not installed); the run
lengths are pinned on the documented examples and everything must round-trip."""
import numpy as np
import pytest
import torch

from oracle import rle_ref
from univs_b200.inference import rle


def test_documented_examples():
    # pycocotools/mask.py docstring: column vectors M=[0 0 1 1 1 0 1] -> [2 3 1 1], M=[1 1 1 1 1 1 0] -> [0 6 1]
    assert rle_ref.rle_counts(np.array([[0, 0, 1, 1, 1, 0, 1]]).T) == [2, 3, 1, 1]
    assert rle_ref.rle_counts(np.array([[1, 1, 1, 1, 1, 1, 0]]).T) == [0, 6, 1]
    assert rle_ref.rle_to_string([4]) == "4"                     # blank 2x2 mask: one run of four zeros


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (7, 5), (33, 64), (90, 160)])
def test_encode_equals_oracle_and_round_trips(shape):
    rng = np.random.default_rng(sum(shape))
    h, w = shape
    masks = [np.zeros(shape, bool), np.ones(shape, bool), rng.random(shape) < 0.5]
    blob = np.zeros(shape, bool)
    blob[h // 4: max(h // 4 + 1, 3 * h // 4), w // 3: max(w // 3 + 1, 2 * w // 3)] = True      # long runs: multi-character counts
    masks.append(blob)
    first_one = np.zeros(shape, bool)
    first_one[0, 0] = True
    masks.append(first_one)
    got = rle.encode(torch.from_numpy(np.stack(masks)))
    for m, g in zip(masks, got):
        want = rle_ref.encode(m)
        assert g == want
        assert np.array_equal(rle_ref.decode(g).astype(bool), m)
    # negative deltas and large counts survive the string codec
    for cnts in ([0, 1, 1000000, 3, 2, 999], [5, 70000, 1, 1, 70000, 2, 31, 32, 33]):
        assert rle_ref.rle_from_string(rle_ref.rle_to_string(cnts)) == cnts
        assert rle.counts_to_string(cnts) == rle_ref.rle_to_string(cnts)


def test_vis_head_rle_output_matches_dense_masks():
    """InferenceVideoVISFast(rle_output=True): the RLEs decode to exactly the dense masks of the default output"""
    from oracle.cpu_backend import oracle_ops
    from tests.golden.make_golden_heads import VIS, video
    from tests.test_heads_golden import _model
    from univs_b200.inference import InferenceVideoVISFast
    s = VIS
    model = _model(s, "cpu")
    inp = lambda: [{"image": video(s), "height": s["out"][0], "width": s["out"][1], "dataset_name": "ytvis21"}]
    kw = dict(num_queries=s["Q"], num_frames=s["T"], test_topk_per_image=s["topk"], num_frames_window_test=s["T"])
    with oracle_ops():
        dense = InferenceVideoVISFast(**kw).eval(model, inp())
        packed = InferenceVideoVISFast(rle_output=True, **kw).eval(model, inp())
    assert packed["pred_masks"] == [] and len(packed["segmentations"]) == len(dense["pred_masks"]) > 0
    assert packed["pred_scores"] == dense["pred_scores"] and packed["pred_labels"] == dense["pred_labels"]
    for rles, m in zip(packed["segmentations"], dense["pred_masks"]):
        assert len(rles) == m.shape[0]
        for r, frame in zip(rles, m):
            assert np.array_equal(rle_ref.decode(r).astype(bool), frame.numpy())
