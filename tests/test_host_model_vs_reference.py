"""Host-logic parity on CPU: the product modules (operators swapped for their CPU oracles, tests/oracle_backend.py)
against the reference's own modules executed by path, with identical key-seeded weights and inputs.
Also checks state_dict key/shape compatibility."""
import pytest
import torch

from oracle import ref_shim
from tests import model_factory as mf
from oracle.cpu_backend import oracle_ops

pytestmark = pytest.mark.reference


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _build_pair(swin, T, Q, **kw):
    clip = mf.make_clip_emb()
    rbb, rpix, rdec = ref_shim.build_reference_model(swin, num_queries=Q, num_frames=T, clip_emb=clip, **kw)
    pbb, ppix, pdec = mf.build_product_model(swin, num_queries=Q, num_frames=T, clip_emb=clip, **kw)
    for r, p in ((rbb, pbb), (rpix, ppix), (rdec, pdec)):
        rsd, psd = r.state_dict(), p.state_dict()
        assert set(rsd) == set(psd), (sorted(set(rsd) - set(psd))[:5], sorted(set(psd) - set(rsd))[:5])
        for k in rsd:
            assert rsd[k].shape == psd[k].shape, k
        sd = mf.keyed_state_dict(rsd)
        r.load_state_dict(sd)
        p.load_state_dict(sd)
    return (rbb, rpix, rdec), (pbb, ppix, pdec)


@pytest.mark.parametrize("swin,H,W", [(mf.TINY_SWIN, 64, 96), (mf.SMALL_SWIN7, 96, 160)])
def test_detection_no_prompts(swin, H, W):
    torch.manual_seed(0)
    T, Q = 2, 10
    ref, prod = _build_pair(swin, T, Q, enc_layers=2, dec_layers=4)
    x = torch.randn(T, 3, H, W)
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual",
                   "frame_indices": torch.tensor([3, 4])}]
    rf, (rmf, rms), rout = ref_shim.reference_clip_forward(*ref, x, tg())
    prod[2].return_aux_outputs = True
    with oracle_ops():
        pf, (pmf, pms), pout = mf.product_clip_forward(*prod, x, tg())
    for k in rf:
        assert pf[k].shape == rf[k].shape
        assert _rel(pf[k], rf[k]) < 1e-4, k
    assert _rel(pmf, rmf) < 1e-4
    for a, b in zip(pms, rms):
        assert _rel(a, b) < 1e-4
    assert pout["pred_masks"].shape == rout["pred_masks"].shape
    assert _rel(pout["pred_masks"], rout["pred_masks"]) < 1e-3
    assert _rel(pout["pred_logits"], rout["pred_logits"]) < 1e-3
    assert _rel(pout["pred_embds"], rout["pred_embds"]) < 1e-3
    assert len(pout["aux_outputs"]) == len(rout["aux_outputs"])
    for pa, ra in zip(pout["aux_outputs"], rout["aux_outputs"]):
        assert _rel(pa["pred_masks"], ra["pred_masks"]) < 1e-3
        assert _rel(pa["pred_logits"], ra["pred_logits"]) < 1e-3
        assert _rel(pa["pred_embds"], ra["pred_embds"]) < 1e-3


@pytest.mark.parametrize("l2v", [False, True])
def test_detection_category_prompts_proca(l2v):
    """task=detection, prompt_type=text: one prompt query per category of the dataset, ProCA with L=1."""
    torch.manual_seed(1)
    T, Q = 2, 6
    ref, prod = _build_pair(mf.TINY_SWIN, T, Q, enc_layers=1, dec_layers=3, text_prompt_to_image_enable=l2v)
    x = torch.randn(T, 3, 64, 64)
    tg = lambda: [{"task": "detection", "dataset_name": "bdd_track", "prompt_type": "text",
                   "frame_indices": torch.arange(T)}]
    _, _, rout = ref_shim.reference_clip_forward(*ref, x, tg())
    with oracle_ops():
        _, _, pout = mf.product_clip_forward(*prod, x, tg())
    assert pout["pred_masks"].shape == rout["pred_masks"].shape == (1, Q + 8, T, 16, 16)
    assert _rel(pout["pred_masks"], rout["pred_masks"]) < 1e-3
    assert _rel(pout["pred_logits"], rout["pred_logits"]) < 1e-3
    assert _rel(pout["pred_embds"], rout["pred_embds"]) < 1e-3


@pytest.mark.parametrize("l2v", [False, True])
def test_grounding_text_prompts(l2v):
    torch.manual_seed(2)
    T, Q, P = 2, 6, 3
    ref, prod = _build_pair(mf.TINY_SWIN, T, Q, enc_layers=1, dec_layers=3, text_prompt_to_image_enable=l2v,
                            self_attn_mask_type="sep-blocked")
    x = torch.randn(T, 3, 64, 64)
    words, sent = torch.randn(P, 77, T, 640), torch.randn(P, T, 640)
    tg = lambda: [{"task": "grounding", "dataset_name": "refytvos", "prompt_type": "text",
                   "frame_indices": torch.arange(T), "exp_word_feats": words.clone(),
                   "exp_sentence_feats": sent.clone(), "exp_word_len": torch.full((P,), 10)}]
    _, _, rout = ref_shim.reference_clip_forward(*ref, x, tg())
    with oracle_ops():
        _, _, pout = mf.product_clip_forward(*prod, x, tg())
    assert pout["pred_masks"].shape == rout["pred_masks"].shape == (1, Q + P, T, 16, 16)
    assert pout["pred_logits"].shape == rout["pred_logits"].shape == (1, Q + P, P)
    assert _rel(pout["pred_masks"], rout["pred_masks"]) < 1e-3
    assert _rel(pout["pred_logits"], rout["pred_logits"]) < 1e-3
    assert _rel(pout["pred_reid_logits"], rout["pred_reid_logits"]) < 1e-3
    assert _rel(pout["pred_embds"], rout["pred_embds"]) < 1e-3


def _rect_masks(P, n, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    masks = torch.zeros(P, n, H, W)
    boxes = torch.zeros(P, n, 4)
    for p in range(P):
        for f in range(n):
            bw, bh = int(W * (0.2 + 0.3 * torch.rand(1, generator=g))), int(H * (0.2 + 0.3 * torch.rand(1, generator=g)))
            x0, y0 = int((W - bw) * torch.rand(1, generator=g)), int((H - bh) * torch.rand(1, generator=g))
            masks[p, f, y0:y0 + bh, x0:x0 + bw] = 1.0
            boxes[p, f] = torch.tensor([x0 / W, y0 / H, (x0 + bw) / W, (y0 + bh) / H])
    return masks, boxes


def test_sot_visual_prompts_two_clips():
    """task=sot: VisualPromptSampler + memory pool + ProCA over two consecutive stride-1 clips (same RNG seed/order)."""
    T, Q, P, H, W = 3, 6, 3, 64, 96
    ref, prod = _build_pair(mf.TINY_SWIN, T, Q, enc_layers=1, dec_layers=3, num_dense_points=8,
                            num_prev_frames_memory=4)
    g = torch.Generator().manual_seed(11)
    frames = torch.randn(T + 1, 3, H, W, generator=g)
    masks, boxes = _rect_masks(P, T + 1, H, W, 5)
    masks[2, 0] = 0        # object 2 is absent in the first frame (blank prompt)

    def targets():
        return [{"task": "sot", "dataset_name": "davis", "prompt_type": "visual", "ids": torch.arange(P),
                 "first_appear_frame_idxs": torch.tensor([0, 0, 1])}]

    def clip_inputs(tg, c):
        tg[0]["first_frame_idx"] = c
        tg[0]["frame_indices"] = torch.arange(c, c + T)
        tg[0]["masks"] = masks[:, : c + T].clone()
        tg[0]["boxes"] = boxes[:, : c + T].clone()
        return frames[c: c + T]

    rtg, ptg = targets(), targets()
    for c in range(2):
        x = clip_inputs(rtg, c)
        torch.manual_seed(100 + c)
        _, _, rout = ref_shim.reference_clip_forward(*ref, x, rtg)
        x = clip_inputs(ptg, c)
        torch.manual_seed(100 + c)
        with oracle_ops():
            _, _, pout = mf.product_clip_forward(*prod, x, ptg)
        assert pout["pred_masks"].shape == rout["pred_masks"].shape == (1, Q + P, T, 16, 24)
        assert _rel(pout["pred_masks"], rout["pred_masks"]) < 1e-3, c
        assert _rel(pout["pred_logits"], rout["pred_logits"]) < 1e-3
        assert _rel(pout["pred_embds"], rout["pred_embds"]) < 1e-3
        for k in ("prompt_feats", "prompt_pe"):
            assert ptg[0][k].shape == rtg[0][k].shape, (k, ptg[0][k].shape, rtg[0][k].shape)
            assert _rel(ptg[0][k], rtg[0][k]) < 1e-4, k
        assert torch.equal(ptg[0]["prompt_attn_masks"], rtg[0]["prompt_attn_masks"])


def test_split_operand_policy_plumbing_on_cpu():
    """The tf32x3 policy's hi|lo operand plumbing (fused LN/GELU/ReLU split outputs, cached split weights, split
    convolution) reproduces the reference when every GEMM is an exact fp32 GEMM (CPU)."""
    from univs_b200 import nn_ops
    torch.manual_seed(0)
    T, Q = 2, 8
    ref, prod = _build_pair(mf.SMALL_SWIN7, T, Q, enc_layers=1, dec_layers=2)
    x = torch.randn(T, 3, 64, 96)
    tg = lambda: [{"task": "detection", "dataset_name": "bdd_track", "prompt_type": "text",
                   "frame_indices": torch.arange(T)}]
    rf, (rmf, rms), rout = ref_shim.reference_clip_forward(*ref, x, tg())
    with oracle_ops(policy="tf32x3"):
        pf, (pmf, pms), pout = mf.product_clip_forward(*prod, x, tg())
    assert nn_ops.policy() == "fp32"
    assert _rel(pf["res5"], rf["res5"]) < 1e-4
    assert _rel(pmf, rmf) < 1e-4
    assert _rel(pout["pred_masks"], rout["pred_masks"]) < 1e-3
    assert _rel(pout["pred_logits"], rout["pred_logits"]) < 1e-3
