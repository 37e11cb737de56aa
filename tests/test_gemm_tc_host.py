"""Host plumbing of the GEMM_TC switch (nn_ops: dense layers through ops.gemm_f16x3_tc -- chunked / compact operand
containers, K-slice accumulation, fused MLP, 3x3 convolution taps on row views) with the CPU oracle operators: the clip
forward under the fp16x3 policy must reproduce the fp32-policy forward that the reference-parity tests pin."""
import pytest
import torch

from oracle.cpu_backend import oracle_ops
from tests import model_factory as mf
from univs_b200 import nn_ops
from univs_b200.meta_arch import UniVS_Prompt
from univs_b200.modeling.head import MaskFormerHead
from univs_b200.registry import ShapeSpec

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _model(T, Q):
    parts = mf.build_product_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(), enc_layers=2,
                                   dec_layers=2)
    mf.load_keyed(parts)
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=parts[1], transformer_predictor=parts[2])
    return UniVS_Prompt(backbone=parts[0], sem_seg_head=head, pixel_mean=MEAN, pixel_std=STD)


@pytest.mark.parametrize("glue", [False, True])
def test_gemm_tc_host_path_equals_fp32_policy(glue):
    T, Q = 2, 6
    model = _model(T, Q)
    g = torch.Generator().manual_seed(4)
    frames = (torch.rand(T, 3, 60, 90, generator=g) * 255).round()
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}]
    with oracle_ops("fp32"):
        want = model.clip_forward(frames, tg())
    nn_ops.set_fused_glue(glue)
    try:
        with oracle_ops("fp16x3"):
            nn_ops.set_gemm_tc(True)
            assert nn_ops.gemm_tc()
            got = model.clip_forward(frames, tg())
    finally:
        nn_ops.set_gemm_tc(False)
        nn_ops.set_fused_glue(False)
    for k in ("pred_masks", "pred_logits", "pred_embds"):
        err = (got[k] - want[k]).abs().max().item() / want[k].abs().max().item()
        assert err < 2e-5, (k, err)


def test_k_slices_accumulate_in_place():
    """K > 1536: one launch per K-chunk, the partial result passed back as addend / out"""
    with oracle_ops("fp16x3"):
        nn_ops.set_gemm_tc(True)
        try:
            g = torch.Generator().manual_seed(1)
            x, lin = torch.randn(7, 3072, generator=g), torch.nn.Linear(3072, 24)
            y = nn_ops.linear(x, lin.weight, lin.bias)
        finally:
            nn_ops.set_gemm_tc(False)
    want = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
    assert float((y - want).abs().max() / want.abs().max()) < 2e-6
