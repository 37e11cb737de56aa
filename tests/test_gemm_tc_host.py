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


@pytest.mark.parametrize("Cin,Cout,k", [(64, 24, 3), (128, 40, 3), (512, 16, 3), (64, 8, 1)])
def test_fused_conv_taps_equal_conv2d(Cin, Cout, k, monkeypatch):
    """CONV_FUSED: a k x k "same" convolution as ONE shifted-row accumulation per tap group (K <= 1536 per chain) of
    ops.gemm_f16x3_tc (univs_gemm_f16x3_tc_taps) -- against F.conv2d (msdeformattn.py:345-360, the FPN output convolutions)."""
    g = torch.Generator().manual_seed(Cin + Cout)
    conv = torch.nn.Conv2d(Cin, Cout, k, padding=k // 2)
    x = torch.randn(2, 5, 7, Cin, generator=g)
    want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), conv.weight.double(), conv.bias.double(),
                                      padding=k // 2).permute(0, 2, 3, 1)
    outs = {}
    with oracle_ops("fp16x3"):
        nn_ops.set_gemm_tc(True)
        try:
            for fused in (False, True):
                monkeypatch.setattr(nn_ops, "_conv_fused", fused)
                calls = []
                from univs_b200 import ops
                real = ops.gemm_f16x3_tc
                monkeypatch.setattr(ops, "gemm_f16x3_tc", lambda *a, **kw: (calls.append(kw.get("tap_rows")), real(*a, **kw))[1])
                outs[fused] = nn_ops.conv2d_cl(x, conv.weight, conv.bias, padding=k // 2)
                monkeypatch.setattr(ops, "gemm_f16x3_tc", real)
                if fused:       # ceil(k*k*Cin / 1536) launches, each with its group of taps
                    assert len(calls) == -(-k * k * Cin // 1536) and all(c is not None for c in calls)
                    assert sum(len(c) for c in calls) == k * k
                else:
                    assert len(calls) == k * k and all(c is None for c in calls)
        finally:
            nn_ops.set_gemm_tc(False)
    for fused in (False, True):
        err = float((outs[fused] - want).abs().max() / want.abs().max())
        assert err < 2e-6, (fused, err)
