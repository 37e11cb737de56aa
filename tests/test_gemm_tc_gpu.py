"""csrc/gemm_tc.cu on the B200 through the C ABI against the oracle restatement (oracle/ops_ref.py::gemm_f16x3) and the
exact product of the original fp32 values; then the whole clip forward with the GEMM_TC switch against the library-GEMM
path (same operands, same split: results agree to rounding)."""
import pytest
import torch

from oracle import ops_ref
from univs_b200 import nn_ops, ops

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("M,N,K,act", [
    (200, 192, 128, 0), (130, 72, 96, 1), (300, 260, 192, 2), (128, 128, 64, 0),
    (5000, 576, 192, 0),        # Swin-L stage-1 qkv shape (fewer tokens)
    (4111, 768, 192, 1),        # fc1 + GELU, ragged token count
    (1000, 3938, 640, 0),       # class logits: odd channel count
    (777, 256, 1536, 0),        # longest single K-chunk
    (64, 48, 48, 0),            # patch embedding: one partial k-block
    (300, 71, 64, 2),           # odd channel count
])
def test_gemm_tc_matches_oracle(M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    s = 8
    x3 = ops.split_operand(x.cuda(), "f16")
    w3 = ops.split_operand((w * 2.0 ** s).cuda(), "f16")
    offs = (2 * K, 0)
    y, y16 = ops.gemm_f16x3_tc(x3, offs, w3, offs, K, 2.0 ** -s, bias.cuda(), add.cuda(), want_f32=True, want_operand=True, act=act)
    torch.cuda.synchronize()
    wy, wy16 = ops_ref.gemm_f16x3(x3.cpu(), offs, w3.cpu(), offs, K, 2.0 ** -s, bias, add, act)
    assert _rel(y, wy) < 4e-6      # K = 1536: 96 truncating accumulation steps of the tensor core (tools/accum_probe.py)
    ref = x.double() @ w.double().t() + bias.double()
    ref = torch.nn.functional.gelu(ref) if act == 1 else (torch.relu(ref) if act == 2 else ref)
    assert _rel(y, ref + add.double()) < 5e-6
    y16 = y16.float().cpu()
    assert _rel(y16[:, :N] + y16[:, N:] / 2048.0, y) < 1e-6
    assert torch.equal(y16[:, :N], y.half().float().cpu())


def test_gemm_tc_mlp_chain_and_k_slices():
    """fc1 + GELU + operand epilogue feeding fc2 over the compact container in K slices accumulated in place"""
    g = torch.Generator().manual_seed(7)
    M, C, Hd = 3000, 768, 3072
    x, w1, w2 = torch.randn(M, C, generator=g), torch.randn(Hd, C, generator=g) * 0.03, torch.randn(C, Hd, generator=g) * 0.02
    b1 = torch.randn(Hd, generator=g) * 0.1
    x3 = ops.split_operand(x.cuda(), "f16")
    w13, w23 = ops.split_operand((w1 * 64).cuda(), "f16"), ops.split_operand((w2 * 64).cuda(), "f16")
    _, h16 = ops.gemm_f16x3_tc(x3, (2 * C, 0), w13, (2 * C, 0), C, 1 / 64, b1.cuda(), None, want_f32=False, want_operand=True,
                               act=ops.ACT_GELU)
    kc = ops.f16_chunk(Hd)
    assert kc == 1536
    out = None
    for c in range(Hd // kc):
        out, _ = ops.gemm_f16x3_tc(h16, (c * kc, Hd + c * kc), w23, (c * 3 * kc + 2 * kc, c * 3 * kc), kc, 1 / 64, None, out, out=out)
    torch.cuda.synchronize()
    hid = torch.nn.functional.gelu(x.double() @ w1.double().t() + b1.double())
    assert _rel(out, hid @ w2.double().t()) < 5e-6


@pytest.mark.parametrize("glue", [False, True])
def test_clip_forward_with_gemm_tc_equals_library_gemm_path(glue):
    from tests import model_factory as mf
    from univs_b200.meta_arch import UniVS_Prompt
    from univs_b200.modeling.head import MaskFormerHead
    from univs_b200.precision import get_precision, set_precision
    from univs_b200.registry import ShapeSpec
    T, Q = 2, 6
    parts = mf.build_product_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(), enc_layers=2, dec_layers=2)
    mf.load_keyed(parts)
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=parts[1], transformer_predictor=parts[2])
    model = UniVS_Prompt(backbone=parts[0], sem_seg_head=head, pixel_mean=[123.675, 116.28, 103.53],
                         pixel_std=[58.395, 57.12, 57.375]).cuda()
    g = torch.Generator().manual_seed(4)
    frames = (torch.rand(T, 3, 96, 160, generator=g) * 255).round().cuda()
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}]
    old, old_glue, old_tc = get_precision(), nn_ops.fused_glue(), nn_ops._gemm_tc
    set_precision("fp16x3")
    nn_ops.set_fused_glue(glue)
    try:
        nn_ops.set_gemm_tc(False)
        want = model.clip_forward(frames, tg())
        nn_ops.set_gemm_tc(True)
        got = model.clip_forward(frames, tg())
        torch.cuda.synchronize()
    finally:
        nn_ops.set_gemm_tc(old_tc)
        nn_ops.set_fused_glue(old_glue)
        set_precision(old)
    for k in ("pred_masks", "pred_logits", "pred_embds"):
        assert _rel(got[k], want[k]) < 5e-5, k


@pytest.mark.parametrize("N,H,W,Cin,Cout,k", [(2, 46, 80, 256, 256, 3), (1, 23, 40, 128, 72, 3), (2, 9, 11, 64, 24, 1)])
def test_fused_conv_taps_on_device(N, H, W, Cin, Cout, k, monkeypatch):
    """CONV_FUSED (univs_gemm_f16x3_tc_taps): the k x k FPN convolution as one shifted-row accumulation per tap group, against
    F.conv2d in float64 (msdeformattn.py:345-360) and against the k*k-GEMM path"""
    from univs_b200.precision import get_precision, set_precision
    g = torch.Generator().manual_seed(H * W + Cin)
    torch.manual_seed(H * W + Cin)              # the layer's init draws from the global generator: same weights in any test order
    conv = torch.nn.Conv2d(Cin, Cout, k, padding=k // 2)
    x = torch.randn(N, H, W, Cin, generator=g)
    want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), conv.weight.double(), conv.bias.double(),
                                      padding=k // 2).permute(0, 2, 3, 1)
    old, old_tc = get_precision(), nn_ops._gemm_tc
    set_precision("fp16x3")
    nn_ops.set_gemm_tc(True)
    conv = conv.cuda()
    outs = {}
    try:
        for fused in (False, True):
            monkeypatch.setattr(nn_ops, "_conv_fused", fused)
            outs[fused] = nn_ops.conv2d_cl(x.cuda(), conv.weight, conv.bias, padding=k // 2).contiguous()
        torch.cuda.synchronize()
    finally:
        nn_ops.set_gemm_tc(old_tc)
        set_precision(old)
    assert _rel(outs[True], want) < 5e-6 and _rel(outs[False], want) < 5e-6
    assert _rel(outs[True], outs[False]) < 6e-6      # two fp32 accumulation orders of the same products (each within 5e-6 of float64)
