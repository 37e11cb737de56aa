"""Opt-in fused glue path (nn_ops.set_fused_glue: channel-last GroupNorm + FPN add / ReLU / operand emission,
PatchMerging gather-LayerNorm, fused frame ingest).  CPU: the host plumbing with oracle operators must reproduce the
default path (which the reference-parity tests pin); GPU op tests are opt-in (UNIVS_GPU_GLUE=1) until validated."""
import os

import pytest
import torch

from oracle import ops_ref
from oracle.cpu_backend import oracle_ops
from tests import model_factory as mf
from univs_b200 import nn_ops
from univs_b200.meta_arch import UniVS_Prompt
from univs_b200.modeling.head import MaskFormerHead
from univs_b200.registry import ShapeSpec

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _model(T=2, Q=6):
    parts = mf.build_product_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(),
                                   enc_layers=2, dec_layers=2)      # two encoder layers: operands carried between them
    mf.load_keyed(parts)
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=parts[1], transformer_predictor=parts[2])
    return UniVS_Prompt(backbone=parts[0], sem_seg_head=head, pixel_mean=MEAN, pixel_std=STD)


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("policy", ["fp32", "tf32x3"])
@pytest.mark.parametrize("size", [(60, 90), (64, 96), (50, 70)])
def test_fused_glue_host_path_equals_default(policy, size):
    T = 2
    model = _model(T)
    g = torch.Generator().manual_seed(4)
    frames = (torch.rand(T, 3, *size, generator=g) * 255).round()
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}]
    outs = {}
    for fused in (False, True):
        nn_ops.set_fused_glue(fused)
        try:
            with oracle_ops(policy):
                x, _ = model.preprocess(frames)
                feats = model.backbone(x) if not fused else model.backbone_from_frames(frames)
                mfeat, _bfe, _enc, ms = model.sem_seg_head.pixel_decoder.forward_features(feats)
                out = model.clip_forward(frames, tg())
            outs[fused] = (feats, mfeat, ms, out)
        finally:
            nn_ops.set_fused_glue(False)
    (f0, m0, s0, o0), (f1, m1, s1, o1) = outs[False], outs[True]
    for k in f0:
        assert f1[k].shape == f0[k].shape and _rel(f1[k], f0[k]) < 2e-5, k
    assert m1.shape == m0.shape and _rel(m1, m0) < 2e-5
    for a, b in zip(s1, s0):
        assert _rel(a, b) < 2e-5
    for k in ("pred_masks", "pred_logits", "pred_embds"):
        assert _rel(o1[k], o0[k]) < 1e-4, k


def test_clip_stream_with_fused_glue():
    from univs_b200.streaming import ClipStream
    T, V = 2, 3
    model = _model(T)
    g = torch.Generator().manual_seed(8)
    video = (torch.rand(V, 3, 60, 90, generator=g) * 255).round()
    mk = lambda s: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}]
    res = {}
    for fused in (False, True):
        nn_ops.set_fused_glue(fused)
        try:
            with oracle_ops():
                res[fused] = {s: o["pred_masks"].clone() for s, o in ClipStream(model, T).run(video, mk)}
        finally:
            nn_ops.set_fused_glue(False)
    for s in res[False]:
        assert _rel(res[True][s], res[False][s]) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# CUDA kernels against the oracle (opt-in until validated on a B200: UNIVS_GPU_GLUE=1)
# ---------------------------------------------------------------------------------------------------------------
_gpu_glue = pytest.mark.filterwarnings("default")      # validated on a B200 (round 2): no gate


def _unsplit(op, fmt, C):
    """GEMM operand -> the fp32 value it represents (hi + lo)."""
    from univs_b200.ops import f16_chunk
    if fmt == "tf32":
        v = op.float().view(*op.shape[:-1], 1, 2, C)
        return (v[..., 0, 0, :] + v[..., 0, 1, :])
    if fmt == "f16u":
        return op[..., :C].float() + op[..., C:].float()
    kc = f16_chunk(C)
    v = op.float().view(*op.shape[:-1], C // kc, 3, kc)
    return (v[..., 2, :] + v[..., 0, :] * 2.0 ** -11).reshape(*op.shape[:-1], C)


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("N,H,W,C,groups", [(2, 23, 37, 256, 32), (1, 184, 320, 256, 32), (3, 8, 12, 64, 8)])
@pytest.mark.parametrize("fmt", [None, "tf32", "f16", "f16u"])
def test_groupnorm_cl_cuda(N, H, W, C, groups, fmt):
    from univs_b200 import ops
    g = torch.Generator().manual_seed(0)
    pad = 1
    # x as a [:, :H, :W] view of a padded convolution output buffer; lowres as a level slice of a token matrix
    buf = torch.randn(N, H + 2, W + 2, C, generator=g) * 2 + 0.5
    h2, w2 = (H + 1) // 2, (W + 1) // 2
    tokens = torch.randn(N, 7 + h2 * w2, C, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    want = ops_ref.groupnorm_cl(buf[:, :H, :W], gamma, beta, groups, 1e-5, tokens[:, 7:].reshape(N, h2, w2, C), relu=True)
    d = "cuda"
    bufd, tokd = buf.to(d), tokens.to(d)
    y, op = ops.groupnorm_cl(bufd[:, :H, :W], gamma.to(d), beta.to(d), groups, 1e-5,
                             lowres=tokd[:, 7:].reshape(N, h2, w2, C), relu=True, want_f32=True, split=fmt,
                             pad=pad if fmt else 0)
    assert _rel(y.cpu(), want) < 2e-5
    if fmt:
        assert op.shape[1:3] == (H + 2 * pad, W + 2 * pad)
        border = op.clone()
        border[:, pad:-pad, pad:-pad] = 0
        assert not border.any()                                   # the zero border is never written
        got = _unsplit(op[:, pad:-pad, pad:-pad], fmt, C).cpu()
        assert _rel(got, want) < (2e-5 if fmt != "f16u" else 1e-4)
    # plain GroupNorm (no top-down add, no ReLU), contiguous input
    y2, _ = ops.groupnorm_cl(bufd[:, :H, :W].contiguous(), gamma.to(d), beta.to(d), groups)
    assert _rel(y2.cpu(), ops_ref.groupnorm_cl(buf[:, :H, :W], gamma, beta, groups)) < 2e-5


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("H,W", [(720, 1280), (61, 95), (64, 96)])
def test_patchify_normalize_cuda(dtype, H, W):
    from univs_b200 import ops
    g = torch.Generator().manual_seed(1)
    frames = (torch.rand(2, 3, H, W, generator=g) * 255).round().to(dtype)
    padded = ((H + 31) // 32 * 32, (W + 31) // 32 * 32)
    want = ops_ref.patchify_normalize(frames, MEAN, STD, padded)
    got = ops.patchify_normalize(frames.cuda(), MEAN, STD, padded)
    assert got.shape == want.shape
    assert (got.cpu() - want).abs().max().item() <= 1e-6
    op = ops.patchify_normalize(frames.cuda(), MEAN, STD, padded, split="f16")
    assert _rel(_unsplit(op, "f16", 48).cpu(), want) < 1e-6


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("N,H,W,C", [(2, 16, 24, 32), (1, 23, 41, 192), (5, 46, 80, 768)])
@pytest.mark.parametrize("fmt", [None, "f16"])
def test_layernorm_merge2x2_cuda(N, H, W, C, fmt):
    from univs_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N, H, W, C, generator=g) * 3 + 1
    gamma, beta = torch.randn(4 * C, generator=g), torch.randn(4 * C, generator=g)
    want = ops_ref.layernorm_merge2x2(x, gamma, beta)
    got = ops.layernorm_merge2x2(x.cuda(), gamma.cuda(), beta.cuda(), split=fmt)
    got = _unsplit(got, fmt, 4 * C) if fmt else got
    assert got.shape == want.shape and _rel(got.cpu(), want) < 2e-5


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("policy", ["fp16x3", "tf32x3", "fp32"])
def test_fused_glue_model_cuda_equals_default(policy):
    from univs_b200 import precision
    T = 2
    model = _model(T).cuda()
    g = torch.Generator().manual_seed(4)
    frames = (torch.rand(T, 3, 60, 90, generator=g) * 255).round()
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}]
    precision.set_precision(policy)
    try:
        outs = {}
        for fused in (False, True):
            nn_ops.set_fused_glue(fused)
            outs[fused] = model.clip_forward(frames.to(torch.uint8 if fused else torch.float32).cuda(), tg())
        for k in ("pred_masks", "pred_logits", "pred_embds"):
            assert _rel(outs[True][k].cpu(), outs[False][k].cpu()) < 1e-3, k
    finally:
        nn_ops.set_fused_glue(False)
        precision.set_precision("fp32")


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("shapes", [[(23, 40), (46, 80), (92, 160)], [(3, 5), (6, 9), (11, 18)]])
@pytest.mark.parametrize("tile", [1, 4, 8, 16, 32])
def test_msda_encoder_tiled_is_bit_identical(shapes, tile):
    """Only the work distribution differs: every (frame, query, head) runs the same instruction sequence."""
    from univs_b200 import ops
    g = torch.Generator().manual_seed(3)
    N, M = 2, 8
    shapes = shapes[::-1]                                   # coarse to fine, as the pixel decoder orders them
    S = sum(h * w for h, w in shapes)
    starts = [0]
    for h, w in shapes[:-1]:
        starts.append(starts[-1] + h * w)
    value = torch.randn(N, S, M, 32, generator=g).cuda()
    ol = torch.randn(N, S, M * 3 * 4 * 3, generator=g).cuda()
    ol[..., : M * 24] *= 3.0                                # offsets reaching across pixels and past the borders
    base = ops.ms_deform_attn_encoder(value, shapes, starts, ol, tile=0)
    tiled = ops.ms_deform_attn_encoder(value, shapes, starts, ol, tile=tile)
    assert torch.equal(base, tiled)


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("fmt", [None, "f16", "tf32"])
def test_msda_encoder_fused_biases_and_operand_cuda(fmt):
    """value_proj / offsets / logits biases folded into the kernel (in-bounds samples only) + operand emission."""
    from univs_b200 import ops
    g = torch.Generator().manual_seed(5)
    N, M = 2, 8
    shapes = [(5, 9), (10, 18), (20, 36)]
    S = sum(h * w for h, w in shapes)
    starts = [0, 45, 45 + 180]
    value = torch.randn(N, S, M, 32, generator=g)
    ol = torch.randn(N, S, M * 36, generator=g)
    ol[..., : M * 24] *= 4.0                       # plenty of samples beyond the borders: bias must not leak into them
    vb, ob = torch.randn(M * 32, generator=g), torch.randn(M * 36, generator=g) * 0.5
    want = ops_ref.ms_deform_attn_fused(value + vb.view(1, 1, M, 32), shapes, starts, ol + ob, M, 3, 4)
    got = ops.ms_deform_attn_encoder(value.cuda(), shapes, starts, ol.cuda(), value_bias=vb.cuda(),
                                     offs_logits_bias=ob.cuda(), split=fmt)
    got = _unsplit(got, fmt, M * 32) if fmt else got
    assert _rel(got.cpu(), want) < 2e-5


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("C,fmt", [(256, "f16"), (256, "tf32"), (192, "f16"), (1024, "f16u")])
def test_layernorm_multi_cuda(C, fmt):
    from univs_b200 import ops
    g = torch.Generator().manual_seed(6)
    N, S = 3, 37
    x, r = torch.randn(N, S, C, generator=g) * 2, torch.randn(N, S, C, generator=g)
    rb, gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g), torch.randn(C, generator=g)
    pos = torch.randn(1, S, C, generator=g)
    want = torch.nn.functional.layer_norm(x + r + rb, (C,), gamma, beta, 1e-5)
    d = "cuda"
    y, op, opp = ops.layernorm_multi(x.to(d), gamma.to(d), beta.to(d), 1e-5, r.to(d), rb.to(d), True, fmt, pos.to(d))
    tol = 2e-5 if fmt != "f16u" else 1e-4
    assert _rel(y.cpu(), want) < 2e-5
    assert _rel(_unsplit(op, fmt, C).cpu(), want) < tol
    assert _rel(_unsplit(opp, fmt, C).cpu(), want + pos) < tol
    y2, op2, opp2 = ops.layernorm_multi(x.to(d), gamma.to(d), beta.to(d), want_f32=False, split=fmt)
    assert y2 is None and opp2 is None
    assert _rel(_unsplit(op2, fmt, C).cpu(), torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)) < tol


# ---------------------------------------------------------------------------------------------------------------
# Pooled-feature attention masks (csrc/decoder_glue.cu, UNIVS_POOLED_MASKS=1)
# ---------------------------------------------------------------------------------------------------------------
def test_pooled_feature_masks_equal_resized_logits_cpu():
    """resize(E.F) == E.resize(F): the oracle's pooled features reproduce the reference's resized-logit decisions, and the
    2x2-centre mean the CUDA kernel computes is torch's bilinear resize for even ratios."""
    torch.manual_seed(2)
    T, Q, C, H, W = 2, 9, 32, 16, 24
    E, Fm = torch.randn(T, Q, C), torch.randn(T, H * W, C)
    logits = ops_ref.mask_einsum(E, Fm.transpose(1, 2))                       # [Q,T,HW]
    for tgt in ((8, 12), (4, 6), (2, 3)):
        want = ops_ref.attn_mask_from_logits(logits, (H, W), tgt)             # reference order: einsum -> resize -> threshold
        pooled = ops_ref.mask_feature_pool(Fm, (H, W), tgt)
        small = ops_ref.mask_einsum(E, pooled.transpose(1, 2))
        got = ops_ref.attn_mask_direct(small)
        resized = torch.nn.functional.interpolate(logits.view(Q, T, H, W), size=tgt, mode="bilinear", align_corners=False)
        near_zero = (resized.permute(1, 0, 2, 3).reshape(T, Q, -1).abs() < 1e-5)
        assert ((got != want) & ~near_zero).sum() == 0
        # closed form used by mask_feature_pool_kernel
        ry, rx = H // tgt[0], W // tgt[1]
        x = Fm.view(T, H, W, C)
        a, b = ry // 2 - 1, rx // 2 - 1
        mean = 0.5 * (0.5 * x[:, a::ry, b::rx] + 0.5 * x[:, a::ry, b + 1::rx]) + 0.5 * (0.5 * x[:, a + 1::ry, b::rx] + 0.5 * x[:, a + 1::ry, b + 1::rx])
        assert _rel(mean.reshape(T, -1, C), pooled) < 1e-6


def test_pooled_masks_decoder_path_equals_default_cpu():
    T = 2
    model = _model(T)
    dec = model.sem_seg_head.predictor
    g = torch.Generator().manual_seed(9)
    frames = (torch.rand(T, 3, 64, 96, generator=g) * 255).round()
    for tg in (lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}],
               lambda: [{"task": "detection", "dataset_name": "bdd_track", "prompt_type": "text"}]):
        outs = {}
        for pooled in (False, True):
            dec.pooled_masks = pooled
            seen = []
            dec.attn_mask_hook = lambda i, bits, ro: (seen.append((bits.clone(), ro.clone())), (bits, ro))[1]
            try:
                with oracle_ops():
                    outs[pooled] = (model.clip_forward(frames, tg()), seen)
            finally:
                dec.pooled_masks = False
                dec.attn_mask_hook = None
        (o0, s0), (o1, s1) = outs[False], outs[True]
        assert len(s0) == len(s1) > 0
        for (b0, r0), (b1, r1) in zip(s0, s1):                 # identical decisions in every intermediate head
            assert torch.equal(b0, b1) and torch.equal(r0, r1)
        for k in ("pred_masks", "pred_logits", "pred_embds"):
            assert o1[k].shape == o0[k].shape and _rel(o1[k], o0[k]) < 1e-5, k


@pytest.mark.gpu
@_gpu_glue
@pytest.mark.parametrize("mode", ["f16x3", "mma3x"])
def test_pooled_mask_kernels_gpu(mode):
    from univs_b200 import ops
    torch.manual_seed(4)
    T, Q, C, H, W = 2, 37, 256, 48, 80
    E, Fm = torch.randn(T, Q, C), torch.randn(T, H * W, C)
    for tgt in ((24, 40), (12, 20), (6, 10)):
        want = ops_ref.mask_feature_pool(Fm, (H, W), tgt)
        got = ops.mask_feature_pool(Fm.cuda(), (H, W), tgt, mode=mode)
        if mode == "f16x3":
            got = got[..., :C].float() + got[..., C:].float()
        assert _rel(got.cpu(), want) < 2e-6
        small = ops_ref.mask_einsum(E, want.transpose(1, 2))
        bits, ro = ops.attn_mask_bits_direct(small.contiguous().cuda())
        m = ops_ref.attn_mask_direct(small).bool()
        assert torch.equal(bits.cpu(), ops.pack_mask_bits(m))
        assert torch.equal(ro.cpu(), (~m.all(-1)).to(torch.int32))
