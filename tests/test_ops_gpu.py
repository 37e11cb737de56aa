"""Parity tests proper: every CUDA kernel, called through the C ABI, against the CPU oracle (oracle/ops_ref.py) on
seeded inputs.  Tolerances: relative max-error  max|a-b| / max|b|  per tensor; 1e-3 is the north-star
bound, the kernels are held to tighter bounds where their arithmetic allows it."""
import pytest
import torch

from oracle import ops_ref

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.fixture(scope="module")
def ops():
    from univs_b200 import ops
    return ops


def _lsi(shapes):
    out = [0]
    for h, w in shapes[:-1]:
        out.append(out[-1] + h * w)
    return out


# ------------------------------------------------------------------ MSDeformAttn
@pytest.mark.parametrize("shapes,N,M,D,P,Lq", [
    ([(6, 4), (3, 2)], 1, 2, 2, 2, 2),          # the reference's own test shapes (ops/test.py:24-31)
    ([(5, 7), (10, 14), (20, 27)], 2, 8, 32, 4, None),
    ([(23, 40), (46, 80), (92, 160)], 1, 8, 32, 4, 300),
    ([(3, 3)], 2, 3, 5, 3, 7),                    # D not a multiple of 4 -> scalar kernel
    ([(4, 4), (2, 2)], 1, 4, 64, 1, 5),
])
def test_msda_forward_reference_abi(ops, shapes, N, M, D, P, Lq):
    torch.manual_seed(3)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    Lq = S if Lq is None else Lq
    value = torch.randn(N, S, M, D)
    loc = torch.rand(N, Lq, M, L, P, 2) * 1.5 - 0.25       # includes samples outside [0,1]
    w = torch.rand(N, Lq, M, L, P) + 1e-5
    w = w / w.sum((-1, -2), keepdim=True)
    want = ops_ref.ms_deform_attn(value, shapes, _lsi(shapes), loc, w)
    got = ops.ms_deform_attn_forward(value.cuda(), shapes, _lsi(shapes), loc.cuda(), w.cuda())
    assert _rel(got, want) < 2e-5
    # device-resident level tables (the reference passes CUDA int64 tensors)
    sh_d = torch.as_tensor(shapes, dtype=torch.long).cuda()
    ls_d = torch.as_tensor(_lsi(shapes), dtype=torch.long).cuda()
    got2 = ops.ms_deform_attn_forward(value.cuda(), sh_d, ls_d, loc.cuda(), w.cuda())
    assert torch.equal(got, got2)


def test_msda_empty_and_bad_args(ops):
    from univs_b200._cabi import UnivsB200Error
    shapes = [(2, 2)]
    v = torch.randn(1, 4, 8, 32).cuda()
    out = ops.ms_deform_attn_forward(v, shapes, [0], torch.zeros(1, 0, 8, 1, 4, 2).cuda(), torch.zeros(1, 0, 8, 1, 4).cuda())
    assert out.shape == (1, 0, 256)
    with pytest.raises(UnivsB200Error):    # sum(H*W) != S
        ops.ms_deform_attn_forward(v, [(3, 3)], [0], torch.zeros(1, 2, 8, 1, 4, 2).cuda(), torch.zeros(1, 2, 8, 1, 4).cuda())


@pytest.mark.parametrize("shapes", [[(3, 5), (6, 10), (12, 20)], [(15, 27), (30, 54), (60, 108)]])
def test_msda_encoder_fused(ops, shapes):
    torch.manual_seed(5)
    N, M = 2, 8
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, 32)
    ol = torch.cat([torch.randn(N, S, M * 3 * 4 * 2) * 3.0, torch.randn(N, S, M * 12) * 2.0], -1).contiguous()
    want = ops_ref.ms_deform_attn_fused(value, shapes, _lsi(shapes), ol, M, 3, 4)
    got = ops.ms_deform_attn_encoder(value.cuda(), shapes, _lsi(shapes), ol.cuda())
    assert _rel(got, want) < 5e-6


# ------------------------------------------------------------------ Swin window attention
@pytest.mark.parametrize("B,H,W,nH,ws,shift", [
    (2, 8, 12, 2, 4, 0), (2, 8, 12, 2, 4, 2), (1, 7, 10, 3, 4, 2),
    (2, 14, 21, 3, 7, 0), (2, 13, 9, 1, 7, 3), (1, 30, 54, 6, 7, 3),
    (1, 24, 36, 2, 12, 0), (2, 24, 27, 4, 12, 6), (1, 46, 80, 6, 12, 6),
])
@pytest.mark.parametrize("prec,tol", [(0, 2e-5), (1, 2e-3)])
def test_swin_window_attention(ops, B, H, W, nH, ws, shift, prec, tol):
    torch.manual_seed(7)
    C = 32 * nH
    qkv = torch.randn(B, H, W, 3 * C)
    bias = torch.randn(3 * C) * 0.3
    table = torch.randn((2 * ws - 1) ** 2, nH) * 0.5
    want = ops_ref.swin_window_attention(qkv, bias, table, nH, ws, shift)
    got = ops.swin_window_attention(qkv.cuda(), bias.cuda(), table.cuda(), nH, ws, shift, precision=prec)
    assert _rel(got, want) < tol
    if prec == 0:      # strict kernel can emit the fp16x3 GEMM operand [lo*2^11 | hi*2^-11 | hi] directly
        op = ops.swin_window_attention_operand(qkv.cuda(), bias.cuda(), table.cuda(), nH, ws, shift).float()
        rec = op[..., 2 * C:] + op[..., :C] * 2.0 ** -11
        assert _rel(rec, want) < tol
        assert torch.equal(op[..., 2 * C:].half(), got.half())


# ------------------------------------------------------------------ mask einsum + attention-mask bits
@pytest.mark.parametrize("T,Q,C,HW", [(1, 20, 256, 64 * 64), (2, 100, 256, 30 * 54), (3, 200, 256, 46 * 80),
                                        (1, 232, 256, 130), (2, 7, 64, 66), (1, 256, 32, 2), (1, 16, 32, 128)])
def test_mask_einsum(ops, T, Q, C, HW):
    torch.manual_seed(11)
    E = torch.randn(T, Q, C)
    F = torch.randn(T, HW, C)
    want = ops_ref.mask_einsum(E.double(), F.transpose(1, 2).double()).float()
    # 1-pass TF32 on the tcgen05 tensor cores: round-to-nearest operands (2^-12 relative each), fp32 accumulation in TMEM
    Fr = ops.prepare_mask_features(F.cuda(), "tf32")
    got = ops.mask_einsum(E.cuda(), Fr, mode="tf32")
    assert got.shape == (Q, T, HW)
    assert _rel(got, want) < 4e-4
    # the register-operand kernel on the same rounded operands must agree with the tensor-memory kernel
    chk = ops.mask_einsum_mma(ops.round_tf32(E.cuda()), Fr, ops.PREC_TF32)
    assert _rel(got, chk) < 2e-6
    # strict: 3xTF32 register kernel, and (C % 64 == 0) the fp16x3 tcgen05 kernel -- both fp32-equivalent
    strict = ops.mask_einsum(E.cuda(), F.cuda(), mode="mma3x")
    assert _rel(strict, want) < 5e-6
    if C % 64 == 0:
        f16 = ops.mask_einsum(E.cuda(), ops.prepare_mask_features(F.cuda(), "f16x3"), mode="f16x3")
        assert _rel(f16, want) < 5e-6


def test_mask_einsum_more_than_256_queries(ops):
    """Category prompts of a large vocabulary (e.g. lvis: 200 + 1203 queries) exceed one launch's 256 queries."""
    torch.manual_seed(14)
    T, Q, C, HW = 2, 600, 256, 500
    E, F = torch.randn(T, Q, C), torch.randn(T, HW, C)
    want = ops_ref.mask_einsum(E.double(), F.transpose(1, 2).double()).float()
    for mode in ("f16x3", "mma3x"):
        got = ops.mask_einsum(E.cuda(), ops.prepare_mask_features(F.cuda(), mode), mode=mode)
        assert got.shape == (Q, T, HW)
        assert _rel(got, want) < 5e-6


def test_mask_einsum_empty(ops):
    out = ops.mask_einsum(torch.zeros(0, 4, 32, device="cuda"), torch.zeros(0, 8, 32, device="cuda"), mode="mma3x")
    assert out.shape == (4, 0, 8)


def test_mask_einsum_linearity_full_size(ops):
    """Size-independent properties at the north-star size (T=5,Q=200,C=256,184x320): linearity in E, a checksum of
    checksums against an fp64 reduction on the device, and agreement of the kernels with each other."""
    torch.manual_seed(12)
    T, Q, C, HW = 5, 200, 256, 184 * 320
    Fraw = torch.randn(T, HW, C, device="cuda")
    E1 = torch.randn(T, Q, C, device="cuda")
    E2 = torch.randn(T, Q, C, device="cuda")
    for mode, tol in (("f16x3", 2e-5), ("tf32", 1e-3)):
        F = ops.prepare_mask_features(Fraw, mode)
        o1, o2, o12 = ops.mask_einsum(E1, F, mode=mode), ops.mask_einsum(E2, F, mode=mode), ops.mask_einsum(E1 + E2, F, mode=mode)
        assert _rel(o12, o1 + o2) < tol
        colsum = Fraw.double().sum(1)                               # [T,C]
        want = torch.einsum("tqc,tc->qt", E1.double(), colsum)
        got = o1.double().sum(-1)
        assert (got - want).abs().max().item() / want.abs().max().item() < tol
    strict = ops.mask_einsum(E1, Fraw, mode="mma3x")
    assert _rel(ops.mask_einsum(E1, ops.prepare_mask_features(Fraw, "f16x3"), mode="f16x3"), strict) < 5e-6


@pytest.mark.parametrize("Q,T,H,W,tgt", [(5, 2, 16, 24, (8, 12)), (5, 2, 16, 24, (4, 6)), (7, 3, 16, 24, (2, 3)),
                                           (20, 1, 64, 64, (32, 32)), (3, 1, 8, 8, (1, 1))])
def test_attn_mask_bits(ops, Q, T, H, W, tgt):
    torch.manual_seed(13)
    logits = torch.randn(Q, T, H * W)
    logits[1] = logits[1].abs() * -1 - 0.1          # a query blocked everywhere -> row_open == 0
    want = ops_ref.attn_mask_from_logits(logits, (H, W), tgt)                 # [T,Q,S] uint8
    bits, row_open = ops.attn_mask_bits(logits.cuda(), (H, W), tgt)
    S = tgt[0] * tgt[1]
    b = bits.cpu().to(torch.int64) & 0xFFFFFFFF
    unpacked = ((b.unsqueeze(-1) >> torch.arange(32)) & 1).flatten(-2)[..., :S].to(torch.uint8)
    assert torch.equal(unpacked, want)
    assert torch.equal(row_open.cpu().bool(), ~(want.bool().all(-1)))
    assert not row_open.cpu()[:, 1].any()


# ------------------------------------------------------------------ MHA core / ProCA
@pytest.mark.parametrize("B,Lq,Lk,masked", [(3, 11, 37, True), (5, 200, 920, True), (2, 100, 6480, True),
                                             (1, 1000, 1000, True), (1, 300, 130, False), (2, 1, 65, True)])
@pytest.mark.parametrize("prec,tol", [(0, 2e-5), (1, 2e-3)])
def test_mha_core(ops, B, Lq, Lk, masked, prec, tol):
    torch.manual_seed(17)
    C = 256
    q, k, v = torch.randn(B, Lq, C), torch.randn(B, Lk, C), torch.randn(B, Lk, C)
    mask = None
    if masked:
        mask = torch.rand(B, Lq, Lk) < 0.7
        mask[0, min(4, Lq - 1)] = True             # fully blocked row -> un-blocked (..._univs.py:390)
        mask[-1, 0, : Lk // 2] = True
    want = ops_ref.mha_core(q, k, v, 8, None if mask is None else mask.to(torch.uint8), unmask_full_rows=True)
    if masked:
        bits = ops.pack_mask_bits(mask.cuda())
        row_open = (~mask.all(-1)).to(torch.int32).cuda()
        got = ops.mha_core(q.cuda(), k.cuda(), v.cuda(), bits, row_open, precision=prec)
    else:
        got = ops.mha_core(q.cuda(), k.cuda(), v.cuda(), precision=prec)
    assert _rel(got, want) < tol


def test_mha_shared_mask_batch1(ops):
    torch.manual_seed(18)
    q, k, v = torch.randn(1, 50, 256), torch.randn(1, 50, 256), torch.randn(1, 50, 256)
    mask = torch.ones(50, 50, dtype=torch.bool)
    mask[:30, :30] = False
    mask[30:, 30:] = False
    want = ops_ref.mha_core(q, k, v, 8, mask[None].to(torch.uint8))
    got = ops.mha_core(q.cuda(), k.cuda(), v.cuda(), ops.pack_mask_bits(mask[None].cuda()))
    assert _rel(got, want) < 2e-5


@pytest.mark.parametrize("P,T,Tm,L", [(4, 3, 3, 6), (10, 5, 1, 1408), (40, 2, 1, 1), (32, 10, 10, 78), (3, 2, 1, 2)])
def test_proca_core(ops, P, T, Tm, L):
    torch.manual_seed(19)
    C = 256
    q, ks, vs = torch.randn(P, T, C), torch.randn(P, T, C), torch.randn(P, T, C)
    km, vm = torch.randn(P, Tm, L, C), torch.randn(P, Tm, L, C)
    want = ops_ref.proca_core(q, ks, vs, km, vm, 8)
    got = ops.proca_core(q.cuda(), ks.cuda(), vs.cuda(), km.cuda(), vm.cuda())
    assert _rel(got, want) < 1e-5


def test_round_tf32(ops):
    x = torch.randn(1000).cuda()
    y = ops.round_tf32(x)
    assert (y.view(torch.int32) & 0x1FFF).eq(0).all()
    assert (y - x).abs().max() <= x.abs().max() * 2 ** -11


# ------------------------------------------------------------------ fused row-wise kernels
@pytest.mark.parametrize("rows,C", [(7, 32), (1000, 192), (333, 256), (65, 768), (40, 1536), (9, 3072), (3, 640)])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_fused(ops, rows, C, with_res):
    torch.manual_seed(21)
    x = torch.randn(rows, C) * 3 + 0.5
    r = torch.randn(rows, C) if with_res else None
    w, b = torch.randn(C), torch.randn(C)
    rb = torch.randn(C) if with_res else None
    s_want, y_want = ops_ref.layernorm(x, w, b, 1e-5, r, rb)
    s, y = ops.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-5, None if r is None else r.cuda(), want_sum=True,
                         residual_bias=None if rb is None else rb.cuda())
    assert _rel(y, y_want) < 2e-6
    assert _rel(s, s_want) < 1e-6
    _, ys = ops.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-5, None if r is None else r.cuda(), split="tf32",
                          residual_bias=None if rb is None else rb.cuda())
    assert ys.shape == (rows, 2 * C)
    assert torch.equal(ys.cpu(), ops_ref.split_tf32(y.cpu()))            # split of exactly the plain output
    kc = C
    v = ys.view(rows, C // kc, 2, kc)
    assert torch.equal((v[:, :, 0] + v[:, :, 1]).reshape(rows, C), y)      # hi + lo == value, bit-exact
    assert (v[:, :, 0].contiguous().view(torch.int32) & 0x1FFF).eq(0).all()   # hi is TF32-representable


def test_gelu_relu_split(ops):
    torch.manual_seed(22)
    x = torch.randn(257, 768) * 2
    assert _rel(ops.gelu(x.cuda()), ops_ref.gelu(x)) < 1e-6
    bb = torch.randn(768)
    assert _rel(ops.gelu(x.cuda(), bias=bb.cuda()), ops_ref.gelu(x, bb)) < 1e-6
    assert torch.equal(ops.relu(x.cuda(), bias=bb.cuda()).cpu(), torch.relu(x + bb))
    gs = ops.gelu(x.cuda(), split="tf32")
    assert torch.equal(gs.cpu(), ops_ref.split_tf32(ops.gelu(x.cuda()).cpu()))
    assert torch.equal(ops.relu(x.cuda()).cpu(), torch.relu(x))
    assert torch.equal(ops.split_tf32(x.cuda()).cpu(), ops_ref.split_tf32(x))
    assert ops.layernorm(torch.zeros(0, 64).cuda(), torch.ones(64).cuda(), torch.zeros(64).cuda())[1].shape == (0, 64)


def test_split_gemm_reproduces_fp32(ops):
    """The 3xTF32 operand split makes a TF32 tensor-core GEMM fp32-equivalent (nn_ops.linear under tf32x3)."""
    from univs_b200 import nn_ops
    from univs_b200.precision import set_precision
    torch.manual_seed(23)
    x, w, b = torch.randn(4096, 768).cuda(), (torch.randn(384, 768) * 0.05).cuda(), torch.randn(384).cuda()
    want = (x.double() @ w.double().t() + b.double()).float()
    try:
        set_precision("tf32x3")
        y3 = nn_ops.linear(x, w, b)
        set_precision("tf32")
        y1 = nn_ops.linear(x, w, b)
    finally:
        set_precision("fp32")
    assert _rel(y3, want) < 2.5e-6
    assert _rel(y1, want) > 1e-5          # plain TF32 is visibly worse: the split is doing the work


def test_split_conv3x3_reproduces_fp32(ops):
    from univs_b200 import nn_ops
    from univs_b200.precision import set_precision
    torch.manual_seed(24)
    x = torch.randn(2, 20, 27, 256).cuda()
    w = (torch.randn(64, 256, 3, 3) * 0.02).cuda()
    want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1).float()
    try:
        set_precision("tf32x3")
        y = nn_ops.conv2d_cl(x, w, None, padding=1)
    finally:
        set_precision("fp32")
    assert y.shape == want.shape
    assert _rel(y, want) < 2.5e-6


def test_fp16_split_formats(ops):
    torch.manual_seed(25)
    x = torch.randn(300, 384) * torch.logspace(-6, 2, 384)          # magnitudes from 1e-6 to 1e2
    s16 = ops.split_operand(x.cuda(), "f16")                         # chunks [lo*2^11 | hi*2^-11 | hi]
    assert s16.dtype == torch.float16 and s16.shape == (300, 3 * 384)
    lo, hs, hi = s16[:, :384].float(), s16[:, 384:768].float(), s16[:, 768:].float()
    rec = hi + lo * 2.0 ** -11
    assert ((rec.cpu() - x).abs() <= x.abs() * 2.0 ** -21 + 3e-11).all()       # ~22-bit reconstruction
    assert ((hs * 2048.0 - hi).abs() <= 2048 * 3.1e-8).all()                    # hi*2^-11 up to the fp16 subnormal floor
    u16 = ops.split_operand(x.cuda(), "f16u")
    rec = u16[:, :384].float() + u16[:, 384:].float()
    assert ((rec.cpu() - x).abs() <= x.abs() * 2.0 ** -21 + 4e-8).all()        # unscaled lo: fp16 subnormal floor
    big = torch.tensor([[1e6, -1e6, 70000.0, 1.0]]).cuda()
    assert torch.isfinite(ops.split_operand(big, "f16").float()).all()          # saturating, never inf
    wide = torch.randn(8, 3072).cuda()                                           # chunked layout (2 chunks of 1536)
    w3 = ops.split_operand(wide, "f16").view(8, 2, 3, 1536).float()
    assert torch.allclose((w3[:, :, 2] + w3[:, :, 0] * 2.0 ** -11).reshape(8, 3072), wide, rtol=1e-6, atol=1e-9)


def test_fp16x3_gemm_reproduces_fp32(ops):
    from univs_b200 import nn_ops
    from univs_b200.precision import set_precision
    torch.manual_seed(26)
    x, w, b = torch.randn(4096, 768).cuda(), (torch.randn(384, 768) * 0.05).cuda(), torch.randn(384).cuda()
    want = (x.double() @ w.double().t() + b.double()).float()
    try:
        set_precision("fp16x3")
        y = nn_ops.linear(x, w, b)
        yc = nn_ops.conv2d_cl(x[:540].view(2, 10, 27, 768)[..., :256].contiguous(), (w.view(384, 768)[:64, :2304 // 9 * 0 + 256].reshape(64, 256, 1, 1)).repeat(1, 1, 3, 3) * 0.1, None, padding=1)
    finally:
        set_precision("fp32")
    assert y.dtype == torch.float32
    assert _rel(y, want) < 2e-6
    xc = x[:540].view(2, 10, 27, 768)[..., :256]
    wc = (w[:64, :256].reshape(64, 256, 1, 1)).repeat(1, 1, 3, 3) * 0.1
    wantc = torch.nn.functional.conv2d(xc.permute(0, 3, 1, 2).double(), wc.double(), padding=1).permute(0, 2, 3, 1).float()
    assert _rel(yc, wantc) < 2e-6
