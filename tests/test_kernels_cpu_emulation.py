"""The plain CUDA kernels of univs_b200/csrc EXECUTED ON THE CPU (tests/emu: the kernel sources compiled unchanged by g++,
every CUDA thread a host thread, warp shuffles / ballots / __syncthreads as barriers) and driven through the real
`univs_b200.ops` wrappers, against the oracle.  Covers the kernels that were written after the round's GPU budget was
spent and have never run on hardware (fused glue: channel-last GroupNorm + FPN add, frame ingest, PatchMerging
gather-LayerNorm, multi-consumer LayerNorm, pooled mask features, tiled MSDeformAttn with fused biases; the 8-wide
GELU / split and wide-store LayerNorm, incl. their BIT equality with the validated kernels) and, as a check of the
emulator itself, kernels that were validated on the B200 (LayerNorm, GELU, split, MSDeformAttn encoder).
Checks index arithmetic, predication, reductions and output formats -- not performance, not the hardware; kernels on
tcgen05 / TMA / mma.sync are out of the emulator's reach."""
import ctypes
import os
import shutil

import pytest
import torch

from oracle import cpu_backend, ops_ref
from tests.emu import build_emu
from univs_b200 import _cabi, ops

if shutil.which("g++") is None or not os.path.exists(os.path.join(build_emu.CUDA_INCLUDE, "cuda_runtime.h")):
    pytest.skip("needs g++ and the CUDA headers", allow_module_level=True)


class _Dev(torch.Tensor):
    """a CPU tensor that reports is_cuda: the wrappers only read metadata and data_ptr()"""

    @staticmethod
    def __new__(cls, t):
        return torch.Tensor._make_subclass(cls, t.contiguous() if not isinstance(t, _Dev) else t)

    is_cuda = property(lambda self: True)


def dev(t):
    return None if t is None else _Dev(t)


def plain(t):
    return None if t is None else torch.Tensor._make_subclass(torch.Tensor, t)


def _load(path):
    lib = ctypes.CDLL(path)
    for name, (res, args) in _cabi.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


@pytest.fixture(scope="module")
def emu_lib_path():
    return build_emu.build()


@pytest.fixture()
def emu(monkeypatch, emu_lib_path):
    monkeypatch.setattr(_cabi, "_lib", _load(emu_lib_path))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_gn_ws", {})
    monkeypatch.setattr(ops, "_ws_cache", {})
    chk = ops._chk            # tensors the wrappers allocate themselves are plain CPU tensors: let them through as well
    monkeypatch.setattr(ops, "_chk", lambda t, name, dtype=torch.float32: chk(torch.Tensor._make_subclass(_Dev, t), name, dtype))
    monkeypatch.setattr(ops, "_on_device", lambda t: True)
    yield ops


def _close(a, b, tol=2e-6):
    a, b = plain(a).float(), plain(b).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)
    assert err <= tol, err


def _same_operand(a, b, fmt):
    """operands in the fp16 formats: a value on a rounding boundary may land on either side, so compare what the GEMM
    reconstructs (hi + lo) to fp32 accuracy; tf32 split: values directly"""
    a, b = plain(a), plain(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    if fmt == "f16u":
        C = a.shape[-1] // 2
        _close(a[..., :C].float() + a[..., C:].float(), b[..., :C].float() + b[..., C:].float(), 4e-6)
    elif fmt == "f16c":                                 # compact [hi | lo * 2^11]
        C = a.shape[-1] // 2
        _close(a[..., :C].float() + a[..., C:].float() / 2048.0, b[..., :C].float() + b[..., C:].float() / 2048.0, 4e-6)
    elif fmt == "f16":
        C = a.shape[-1] // 3
        kc = ops.f16_chunk(C)
        ra = a.float().reshape(*a.shape[:-1], C // kc, 3, kc)
        rb = b.float().reshape(*b.shape[:-1], C // kc, 3, kc)
        _close(ra[..., 2, :] + ra[..., 0, :] / 2048.0, rb[..., 2, :] + rb[..., 0, :] / 2048.0, 4e-6)
        _close(ra[..., 1, :], rb[..., 1, :], 1e-3)       # hi * 2^-11 (lands in the fp16 subnormals for tiny hi)
    elif not fmt:
        _close(a, b, 4e-6)
    else:                                               # fp32 [hi | lo] (tf32 split, chunk = C)
        C = a.shape[-1] // 2
        _close(a[..., :C] + a[..., C:], b[..., :C] + b[..., C:], 4e-6)
        assert torch.equal(plain(a[..., :C]).view(torch.int32) & 0x1FFF, torch.zeros_like(plain(a[..., :C]).view(torch.int32)))


# ------------------------------------------------------------------ validated kernels: a check of the emulator itself
@pytest.mark.parametrize("rows,C", [(37, 192), (5, 48), (9, 1024)])
@pytest.mark.parametrize("fmt", [None, "tf32", "f16", "f16u", "f16c"])
def test_validated_rowwise_kernels_on_the_emulator(emu, rows, C, fmt):
    g = torch.Generator().manual_seed(rows + C)
    x, r = torch.randn(rows, C, generator=g) * 2, torch.randn(rows, C, generator=g)
    w, b, rb = torch.randn(C, generator=g), torch.randn(C, generator=g), torch.randn(C, generator=g)
    s, y = emu.layernorm(dev(x), dev(w), dev(b), 1e-5, dev(r), True, fmt, dev(rb))
    ws, wy = cpu_backend._layernorm(x, w, b, 1e-5, r, True, fmt, rb)
    _close(s, ws)
    _same_operand(y, wy, fmt)
    if fmt:
        _same_operand(emu.gelu(dev(x), split=fmt, bias=dev(rb)), cpu_backend._PATCH["gelu"](x, fmt, rb), fmt)
        _same_operand(emu.relu(dev(x), split=fmt), cpu_backend._PATCH["relu"](x, fmt), fmt)
        _same_operand(emu.split_operand(dev(x), fmt), cpu_backend._maybe_split(x, fmt), fmt)


def _msda_case(seed=0, N=2):
    g = torch.Generator().manual_seed(seed)
    shapes = [(6, 8), (3, 4), (2, 2)]
    starts = [0, 48, 60]
    S, M, L, P = 64, 8, 3, 4
    value = torch.randn(N, S, M, 32, generator=g)
    offs_logits = torch.randn(N, S, M * L * P * 3, generator=g)
    offs_logits[0, 0, :16] = 40.0                       # far outside the maps: zero-padding branch
    return value, shapes, starts, offs_logits, (g, S, M, L, P)


def test_msda_encoder_validated_kernel_on_the_emulator(emu):
    value, shapes, starts, ol, _ = _msda_case()
    got = emu.ms_deform_attn_encoder(dev(value), shapes, starts, dev(ol), 3, 4, tile=0)
    _close(got, ops_ref.ms_deform_attn_fused(value, shapes, starts, ol, 8, 3, 4), 2e-5)


# ------------------------------------------------------------------ never-run kernels
@pytest.mark.parametrize("tile", [1, 4, 8, 32])
def test_msda_tiled_kernel_is_bit_identical_to_the_untiled_one(emu, tile):
    value, shapes, starts, ol, _ = _msda_case(1)
    base = emu.ms_deform_attn_encoder(dev(value), shapes, starts, dev(ol), 3, 4, tile=0)
    got = emu.ms_deform_attn_encoder(dev(value), shapes, starts, dev(ol), 3, 4, tile=tile)
    assert torch.equal(plain(got), plain(base))


@pytest.mark.parametrize("fmt", [None, "f16", "tf32"])
def test_msda_tiled_kernel_with_fused_biases_and_operand_output(emu, fmt):
    value, shapes, starts, ol, (g, S, M, L, P) = _msda_case(2)
    vb, ob = torch.randn(M * 32, generator=g), torch.randn(M * L * P * 3, generator=g) * 0.3
    got = emu.ms_deform_attn_encoder(dev(value), shapes, starts, dev(ol), L, P, tile=8, value_bias=dev(vb),
                                     offs_logits_bias=dev(ob), split=fmt)
    want = cpu_backend._msda_enc(value, shapes, starts, ol, L, P, tile=8, value_bias=vb, offs_logits_bias=ob, split=fmt)
    if fmt:
        _same_operand(got, want, fmt)
    else:
        _close(got, want, 2e-5)


@pytest.mark.parametrize("case", ["plain", "lowres_relu", "operand_pad", "strided_rows"])
def test_groupnorm_channel_last(emu, case):
    g = torch.Generator().manual_seed(3)
    N, H, W, C, G = 2, 6, 10, 64, 8
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    x = torch.randn(N, H, W, C, generator=g) * 2 + 0.5
    if case == "plain":
        y, op = emu.groupnorm_cl(dev(x), dev(w), dev(b), G)
        _close(y, cpu_backend._groupnorm_cl(x, w, b, G)[0], 4e-6)
        assert op is None
    elif case == "lowres_relu":
        low = torch.randn(N, 3, 5, C, generator=g)
        y, _ = emu.groupnorm_cl(dev(x), dev(w), dev(b), G, lowres=dev(low), relu=True)
        _close(y, cpu_backend._groupnorm_cl(x, w, b, G, lowres=low, relu=True)[0], 4e-6)
        # a lowres map of another aspect ratio (different x / y scales)
        low = torch.randn(N, 2, 5, C, generator=g)
        _close(emu.groupnorm_cl(dev(x), dev(w), dev(b), G, lowres=dev(low))[0],
               cpu_backend._groupnorm_cl(x, w, b, G, lowres=low)[0], 4e-6)
    elif case == "operand_pad":
        for fmt in ("f16", "tf32"):
            y, op = emu.groupnorm_cl(dev(x), dev(w), dev(b), G, relu=True, want_f32=False, split=fmt, pad=1)
            wy, wop = cpu_backend._groupnorm_cl(x, w, b, G, relu=True, want_f32=False, split=fmt, pad=1)
            assert y is None and wy is None
            _same_operand(op, wop, fmt)
            assert not plain(op)[:, 0].any() and not plain(op)[:, :, -1].any()          # the zero border is untouched
    else:   # the convolution's output: rows and images strided inside a padded buffer
        buf = torch.randn(N, H + 2, W + 2, C, generator=g)
        view = buf[:, :H, :W]
        y, _ = emu.groupnorm_cl(torch.Tensor._make_subclass(_Dev, view), dev(w), dev(b), G)      # keep the strides
        _close(y, cpu_backend._groupnorm_cl(view.contiguous(), w, b, G)[0], 4e-6)


@pytest.mark.parametrize("dtype,H,W", [(torch.uint8, 50, 75), (torch.float32, 64, 96), (torch.uint8, 33, 62)])
@pytest.mark.parametrize("fmt", [None, "f16"])
def test_patchify_normalize(emu, dtype, H, W, fmt):
    g = torch.Generator().manual_seed(4)
    frames = (torch.rand(2, 3, H, W, generator=g) * 255).round().to(dtype)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    padded = ((H + 31) // 32 * 32, (W + 31) // 32 * 32)
    got = emu.patchify_normalize(dev(frames), mean, std, padded, 4, fmt)
    want = cpu_backend._PATCH["patchify_normalize"](frames, mean, std, padded, 4, fmt)
    if fmt:
        _same_operand(got, want, fmt)
    else:
        assert torch.equal(plain(got), want)            # IEEE subtract / divide: exact agreement


@pytest.mark.parametrize("H,W,C", [(6, 8, 32), (7, 9, 48), (5, 4, 192)])
@pytest.mark.parametrize("fmt", [None, "f16"])
def test_layernorm_merge2x2(emu, H, W, C, fmt):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, H, W, C, generator=g)
    w, b = torch.randn(4 * C, generator=g), torch.randn(4 * C, generator=g)
    got = emu.layernorm_merge2x2(dev(x), dev(w), dev(b), 1e-5, fmt)
    want = cpu_backend._PATCH["layernorm_merge2x2"](x, w, b, 1e-5, fmt)
    if fmt:
        _same_operand(got, want, fmt)
    else:
        _close(got, want, 4e-6)


@pytest.mark.parametrize("fmt", ["f16", "tf32"])
def test_layernorm_multi(emu, fmt):
    g = torch.Generator().manual_seed(6)
    N, S, C = 2, 21, 256
    x, r = torch.randn(N, S, C, generator=g), torch.randn(N, S, C, generator=g)
    w, b, rb = torch.randn(C, generator=g), torch.randn(C, generator=g), torch.randn(C, generator=g)
    pos = torch.randn(S, C, generator=g)
    y, op, opp = emu.layernorm_multi(dev(x), dev(w), dev(b), 1e-5, dev(r), dev(rb), True, fmt, dev(pos), True)
    wy, wop, wopp = cpu_backend._layernorm_multi(x, w, b, 1e-5, r, rb, True, fmt, pos, True)
    _close(y, wy, 4e-6)
    _same_operand(op, wop, fmt)
    _same_operand(opp, wopp, fmt)
    y, op, opp = emu.layernorm_multi(dev(x), dev(w), dev(b), 1e-5, None, None, True, fmt, None, False)
    assert op is None and opp is None
    _close(y, cpu_backend._layernorm_multi(x, w, b, 1e-5)[0], 4e-6)


def test_pooled_mask_features_and_direct_bits(emu):
    g = torch.Generator().manual_seed(7)
    T, H, W, C = 2, 8, 12, 32
    feats = torch.randn(T, H * W, C, generator=g)
    for target in ((4, 6), (2, 3), (1, 1)):
        _close(emu.mask_feature_pool(dev(feats), (H, W), target, mode="mma3x"), ops_ref.mask_feature_pool(feats, (H, W), target), 1e-6)
        got = emu.mask_feature_pool(dev(feats), (H, W), target, mode="f16x3")
        _same_operand(got, cpu_backend._split16(ops_ref.mask_feature_pool(feats, (H, W), target), False), "f16u")
    Q, S = 5, 70                                        # 70 keys: a partial last word
    logits = torch.randn(Q, T, S, generator=g)
    logits[1, 0] = -3.0                                 # a fully blocked row: row_open must stay 0
    logits[2, 1, 64:] = 5.0
    bits, row_open = emu.attn_mask_bits_direct(dev(logits))
    blocked = ops_ref.attn_mask_direct(logits).bool()   # [T, Q, S], True = blocked
    got = cpu_backend.unpack_bits(plain(bits), S)
    assert torch.equal(got.bool(), blocked)
    assert torch.equal(plain(row_open) != 0, ~blocked.all(-1))
    assert row_open[0, 1] == 0


def test_rowwise_v2_kernels_are_bit_identical_on_the_emulator(emu_lib_path, monkeypatch, tmp_path):
    """UNIVS_ROWWISE_V2 is read once per library instance: a second copy of the emulated library runs the v2 kernels
    (bit 0: 8-wide GELU / ReLU / split, bit 1: wide-store LayerNorm, bit 2: streaming LayerNorm of the plain compact-operand
    case); outputs must equal the validated kernels' bytes"""
    v2_path = str(tmp_path / "libunivs_emu_v2.so")
    shutil.copy(emu_lib_path, v2_path)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    g = torch.Generator().manual_seed(8)
    cases = []
    for rows, C in [(1, 8), (37, 48), (19, 192), (6, 768), (3, 2048), (83, 384)]:      # 83 rows: several rows per warp in the streaming kernel
        x = torch.randn(rows, C, generator=g) * 3
        x[0, 0] = 70000.0                               # saturates the fp16 hi part
        cases.append((x, torch.randn(C, generator=g), torch.randn(C, generator=g), torch.randn(C, generator=g),
                      torch.randn(rows, C, generator=g)))

    def run():
        out = []
        for x, b, w, bb, r in cases:
            for fmt in ("f16", "f16u", "f16c"):
                out.append(plain(ops.gelu(dev(x), split=fmt, bias=dev(b))))
                out.append(plain(ops.relu(dev(x), split=fmt)))
                out.append(plain(ops.split_operand(dev(x), fmt)))
                s, y = ops.layernorm(dev(x), dev(w), dev(bb), 1e-5, dev(r), True, fmt, dev(b))
                out += [plain(s), plain(y), plain(ops.layernorm(dev(x), dev(w), dev(bb), 1e-5, None, False, fmt, None)[1])]
        return out

    monkeypatch.setenv("UNIVS_ROWWISE_V2", "0")
    monkeypatch.setattr(_cabi, "_lib", _load(emu_lib_path))
    base = run()
    for bits in ("3", "7"):
        path = str(tmp_path / f"libunivs_emu_v{bits}.so")
        shutil.copy(emu_lib_path, path)
        monkeypatch.setenv("UNIVS_ROWWISE_V2", bits)
        monkeypatch.setenv("UNIVS_EMU_SMS", "2")          # few resident blocks: the streaming kernel's warps walk several rows
        monkeypatch.setattr(_cabi, "_lib", _load(path))
        v2 = run()
        assert len(base) == len(v2) == 108
        for a, b in zip(base, v2):
            assert a.dtype == b.dtype and a.shape == b.shape
            assert torch.equal(a.contiguous().view(torch.int16), b.contiguous().view(torch.int16)), bits


# ------------------------------------------------------------------ the mma.sync / cp.async kernels (validated on the B200)
# Their warp-level MMAs run on the emulator too (fragments exchanged between the 32 host threads of a warp, PTX fragment
# layouts): a CPU regression harness for the default path, and the second calibration of the emulator against code whose
# behaviour on the hardware is known.
@pytest.mark.parametrize("window,H,W,nH,shift", [(7, 14, 21, 2, 3), (7, 10, 9, 1, 0), (12, 24, 20, 1, 6), (4, 8, 8, 3, 2)])
@pytest.mark.parametrize("precision", [ops.PREC_TF32X3, ops.PREC_TF32])
def test_default_window_attention_kernels_on_the_emulator(emu, window, H, W, nH, shift, precision):
    g = torch.Generator().manual_seed(window * 100 + H)
    C = nH * 32
    qkv = torch.randn(1, H, W, 3 * C, generator=g)
    bias, table = torch.randn(3 * C, generator=g) * 0.2, torch.randn((2 * window - 1) ** 2, nH, generator=g) * 0.5
    want = ops_ref.swin_window_attention(qkv, bias, table, nH, window, shift)
    got = emu.swin_window_attention(dev(qkv), dev(bias), dev(table), nH, window, shift, precision)
    _close(got, want, 2e-5 if precision == ops.PREC_TF32X3 else 4e-3)
    if precision == ops.PREC_TF32X3:
        op = plain(emu.swin_window_attention_operand(dev(qkv), dev(bias), dev(table), nH, window, shift)).float()
        _close(op[..., 2 * C:] + op[..., :C] / 2048.0, want, 2e-5)


@pytest.mark.parametrize("precision", [ops.PREC_TF32X3, ops.PREC_TF32])
def test_default_attention_core_and_einsum_kernels_on_the_emulator(emu, precision):
    from tests.test_tc_kernels_cpu_emulation import _mask_case
    g = torch.Generator().manual_seed(31)
    B, Lq, Lk, heads = 2, 37, 150, 2
    q, k, v = (torch.randn(B, L, heads * 32, generator=g) for L in (Lq, Lk, Lk))
    mask, bits, row_open = _mask_case(g, B, Lq, Lk)
    got = emu.mha_core(dev(q), dev(k), dev(v), dev(bits), dev(row_open), precision)
    want = ops_ref.mha_core(q, k, v, heads, mask, unmask_full_rows=True)
    _close(got, want, 2e-5 if precision == ops.PREC_TF32X3 else 4e-3)
    E, F = torch.randn(2, 20, 64, generator=g), torch.randn(2, 130, 64, generator=g)
    got = emu.mask_einsum_mma(dev(E), dev(F), precision)
    _close(got, ops_ref.mask_einsum(E, F.transpose(1, 2)), 5e-6 if precision == ops.PREC_TF32X3 else 2e-3)


def test_default_decoder_glue_kernels_on_the_emulator(emu):
    g = torch.Generator().manual_seed(32)
    Q, T, H, W = 5, 2, 8, 12
    logits = torch.randn(Q, T, H * W, generator=g)
    logits[1, 0] = -2.0
    bits, row_open = emu.attn_mask_bits(dev(logits), (H, W), (4, 6))
    want = ops_ref.attn_mask_from_logits(logits, (H, W), (4, 6)).bool()
    assert torch.equal(cpu_backend.unpack_bits(plain(bits), 24).bool(), want)
    assert torch.equal(plain(row_open) != 0, ~want.all(-1))
    P, L, C = 3, 5, 64
    qq, ks, vs = (torch.randn(P, T, C, generator=g) for _ in range(3))
    km, vm = torch.randn(P, T, L, C, generator=g), torch.randn(P, T, L, C, generator=g)
    _close(emu.proca_core(dev(qq), dev(ks), dev(vs), dev(km), dev(vm)), ops_ref.proca_core(qq, ks, vs, km, vm, 2), 2e-5)
    x = torch.randn(7, 33, generator=g)
    assert torch.equal(plain(emu.round_tf32(dev(x))), cpu_backend._split(x)[..., :33].contiguous())


def test_query_chunking_beyond_256_queries_on_the_emulator(emu):
    """the body of tests/test_ops_gpu.py::test_mask_einsum_more_than_256_queries (added after the last GPU run of round 1),
    on the emulator: 600 queries run as 256 + 256 + 88, tcgen05 and register kernels"""
    torch.manual_seed(14)
    T, Q, C, HW = 1, 600, 64, 260
    E, F = torch.randn(T, Q, C), torch.randn(T, HW, C)
    want = ops_ref.mask_einsum(E.double(), F.transpose(1, 2).double()).float()
    for mode in ("f16x3", "mma3x"):
        feats = emu.prepare_mask_features(dev(F), mode)
        got = emu.mask_einsum(dev(E), dev(feats), mode=mode)
        assert plain(got).shape == (Q, T, HW)
        _close(got, want, 5e-6)


def test_graft_entry_smoke_runs_on_the_emulator(emu, monkeypatch):
    """__graft_entry__.smoke() -- what the driver runs on cuda:0 before the bench -- with `.cuda()` handing out emulator-backed
    tensors: the tcgen05 einsum, mask bits, the masked attention core, window attention and MSDeformAttn, each against the
    oracle inside smoke() itself"""
    import __graft_entry__ as entry
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: dev(self))
    monkeypatch.setattr(torch.Tensor, "cpu", lambda self, *a, **k: plain(self))
    entry.smoke(clip=False)       # the clip-forward part builds a model on a real device
