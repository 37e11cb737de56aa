"""csrc/gemm_tc.cu (tcgen05 dense layer with fused epilogue) EXECUTED ON THE CPU emulator of tests/emu (the kernel source
compiled by g++ against the CPU implementation of tc05.cuh: TMA boxes with zero fill and SWIZZLE_128B, K-major descriptors,
in-order asynchronous tensor pipe, TMEM, mbarrier protocol), driven through the real ops wrapper, against the oracle
restatement (oracle/ops_ref.py::gemm_f16x3) and against the fp64 product of the original fp32 values."""
import os
import shutil

import pytest
import torch

from oracle import cpu_backend, ops_ref
from tests.emu import build_emu
from tests.test_kernels_cpu_emulation import _Dev, _load, dev, plain
from univs_b200 import _cabi, ops

if shutil.which("g++") is None or not os.path.exists(os.path.join(build_emu.CUDA_INCLUDE, "cuda_runtime.h")):
    pytest.skip("needs g++ and the CUDA headers", allow_module_level=True)


@pytest.fixture(scope="module")
def emu_lib_path():
    return build_emu.build()


def _use(monkeypatch, path, sms, tmp_path):
    copy = str(tmp_path / f"libunivs_emu_{sms}.so")
    shutil.copy(path, copy)
    monkeypatch.setenv("UNIVS_EMU_SMS", str(sms))
    monkeypatch.setattr(_cabi, "_lib", _load(copy))
    monkeypatch.setattr(ops, "_stream", lambda: 0)


def _view(t):
    return torch.Tensor._make_subclass(_Dev, t)        # keeps the strides (row views of wider containers)


def _rel(a, b):
    a, b = plain(a).double(), plain(b).double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)


def _chunk3(x):
    """the [lo*2^11 | hi*2^-11 | hi] K-chunk container the row-wise kernels write (split = "f16")"""
    return cpu_backend._maybe_split(x, "f16")


@pytest.mark.parametrize("M,N,K,sms,act", [
    (200, 192, 128, 148, 0),      # 2 token tiles x 2 channel tiles (second one half empty), one CTA per tile
    (130, 72, 96, 148, 1),        # channel tile mostly empty, K tail (96 = 64 + 32: zero-filled half box), GELU
    (300, 260, 192, 2, 2),        # 9 tiles on 2 persistent CTAs: ring wrap-around, both accumulator stages, ReLU
    (128, 128, 64, 1, 0),         # exactly one full tile, one k-block
    (70, 71, 64, 1, 1),           # odd channel count: the last channel of the operand output is stored alone
])
def test_gemm_tc_kernel_chunk3_operands(monkeypatch, emu_lib_path, tmp_path, M, N, K, sms, act):
    _use(monkeypatch, emu_lib_path, sms, tmp_path)
    g = torch.Generator().manual_seed(M + N + K)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    s = 8                                                    # weight scale 2^s, undone by alpha
    x3, w3 = _chunk3(x), _chunk3(w * 2.0 ** s)
    kc = ops.f16_chunk(K)
    assert kc == K
    offs = (2 * K, 0)
    y, y16 = ops.gemm_f16x3_tc(dev(x3), offs, dev(w3), offs, K, 2.0 ** -s, dev(bias), dev(add), want_f32=True,
                               want_operand=True, act=act)
    wy, wy16 = ops_ref.gemm_f16x3(x3, offs, w3, offs, K, 2.0 ** -s, bias, add, act)
    assert _rel(y, wy) < 2e-6
    # against the exact product of the original values: the two-term split carries ~22 bits
    ref = x.double() @ w.double().t() + bias.double()
    ref = torch.nn.functional.gelu(ref) if act == 1 else (torch.relu(ref) if act == 2 else ref)
    assert _rel(y, ref + add.double()) < 5e-6
    # operand output: compact [hi | lo*2^11]; the value it carries equals the fp32 output to ~2^-22
    y16 = plain(y16).float()
    assert y16.shape == (M, 2 * N)
    assert _rel(y16[:, :N] + y16[:, N:] / 2048.0, y) < 1e-6
    assert torch.equal(y16[:, :N], plain(y).half().float())


@pytest.mark.parametrize("M,N,K,sms", [
    (200, 256, 128, 4),       # 2 token tiles x 1 channel pair on 2 clusters
    (300, 260, 192, 2),       # 3 channel tiles: the second CTA of the last pair repeats tile 2 and stores nothing; one cluster
    (130, 520, 320, 6),       # 5 channel tiles, K tail, ring wrap-around on 3 clusters
])
def test_gemm_tc_cta_pair_variant(monkeypatch, emu_lib_path, tmp_path, M, N, K, sms):
    """UNIVS_GEMM_PAIR=1: two CTAs of a cluster run one tcgen05.mma.cta_group::2 (M = 256) per product -- TMA boxes of both CTAs
    credited to the leader's barrier, multicast commits, remote accumulator release; results equal the one-CTA kernel's bit for
    bit (same products in the same order per accumulator element)."""
    _use(monkeypatch, emu_lib_path, sms, tmp_path)
    g = torch.Generator().manual_seed(M * N + K)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    x3, w3 = _chunk3(x), _chunk3(w * 256.0)
    offs = (2 * K, 0)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("UNIVS_GEMM_PAIR", mode)
        y, y16 = ops.gemm_f16x3_tc(dev(x3), offs, dev(w3), offs, K, 2.0 ** -8, dev(bias), dev(add), want_f32=True,
                                   want_operand=True, act=1)
        outs[mode] = (plain(y).clone(), plain(y16).clone())
    assert torch.equal(outs["0"][0], outs["1"][0]) and torch.equal(outs["0"][1], outs["1"][1])
    wy, _ = ops_ref.gemm_f16x3(x3, offs, w3, offs, K, 2.0 ** -8, bias, add, 1)
    assert _rel(outs["1"][0], wy) < 2e-6


def test_gemm_tc_compact_operand_k_slices_and_row_views(monkeypatch, emu_lib_path, tmp_path):
    """fc1 -> GELU -> operand, then fc2 over the compact container in two K slices accumulated in place, on row views"""
    _use(monkeypatch, emu_lib_path, 3, tmp_path)
    g = torch.Generator().manual_seed(7)
    M, C, Hd = 150, 64, 256
    x, w1, w2 = torch.randn(M + 20, C, generator=g), torch.randn(Hd, C, generator=g) * 0.1, torch.randn(C, Hd, generator=g) * 0.1
    b1 = torch.randn(Hd, generator=g) * 0.1
    x3, w13, w23 = _chunk3(x), _chunk3(w1 * 16.0), _chunk3(w2 * 16.0)
    xv = x3[20:]                                             # a row view (the 3x3 convolution taps address rows like this)
    _, h16 = ops.gemm_f16x3_tc(_view(xv), (2 * C, 0), dev(w13), (2 * C, 0), C, 1 / 16.0, dev(b1), None, want_f32=False,
                               want_operand=True, act=ops.ACT_GELU)
    hid = torch.nn.functional.gelu(x[20:].double() @ w1.double().t() + b1.double())
    h16p = plain(h16).float()
    assert _rel(h16p[:, :Hd] + h16p[:, Hd:] / 2048.0, hid) < 5e-6
    out = torch.zeros(M, C)
    half = Hd // 2
    for i, k0 in enumerate((0, half)):                       # weights: chunk3 container of the full K, sliced by offsets
        ops.gemm_f16x3_tc(_view(h16), (k0, Hd + k0), dev(w23), (2 * Hd + k0, k0), half, 1 / 16.0, None, dev(out) if i else None,
                          out=dev(out), want_f32=True)
    want = hid @ w2.double().t()
    assert _rel(out, want) < 5e-6


@pytest.mark.parametrize("pair", ["0", "1"])
def test_gemm_tc_shifted_row_taps(monkeypatch, emu_lib_path, tmp_path, pair):
    """univs_gemm_f16x3_tc_taps: y[m] = sum_t x[m + off_t] w_t^T in one accumulation (the 3x3 convolution over a padded
    channel-last activation), in two tap groups accumulated through `addend`, one-CTA and CTA-pair kernels"""
    _use(monkeypatch, emu_lib_path, 4, tmp_path)
    monkeypatch.setenv("UNIVS_GEMM_PAIR", pair)
    g = torch.Generator().manual_seed(11)
    Cin, Cout, Wp, rows_total = 64, 200, 9, 300
    taps = [(t // 3) * Wp + (t % 3) for t in range(9)]
    R = rows_total - taps[-1]
    x = torch.randn(rows_total, Cin, generator=g)
    w = torch.randn(Cout, 9, Cin, generator=g) * 0.05           # tap-major columns
    bias = torch.randn(Cout, generator=g)
    xc = cpu_backend._maybe_split(x, "f16c")                    # compact [hi | lo']
    wc = cpu_backend._maybe_split(w.reshape(Cout, 9 * Cin) * 64.0, "f16c")
    out = torch.zeros(R, Cout)
    for g0, g1 in ((0, 5), (5, 9)):
        ops.gemm_f16x3_tc(dev(xc), (0, Cin), dev(wc), (g0 * Cin, 9 * Cin + g0 * Cin), Cin, 1 / 64.0, dev(bias) if g0 == 0 else None,
                          dev(out) if g0 else None, out=dev(out), tap_rows=taps[g0:g1], rows=R)
    want = sum(x[o:o + R].double() @ w[:, t].double().t() for t, o in enumerate(taps)) + bias.double()
    assert _rel(out, want) < 5e-6
    wy = ops_ref.gemm_f16x3(xc, (0, Cin), wc, (0, 9 * Cin), Cin, 1 / 64.0, bias, None, 0, tap_rows=taps, rows=R)[0]
    assert _rel(out, wy) < 2e-6
    # rows beyond the end of x16 read as zeros: ask for all rows_total output rows
    full, _ = ops.gemm_f16x3_tc(dev(xc), (0, Cin), dev(wc), (0, 9 * Cin), Cin, 1 / 64.0, None, None, tap_rows=taps[:4], rows=rows_total)
    wf = ops_ref.gemm_f16x3(xc, (0, Cin), wc, (0, 9 * Cin), Cin, 1 / 64.0, None, None, 0, tap_rows=taps[:4], rows=rows_total)[0]
    assert _rel(full, wf) < 2e-6


def test_gemm_tc_argument_validation(monkeypatch, emu_lib_path, tmp_path):
    _use(monkeypatch, emu_lib_path, 1, tmp_path)
    x3, w3 = torch.zeros(8, 3 * 64, dtype=torch.float16), torch.zeros(8, 3 * 64, dtype=torch.float16)
    with pytest.raises(_cabi.UnivsB200Error, match="exceed the row pitch"):
        ops.gemm_f16x3_tc(dev(x3), (160, 0), dev(w3), (128, 0), 64)
    with pytest.raises(_cabi.UnivsB200Error, match="multiples of 8"):
        ops.gemm_f16x3_tc(dev(x3), (4, 0), dev(w3), (128, 0), 64)
    y, _ = ops.gemm_f16x3_tc(dev(x3[:0]), (128, 0), dev(w3), (128, 0), 64)      # empty token matrix
    assert y.shape == (0, 8)
