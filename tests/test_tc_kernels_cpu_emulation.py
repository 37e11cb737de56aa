"""The tcgen05 attention kernels (csrc/swin_window_attn_tc.cu, csrc/mha_tc.cu -- written after the round's GPU budget was
spent, never run on hardware) EXECUTED ON THE CPU: their sources are compiled unchanged by g++ against a CPU implementation
of csrc/tc05.cuh (tests/emu/tc05.cuh: mbarriers with phases and transaction counts, tcgen05.mma reading emulated shared
memory through the matrix descriptors -- start address, stride byte offset, swizzle on the absolute address, K-major and
MN-major operands --, TMEM with the 32-lanes-per-warp access rule, named barriers) and driven through the real
`univs_b200.ops` wrappers, against the oracle.  512 / 448 host threads play the CUDA threads of a CTA, so the kernels'
own warp specialisation, pipelining and barrier protocol run as written (shared memory and TMEM start NaN-filled: using
anything that was not produced first is visible).

This checks the kernels' logic under the modelled hardware semantics (the ones the kernels were written against, taken from
the PTX ISA tables restated in CuTe) -- loader addressing, operand tiles and descriptors, accumulator placement, softmax,
tail-tile handling, split-K partials, the barrier protocol -- not the hardware itself and not performance."""
import os
import shutil

import pytest
import torch

from oracle import cpu_backend, ops_ref
from tests.emu import build_emu
from tests.test_kernels_cpu_emulation import _load, dev, plain
from univs_b200 import _cabi, ops

if shutil.which("g++") is None or not os.path.exists(os.path.join(build_emu.CUDA_INCLUDE, "cuda_runtime.h")):
    pytest.skip("needs g++ and the CUDA headers", allow_module_level=True)


@pytest.fixture(scope="module")
def emu_lib_path():
    return build_emu.build()


def _use(monkeypatch, path, sms, tmp_path):
    """a private copy of the emulated library whose launchers see `sms` SMs (they cache the count): few SMs make the
    persistent CTAs walk many work units each"""
    copy = str(tmp_path / f"libunivs_emu_{sms}.so")
    shutil.copy(path, copy)
    monkeypatch.setenv("UNIVS_EMU_SMS", str(sms))
    monkeypatch.setattr(_cabi, "_lib", _load(copy))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_ws_cache", {})


def _rel(a, b):
    a, b = plain(a).float(), plain(b).float()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)


@pytest.mark.parametrize("ver", [0, 1])      # flags bit 0: the second version of the kernel (two tail warps)
@pytest.mark.parametrize("B,H,W,nH,shift,sms", [
    (1, 24, 24, 2, 0, 148),      # 8 units, one per CTA
    (2, 24, 36, 2, 6, 3),        # 24 units over 3 persistent CTAs: ring reuse, barrier phases beyond the first wrap
    (1, 20, 30, 1, 6, 2),        # padded grid (20x30 -> 24x36): pad tokens carry the bias only; shifted + masked windows
    (1, 12, 12, 3, 0, 1),        # a single window, three heads on one CTA
    (2, 36, 48, 1, 6, 5),        # 24 units over 5 CTAs: a CTA's next window is 5 further (carries into the window row and the frame)
])
def test_window_attention_tc_kernel(monkeypatch, emu_lib_path, tmp_path, B, H, W, nH, shift, sms, ver):
    _use(monkeypatch, emu_lib_path, sms, tmp_path)
    g = torch.Generator().manual_seed(B * 1000 + H + shift)
    C = nH * 32
    qkv = torch.randn(B, H, W, 3 * C, generator=g)
    bias = torch.randn(3 * C, generator=g) * 0.2
    table = torch.randn(23 * 23, nH, generator=g) * 0.5
    out, op = ops.swin_window_attention_tc(dev(qkv), dev(bias), dev(table), nH, shift, want_f32=True, want_operand=True, flags=ver)
    want = ops_ref.swin_window_attention(qkv, bias, table, nH, 12, shift)
    assert _rel(out, want) < 5e-6
    # the fp16x3 operand of the projection GEMM [lo * 2^11 | hi * 2^-11 | hi]
    op = plain(op).float()
    rec = op[..., 2 * C:] + op[..., :C] / 2048.0
    assert _rel(rec, want) < 5e-6
    assert torch.equal(plain(out).half().float(), op[..., 2 * C:])
    # operand-only and fp32-only calls produce the same values
    only16 = ops.swin_window_attention_tc(dev(qkv), dev(bias), dev(table), nH, shift, want_f32=False, want_operand=True, flags=ver)[1]
    assert torch.equal(plain(only16).float(), op)
    if ver == 1:     # flags bit 1: the compact operand [hi | lo * 2^11] of the own GEMM
        c16 = plain(ops.swin_window_attention_tc(dev(qkv), dev(bias), dev(table), nH, shift, want_f32=False, want_operand=True,
                                                 flags=1, compact=True)[1]).float()
        assert c16.shape[-1] == 2 * C
        assert torch.equal(c16[..., :C], op[..., 2 * C:]) and torch.equal(c16[..., C:], op[..., :C])


@pytest.mark.parametrize("ver", [0, 1])
def test_window_attention_tc_score_dump(monkeypatch, emu_lib_path, tmp_path, ver):
    """the staged diagnostic of tests/tools/win_tc_check.py: the debug dump holds the raw scores q.k * scale per unit"""
    _use(monkeypatch, emu_lib_path, 2, tmp_path)
    g = torch.Generator().manual_seed(3)
    nH, C = 1, 32
    qkv, bias, table = torch.randn(1, 12, 24, 3 * C, generator=g), torch.randn(3 * C, generator=g) * 0.2, torch.randn(529, nH, generator=g)
    out, _, dbg = ops.swin_window_attention_tc(dev(qkv), dev(bias), dev(table), nH, 0, debug_scores=True, flags=ver)
    want, scores = ops_ref.swin_window_attention(qkv, bias, table, nH, 12, 0, return_scores=True)
    assert _rel(out, want) < 5e-6
    assert plain(dbg).shape == (2, 144, 144)
    assert _rel(plain(dbg).reshape(scores.shape), scores) < 5e-6


def _mask_case(g, B, Lq, Lk):
    mask = torch.rand(B, Lq, Lk, generator=g) < 0.4
    mask[0, 1] = True                                   # fully blocked row: ignores its mask (row_open = 0)
    mask[0, 2, :-1] = True                              # only the last key open
    words = (Lk + 31) // 32
    padded = torch.zeros(B, Lq, words * 32, dtype=torch.int64)
    padded[..., :Lk] = mask.long()
    bits = (padded.view(B, Lq, words, 32) << torch.arange(32)).sum(-1)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
    row_open = (~mask.all(-1)).to(torch.int32)
    return mask, bits, row_open


@pytest.mark.parametrize("B,Lq,Lk,heads,sms,masked", [
    (1, 40, 300, 2, 148, True),      # one row tile in use, 3 key blocks (the last partial), split over many CTAs + combine
    (2, 200, 520, 1, 2, True),       # both row tiles, 5 key blocks on one CTA each: the 3-stage ring wraps
    (1, 256, 128, 2, 1, False),      # Lq at the limit, one key block, no mask
    (1, 7, 1100, 1, 4, True),        # few queries, 9 key blocks, split-K with uneven splits
])
@pytest.mark.parametrize("flags", [0, 1])        # 1 = transposed-V (K-major) diagnostic variant
def test_cross_attention_tc_kernel(monkeypatch, emu_lib_path, tmp_path, B, Lq, Lk, heads, sms, masked, flags):
    _use(monkeypatch, emu_lib_path, sms, tmp_path)
    g = torch.Generator().manual_seed(Lq + Lk)
    C = heads * 32
    q, k, v = (torch.randn(B, L, C, generator=g) for L in (Lq, Lk, Lk))
    k = k * 1.5
    mask = bits = row_open = None
    if masked:
        mask, bits, row_open = _mask_case(g, B, Lq, Lk)
    got = ops.mha_core_tc(dev(q), dev(k), dev(v), dev(bits), dev(row_open), flags=flags)
    want = ops_ref.mha_core(q, k, v, heads, mask, unmask_full_rows=masked)
    assert _rel(got, want) < 5e-6
    if masked:      # the two special rows, exactly as the reference's rule (..._univs.py:390) treats them
        full = ops_ref.mha_core(q[:1, 1:2], k[:1], v[:1], heads)
        assert _rel(plain(got)[:1, 1:2], full) < 5e-6


@pytest.mark.parametrize("T,Q,C,HW,sms", [
    (1, 20, 64, 300, 4),         # 3 tiles over 2 clusters; queries split 16 / 16 (4 of the second half exist)
    (2, 200, 256, 700, 4),       # north-star query / channel counts: split 112 / 96, E reloaded at the frame boundary
    (3, 40, 128, 129, 2),        # one cluster walks all 6 tiles: every other tile has a single valid pixel, 3 frames
    (1, 256, 64, 128, 148),      # Q at the limit (128 / 128), exactly one tile
    (2, 17, 64, 1000, 148),      # the smallest query count the split accepts (16 + 1), one cluster per tile
])
def test_cluster_einsum_kernel(monkeypatch, emu_lib_path, tmp_path, T, Q, C, HW, sms):
    """csrc/mask_einsum_mc.cu on the emulator: two CTAs of a cluster run concurrently, E halves resident per frame, F stages
    delivered to both CTAs by (emulated) TMA multicast, stage release by multicast commits, contiguous tile ranges"""
    _use(monkeypatch, emu_lib_path, sms, tmp_path)
    chk = ops._chk
    monkeypatch.setattr(ops, "_chk", lambda t, name, dtype=torch.float32: chk(dev(t), name, dtype))
    monkeypatch.setattr(ops, "_einsum_mc", 1)
    g = torch.Generator().manual_seed(T * 100 + Q)
    E, F = torch.randn(T, Q, C, generator=g), torch.randn(T, HW, C, generator=g)
    F16 = cpu_backend._split16(F, False)                    # fp16 [hi | lo], what prepare_mask_features("f16x3") makes
    guard = torch.full((Q + 2, T, HW), 7.0)
    got = ops.mask_einsum(dev(E), dev(F16), out=guard[1:Q + 1], mode="f16x3")
    want = ops_ref.mask_einsum(E.double(), F.transpose(1, 2).double()).float()
    assert not torch.isnan(plain(got)).any()
    assert _rel(got, want) < 5e-6
    assert bool((guard[0] == 7.0).all()) and bool((guard[Q + 1] == 7.0).all())       # nothing written outside [Q, T, HW]


@pytest.mark.parametrize("T,Q,C,HW", [(1, 20, 256, 300), (2, 200, 256, 400), (1, 232, 64, 130), (2, 7, 64, 66)])
def test_calibration_the_hardware_validated_einsum_kernel_runs_on_the_emulator(monkeypatch, emu_lib_path, tmp_path, T, Q, C, HW):
    """csrc/mask_einsum_tc.cu is the tcgen05 kernel that HAS run on the B200 (bit-identical there to the mma.sync kernel, 0.73
    of the HBM roofline): on the emulator it must give the same answers.  This pins the emulated semantics the never-run
    kernels rely on -- TMA boxes with SWIZZLE_64B / SWIZZLE_128B, K-major matrix descriptors with the 32-byte k-step
    advance, kind::f16 and kind::tf32 instruction descriptors, TMEM lane = M row / column = N, 32x32b loads, the mbarrier
    pipeline -- to code whose behaviour on the hardware is known."""
    _use(monkeypatch, emu_lib_path, 3, tmp_path)
    chk = ops._chk
    monkeypatch.setattr(ops, "_chk", lambda t, name, dtype=torch.float32: chk(dev(t), name, dtype))
    monkeypatch.setattr(ops, "_einsum_mc", 0)
    g = torch.Generator().manual_seed(T * 100 + Q)
    E, F = torch.randn(T, Q, C, generator=g), torch.randn(T, HW, C, generator=g)
    want = ops_ref.mask_einsum(E.double(), F.transpose(1, 2).double()).float()
    # fp16 hi|lo operands, three kind::f16 MMAs per k-step, SWIZZLE_64B
    got = ops.mask_einsum(dev(E), dev(cpu_backend._split16(F, False)), mode="f16x3")
    assert not torch.isnan(plain(got)).any() and _rel(got, want) < 5e-6
    # fp32 operands consumed as TF32 (pre-rounded to nearest, as the decoder does), SWIZZLE_128B
    rnd = lambda x: cpu_backend._split(x)[..., :x.shape[-1]].contiguous()
    Er, Fr = rnd(E), rnd(F)
    out = torch.empty(Q, T, HW)
    rc = _cabi.lib().univs_mask_einsum_f32(0, Er.data_ptr(), Fr.data_ptr(), T, Q, C, HW, out.data_ptr())
    assert rc == 0
    exact = ops_ref.mask_einsum(Er.double(), Fr.transpose(1, 2).double()).float()
    assert _rel(out, exact) < 5e-6                           # exact products of the rounded operands, fp32 accumulation
    assert 1e-5 < _rel(out, want) < 2e-3                     # and it IS the one-pass TF32 result, not fp32


@pytest.mark.parametrize("kernel", ["window", "window_v2", "cross_attention", "cluster_einsum", "einsum"])
def test_tc_kernels_under_randomised_scheduling(monkeypatch, emu_lib_path, tmp_path, kernel):
    """UNIVS_EMU_CHAOS: random delays before every barrier operation / MMA issue / TMEM load / TMA copy of every thread, so
    producers, the MMA lane, softmax and epilogue warps overtake each other in unusual orders; results must not change"""
    monkeypatch.setenv("UNIVS_EMU_CHAOS", "300")
    if kernel.startswith("window"):
        test_window_attention_tc_kernel(monkeypatch, emu_lib_path, tmp_path, 1, 24, 36, 1, 6, 1, int(kernel == "window_v2"))
    elif kernel == "cross_attention":
        test_cross_attention_tc_kernel(monkeypatch, emu_lib_path, tmp_path, 1, 150, 700, 1, 1, True, 0)
    elif kernel == "cluster_einsum":
        test_cluster_einsum_kernel(monkeypatch, emu_lib_path, tmp_path, 2, 40, 128, 520, 2)
    else:
        test_calibration_the_hardware_validated_einsum_kernel_runs_on_the_emulator(monkeypatch, emu_lib_path, tmp_path, 2, 40, 64, 520)


def test_cross_attention_tc_kernel_edge_cases(monkeypatch, emu_lib_path, tmp_path):
    """single query / single key / a mask shared by the batch / key counts around the 128-key block and 32-bit word sizes"""
    _use(monkeypatch, emu_lib_path, 2, tmp_path)
    g = torch.Generator().manual_seed(77)
    for B, Lq, Lk, shared in [(1, 1, 1, False), (2, 3, 31, True), (1, 129, 128, False), (2, 5, 129, True), (1, 2, 257, False)]:
        C = 32
        q, k, v = (torch.randn(B, L, C, generator=g) for L in (Lq, Lk, Lk))
        mask, bits, row_open = _mask_case(g, B, max(Lq, 3), Lk)
        mask, bits, row_open = mask[:, :Lq], bits[:, :Lq].contiguous(), row_open[:, :Lq].contiguous()
        if shared:                                   # one mask for every batch element (mask_batch = 1)
            mask, bits, row_open = mask[:1], bits[:1].contiguous(), row_open[:1].contiguous()
        got = ops.mha_core_tc(dev(q), dev(k), dev(v), dev(bits), dev(row_open), flags=0)
        want = ops_ref.mha_core(q, k, v, 1, mask.expand(B, -1, -1), unmask_full_rows=True)
        assert not torch.isnan(plain(got)).any(), (B, Lq, Lk)
        assert _rel(got, want) < 5e-6, (B, Lq, Lk, shared)


@pytest.mark.parametrize("ver", [0, 1])
def test_window_attention_tc_kernel_stage1_heads(monkeypatch, emu_lib_path, tmp_path, ver):
    """Swin-L stage 1 head count (6 heads, C = 192), one shifted window row with padding in x"""
    _use(monkeypatch, emu_lib_path, 3, tmp_path)
    g = torch.Generator().manual_seed(5)
    nH, C = 6, 192
    qkv = torch.randn(1, 12, 20, 3 * C, generator=g)
    bias, table = torch.randn(3 * C, generator=g) * 0.2, torch.randn(529, nH, generator=g) * 0.5
    out, _ = ops.swin_window_attention_tc(dev(qkv), dev(bias), dev(table), nH, 6, flags=ver)
    assert _rel(out, ops_ref.swin_window_attention(qkv, bias, table, nH, 12, 6)) < 5e-6


def test_cross_attention_tc_kernel_north_star_level(monkeypatch, emu_lib_path, tmp_path):
    """the decoder's cross-attention at the north-star size, coarsest level: 200 queries x 920 keys (23 x 40), 8 heads,
    per-frame mask bits shared by the heads"""
    _use(monkeypatch, emu_lib_path, 16, tmp_path)
    g = torch.Generator().manual_seed(920)
    B, Lq, Lk, heads = 1, 200, 920, 8
    q, k, v = (torch.randn(B, L, heads * 32, generator=g) for L in (Lq, Lk, Lk))
    mask, bits, row_open = _mask_case(g, B, Lq, Lk)
    got = ops.mha_core_tc(dev(q), dev(k), dev(v), dev(bits), dev(row_open), flags=0)
    assert _rel(got, ops_ref.mha_core(q, k, v, heads, mask, unmask_full_rows=True)) < 5e-6
