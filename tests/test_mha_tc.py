"""tcgen05 masked cross-attention (univs_b200/csrc/mha_tc.cu).

CPU part: a byte-level data-flow model of one CTA (swizzled operand tiles, descriptor reads, per-thread key ownership,
mask-bit / key-bound handling, running-max merge, split-K partial format + combine) against the oracle attention.
GPU part (opt-in until it has run on a B200: UNIVS_GPU_MHATC=1): parity against the oracle and the mma.sync kernel."""
import os

import numpy as np
import pytest
import torch

from oracle import ops_ref
from tests.test_window_attn_tc import _Smem, _split16, _swz_off, _umma

ROW, TILE, BLK, STAGES = 64, 8192, 128, 3
OFF_QH, OFF_QL, OFF_KV = 0, 2 * TILE, 4 * TILE
STAGE = 4 * TILE
OFF_PH = OFF_KV + STAGES * STAGE
OFF_PL = OFF_PH + 4 * TILE
SMEM = OFF_PL + 4 * TILE + 4096
LOG2E = np.float32(1.4426950408889634)


def _store_row(sm, th, tl, row, vals):
    for d in range(32):
        lane8, e = d // 4, d % 4
        off = _swz_off(row, lane8 >> 1) + (lane8 & 1) * 8 + 2 * e
        hi, lo = _split16(vals[d])
        sm.st(th + off, hi)
        sm.st(tl + off, lo)


def _cta(rng, q, k, v, bits, row_open, kb_begin, kb_end, vt=False):
    """One CTA of mha_tc_kernel for one (batch, head): returns (o [Lq,32] unnormalised, m [Lq], l [Lq])."""
    Lq, Lk = q.shape[0], k.shape[0]
    words = (Lk + 31) // 32
    sm = _Smem(rng, SMEM)
    scale = np.float32(0.17677669529663687)
    for r in range(256):
        _store_row(sm, OFF_QH, OFF_QL, r, q[r] * scale if r < Lq else np.zeros(32, np.float32))
    run = {(tile, trow, half): dict(m=-np.inf, l=np.float32(0), alpha=np.float32(0), o=np.zeros(16, np.float32))
           for tile in range(2) for trow in range(128) for half in range(2)}
    tmem = rng.standard_normal((128, 256)).astype(np.float32)
    for it, kb in enumerate(range(kb_begin, kb_end)):
        st = OFF_KV + (it % STAGES) * STAGE
        for r in range(BLK):
            key = kb * BLK + r
            _store_row(sm, st, st + TILE, r, k[key] if key < Lk else np.zeros(32, np.float32))
            vrow = v[key] if key < Lk else np.zeros(32, np.float32)
            if not vt:
                _store_row(sm, st + 2 * TILE, st + 3 * TILE, r, vrow)
            else:       # store_row_vt: row = dim, 32-key atoms of 32 rows x 64 B
                for d in range(32):
                    o = (r >> 5) * 2048 + _swz_off(d, (r & 31) >> 3) + (r & 7) * 2
                    hi, lo = _split16(vrow[d])
                    sm.st(st + 2 * TILE + o, hi)
                    sm.st(st + 3 * TILE + o, lo)
        for tile in range(2):
            for kk in range(2):
                _umma(sm, tmem, 0, OFF_QL + tile * TILE + 32 * kk, st + 32 * kk, BLK, False, kk > 0)
                _umma(sm, tmem, 0, OFF_QH + tile * TILE + 32 * kk, st + TILE + 32 * kk, BLK, False, True)
                _umma(sm, tmem, 0, OFF_QH + tile * TILE + 32 * kk, st + 32 * kk, BLK, False, True)
            sc = {}
            for trow in range(128):
                row = tile * 128 + trow
                masked_row = bits is not None and row < Lq and (row_open is None or row_open[row] != 0)
                for half in range(2):
                    key0 = kb * BLK + half * 64
                    s = tmem[trow, half * 64:half * 64 + 64].copy()
                    w = [0, 0]
                    if masked_row:
                        wi = key0 >> 5
                        w = [int(bits[row, wi + x]) & 0xFFFFFFFF if wi + x < words else 0 for x in range(2)]
                    valid = Lk - key0
                    if valid < 64:
                        if valid <= 0:
                            w = [0xFFFFFFFF, 0xFFFFFFFF]
                        elif valid < 32:
                            w = [w[0] | (~((1 << valid) - 1) & 0xFFFFFFFF), 0xFFFFFFFF]
                        elif valid == 32:
                            w[1] = 0xFFFFFFFF
                        else:
                            w[1] |= ~((1 << (valid - 32)) - 1) & 0xFFFFFFFF
                    for j in range(64):
                        if (w[j // 32] >> (j % 32)) & 1:
                            s[j] = -np.inf
                    sc[(trow, half)] = s
            for (trow, half), s in sc.items():
                st_ = run[(tile, trow, half)]
                bm = max(s.max(), sc[(trow, half ^ 1)].max())
                m_new = max(st_["m"], bm)
                e = np.float32(0) if m_new == -np.inf else np.float32(m_new)
                st_["alpha"] = np.float32(0) if st_["m"] == -np.inf else np.exp2((np.float32(st_["m"]) - e) * LOG2E)
                st_["m"] = m_new
                with np.errstate(all="ignore"):
                    p = np.exp2((s - e) * LOG2E).astype(np.float32)
                st_["l"] = st_["l"] * st_["alpha"] + p.sum()
                swz = (trow >> 1) & 3
                for cc in range(8):
                    g = half * 8 + cc
                    off = (g >> 2) * TILE + trow * 64 + (((g & 3) ^ swz) << 4)
                    for x in range(8):
                        hi, lo = _split16(p[cc * 8 + x])
                        sm.st(OFF_PH + off + 2 * x, hi)
                        sm.st(OFF_PL + off + 2 * x, lo)
            for ks in range(8):
                aoff = (ks >> 1) * TILE + (ks & 1) * 32
                boff = (ks >> 1) * 2048 + (ks & 1) * 32 if vt else ks * 16 * ROW
                _umma(sm, tmem, 128, OFF_PL + aoff, st + 2 * TILE + boff, 32, not vt, ks > 0)
                _umma(sm, tmem, 128, OFF_PH + aoff, st + 3 * TILE + boff, 32, not vt, True)
                _umma(sm, tmem, 128, OFF_PH + aoff, st + 2 * TILE + boff, 32, not vt, True)
            for trow in range(128):
                for half in range(2):
                    st_ = run[(tile, trow, half)]
                    st_["o"] = st_["o"] * st_["alpha"] + tmem[trow, 128 + half * 16:128 + half * 16 + 16]
    o, m, l = np.zeros((Lq, 32), np.float32), np.zeros(Lq, np.float32), np.zeros(Lq, np.float32)
    for row in range(Lq):
        tile, trow = divmod(row, 128)
        for half in range(2):
            o[row, half * 16:half * 16 + 16] = run[(tile, trow, half)]["o"]
        m[row] = run[(tile, trow, 0)]["m"]
        assert run[(tile, trow, 0)]["m"] == run[(tile, trow, 1)]["m"]
        l[row] = run[(tile, trow, 0)]["l"] + run[(tile, trow, 1)]["l"]
    return o, m, l


@pytest.mark.parametrize("nsplit,vt", [(1, False), (2, False), (1, True)])
def test_dataflow_model_one_head(nsplit, vt):
    rng = np.random.default_rng(5)
    Lq, Lk = 150, 300                      # two row tiles (second partly empty), three key blocks (last: 44 keys)
    q = rng.standard_normal((Lq, 32)).astype(np.float32)
    k = rng.standard_normal((Lk, 32)).astype(np.float32)
    v = rng.standard_normal((Lk, 32)).astype(np.float32)
    mask = rng.random((Lq, Lk)) < 0.6
    mask[3] = True                          # fully blocked row, re-opened by the row flag (..._univs.py:390)
    mask[140, :256] = True                  # everything blocked except the last block
    mask[7, 128:] = True                    # nothing visible after the first block
    from univs_b200.ops import pack_mask_bits
    bits = pack_mask_bits(torch.from_numpy(mask)[None])[0].numpy()
    row_open = (~mask.all(1)).astype(np.int32)
    nkb = (Lk + BLK - 1) // BLK
    bps = (nkb + nsplit - 1) // nsplit
    parts = [_cta(rng, q, k, v, bits, row_open, s * bps, min(nkb, (s + 1) * bps), vt) for s in range(nsplit)]
    # combine (mha_combine_kernel)
    M = np.max([p[1] for p in parts], 0)
    acc, l = np.zeros((Lq, 32)), np.zeros(Lq)
    for o_, m_, l_ in parts:
        with np.errstate(all="ignore"):
            w = np.where(m_ == -np.inf, 0.0, np.exp(m_.astype(np.float64) - M))
        acc += w[:, None] * o_
        l += w * l_
    out = acc / l[:, None]
    want = ops_ref.mha_core(torch.from_numpy(q)[None], torch.from_numpy(k)[None], torch.from_numpy(v)[None], 1,
                            torch.from_numpy(mask)[None], unmask_full_rows=True)[0].numpy()
    assert np.abs(out - want).max() / np.abs(want).max() < 5e-6


# ---- GPU parity ----------------------------------------------------------------------------------------------------
_gpu_mhatc = pytest.mark.filterwarnings("default")      # validated on a B200 (round 2): no gate


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.gpu
@_gpu_mhatc
@pytest.mark.parametrize("flags", [0, 1])       # 1: transposed-V diagnostic variant
@pytest.mark.parametrize("B,Lq,Lk,C,masked", [
    (1, 20, 96, 64, False), (1, 7, 128, 32, True), (2, 200, 920, 256, True), (3, 256, 130, 256, True),
    (5, 200, 3680, 256, True), (1, 232, 14720, 256, True), (5, 200, 14720, 256, False), (40, 100, 777, 256, True),
])
def test_mha_tc(B, Lq, Lk, C, masked, flags):
    from univs_b200 import ops
    torch.manual_seed(B * 1000 + Lq)
    q, k, v = torch.randn(B, Lq, C), torch.randn(B, Lk, C), torch.randn(B, Lk, C)
    mask = bits = row_open = None
    if masked:
        mask = torch.rand(B, Lq, Lk) < 0.7
        mask[:, 0] = True                      # fully blocked row -> un-blocked by the row flag
        if Lk > 128:
            mask[:, 1, :128] = True
            mask[:, 2, 128:] = True
        bits = ops.pack_mask_bits(mask).cuda()
        row_open = (~mask.all(-1)).to(torch.int32).cuda()
    want = ops_ref.mha_core(q, k, v, C // 32, None if mask is None else mask.to(torch.uint8), unmask_full_rows=masked)
    got = ops.mha_core_tc(q.cuda(), k.cuda(), v.cuda(), bits, row_open, flags=flags)
    torch.cuda.synchronize()
    assert _rel(got, want) < 2e-5
    ref = ops.mha_core(q.cuda(), k.cuda(), v.cuda(), bits, row_open, precision=0)
    assert _rel(got, ref) < 2e-5


@pytest.mark.gpu
@_gpu_mhatc
@pytest.mark.parametrize("Lq,masked", [(1000, False), (1000, True), (1160, True), (600, True)])
def test_self_attention_shape_through_query_chunks(Lq, masked, monkeypatch):
    """the decoder's Q*T self-attention (one batch element, Lq = Q*T query tokens = keys, transformer_layers.py:34-44) on
    the tcgen05 kernel: ops.mha_core cuts the queries into <= 256-row chunks that become its batch dimension"""
    from univs_b200 import ops
    monkeypatch.setattr(ops, "_mha_tc", 1)
    torch.manual_seed(Lq)
    C = 256
    q, k, v = torch.randn(1, Lq, C), torch.randn(1, Lq, C), torch.randn(1, Lq, C)
    mask = bits = row_open = None
    if masked:
        mask = torch.rand(1, Lq, Lq) < 0.6
        mask[:, 3] = True
        bits = ops.pack_mask_bits(mask).cuda()
        row_open = (~mask.all(-1)).to(torch.int32).cuda()
    want = ops_ref.mha_core(q, k, v, 8, None if mask is None else mask.to(torch.uint8), unmask_full_rows=masked)
    calls = []
    real = ops.mha_core_tc
    monkeypatch.setattr(ops, "mha_core_tc", lambda *a, **kw: (calls.append(a[0].shape), real(*a, **kw))[1])
    got = ops.mha_core(q.cuda(), k.cuda(), v.cuda(), bits, row_open)
    torch.cuda.synchronize()
    assert len(calls) == 1 and calls[0][1] <= 256 and calls[0][0] * calls[0][1] >= Lq
    assert _rel(got, want) < 2e-5
