"""tcgen05 window attention for 12x12 windows (univs_b200/csrc/swin_window_attn_tc.cu).

CPU part: the kernel replaces table lookups of the reference by closed forms (relative-position index with a per-column
compile-time constant, shift-mask region bits, roll/pad source addressing); the same closed forms are restated here in
Python and checked against the oracle's explicit construction (oracle/ops_ref.py::swin_window_attention, which follows
swin.py:108-121, 247-289, 413-440).
GPU part (opt-in until it has run on a B200: UNIVS_GPU_WINTC=1): parity of scores, output and GEMM-operand output
against the oracle and the mma.sync kernel."""
import os

import pytest
import torch

from oracle import ops_ref

WS, N = 12, 144


# ---- Python restatement of the kernel's closed forms ---------------------------------------------------------------
def _bias_index(row, half, j):
    qy, qx = divmod(row, WS)
    return (qy + 11) * 23 + (qx + 11) - half * 6 * 23 - ((j // WS) * 23 + (j % WS))


def _masked(row, half, j, wy, wx, nWh, nWw, shift):
    qy, qx = divmod(row, WS)
    mh = shift > 0 and wy == nWh - 1
    mw = shift > 0 and wx == nWw - 1
    thr = WS - shift
    low = (1 << thr) - 1
    dh = ((low if qy >= thr else (0xFFF & ~low)) if mh else 0) >> (half * 6)
    dw = (low if qx >= thr else (0xFFF & ~low)) if mw else 0
    return bool(((dh >> (j // WS)) | (dw >> (j % WS))) & 1)


def _source_token(b, wy, wx, i, H, W, Hp, Wp, shift):
    iy, ix = divmod(i, WS)
    hs, ws_ = wy * WS + iy + shift, wx * WS + ix + shift
    if hs >= Hp:
        hs -= Hp
    if ws_ >= Wp:
        ws_ -= Wp
    return (b * H + hs) * W + ws_ if (hs < H and ws_ < W) else -1


def test_bias_index_closed_form():
    ar = torch.arange(WS)
    cy, cx = torch.meshgrid(ar, ar, indexing="ij")
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    idx = (cy[:, None] - cy[None, :] + WS - 1) * (2 * WS - 1) + (cx[:, None] - cx[None, :] + WS - 1)   # swin.py:108-121
    for row in range(N):
        for half in range(2):
            for j in range(72):
                assert _bias_index(row, half, j) == idx[row, half * 72 + j].item()


@pytest.mark.parametrize("H,W,shift", [(24, 36, 6), (46, 80, 6), (23, 40, 6), (30, 25, 3), (24, 24, 0)])
def test_shift_mask_and_source_closed_forms(H, W, shift):
    Hp, Wp = -(-H // WS) * WS, -(-W // WS) * WS
    nWh, nWw = Hp // WS, Wp // WS
    B = 2
    # oracle construction of the region labels (ops_ref / swin.py:413-440)
    if shift > 0:
        lab = torch.zeros(Hp, Wp)
        cnt = 0
        for hs in (slice(0, -WS), slice(-WS, -shift), slice(-shift, None)):
            for wsl in (slice(0, -WS), slice(-WS, -shift), slice(-shift, None)):
                lab[hs, wsl] = cnt
                cnt += 1
        lab = lab.view(nWh, WS, nWw, WS).permute(0, 2, 1, 3).reshape(nWh * nWw, N)
        m = lab[:, None, :] != lab[:, :, None]          # [nW, query, key]
    else:
        m = torch.zeros(nWh * nWw, N, N, dtype=torch.bool)
    for wy in range(nWh):
        for wx in range(nWw):
            got = torch.tensor([[_masked(r, (k // 72), k % 72, wy, wx, nWh, nWw, shift) for k in range(N)] for r in range(N)])
            assert torch.equal(got, m[wy * nWw + wx]), (wy, wx)
    # roll / pad / partition source addressing
    tok = torch.full((B, Hp, Wp), -1, dtype=torch.long)
    tok[:, :H, :W] = torch.arange(B * H * W).view(B, H, W)
    if shift > 0:
        tok = torch.roll(tok, shifts=(-shift, -shift), dims=(1, 2))
    win = tok.view(B, nWh, WS, nWw, WS).permute(0, 1, 3, 2, 4).reshape(B, nWh, nWw, N)
    for b in range(B):
        for wy in range(nWh):
            for wx in range(nWw):
                got = [_source_token(b, wy, wx, i, H, W, Hp, Wp, shift) for i in range(N)]
                assert got == win[b, wy, wx].tolist()


def test_oracle_scores_are_consistent():
    torch.manual_seed(0)
    qkv = torch.randn(1, 13, 24, 3 * 64)
    bias, table = torch.randn(3 * 64) * 0.3, torch.randn(23 * 23, 2) * 0.5
    o, s = ops_ref.swin_window_attention(qkv, bias, table, 2, 12, 6, return_scores=True)
    assert s.shape == (1, 4, 2, N, N) and o.shape == (1, 13, 24, 64)
    assert torch.equal(o, ops_ref.swin_window_attention(qkv, bias, table, 2, 12, 6))


# ---- GPU parity ----------------------------------------------------------------------------------------------------
_gpu_wintc = pytest.mark.filterwarnings("default")      # validated on a B200 (round 2): no gate


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.gpu
@_gpu_wintc
@pytest.mark.parametrize("ver", [0, 1])      # flags bit 0: the second version of the kernel (swin_window_attn_tc2.cu)
@pytest.mark.parametrize("B,H,W,nH,shift", [
    (1, 12, 12, 1, 0),          # one unit
    (1, 24, 36, 2, 0), (2, 24, 27, 4, 6), (1, 46, 80, 6, 6),
    (3, 23, 40, 24, 6),         # more units than SMs: persistent loop, both ring stages, stage reuse
    (1, 92, 160, 12, 6),
])
def test_window_attention_tc(B, H, W, nH, shift, ver):
    from univs_b200 import ops
    torch.manual_seed(7)
    C = 32 * nH
    qkv = torch.randn(B, H, W, 3 * C)
    bias = torch.randn(3 * C) * 0.3
    table = torch.randn(23 * 23, nH) * 0.5
    want, scores = ops_ref.swin_window_attention(qkv, bias, table, nH, 12, shift, return_scores=True)
    out, op, dbg = ops.swin_window_attention_tc(qkv.cuda(), bias.cuda(), table.cuda(), nH, shift, want_f32=True,
                                                want_operand=True, debug_scores=True, flags=ver)
    torch.cuda.synchronize()
    assert _rel(dbg.view(scores.shape), scores) < 2e-5, "QK^T + bias + mask"
    assert _rel(out, want) < 2e-5, "softmax / PV / store"
    opf = op.float()
    assert _rel(opf[..., 2 * C:] + opf[..., :C] * 2.0 ** -11, want) < 2e-5
    assert torch.equal(opf[..., 2 * C:].half(), out.half())
    assert torch.equal(opf[..., C:2 * C], (opf[..., 2 * C:] * 2.0 ** -11).half().float())
    # against the validated mma.sync kernel: same arithmetic contract
    ref = ops.swin_window_attention(qkv.cuda(), bias.cuda(), table.cuda(), nH, 12, shift, precision=0)
    assert _rel(out, ref) < 2e-5
    if ver == 1:     # flags bit 1: the compact operand [hi | lo * 2^11] the own GEMM reads (production variant of the kernel)
        c16 = ops.swin_window_attention_tc(qkv.cuda(), bias.cuda(), table.cuda(), nH, shift, want_f32=False, want_operand=True,
                                           flags=1, compact=True)[1].float()
        assert c16.shape[-1] == 2 * C
        assert torch.equal(c16[..., :C], opf[..., 2 * C:]) and torch.equal(c16[..., C:], opf[..., :C])


# ---- data-flow model of the kernel (CPU) -----------------------------------------------------------------------------
# A byte-level restatement of what swin_window_attn_tc.cu does for ONE unit: the loader's swizzled stores, the operand
# reads the tcgen05 descriptors describe (K-major / MN-major SWIZZLE_64B canonical layouts, cute/atom/mma_traits_sm100.hpp;
# swizzle = XOR of byte-address bits 4-5 with bits 7-8), the TMEM lane/column placement of the two row tiles including the
# zero-row trick of the tail tile, the per-thread softmax column ownership, the P tile stores and the PV step.  It checks
# that these pieces are mutually consistent and add up to the oracle's attention -- the part of the kernel that can be
# wrong without any hardware being involved.
import numpy as np

ROW = 64
Q_TAIL_OFF = 128 * ROW
Q_BYTES = Q_TAIL_OFF + 48 * ROW
K_BYTES = 144 * ROW
OFF_QH, OFF_QL = 0, Q_BYTES
OFF_KH, OFF_KL = 2 * Q_BYTES, 2 * Q_BYTES + K_BYTES
OFF_VH, OFF_VL = OFF_KL + K_BYTES, OFF_KL + 2 * K_BYTES
STAGE = OFF_VL + K_BYTES
P1_ATOM, P0_ATOM = 32 * ROW, 128 * ROW
OFF_P1H, OFF_P1L = STAGE, STAGE + 5 * P1_ATOM
OFF_P0H, OFF_P0L = OFF_P1L + 5 * P1_ATOM, OFF_P1L + 5 * P1_ATOM + 5 * P0_ATOM
SMEM = OFF_P0L + 5 * P0_ATOM + 8192
TAIL_KEYS0 = 64


def _swz_off(row, chunk):
    return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)


def _split16(x):
    hi = np.float16(x)
    lo = np.float16(np.float32(x) - np.float32(hi))
    return hi, lo


class _Smem:
    def __init__(self, rng, nbytes=SMEM):
        self.h = rng.standard_normal(nbytes // 2).astype(np.float16)    # garbage everywhere a store does not reach

    def st(self, byte_addr, v16):
        assert byte_addr % 2 == 0
        self.h[byte_addr // 2] = v16

    def operand(self, start, rows, kstep_elems, mn_major):
        """16-bit operand tile as the MMA sees it through a SWIZZLE_64B descriptor with SBO = 512:
        K-major: [rows (M/N), 16 (K)], element (r, k) at start + r*64 + 2k;  MN-major: [16 (K), rows (N)], element (k, n)
        at start + (k//8)*512 + (k%8)*64 + 2n;  both through the address swizzle."""
        out = np.zeros((rows, kstep_elems), np.float32)
        for r in range(rows):
            for k in range(kstep_elems):
                a = start + ((k // 8) * 512 + (k % 8) * 64 + 2 * r if mn_major else r * 64 + 2 * k)
                a ^= ((a >> 7) & 3) << 4
                out[r, k] = np.float32(self.h[a // 2])
        return out


def _umma(sm, D, col0, a_start, b_start, n, mn_major_b, accumulate):
    """D[0:128, col0:col0+n] (+)= A[128,16] @ B[n,16]^T   (one kind::f16 MMA, fp32 accumulate)"""
    A = sm.operand(a_start, 128, 16, False)
    B = sm.operand(b_start, n, 16, mn_major_b)
    prod = A @ B.T
    with np.errstate(all="ignore"):
        D[:, col0:col0 + n] = D[:, col0:col0 + n] + prod if accumulate else prod


def _issue_qk(sm, D, col0, a_hi, a_lo, b_hi, b_lo, n, accumulate):
    for k in range(2):
        _umma(sm, D, col0, a_lo + 32 * k, b_hi + 32 * k, n, False, accumulate or k > 0)
        _umma(sm, D, col0, a_hi + 32 * k, b_lo + 32 * k, n, False, True)
        _umma(sm, D, col0, a_hi + 32 * k, b_hi + 32 * k, n, False, True)


def test_dataflow_model_one_unit():
    rng = np.random.default_rng(3)
    shift, nWh, nWw, wy, wx = 6, 2, 3, 1, 2          # a corner window of a shifted block: both mask directions active
    q = (rng.standard_normal((144, 32)) * 0.6).astype(np.float32)
    k = rng.standard_normal((144, 32)).astype(np.float32)
    v = rng.standard_normal((144, 32)).astype(np.float32)
    table = (rng.standard_normal(529) * 0.5).astype(np.float32)
    sm = _Smem(rng)
    # ---- loader (store_token + zero rows)
    for hl_off in (OFF_QH, OFF_QL):
        for z in range(2):
            for row in range(16):
                for b in range(0, 64, 2):
                    sm.st(hl_off + Q_TAIL_OFF + (z * 32 + row) * 64 + b, np.float16(0))
    for i in range(144):
        for lane8 in range(8):
            sub = (lane8 & 1) * 8
            off = _swz_off(i, lane8 >> 1) + sub
            qoff = off if i < 128 else Q_TAIL_OFF + _swz_off(16 + (i - 128), lane8 >> 1) + sub
            for e in range(4):
                d = lane8 * 4 + e
                for val, oh, ol, o in ((q[i, d], OFF_QH, OFF_QL, qoff), (k[i, d], OFF_KH, OFF_KL, off), (v[i, d], OFF_VH, OFF_VL, off)):
                    hi, lo = _split16(val)
                    sm.st(oh + o + 2 * e, hi)
                    sm.st(ol + o + 2 * e, lo)
    # ---- issue_scores: TMEM as [128 lanes, 512 columns]
    tmem = rng.standard_normal((128, 512)).astype(np.float32)
    COL_S0, COL_S1, COL_O0, COL_O1 = 0, 160, 320, 352
    _issue_qk(sm, tmem, COL_S0, OFF_QH, OFF_QL, OFF_KH, OFF_KL, 144, False)
    k1 = TAIL_KEYS0 * ROW
    _issue_qk(sm, tmem, COL_S1, OFF_QH + Q_TAIL_OFF, OFF_QL + Q_TAIL_OFF, OFF_KH + k1, OFF_KL + k1, 144 - TAIL_KEYS0, False)
    _issue_qk(sm, tmem, COL_S1, OFF_QH + Q_TAIL_OFF + 16 * ROW, OFF_QL + Q_TAIL_OFF + 16 * ROW, OFF_KH, OFF_KL, TAIL_KEYS0, True)
    # ---- softmax threads
    kidx = [(kk // 12) * 23 + kk % 12 for kk in range(144)]
    sums = {}
    scores = np.full((144, 144), np.nan, np.float32)

    def softmax_thread(tail, quarter, half, lane):
        row = 128 + (lane & 15) if tail else quarter * 32 + lane
        prow = lane if tail else row
        ncol = 80 if tail else 72
        tl = lane if tail else quarter * 32 + lane            # TMEM lane the 32x32b load hands to this thread
        c0 = COL_S1 if tail else COL_S0 + half * 72           # warp-uniform column address
        key0 = half * TAIL_KEYS0 if tail else half * 72
        sc = tmem[tl, c0:c0 + ncol].copy()
        for j in range(ncol):
            if tail and half == 0 and j >= TAIL_KEYS0:
                sc[j] = -np.inf
                continue
            key = key0 + j
            sc[j] += table[_bias_index(row, 0, 0) - kidx[key]] if tail else table[_bias_index(row, half, j)]
            if _masked(row, key // 72, key % 72, wy, wx, nWh, nWw, shift):
                sc[j] += -100.0
            scores[row, key] = sc[j]
        return row, prow, key0, ncol, sc

    parts = {}
    for tail, quarter, half, lane in [(False, qd, h, l) for qd in range(4) for h in range(2) for l in range(32)] + \
                                     [(True, 0, l >> 4, l) for l in range(32)]:
        parts[(tail, quarter, half, lane)] = softmax_thread(tail, quarter, half, lane)
    for key_t, (row, prow, key0, ncol, sc) in parts.items():
        tail, quarter, half, lane = key_t
        partner = (tail, quarter, half ^ 1, lane ^ 16) if tail else (tail, quarter, half ^ 1, lane)
        mx = max(sc.max(), parts[partner][4].max())
        with np.errstate(all="ignore"):
            p = np.exp2((sc - mx) * np.float32(1.4426950408889634)).astype(np.float32)
        sums[key_t] = p.sum()
        ph, pl = (OFF_P1H, OFF_P1L) if tail else (OFF_P0H, OFF_P0L)
        atom = P1_ATOM if tail else P0_ATOM
        swz = (prow >> 1) & 3
        chunk0 = key0 >> 3
        for cc in range(ncol // 8):
            g = chunk0 + cc
            off = (g >> 2) * atom + prow * 64 + (((g & 3) ^ swz) << 4)
            for e in range(8):
                hi, lo = _split16(p[cc * 8 + e])
                sm.st(ph + off + 2 * e, hi)
                sm.st(pl + off + 2 * e, lo)
        if tail:
            for cz in range(8):
                g = (0 if half else 10) + cz
                off = (g >> 2) * atom + prow * 64 + (((g & 3) ^ swz) << 4)
                for e in range(8):
                    sm.st(ph + off + 2 * e, np.float16(0))
                    sm.st(pl + off + 2 * e, np.float16(0))
    # ---- issue_pv
    for tile, (ph, pl, atom, col) in enumerate([(OFF_P0H, OFF_P0L, P0_ATOM, COL_O0), (OFF_P1H, OFF_P1L, P1_ATOM, COL_O1)]):
        for s in range(9):
            aoff = (s >> 1) * atom + (s & 1) * 32
            boff = s * 16 * ROW
            _umma(sm, tmem, col, pl + aoff, OFF_VH + boff, 32, True, s > 0)
            _umma(sm, tmem, col, ph + aoff, OFF_VL + boff, 32, True, True)
            _umma(sm, tmem, col, ph + aoff, OFF_VH + boff, 32, True, True)
    # ---- epilogue
    out = np.zeros((144, 32), np.float32)
    for (tail, quarter, half, lane), (row, prow, key0, ncol, sc) in parts.items():
        if not tail:
            o = tmem[quarter * 32 + lane, COL_O0 + half * 16:COL_O0 + half * 16 + 16]
            total = sums[(tail, quarter, half, lane)] + sums[(tail, quarter, half ^ 1, lane)]
        else:
            full = tmem[lane, COL_O1:COL_O1 + 32] + tmem[lane ^ 16, COL_O1:COL_O1 + 32]
            o = full[half * 16:half * 16 + 16]
            total = sums[(tail, quarter, half, lane)] + sums[(tail, quarter, half ^ 1, lane ^ 16)]
        out[row, half * 16:half * 16 + 16] = o / total
    # ---- oracle for the same unit (swin.py:131-171 with the mask of this window)
    ar = np.arange(12)
    cy, cx = np.repeat(ar, 12), np.tile(ar, 12)
    idx = (cy[:, None] - cy[None, :] + 11) * 23 + (cx[:, None] - cx[None, :] + 11)
    s_ref = q.astype(np.float64) @ k.astype(np.float64).T + table[idx]
    lab_h = np.where(cy >= 12 - shift, 2, 1) if wy == nWh - 1 else np.zeros(144, int)
    lab_w = np.where(cx >= 12 - shift, 2, 1) if wx == nWw - 1 else np.zeros(144, int)
    lab = lab_h * 3 + lab_w
    s_ref = s_ref + np.where(lab[:, None] != lab[None, :], -100.0, 0.0)
    assert not np.isnan(scores).any(), "every (row, key) score is owned by exactly one thread"
    assert np.abs(scores - s_ref).max() / np.abs(s_ref).max() < 2e-6
    p_ref = np.exp(s_ref - s_ref.max(1, keepdims=True))
    o_ref = (p_ref / p_ref.sum(1, keepdims=True)) @ v.astype(np.float64)
    assert np.abs(out - o_ref).max() / np.abs(o_ref).max() < 5e-6
