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
_gpu_wintc = pytest.mark.skipif(os.environ.get("UNIVS_GPU_WINTC") != "1",
                                reason="opt-in (UNIVS_GPU_WINTC=1): tcgen05 window attention not yet validated on a B200")


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.gpu
@_gpu_wintc
@pytest.mark.parametrize("B,H,W,nH,shift", [
    (1, 12, 12, 1, 0),          # one unit
    (1, 24, 36, 2, 0), (2, 24, 27, 4, 6), (1, 46, 80, 6, 6),
    (3, 23, 40, 24, 6),         # more units than SMs: persistent loop, both ring stages, stage reuse
    (1, 92, 160, 12, 6),
])
def test_window_attention_tc(B, H, W, nH, shift):
    from univs_b200 import ops
    torch.manual_seed(7)
    C = 32 * nH
    qkv = torch.randn(B, H, W, 3 * C)
    bias = torch.randn(3 * C) * 0.3
    table = torch.randn(23 * 23, nH) * 0.5
    want, scores = ops_ref.swin_window_attention(qkv, bias, table, nH, 12, shift, return_scores=True)
    out, op, dbg = ops.swin_window_attention_tc(qkv.cuda(), bias.cuda(), table.cuda(), nH, shift, want_f32=True,
                                                want_operand=True, debug_scores=True)
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
