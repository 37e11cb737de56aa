"""Host-side check of every tensor-level wrapper in univs_b200/ops.py WITHOUT a GPU: the C-ABI library is replaced by a
recorder that validates each call against the ctypes signature table (argument count, pointer / integer / float kinds)
and returns success, device checks are relaxed to CPU tensors, and the wrappers run on small CPU inputs.  This catches
argument-order / count mistakes, which otherwise only show up as a ctypes TypeError on the GPU box.  No arithmetic runs
(the outputs are uninitialised memory): nothing here is a parity claim."""
import ctypes as C

import numpy as np
import pytest
import torch

from univs_b200 import _cabi, ops


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name not in _cabi.SIGNATURES:
            raise AttributeError(name)
        res, argtypes = _cabi.SIGNATURES[name]

        def fn(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} arguments, signature has {len(argtypes)}"
            for i, (a, t) in enumerate(zip(args, argtypes)):
                if t is C.c_void_p:
                    assert a is None or isinstance(a, int), f"{name} arg {i}: pointer expected, got {type(a)}"
                elif t in (C.c_int, C.c_int64):
                    assert isinstance(a, (int, np.integer)) and not isinstance(a, bool), f"{name} arg {i}: int expected, got {type(a)} {a!r}"
                elif t is C.c_float:
                    assert isinstance(a, float), f"{name} arg {i}: float expected, got {type(a)}"
                # POINTER(c_float) arrays etc. are passed through
            self.calls.append(name)
            if res is C.c_int64:
                return 1 << 16
            return 0

        return fn


@pytest.fixture()
def fake(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(ops, "lib", lambda: rec)
    monkeypatch.setattr(ops, "_stream", lambda: 0)

    def chk(t, name, dtype=torch.float32):
        assert t.dtype == dtype, f"{name}: expected {dtype}, got {t.dtype}"
        assert t.is_contiguous(), name
        return t.data_ptr()

    monkeypatch.setattr(ops, "_chk", chk)
    monkeypatch.setattr(ops, "_workspace", lambda nbytes, device: torch.empty(max(int(nbytes), 16), dtype=torch.uint8))
    monkeypatch.setattr(ops, "_gn_workspace", lambda device, nbytes: torch.empty(max(int(nbytes), 16), dtype=torch.uint8))
    return rec


def test_attention_wrappers(fake, monkeypatch):
    monkeypatch.setattr(ops, "_mha_tc", 0)        # explicit calls below reach both kernel families whatever tuned.json promotes
    monkeypatch.setattr(ops, "_win_tc", 0)
    q = torch.zeros(2, 10, 64)
    k = v = torch.zeros(2, 600, 64)
    bits = torch.zeros(2, 10, 19, dtype=torch.int32)
    ro = torch.ones(2, 10, dtype=torch.int32)
    assert ops.mha_core(q, k, v, bits, ro, precision=0).shape == q.shape
    assert ops.mha_core_tc(q, k, v, bits, ro, flags=1).shape == q.shape
    assert ops.mha_core_tc(q, k, v).shape == q.shape
    p = torch.zeros(3, 2, 64)
    assert ops.proca_core(p, p, p, torch.zeros(3, 1, 5, 64), torch.zeros(3, 1, 5, 64)).shape == p.shape
    qkv, b, tab = torch.zeros(1, 24, 36, 192), torch.zeros(192), torch.zeros(529, 2)
    assert ops.swin_window_attention(qkv, b, tab, 2, 12, 6, precision=0).shape == (1, 24, 36, 64)
    assert ops.swin_window_attention_operand(qkv, b, tab, 2, 12, 6).shape == (1, 24, 36, 192)
    out, op = ops.swin_window_attention_tc(qkv, b, tab, 2, 6, want_f32=True, want_operand=True)
    assert out.shape == (1, 24, 36, 64) and op.shape == (1, 24, 36, 192) and op.dtype == torch.float16
    out, op, dbg = ops.swin_window_attention_tc(qkv, b, tab, 2, 0, True, False, 0, True)
    assert op is None and dbg.shape == (2 * 3 * 2, 144, 144)
    assert {"univs_mha_tc_forward_f32", "univs_swin_window_attention_tc", "univs_mha_forward_f32"} <= set(fake.calls)


def test_opt_in_routing(fake, monkeypatch):
    qkv, b, tab = torch.zeros(1, 24, 36, 192), torch.zeros(192), torch.zeros(529, 2)
    monkeypatch.setattr(ops, "_win_tc", 1)
    ops.swin_window_attention(qkv, b, tab, 2, 12, 6, precision=0)
    ops.swin_window_attention_operand(qkv, b, tab, 2, 12, 6)
    assert fake.calls == ["univs_swin_window_attention_tc"] * 2
    ops.swin_window_attention(torch.zeros(1, 14, 21, 96), torch.zeros(96), torch.zeros(169, 1), 1, 7, 3, precision=0)
    assert fake.calls[-1] == "univs_swin_window_attention_f32"             # 7x7 windows stay on the mma.sync kernel
    monkeypatch.setattr(ops, "_mha_tc", 1)
    fake.calls.clear()
    ops.mha_core(torch.zeros(1, 200, 256), torch.zeros(1, 920, 256), torch.zeros(1, 920, 256), precision=0)
    ops.mha_core(torch.zeros(1, 1000, 256), torch.zeros(1, 1000, 256), torch.zeros(1, 1000, 256), precision=0)   # Q*T self-attention
    ops.mha_core(torch.zeros(1, 200, 256), torch.zeros(1, 78, 256), torch.zeros(1, 78, 256), precision=0)        # short memory
    # the self-attention goes to the tcgen05 kernel too (query chunks as the batch dimension); short memories stay on mma.sync
    assert [c for c in fake.calls if "forward" in c] == ["univs_mha_tc_forward_f32", "univs_mha_tc_forward_f32", "univs_mha_forward_f32"]


def test_einsum_and_mask_wrappers(fake):
    T, Q, Cc, H, W = 2, 5, 32, 8, 12
    feats = torch.zeros(T, H * W, Cc)
    e = torch.zeros(T, Q, Cc)
    for mode in ("f16x3", "tf32", "mma3x"):
        fp = ops.prepare_mask_features(feats, mode)
        assert ops.mask_einsum(e, fp, mode=mode).shape == (Q, T, H * W)
        pooled = ops.mask_feature_pool(feats, (H, W), (4, 6), mode=mode)
        assert pooled.shape == (T, 24, 2 * Cc if mode == "f16x3" else Cc)
        assert ops.mask_einsum(e, pooled, mode=mode, tag="mask_einsum_pooled").shape == (Q, T, 24)
    logits = torch.zeros(Q, T, H * W)
    bits, ro = ops.attn_mask_bits(logits, (H, W), (4, 6))
    assert bits.shape == (T, Q, 1) and ro.shape == (T, Q)
    bits, ro = ops.attn_mask_bits_direct(torch.zeros(Q, T, 40))
    assert bits.shape == (T, Q, 2)
    assert ops.mask_einsum_mma(e, feats, 0).shape == (Q, T, H * W)


def test_rowwise_and_glue_wrappers(fake):
    x = torch.zeros(6, 64)
    w = torch.ones(64)
    for fmt in (None, "tf32", "f16", "f16u"):
        s, y = ops.layernorm(x, w, w, 1e-5, residual=x, want_sum=True, split=fmt, residual_bias=w)
        assert ops.gelu(x, fmt, bias=w).shape[0] == 6 and ops.relu(x, fmt).shape[0] == 6
    assert ops.split_operand(x, "f16").shape == (6, 192)
    y, op, opp = ops.layernorm_multi(x, w, w, 1e-5, residual=x, residual_bias=w, want_f32=True, split="f16", pos=torch.zeros(3, 64))
    assert y.shape == x.shape and op.shape == (6, 192) and opp.shape == (6, 192)
    xcl = torch.zeros(2, 5, 7, 64)
    assert ops.layernorm_merge2x2(xcl, torch.ones(256), torch.ones(256), 1e-5, "f16").shape == (2, 3, 4, 768)
    # groupnorm_cl / patchify_normalize check `.is_cuda` inline (they read strides of the live tensor): GPU tests only
    val = torch.zeros(1, 20, 8, 32)
    ol = torch.zeros(1, 20, 288)
    assert ops.ms_deform_attn_encoder(val, [(2, 2), (2, 4), (2, 4)], [0, 4, 12], ol).shape == (1, 20, 256)
    assert ops.ms_deform_attn_encoder(val, [(2, 2), (2, 4), (2, 4)], [0, 4, 12], ol, tile=8, value_bias=torch.zeros(256),
                                      offs_logits_bias=torch.zeros(288), split="f16").shape == (1, 20, 768)
    assert ops.round_tf32(x).shape == x.shape
