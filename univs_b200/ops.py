"""Tensor-level wrappers of the hot-path kernels.  Inputs must be contiguous fp32 CUDA tensors; each wrapper
allocates the output with torch (device memory plumbing), passes raw pointers and torch's current CUDA stream
through the C ABI and returns immediately (stream-ordered, no synchronisation).  No CPU path exists."""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _cabi, switches
from ._cabi import check, lib

PREC_TF32X3 = 0
PREC_TF32 = 1

_default_precision = PREC_TF32X3

# ---- optional instrumentation (bench.py): launch counter + per-kernel CUDA-event brackets --------------------
launch_count = 0          # number of univs_b200 kernels launched through this module
_event_sink = None        # None, or dict name -> list[(start_event, end_event)]
flop_count = {}           # while the event brackets are on: name -> algorithmic flops (2*M*N*K per dense-layer launch)


def profile_events(enable: bool):
    """When enabled every wrapper brackets its launch(es) with CUDA events on the current stream."""
    global _event_sink
    _event_sink = {} if enable else None
    flop_count.clear()
    return _event_sink


class _Bracket:
    __slots__ = ("name", "n", "ev")

    def __init__(self, name, n=1):
        self.name, self.n = name, n

    def __enter__(self):
        global launch_count
        launch_count += self.n
        if _event_sink is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if _event_sink is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _event_sink.setdefault(self.name, []).append((self.ev, e1))
        return False


_einsum_mode = "mma3x"     # "f16x3" (tcgen05, fp16 hi|lo operands) | "tf32" (tcgen05, 1-pass) | "mma3x" (register 3xTF32)


def set_einsum_mode(m: str):
    global _einsum_mode
    assert m in ("f16x3", "tf32", "mma3x")
    _einsum_mode = m


def set_attention_precision(p: int):
    global _default_precision
    assert p in (PREC_TF32X3, PREC_TF32)
    _default_precision = p


def _stream():
    if not torch.cuda.is_available():
        raise _cabi.UnivsB200Error("no CUDA device: the hot-path operators have no CPU fallback")
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise _cabi.UnivsB200Error(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise _cabi.UnivsB200Error(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _cabi.UnivsB200Error(f"{name}: expected a contiguous tensor")
    return t.data_ptr()


def _chk_t(t: torch.Tensor, name: str, dtype=torch.float32):
    _chk(t, name, dtype)
    return t


def _levels(spatial_shapes, level_start_index):
    sh = np.ascontiguousarray(np.asarray(spatial_shapes, dtype=np.int64).reshape(-1, 2))
    ls = np.ascontiguousarray(np.asarray(level_start_index, dtype=np.int64).reshape(-1))
    return sh, ls


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    """value [N,S,M,D], sampling_loc [N,Lq,M,L,P,2], attn_weight [N,Lq,M,L,P] -> [N,Lq,M*D]
    (MSDA.ms_deform_attn_forward, reference ops/src/vision.cpp:18-21).  spatial_shapes / level_start_index may
    be CUDA int64 tensors (reference ABI) or host sequences."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    out = torch.empty((N, Lq, M * D), device=value.device, dtype=torch.float32)
    if torch.is_tensor(spatial_shapes) and spatial_shapes.is_cuda:
        shp, lsp = _chk(spatial_shapes, "spatial_shapes", torch.int64), _chk(level_start_index, "level_start_index", torch.int64)
        keep = None
    else:
        sh, ls = _levels(torch.as_tensor(spatial_shapes).cpu().numpy() if torch.is_tensor(spatial_shapes) else spatial_shapes,
                         torch.as_tensor(level_start_index).cpu().numpy() if torch.is_tensor(level_start_index) else level_start_index)
        shp, lsp, keep = sh.ctypes.data, ls.ctypes.data, (sh, ls)
    with _Bracket("ms_deform_attn_forward", 1):
        rc = lib().univs_ms_deform_attn_forward_f32(_stream(), _chk(value, "value"), shp, lsp, _chk(sampling_loc, "sampling_loc"),
                                                _chk(attn_weight, "attn_weight"), N, S, M, D, L, Lq, P, out.data_ptr())
    check(rc, "ms_deform_attn_forward")
    del keep
    return out


# Work distribution of the encoder MSDeformAttn kernel: 0 = 4 consecutive queries x 8 heads per CTA (validated in round 1),
# w in {1,..,32} = w x (32/w) query tiles of one head per CTA (bit-identical results, better L1 reuse; opt-in until it
# has been run on a B200).
_msda_tile = switches.get("MSDA_TILE")


def set_msda_tile(width: int):
    global _msda_tile
    _msda_tile = int(width)


def ms_deform_attn_encoder(value, spatial_shapes, level_start_index, offs_logits, num_levels=3, num_points=4, tile=None,
                           value_bias=None, offs_logits_bias=None, split=None):
    """value [N,S,M,32]; offs_logits [N,S,M*L*P*3] raw linear outputs -> [N,S,M*32] (fp32, or a GEMM operand if `split`).
    value_bias / offs_logits_bias: biases of the producing linears, folded into the kernel (tiled variant only)."""
    N, S, M, D = value.shape
    assert D == 32 and offs_logits.shape == (N, S, M * num_levels * num_points * 3)
    sh, ls = _levels(spatial_shapes, level_start_index)
    tile = _msda_tile if tile is None else tile
    fused = value_bias is not None or offs_logits_bias is not None or bool(split)
    if fused and not tile:
        tile = 8
    if tile:
        code, mult, dt = _split_code(split, M * D)
        out = torch.empty((N, S, mult * M * D), device=value.device, dtype=dt)
        with _Bracket("ms_deform_attn_encoder", 1):
            rc = lib().univs_ms_deform_attn_encoder_tiled_f32(
                _stream(), _chk(value, "value"), sh.ctypes.data, ls.ctypes.data, _chk(offs_logits, "offs_logits"), N, S, M,
                num_levels, num_points, tile, None if value_bias is None else _chk(value_bias, "value_bias"),
                None if offs_logits_bias is None else _chk(offs_logits_bias, "offs_logits_bias"), code, out.data_ptr())
        check(rc, "ms_deform_attn_encoder_tiled")
        return out
    out = torch.empty((N, S, M * D), device=value.device, dtype=torch.float32)
    with _Bracket("ms_deform_attn_encoder", 1):
        rc = lib().univs_ms_deform_attn_encoder_f32(_stream(), _chk(value, "value"), sh.ctypes.data, ls.ctypes.data,
                                                _chk(offs_logits, "offs_logits"), N, S, M, num_levels, num_points,
                                                out.data_ptr())
    check(rc, "ms_deform_attn_encoder")
    return out


_win_tc = switches.get("WIN_TC")   # 1 = tcgen05 kernel for 12x12 windows, 2 = its second version (two tail warps; flags bit 0)


def swin_window_attention_tc(qkv, qkv_bias, rel_bias_table, num_heads, shift, want_f32=True, want_operand=False,
                             flags=None, debug_scores=False, compact=False):
    """12x12-window attention on the tcgen05 tensor cores (univs_swin_window_attention_tc).  Returns
    (out f32 [B,H,W,C] | None, operand f16 [B,H,W,3C] | None[, scores f32 [units,144,144]]); with `compact` (second
    kernel version only) the operand is [B,H,W,2C] = [hi | lo*2^11], the format the own GEMM reads."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if flags is None:
        flags = 1 if _win_tc >= 2 else 0
    compact = bool(compact) and bool(flags & 1)
    if compact:
        flags |= 2
    out = torch.empty((B, H, W, C), device=qkv.device, dtype=torch.float32) if want_f32 else None
    op = torch.empty((B, H, W, (2 if compact else 3) * C), device=qkv.device, dtype=torch.float16) if want_operand else None
    dbg = None
    if debug_scores:
        units = B * (-(-H // 12)) * (-(-W // 12)) * num_heads
        dbg = torch.zeros((units, 144, 144), device=qkv.device, dtype=torch.float32)
    with _Bracket("swin_window_attention", 1):
        rc = lib().univs_swin_window_attention_tc(_stream(), _chk(qkv, "qkv"), _chk(qkv_bias, "qkv_bias"),
                                                  _chk(rel_bias_table, "rel_bias_table"), B, H, W, C, num_heads, 12,
                                                  shift, flags, None if out is None else out.data_ptr(),
                                                  None if op is None else op.data_ptr(),
                                                  None if dbg is None else dbg.data_ptr())
    check(rc, "swin_window_attention_tc")
    return (out, op, dbg) if debug_scores else (out, op)


def swin_window_attention(qkv, qkv_bias, rel_bias_table, num_heads, window, shift, precision=None):
    """qkv [B,H,W,3C] (bias-free GEMM output; the kernel adds qkv_bias) -> [B,H,W,C] (see include/univs_b200.h)"""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if _win_tc and window == 12 and (_default_precision if precision is None else precision) == PREC_TF32X3:
        return swin_window_attention_tc(qkv, qkv_bias, rel_bias_table, num_heads, shift)[0]
    out = torch.empty((B, H, W, C), device=qkv.device, dtype=torch.float32)
    with _Bracket("swin_window_attention", 1):
        rc = lib().univs_swin_window_attention_f32(_stream(), _chk(qkv, "qkv"), _chk(qkv_bias, "qkv_bias"),
                                               _chk(rel_bias_table, "rel_bias_table"), B, H, W, C, num_heads, window,
                                               shift, _default_precision if precision is None else precision,
                                               out.data_ptr())
    check(rc, "swin_window_attention")
    return out


def swin_window_attention_operand(qkv, qkv_bias, rel_bias_table, num_heads, window, shift, compact=False):
    """Strict-precision window attention emitting the fp16x3 GEMM operand directly: fp16 [B,H,W,3C] =
    [lo*2^11 | hi*2^-11 | hi] (the A operand of the projection GEMM; C <= 1536).  `compact` (the consumer is the own
    GEMM): [B,H,W,2C] = [hi | lo*2^11] where the kernel in use can write it (callers tell the two apart by the width)."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if _win_tc and window == 12:
        return swin_window_attention_tc(qkv, qkv_bias, rel_bias_table, num_heads, shift, want_f32=False,
                                        want_operand=True, compact=compact)[1]
    out = torch.empty((B, H, W, 3 * C), device=qkv.device, dtype=torch.float16)
    with _Bracket("swin_window_attention", 1):
        rc = lib().univs_swin_window_attention_f16x3out(_stream(), _chk(qkv, "qkv"), _chk(qkv_bias, "qkv_bias"),
                                                        _chk(rel_bias_table, "rel_bias_table"), B, H, W, C, num_heads,
                                                        window, shift, out.data_ptr())
    check(rc, "swin_window_attention_f16x3out")
    return out


def prepare_mask_features(mask_features_cl, mode=None):
    """Once per clip: mask features in the operand format the mask einsum of the active policy consumes.
    "f16x3": fp16 [T,HW,2C] = [hi | lo] (same bytes as fp32);  "tf32": round-to-nearest TF32 copy (the tcgen05 kernel
    truncates, pre-rounding makes it round-to-nearest overall);  "mma3x": the tensor itself."""
    mode = _einsum_mode if mode is None else mode
    if mode == "f16x3":
        return split_operand(mask_features_cl, "f16u")
    if mode == "tf32":
        return round_tf32(mask_features_cl)
    return mask_features_cl


_einsum_mc = switches.get("EINSUM_MC")   # opt-in: 1 = cluster / multicast kernel for the f16x3 mode


def mask_einsum(mask_embed, mask_features_prepared, out=None, mode=None, tag="mask_einsum"):
    """mask_embed [T,Q,C] fp32, mask_features_prepared from prepare_mask_features (channel-last) -> [Q,T,HW] fp32.
    `tag` names the launch in the per-kernel event brackets (bench.py)."""
    mode = _einsum_mode if mode is None else mode
    T, Q, Cc = mask_embed.shape
    HW = mask_features_prepared.shape[1]
    if out is None:
        out = torch.empty((Q, T, HW), device=mask_embed.device, dtype=torch.float32)
    if Q > 256:      # one launch covers <= 256 queries (TMEM columns); large prompt vocabularies run in query chunks
        for q0 in range(0, Q, 256):
            mask_einsum(mask_embed[:, q0:q0 + 256].contiguous(), mask_features_prepared, out=out[q0:q0 + 256], mode=mode, tag=tag)
        return out
    if mode == "f16x3":
        e = split_operand(mask_embed if mask_embed.is_contiguous() else mask_embed.contiguous(), "f16u")
        entry = lib().univs_mask_einsum_f16x3_cluster if (_einsum_mc and Q > 16 and Cc <= 256) else lib().univs_mask_einsum_f16x3
        with _Bracket(tag, 1):
            rc = entry(_stream(), _chk(e, "mask_embed", torch.float16),
                       _chk(mask_features_prepared, "mask_features", torch.float16), T, Q, Cc, HW, _chk(out, "out"))
    elif mode == "tf32":
        e = round_tf32(mask_embed)
        with _Bracket(tag, 1):
            rc = lib().univs_mask_einsum_f32(_stream(), _chk(e, "mask_embed"), _chk(mask_features_prepared, "mask_features"),
                                             T, Q, Cc, HW, _chk(out, "out"))
    else:
        with _Bracket(tag, 1):
            rc = lib().univs_mask_einsum_mma_f32(_stream(), _chk(mask_embed, "mask_embed"),
                                                 _chk(mask_features_prepared, "mask_features"), T, Q, Cc, HW, PREC_TF32X3,
                                                 _chk(out, "out"))
    check(rc, "mask_einsum")
    return out


def mask_einsum_mma(mask_embed, mask_features_cl, precision, out=None):
    """register-operand variant, explicit precision (cross-check of the tcgen05 kernel)."""
    T, Q, Cc = mask_embed.shape
    HW = mask_features_cl.shape[1]
    if out is None:
        out = torch.empty((Q, T, HW), device=mask_embed.device, dtype=torch.float32)
    with _Bracket("mask_einsum", 1):
        rc = lib().univs_mask_einsum_mma_f32(_stream(), _chk(mask_embed, "mask_embed"),
                                             _chk(mask_features_cl, "mask_features"), T, Q, Cc, HW, precision,
                                             _chk(out, "out"))
    check(rc, "mask_einsum_mma")
    return out


def attn_mask_bits(mask_logits, hw, target_hw):
    """mask_logits [Q,T,H*W] -> (bits uint32-as-int32 [T,Q,words], row_open int32 [T,Q])"""
    Q, T, _ = mask_logits.shape
    H, W = hw
    h, w = target_hw
    words = (h * w + 31) // 32
    bits = torch.empty((T, Q, words), device=mask_logits.device, dtype=torch.int32)
    row_open = torch.empty((T, Q), device=mask_logits.device, dtype=torch.int32)
    with _Bracket("attn_mask_bits", 1):
        rc = lib().univs_attn_mask_bits_f32(_stream(), _chk(mask_logits, "mask_logits"), Q, T, H, W, h, w,
                                        bits.data_ptr(), row_open.data_ptr())
    check(rc, "attn_mask_bits")
    return bits, row_open


def mask_feature_pool(feats_cl, hw, target_hw, mode=None):
    """Mask features [T, H*W, C] fp32 (channel-last) pooled to `target_hw` in the operand format of the einsum mode
    (see prepare_mask_features): [T, h*w, C or 2C]."""
    mode = _einsum_mode if mode is None else mode
    T, _, C = feats_cl.shape
    (H, W), (h, w) = hw, target_hw
    f16 = mode == "f16x3"
    out = torch.empty((T, h * w, 2 * C if f16 else C), device=feats_cl.device, dtype=torch.float16 if f16 else torch.float32)
    with _Bracket("mask_feature_pool", 1):
        rc = lib().univs_mask_feature_pool_f32(_stream(), _chk(feats_cl, "mask_features"), T, H, W, C, h, w, out.data_ptr(),
                                               -2 if f16 else 0)
    check(rc, "mask_feature_pool")
    return round_tf32(out) if mode == "tf32" else out


def attn_mask_bits_direct(mask_logits):
    """mask_logits [Q,T,S] at the memory resolution -> (bits [T,Q,words], row_open [T,Q])"""
    Q, T, S = mask_logits.shape
    bits = torch.empty((T, Q, (S + 31) // 32), device=mask_logits.device, dtype=torch.int32)
    row_open = torch.empty((T, Q), device=mask_logits.device, dtype=torch.int32)
    with _Bracket("attn_mask_bits", 1):
        rc = lib().univs_attn_mask_bits_direct_f32(_stream(), _chk(mask_logits, "mask_logits"), Q, T, S, bits.data_ptr(),
                                                   row_open.data_ptr())
    check(rc, "attn_mask_bits_direct")
    return bits, row_open


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, _stream())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), device=device, dtype=torch.uint8)
        _ws_cache[key] = ws
    return ws


_mha_tc = switches.get("MHA_TC")   # opt-in: 1 = tcgen05 kernel for the cross-attention shape (Lq <= 256),
                                                      # 3 = the same with the transposed-V (K-major) diagnostic variant
MHA_TC_MIN_KEYS = 512


def mha_core_tc(q, k, v, mask_bits=None, row_open=None, flags=None):
    """Strict-precision attention core on the tcgen05 tensor cores (univs_mha_tc_forward_f32): Lq <= 256."""
    B, Lq, Cc = q.shape
    Lk = k.shape[1]
    out = torch.empty_like(q)
    nbytes = lib().univs_mha_tc_workspace_bytes(B, Lq, Lk, Cc)
    ws = _workspace(nbytes, q.device)
    mb = 0 if mask_bits is None else mask_bits.shape[0]
    with _Bracket("mha", 1 if nbytes <= 16 else 2):
        rc = lib().univs_mha_tc_forward_f32(
            _stream(), _chk(q, "q"), _chk(k, "k"), _chk(v, "v"),
            None if mask_bits is None else _chk(mask_bits, "mask_bits", torch.int32),
            None if row_open is None else _chk(row_open, "row_open", torch.int32),
            mb, B, Lq, Lk, Cc, (_mha_tc >> 1) if flags is None else flags, ws.data_ptr(), out.data_ptr())
    check(rc, "mha_tc_forward")
    return out


def mha_core(q, k, v, mask_bits=None, row_open=None, precision=None):
    """q [B,Lq,C], k,v [B,Lk,C] (projected, q unscaled); mask_bits int32 [Bm,Lq,ceil(Lk/32)] (bit set = blocked)."""
    B, Lq, Cc = q.shape
    Lk = k.shape[1]
    if (_mha_tc and Lq <= 256 and Lk >= MHA_TC_MIN_KEYS
            and (_default_precision if precision is None else precision) == PREC_TF32X3):
        return mha_core_tc(q, k, v, mask_bits, row_open)
    if (_mha_tc and B == 1 and Lq > 256 and Lk >= MHA_TC_MIN_KEYS and (mask_bits is None or mask_bits.shape[0] == 1)
            and (_default_precision if precision is None else precision) == PREC_TF32X3):
        # the Q*T self-attention of the decoder (one batch element, ~1000 query tokens, transformer_layers.py:34-44): the
        # tcgen05 kernel takes <= 256 query rows per batch element, so the queries are cut into equal chunks that become the
        # batch dimension (keys / values repeated: ~1 MB each); padding rows attend everywhere and are dropped
        n = -(-Lq // 256)
        chunk = -(-Lq // n)
        pad = n * chunk - Lq
        qq = (torch.nn.functional.pad(q, (0, 0, 0, pad)) if pad else q).reshape(n, chunk, Cc)
        kk, vv = k.expand(n, Lk, Cc).contiguous(), v.expand(n, Lk, Cc).contiguous()
        mb = ro = None
        if mask_bits is not None:
            mb = (torch.nn.functional.pad(mask_bits, (0, 0, 0, pad)) if pad else mask_bits).reshape(n, chunk, -1).contiguous()
            if row_open is not None:
                ro = (torch.nn.functional.pad(row_open, (0, pad), value=1) if pad else row_open).reshape(n, chunk).contiguous()
        return mha_core_tc(qq.contiguous(), kk, vv, mb, ro).reshape(1, n * chunk, Cc)[:, :Lq].contiguous()
    out = torch.empty_like(q)
    nbytes = lib().univs_mha_workspace_bytes(B, Lq, Lk, Cc)
    ws = _workspace(nbytes, q.device)
    mb = 0 if mask_bits is None else mask_bits.shape[0]
    with _Bracket("mha", 1 if nbytes <= 16 else 2):
        rc = lib().univs_mha_forward_f32(
            _stream(), _chk(q, "q"), _chk(k, "k"), _chk(v, "v"),
            None if mask_bits is None else _chk(mask_bits, "mask_bits", torch.int32),
            None if row_open is None else _chk(row_open, "row_open", torch.int32),
            mb, B, Lq, Lk, Cc, _default_precision if precision is None else precision, ws.data_ptr(), out.data_ptr())
    check(rc, "mha_forward")
    return out


def proca_core(q, k_self, v_self, k_mem, v_mem):
    """q,k_self,v_self [P,T,C]; k_mem,v_mem [P,Tm,L,C] -> [P,T,C]"""
    P, T, Cc = q.shape
    Tm, L = k_mem.shape[1], k_mem.shape[2]
    out = torch.empty_like(q)
    with _Bracket("proca", 1):
        rc = lib().univs_proca_forward_f32(_stream(), _chk(q, "q"), _chk(k_self, "k_self"), _chk(v_self, "v_self"),
                                       _chk(k_mem, "k_mem"), _chk(v_mem, "v_mem"), P, T, Tm, L, Cc, out.data_ptr())
    check(rc, "proca_forward")
    return out


def _on_device(t: torch.Tensor) -> bool:
    return t.is_cuda


def _chk_rows16(t: torch.Tensor, name: str):
    """fp16 CUDA matrix whose rows are contiguous (a row pitch is allowed: views of wider operand containers)"""
    if not _on_device(t):
        raise _cabi.UnivsB200Error(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != torch.float16 or t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _cabi.UnivsB200Error(f"{name}: expected an fp16 matrix with contiguous rows")
    return t.data_ptr()


ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def gemm_f16x3_tc(x16, x_offs, w16, w_offs, k, alpha=1.0, bias=None, addend=None, out=None, want_f32=True,
                  want_operand=False, act=ACT_NONE, tap_rows=None, rows=None):
    """Dense layer on the tcgen05 tensor cores with fused epilogue (univs_gemm_f16x3_tc, csrc/gemm_tc.cu):
        y = act(alpha * x w^T + bias) + addend
    x16 [tokens, ldx] / w16 [channels, ldw]: fp16 operand containers (row views allowed); `x_offs` / `w_offs` = (column of
    the hi block, column of the lo*2^11 block), `k` columns each.  addend fp32 [tokens, channels] (rows may be strided; may
    be `out`).  Returns (y fp32 [tokens, channels] or None, y as compact operand fp16 [tokens, 2*channels] = [hi | lo*2^11]
    or None).
    tap_rows (univs_gemm_f16x3_tc_taps): row offsets of shifted-row taps -- y[m] = sum_t x[m + tap_rows[t]] w_t^T with the hi / lo'
    blocks of w16 `len(tap_rows) * k` columns wide (tap-major) and `rows` output rows (x16 rows beyond its end read as zeros):
    a k x k convolution over a zero-padded channel-last activation as ONE accumulation."""
    M, N = x16.shape[0], w16.shape[0]
    taps = 1 if tap_rows is None else len(tap_rows)
    x_rows = M
    if tap_rows is not None:
        M = int(rows)
    xp, wp = _chk_rows16(x16, "x16"), _chk_rows16(w16, "w16")
    if out is not None and (out.dtype != torch.float32 or not _on_device(out) or out.shape != (M, N) or out.stride(1) != 1):
        raise _cabi.UnivsB200Error("gemm_f16x3_tc: out must be an fp32 CUDA [tokens, channels] matrix with contiguous rows")
    if want_f32 and out is None:
        out = torch.empty((M, N), device=x16.device, dtype=torch.float32)
    if addend is not None and (addend.dtype != torch.float32 or not _on_device(addend) or addend.shape != (M, N) or addend.stride(1) != 1):
        raise _cabi.UnivsB200Error("gemm_f16x3_tc: addend must be an fp32 CUDA [tokens, channels] matrix with contiguous rows")
    out16 = torch.empty((M, 2 * N), device=x16.device, dtype=torch.float16) if want_operand else None
    if M == 0:
        return out, out16
    if _event_sink is not None:
        flop_count["gemm_f16x3_tc"] = flop_count.get("gemm_f16x3_tc", 0) + 2 * M * N * int(k) * taps
    offs = (ctypes.c_int64 * taps)(*([0] if tap_rows is None else [int(v) for v in tap_rows]))
    with _Bracket("gemm_f16x3_tc", 1):
        rc = lib().univs_gemm_f16x3_tc_taps(
            _stream(), xp, x16.stride(0), int(x_offs[0]), int(x_offs[1]), x_rows, wp, w16.stride(0), int(w_offs[0]), int(w_offs[1]),
            M, N, int(k), taps, ctypes.cast(offs, ctypes.c_void_p), float(alpha), None if bias is None else _chk(bias, "bias"),
            None if addend is None else addend.data_ptr(), 0 if addend is None else addend.stride(0),
            None if out is None else out.data_ptr(), 0 if out is None else out.stride(0),
            None if out16 is None else out16.data_ptr(), 2 * N, N, int(act))
    check(rc, "gemm_f16x3_tc")
    return out, out16


def f16_chunk(K: int) -> int:
    """K-chunk of the fp16x3 operand layout: the largest divisor of K that is <= 1536 and a multiple of 32 (bounds the
    main-term accumulation chain of one GEMM to 96 MMA steps; chunks are summed by fp32 GEMM epilogues)."""
    if K <= 1536:
        return K
    for c in range(1536, 31, -32):
        if K % c == 0:
            return c
    return K


def _split_code(split, C):
    """split: None/False plain | "tf32"/True fp32 [hi|lo] | "f16" fp16 chunks [lo*2^11|hi*2^-11|hi] | "f16u" fp16 [hi|lo]"""
    if not split:
        return 0, 1, torch.float32
    if split in ("tf32", True):
        return C, 2, torch.float32
    if split == "f16":
        return -f16_chunk(C), 3, torch.float16
    if split == "f16u":
        return -2, 2, torch.float16
    if split == "f16c":
        return -3, 2, torch.float16
    raise ValueError(split)


def _split_out(x, split):
    C = x.shape[-1]
    code, mult, dt = _split_code(split, C)
    return torch.empty((*x.shape[:-1], mult * C), device=x.device, dtype=dt), code


def layernorm(x, weight, bias, eps=1e-5, residual=None, want_sum=False, split=None, residual_bias=None):
    """Fused (residual add +) LayerNorm over the last dim.  Returns (sum_or_None, out); `out` is fp32 [..., C] or a
    split GEMM operand [..., 2C] (see _split_code)."""
    C = x.shape[-1]
    rows = x.numel() // C
    out, code = _split_out(x, split)
    s = torch.empty_like(x) if (want_sum and residual is not None) else None
    with _Bracket("layernorm", 1):
        rc = lib().univs_layernorm_f32(_stream(), _chk(x, "x"), None if residual is None else _chk(residual, "residual"),
                                       None if residual_bias is None else _chk(residual_bias, "residual_bias"),
                                       _chk(weight, "weight"), _chk(bias, "bias"), rows, C, float(eps),
                                       None if s is None else s.data_ptr(), out.data_ptr(), code)
    check(rc, "layernorm")
    if want_sum and residual is None:
        s = x
    return s, out


def gelu(x, split=None, bias=None):
    C = x.shape[-1]
    out, code = _split_out(x, split)
    with _Bracket("gelu", 1):
        rc = lib().univs_gelu_f32(_stream(), _chk(x, "x"), None if bias is None else _chk(bias, "bias"),
                                  x.numel() // C, C, out.data_ptr(), code)
    check(rc, "gelu")
    return out


def relu(x, split=None, bias=None):
    C = x.shape[-1]
    out, code = _split_out(x, split)
    with _Bracket("relu", 1):
        rc = lib().univs_relu_f32(_stream(), _chk(x, "x"), None if bias is None else _chk(bias, "bias"),
                                  x.numel() // C, C, out.data_ptr(), code)
    check(rc, "relu")
    return out


def split_operand(x, split="tf32"):
    """[..., C] fp32 -> split GEMM operand ("tf32": fp32 [..,2C] = [hi|lo]; "f16": fp16 [..,3C] chunks [lo*2^11|hi*2^-11|hi];
    "f16u": fp16 [..,2C] = [hi|lo])"""
    C = x.shape[-1]
    out, code = _split_out(x, split)
    with _Bracket("split", 1):
        rc = lib().univs_split_tf32_f32(_stream(), _chk(x, "x"), x.numel() // C, C, code, out.data_ptr())
    check(rc, "split")
    return out


# ---- channel-last GroupNorm fused with the FPN glue (csrc/groupnorm.cu) ------------------------------------------
_gn_ws = {}


# Scratch buffers that persist between calls (GroupNorm partial sums here, the zero-bordered operand buffers in nn_ops) are
# keyed by this slot: meta_arch._grouped_forward issues frame group g with scratch_slot = g, so concurrently running groups
# never share one.  (Not keyed by the stream id: under CUDA-graph capture the stream differs from the warm-up stream and a
# fresh buffer -- with its zero fill -- would be captured into the graph.)
scratch_slot = 0


def _gn_workspace(device, nbytes):
    key = (device, scratch_slot)
    buf = _gn_ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), device=device, dtype=torch.uint8)
        _gn_ws[key] = buf
    return buf


def _strided_image(x):
    """x [N,H,W,C] fp32 with contiguous pixels (stride(3) == 1, stride(2) == C); rows / images may be strided (a
    [:, :H, :W] view of a padded buffer).  Returns (img_stride, row_stride) in elements."""
    N, H, W, C = x.shape
    if x.stride(3) != 1 or x.stride(2) != C:
        raise _cabi.UnivsB200Error("groupnorm_cl: pixels must be contiguous channel-last rows")
    return x.stride(0), x.stride(1)


def groupnorm_cl(x, weight, bias, groups, eps=1e-5, lowres=None, relu=False, want_f32=True, split=None, pad=0,
                 out_split=None):
    """GroupNorm over a channel-last activation x [N,H,W,C] (+ bilinear-upsampled `lowres` [N,h2,w2,C]) (+ ReLU).
    Returns (y_f32 or None, operand or None): y_f32 [N,H,W,C]; operand = the result in the GEMM operand format `split`
    ("tf32" | "f16" | "f16u") written at the interior of a spatially zero-padded [N,H+2*pad,W+2*pad,width] buffer
    (`out_split`, allocated zeroed if not given -- only the interior is ever written, so a cached buffer stays valid)."""
    N, H, W, C = x.shape
    if x.dtype != torch.float32 or not x.is_cuda:
        raise _cabi.UnivsB200Error("groupnorm_cl: x must be a CUDA float32 tensor")
    img_stride, row_stride = _strided_image(x)
    dev = x.device
    ws = _gn_workspace(dev, int(lib().univs_groupnorm_workspace_bytes(N, H, W, groups)))
    stats = torch.empty((N, groups, 2), device=dev, dtype=torch.float32)
    with _Bracket("groupnorm_stats", 2):
        rc = lib().univs_groupnorm_stats_f32(_stream(), x.data_ptr(), N, H, W, C, img_stride, row_stride, groups, float(eps),
                                             ws.data_ptr(), stats.data_ptr())
    check(rc, "groupnorm_stats")
    y = torch.empty((N, H, W, C), device=dev, dtype=torch.float32) if want_f32 else None
    code = 0
    if split:
        code, mult, dt = _split_code(split, C)
        shape = (N, H + 2 * pad, W + 2 * pad, mult * C)
        if out_split is None:
            out_split = torch.zeros(shape, device=dev, dtype=dt) if pad else torch.empty(shape, device=dev, dtype=dt)
        elif tuple(out_split.shape) != shape or out_split.dtype != dt or not out_split.is_contiguous():
            raise _cabi.UnivsB200Error(f"groupnorm_cl: out_split must be a contiguous {dt} tensor of shape {shape}")
    else:
        out_split = None
    low_stride = 0
    if lowres is not None:
        if lowres.dtype != torch.float32 or not lowres.is_cuda or lowres.shape[0] != N or lowres.shape[3] != C:
            raise _cabi.UnivsB200Error("groupnorm_cl: lowres must be a CUDA float32 [N,h2,w2,C] tensor")
        if lowres.stride(3) != 1 or lowres.stride(2) != C or lowres.stride(1) != C * lowres.shape[2]:
            lowres = lowres.contiguous()        # frames may be strided (a level slice of the token matrix), pixels not
        low_stride = lowres.stride(0) if N > 1 else lowres.shape[1] * lowres.shape[2] * C
    with _Bracket("groupnorm_apply", 1):
        rc = lib().univs_groupnorm_apply_f32(
            _stream(), x.data_ptr(), N, H, W, C, img_stride, row_stride, stats.data_ptr(), _chk(weight, "weight"),
            _chk(bias, "bias"), groups, None if lowres is None else lowres.data_ptr(), low_stride,
            0 if lowres is None else lowres.shape[1], 0 if lowres is None else lowres.shape[2], 1 if relu else 0,
            None if y is None else y.data_ptr(), None if out_split is None else out_split.data_ptr(), code, pad)
    check(rc, "groupnorm_apply")
    return y, out_split


# ---- gather-fused kernels at the edges of the Swin stages (csrc/swin_glue.cu) ------------------------------------
def patchify_normalize(frames, pixel_mean, pixel_std, padded_size, patch=4, split=None):
    """frames [N,3,H,W] uint8 / float32 (0..255) on the device -> [N, Hp/4, Wp/4, 48 (* split width)]: normalised,
    zero-padded 4x4 patches, column = c*16 + ky*4 + kx (PatchEmbed.proj weight.view(E, 48))."""
    if frames.dtype not in (torch.uint8, torch.float32) or not frames.is_cuda or not frames.is_contiguous():
        raise _cabi.UnivsB200Error("patchify_normalize: frames must be a contiguous CUDA uint8 / float32 tensor")
    N, c3, H, W = frames.shape
    Hp, Wp = padded_size
    if c3 != 3:
        raise _cabi.UnivsB200Error("patchify_normalize: frames must be [N,3,H,W]")
    import ctypes
    K = 3 * patch * patch
    code, mult, dt = _split_code(split, K)
    out = torch.empty((N, Hp // patch, Wp // patch, mult * K), device=frames.device, dtype=dt)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in pixel_mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in pixel_std])
    with _Bracket("patchify", 1):
        rc = lib().univs_patchify_normalize(_stream(), frames.data_ptr(), 1 if frames.dtype == torch.uint8 else 0, N, H, W,
                                            Hp, Wp, patch, m3, s3, out.data_ptr(), code)
    check(rc, "patchify_normalize")
    return out


def layernorm_multi(x, weight, bias, eps=1e-5, residual=None, residual_bias=None, want_f32=True, split=None, pos=None,
                    want_operand=True):
    """LayerNorm(x + residual + residual_bias) with several consumers.  Returns (y fp32 or None, operand(y) or None,
    operand(y + pos[row % pos_rows]) or None); operands in format `split` ("tf32" | "f16" | "f16u")."""
    x = _chk_t(x, "x")
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x) if want_f32 else None
    code, mult, dt = _split_code(split, C)
    op = torch.empty((*x.shape[:-1], mult * C), device=x.device, dtype=dt) if (split and want_operand) else None
    op_pos = None
    pos_rows = 0
    if pos is not None:
        if not split:
            raise _cabi.UnivsB200Error("layernorm_multi: pos needs an operand format")
        pos = _chk_t(pos, "pos")
        pos_rows = pos.numel() // C
        op_pos = torch.empty((*x.shape[:-1], mult * C), device=x.device, dtype=dt)
    with _Bracket("layernorm_multi", 1):
        rc = lib().univs_layernorm_multi_f32(
            _stream(), x.data_ptr(), None if residual is None else _chk(residual, "residual"),
            None if residual_bias is None else _chk(residual_bias, "residual_bias"), _chk(weight, "weight"),
            _chk(bias, "bias"), rows, C, float(eps), None if y is None else y.data_ptr(),
            None if op is None else op.data_ptr(), code, None if pos is None else pos.data_ptr(), pos_rows,
            None if op_pos is None else op_pos.data_ptr())
    check(rc, "layernorm_multi")
    return y, op, op_pos


def layernorm_merge2x2(x, weight, bias, eps=1e-5, split=None):
    """PatchMerging gather + LayerNorm: x [N,H,W,C] -> [N, ceil(H/2), ceil(W/2), 4C (* split width)]."""
    x = _chk_t(x, "x")
    N, H, W, C = x.shape
    code, mult, dt = _split_code(split, 4 * C)
    out = torch.empty((N, (H + 1) // 2, (W + 1) // 2, mult * 4 * C), device=x.device, dtype=dt)
    with _Bracket("layernorm_merge2x2", 1):
        rc = lib().univs_layernorm_merge2x2_f32(_stream(), x.data_ptr(), N, H, W, C, _chk(weight, "weight"),
                                                _chk(bias, "bias"), float(eps), out.data_ptr(), code)
    check(rc, "layernorm_merge2x2")
    return out


def split_tf32(x, chunk=None):
    return split_operand(x, "tf32")


def round_tf32(x, out=None):
    if out is None:
        out = torch.empty_like(x)
    with _Bracket("round_tf32", 1):
        rc = lib().univs_round_tf32_f32(_stream(), _chk(x, "x"), _chk(out, "out"), x.numel())
    check(rc, "round_tf32")
    return out


def pack_mask_bits(mask_bool):
    """Host-side helper: bool [..., Lk] (True = blocked) -> int32 words [..., ceil(Lk/32)] (torch ops; used once per
    clip for the static self-attention mask)."""
    Lk = mask_bool.shape[-1]
    words = (Lk + 31) // 32
    pad = words * 32 - Lk
    m = torch.nn.functional.pad(mask_bool.to(torch.int64), (0, pad), value=1)
    m = m.view(*mask_bool.shape[:-1], words, 32)
    weights = (1 << torch.arange(32, device=mask_bool.device, dtype=torch.int64))
    packed = (m * weights).sum(-1)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)
    return packed.to(torch.int32).contiguous()
