"""Dense-layer plumbing around the library GEMMs, parameterised by the arithmetic policy (precision.py).

Policy "tf32x3" (strict, default): every GEMM / convolution runs on the TF32 tensor cores (cuBLAS / cuDNN through
torch) but on *split* operands, X*W^T = [Xh|Xl]*[Wh|Wh]^T + Xh*Wl^T, where hi = upper 19 bits (exact in TF32, so the
tensor core's truncation loses nothing) and lo = x - hi: products are fp32-equivalent (the dropped Xl*Wl term is
2^-22 relative), accumulation is fp32.  Activations are produced already split by the fused kernels that feed the
GEMMs (LayerNorm, GELU; csrc/elementwise.cu); weights are split once and cached.
Policy "tf32": plain operands, single TF32 GEMM.  Policy "fp32": plain operands, IEEE fp32 GEMMs (slow reference)."""
from __future__ import annotations

import contextlib
import os
import weakref

import torch
import torch.nn.functional as F

from . import ops, switches

_policy = "fp32"
_wcache = {}
# Fused glue kernels (channel-last GroupNorm + FPN add / ReLU / operand emission, PatchMerging gather-LayerNorm, fused
# frame ingest; csrc/groupnorm.cu, csrc/swin_glue.cu).  Written at the end of round 1 without GPU time left to validate
# them, hence opt-in (UNIVS_FUSED_GLUE=1 or set_fused_glue(True)); the default path is the one measured in round 1.
_fused_glue = switches.get("FUSED_GLUE") == 1


# Dense layers on the own tcgen05 GEMM (csrc/gemm_tc.cu) instead of the library GEMM: fp16x3 policy only.  Reads four operand
# tiles per k-step instead of six, keeps the correction terms in their own accumulator, and fuses bias / GELU / ReLU /
# K-slice accumulation / operand emission into the epilogue (the Swin MLP never materialises its fp32 hidden activation).
_gemm_tc = switches.get("GEMM_TC") == 1


def set_gemm_tc(on: bool):
    global _gemm_tc
    _gemm_tc = bool(on)


def gemm_tc() -> bool:
    return _gemm_tc and _policy == "fp16x3"


def _tc_linear(h3, K, b3, alpha, bias, act=0, want_f32=True, want_operand=False, addend=None, inplace=False):
    """y = act(x W^T + bias) (+ addend) through ops.gemm_f16x3_tc.  h3: activation operand, either [rows, 3K] in the K-chunk
    container [lo' | hi_s | hi] (attention / MSDeformAttn epilogues, split "f16") or the compact [rows, 2K] = [hi | lo']
    (row-wise kernels with split "f16c", the GEMM's own operand epilogue) -- told apart by the width; b3: weight operand
    [N, 3K] chunks [hi_s | lo' | hi] (_split_weight).  K-chunks beyond the first accumulate into the fp32 result in place.
    `addend` [rows, N] fp32 is added by the epilogue (residual); with `inplace` the result overwrites it."""
    compact = h3.shape[-1] == 2 * K
    assert compact or h3.shape[-1] == 3 * K, (tuple(h3.shape), K)
    kc = ops.f16_chunk(K)
    n = K // kc
    assert n == 1 or (act == 0 and not want_operand), "activation / operand epilogues need a single K-chunk"
    y, y16 = addend, None
    for c in range(n):
        x_offs = (c * kc, K + c * kc) if compact else (c * 3 * kc + 2 * kc, c * 3 * kc)
        w_offs = (c * 3 * kc + 2 * kc, c * 3 * kc + kc)
        y, y16 = ops.gemm_f16x3_tc(h3, x_offs, b3, w_offs, kc, alpha, bias if c == 0 else None, y,
                                   out=y if (c > 0 or addend is None or inplace) else None, want_f32=want_f32,
                                   want_operand=want_operand, act=act)
    return y, y16


_conv_fused = switches.get("CONV_FUSED") == 1     # k x k convolutions as one shifted-row accumulation (gemm_tc taps)


def set_fused_glue(on: bool):
    global _fused_glue
    _fused_glue = bool(on)


def fused_glue() -> bool:
    return _fused_glue


def set_policy(p: str):
    global _policy
    assert p in ("fp32", "tf32", "tf32x3", "fp16x3")
    _policy = p
    _wcache.clear()


def policy() -> str:
    return _policy


def splitting() -> bool:
    return _policy in ("tf32x3", "fp16x3")


def _fmt():
    """operand format produced by the fused kernels for the active policy: with the own tcgen05 GEMM the compact
    [hi | lo*2^11] container (4 bytes per element; the library-GEMM formulation needs the 6-byte [lo' | hi_s | hi] one)"""
    if _policy == "fp16x3" and _gemm_tc:
        return "f16c"
    return {"tf32x3": "tf32", "fp16x3": "f16"}.get(_policy)


def g1_chunk(K: int) -> int:
    """K-slice of the hi*hi GEMM.  The tensor cores accumulate with truncation, a bias of ~ -1.65e-8 per MMA step
    (tools/accum_probe.py); slicing K at <= 1024 keeps it <= 2.1e-6 per GEMM; slices are summed by fp32 epilogues."""
    if K <= 1024:
        return K
    for c in range(1024, 255, -32):
        if K % c == 0:
            return c
    return K


def _owner(w: torch.Tensor) -> torch.Tensor:
    """the tensor whose lifetime bounds the validity of a cache entry made from `w`: the base of a view (in_proj slices are
    fresh view objects on every call), else `w` itself"""
    return w._base if w._base is not None else w


def _owner_ref(w):
    return weakref.ref(_owner(w))


def _same_owner(ref, w) -> bool:
    """A cache hit needs the SAME live tensor object, not just the same address: the allocator hands a freed model's
    addresses to the next model of the same architecture (same shape, same _version), whose weights differ."""
    return ref() is _owner(w)


def _split_weight(w: torch.Tensor, cache=True):
    """2-D weight [N,K] -> policy-specific B operands, cached per parameter version.
    tf32x3: (Wh [N,K], Wlh [N,2K] = [Wl | Wh]).
    fp16x3: (B3 [N,3K] fp16 in K-chunks [Wh*2^-11 | Wl*2^11 | Wh] of W*2^s, alpha = 2^-s); s is a per-tensor power of
            two that lifts the weights into fp16's normal range."""
    key = (w.data_ptr(), tuple(w.shape), tuple(w.stride()))        # views of packed parameters (in_proj slices) hit too
    sig = (w.data_ptr(), w._version, tuple(w.shape), _policy)
    ent = _wcache.get(key) if cache else None
    if ent is None or ent[0] != sig or not _same_owner(ent[3], w):
        w2d = w.detach().float().reshape(w.shape[0], -1).contiguous()
        N, K = w2d.shape
        if _policy == "tf32x3":
            hl = ops.split_operand(w2d, "tf32")                              # [N,2K] = [hi | lo]
            hi, lo = hl[:, :K], hl[:, K:]
            ent = (sig, hi.contiguous(), torch.cat([lo, hi], 1).contiguous(), _owner_ref(w))
        else:
            amax = float(w2d.abs().max())
            s = 0 if amax == 0.0 else max(-24, min(24, int(torch.floor(torch.log2(torch.tensor(8192.0 / amax))))))
            a3 = ops.split_operand(w2d * (2.0 ** s), "f16")                   # chunks [lo' | hi_s | hi]
            kc = ops.f16_chunk(K)
            v = a3.view(N, K // kc, 3, kc)
            b3 = torch.stack([v[:, :, 1], v[:, :, 0], v[:, :, 2]], 2).reshape(N, 3 * K).contiguous()   # [hi_s | lo' | hi]
            ent = (sig, b3, 2.0 ** -s, _owner_ref(w))
        if cache:
            _wcache[key] = ent
    return ent[1], ent[2]


_zero_cache = {}


def _zeros(n, device):
    z = _zero_cache.get((n, device))
    if z is None:
        z = torch.zeros(n, device=device, dtype=torch.float32)
        _zero_cache[(n, device)] = z
    return z


_inplace16 = [None]     # whether torch.addmm(..., out_dtype=f32, out=acc) accepts acc as both input and output


def _addmm16(acc, a, bt, alpha):
    """acc (fp32) + alpha * a @ bt with fp16 operands and fp32 accumulate/output."""
    if _inplace16[0] is not False:
        try:
            y = torch.addmm(acc, a, bt, alpha=alpha, out_dtype=torch.float32, out=acc)
            _inplace16[0] = True
            return y
        except (RuntimeError, TypeError):
            if _inplace16[0] is True:
                raise
            _inplace16[0] = False
    return torch.addmm(acc, a, bt, alpha=alpha, out_dtype=torch.float32)


def prep(x):
    """activation -> GEMM operand format of the active policy"""
    return ops.split_operand(x if x.is_contiguous() else x.contiguous(), _fmt()) if splitting() else x


def layernorm(x, norm, residual=None, want_sum=False, for_gemm=True, residual_bias=None):
    """(sum | None, LN(x + residual + residual_bias)); the LN output is emitted in the GEMM operand format of the active
    policy iff it feeds a GEMM.  `residual_bias` = the deferred bias of the GEMM that produced `residual`."""
    return ops.layernorm(x.contiguous(), norm.weight, norm.bias, norm.eps,
                         None if residual is None else residual.contiguous(), want_sum, _fmt() if for_gemm else None,
                         residual_bias)


def layernorm_multi(x, norm, residual=None, residual_bias=None, pos=None, want_operand=True):
    """Post-norm LayerNorm whose result is both the residual stream and GEMM input(s).  Returns
    (y fp32, operand(y), operand(y + pos) or None); under non-splitting policies the operands are plain fp32 tensors."""
    if splitting():
        return ops.layernorm_multi(x.contiguous(), norm.weight, norm.bias, norm.eps,
                                   None if residual is None else residual.contiguous(), residual_bias, True, _fmt(),
                                   pos, want_operand)
    y = ops.layernorm(x.contiguous(), norm.weight, norm.bias, norm.eps,
                      None if residual is None else residual.contiguous(), False, None, residual_bias)[1]
    return y, (y if want_operand else None), (None if pos is None else y + pos.view(1, -1, y.shape[-1]).expand(
        y.numel() // pos.numel(), -1, -1).reshape(y.shape))


def layernorm_merge2x2(x_cl, norm):
    """PatchMerging gather + LayerNorm (ops.layernorm_merge2x2), emitted as the operand of the reduction GEMM."""
    return ops.layernorm_merge2x2(x_cl if x_cl.is_contiguous() else x_cl.contiguous(), norm.weight, norm.bias, norm.eps,
                                  _fmt())


def patchify(frames, pixel_mean, pixel_std, padded_size, patch):
    """Frame ingest: normalise + zero-pad + 4x4 patch gather in one pass, emitted as the operand of the patch-embedding GEMM."""
    return ops.patchify_normalize(frames, pixel_mean, pixel_std, padded_size, patch, _fmt())


def gelu(x, for_gemm=True, bias=None):
    return ops.gelu(x.contiguous(), _fmt() if for_gemm else None, bias)


def relu(x, for_gemm=True, bias=None):
    return ops.relu(x.contiguous(), _fmt() if for_gemm else None, bias)


def linear_prepped(h, weight, bias=None, cache=True):
    """h: operand from prep()/layernorm()/gelu() ([..., 2K] = [Xh | Xl] when splitting, else [..., K]).
    Splitting:  y = Xh Wh^T (+ bias)            -- main term, K sliced at <= 1024 to bound the accumulation chain
                  + [Xh | Xl] [Wl | Wh]^T       -- both correction terms in one 2K-wide GEMM with its own (small)
                                                   accumulator, added by the GEMM epilogue (beta = 1, fp32 RN)."""
    if not splitting():
        return F.linear(h, weight, bias)
    N, K = weight.shape
    wh, wlh = _split_weight(weight, cache)
    kc = g1_chunk(K)
    if _policy == "tf32x3":
        h2 = h.reshape(-1, 2 * K)
        y = F.linear(h2[:, :kc], wh[:, :kc], bias)
        for k0 in range(kc, K, kc):
            y.addmm_(h2[:, k0:k0 + kc], wh[:, k0:k0 + kc].t())
        y.addmm_(h2, wlh.t())
    elif gemm_tc() and K % 8 == 0:
        y = _tc_linear(h.reshape(-1, h.shape[-1]), K, wh, wlh, None if bias is None else bias.float().contiguous())[0]
    elif gemm_tc():      # K % 8 != 0: tiny prompt-side products; rebuild the fp32 value (hi + lo' * 2^-11, exact) -> IEEE GEMM
        h2 = h.reshape(-1, h.shape[-1]).float()
        x = (h2[:, :K] + h2[:, K:] * 2.0 ** -11) if h2.shape[-1] == 2 * K else \
            (h2.view(-1, 1, 3, K)[:, :, 2] + h2.view(-1, 1, 3, K)[:, :, 0] * 2.0 ** -11).reshape(-1, K)
        with ieee_fp32():
            y = F.linear(x, weight.float(), bias)
    else:  # fp16x3: ONE fp16 GEMM per K-chunk over [Xl' | Xh_s | Xh] x [Wh_s | Wl' | Wh], fp32 accumulate and output
        f32 = torch.float32
        b3, alpha = wh, wlh
        kc = ops.f16_chunk(K)
        kc3 = 3 * kc
        h3 = h.reshape(-1, 3 * K)
        try:
            if bias is not None:                  # small GEMMs only: hot-path callers defer the bias into the consumer kernel
                y = torch.addmm(bias.float(), h3[:, :kc3], b3[:, :kc3].t(), alpha=alpha, out_dtype=f32)
            else:                                 # beta = 0: `input` is ignored (no broadcast copy), alpha undoes the weight scale
                y = torch.addmm(_zeros(N, h.device), h3[:, :kc3], b3[:, :kc3].t(), beta=0, alpha=alpha, out_dtype=f32)
            for k0 in range(kc3, 3 * K, kc3):
                y = _addmm16(y, h3[:, k0:k0 + kc3], b3[:, k0:k0 + kc3].t(), alpha)
        except NotImplementedError:
            # shapes the fp16->fp32 library GEMM does not cover (e.g. degenerate row counts): rebuild the fp32 operand
            # from its split (hi + lo*2^-11, exact) and use an IEEE fp32 GEMM -- these are tiny prompt-side products
            v = h3.view(-1, K // kc, 3, kc).float()
            x = (v[:, :, 2] + v[:, :, 0] * 2.0 ** -11).reshape(-1, K)
            with ieee_fp32():
                y = F.linear(x, weight.float(), bias)
    return y.view(*h.shape[:-1], N)


def linear(x, weight, bias=None, cache=True):
    return linear_prepped(prep(x), weight, bias, cache)


# L2-resident MLP (opt-in: UNIVS_MLP_CHUNK_MB = working set in MB, e.g. 96).  fc1 -> GELU -> fc2 over the whole token matrix
# streams the 4C-wide hidden activation through HBM twice as fp32 and twice as operand (4.5 GB per Swin-L stage-1 block at
# the north-star size).  The three steps are row-independent, so they can run chunk by chunk over the rows with the chunk
# sized so that its fp32 hidden + operand fit in the 126 MB L2: the GELU pass then reads what the GEMM just wrote from L2,
# fc2 reads its operand from L2, and because the allocator hands every chunk the same two scratch blocks the dirty lines
# are overwritten in L2 instead of being written back.  Same kernels, same arithmetic per row; only the schedule changes
# (the library may pick another tile shape for the smaller M: ~1e-7 differences).  The price is GEMM wave quantisation on
# small chunks -- to be swept on the GPU (tools/gpu_round2_sweep.sh).
_mlp_chunk_mb = switches.get("MLP_CHUNK_MB")


def set_mlp_chunk_mb(mb: int):
    global _mlp_chunk_mb
    _mlp_chunk_mb = int(mb)


def mlp_rows_per_chunk(rows: int, hidden: int) -> int:
    """rows per chunk such that fp32 hidden + GEMM operand of the chunk take about _mlp_chunk_mb MB (multiple of 256);
    `rows` when chunking is off or would not split"""
    if _mlp_chunk_mb <= 0:
        return rows
    operand_bytes = {"fp16x3": 6, "tf32x3": 8}.get(_policy, 4)
    per_row = hidden * (4 + operand_bytes)
    chunk = max(256, (_mlp_chunk_mb << 20) // per_row // 256 * 256)
    return rows if chunk >= rows else chunk


def linear_residual(h, weight, bias, residual):
    """residual + h W^T + bias, written over `residual` (own GEMM only: the residual add and the bias ride in the epilogue,
    the consumer LayerNorm then reads one tensor instead of two and writes no sum).  h: operand of the layer."""
    assert gemm_tc()
    N, K = weight.shape
    b3, alpha = _split_weight(weight)
    r2 = residual.view(-1, N)
    _tc_linear(h.reshape(-1, h.shape[-1]), K, b3, alpha, None if bias is None else bias, addend=r2, inplace=True)
    return residual


def linear_act_operand(h, lin, act="relu"):
    """act(lin(h)) emitted as the next layer's GEMM operand.  Own GEMM: one kernel (bias + activation + operand emission in the
    epilogue, no fp32 intermediate); otherwise the library GEMM followed by the row-wise activation kernel."""
    K = lin.in_features
    if gemm_tc() and K % 8 == 0 and K <= 1536:
        b3, alpha = _split_weight(lin.weight)
        y16 = _tc_linear(h.reshape(-1, h.shape[-1]), K, b3, alpha, lin.bias, act=ops.ACT_GELU if act == "gelu" else ops.ACT_RELU,
                         want_f32=False, want_operand=True)[1]
        return y16.view(*h.shape[:-1], y16.shape[-1])
    f = linear_prepped(h, lin.weight, None)
    return gelu(f, bias=lin.bias) if act == "gelu" else relu(f, bias=lin.bias)


def mlp(h, fc1, fc2, residual=None):
    """fc2(GELU(fc1(h) + b1)) without b2 (deferred into the consumer, like everywhere on this path); h is the operand of fc1.
    With `residual` (own GEMM only): residual + fc2(...) + b2, written over `residual`."""
    width = h.shape[-1]
    rows = h.numel() // width
    if residual is not None:
        assert gemm_tc() and fc1.in_features % 8 == 0 and fc1.in_features <= 1536
        b31, a1 = _split_weight(fc1.weight)
        b32, a2 = _split_weight(fc2.weight)
        hid16 = _tc_linear(h.reshape(rows, width), fc1.in_features, b31, a1, fc1.bias, act=ops.ACT_GELU, want_f32=False,
                           want_operand=True)[1]
        _tc_linear(hid16, fc1.out_features, b32, a2, fc2.bias, addend=residual.view(rows, fc2.out_features), inplace=True)
        return residual
    if gemm_tc() and fc1.in_features % 8 == 0 and fc1.in_features <= 1536:
        # fc1 + bias + GELU + operand emission in ONE kernel: the fp32 hidden activation never exists
        b31, a1 = _split_weight(fc1.weight)
        b32, a2 = _split_weight(fc2.weight)
        hid16 = _tc_linear(h.reshape(rows, width), fc1.in_features, b31, a1, fc1.bias, act=ops.ACT_GELU, want_f32=False,
                           want_operand=True)[1]
        y = _tc_linear(hid16, fc1.out_features, b32, a2, None)[0]
        return y.view(*h.shape[:-1], fc2.out_features)
    chunk = mlp_rows_per_chunk(rows, fc1.out_features)
    if chunk >= rows:
        return linear_prepped(gelu(linear_prepped(h, fc1.weight, None), bias=fc1.bias), fc2.weight, None)
    h2 = h.reshape(rows, width)
    out = torch.empty((rows, fc2.out_features), device=h.device, dtype=torch.float32)
    for r0 in range(0, rows, chunk):
        f = linear_prepped(h2[r0:r0 + chunk], fc1.weight, None)
        out[r0:r0 + chunk] = linear_prepped(gelu(f, bias=fc1.bias), fc2.weight, None)
        del f
    return out.view(*h.shape[:-1], fc2.out_features)


_pad_cache = {}


def padded_operand_buffer(N, H, W, C, pad, device):
    """Zeroed [N, H+2p, W+2p, width] buffer in the operand format of the active policy, cached per shape: producers
    only ever write its interior, so the zero border survives from call to call (and across CUDA-graph replays)."""
    _code, mult, dt = ops._split_code(_fmt(), C)
    key = (N, H, W, C, pad, mult, str(device), dt, ops.scratch_slot)   # per frame group: groups may run concurrently
    buf = _pad_cache.get(key)
    if buf is None:
        buf = torch.zeros((N, H + 2 * pad, W + 2 * pad, mult * C), device=device, dtype=dt)
        _pad_cache[key] = buf
    return buf


def groupnorm_cl(x_cl, gn, lowres=None, relu=False, want_f32=True, for_gemm=False, pad=0):
    """GroupNorm module `gn` on a channel-last activation [N,H,W,C] (rows / frames may be strided) + optional bilinear
    top-down add + ReLU, in one pass (ops.groupnorm_cl).  Returns (fp32 [N,H,W,C] or None, GEMM operand or None);
    with pad > 0 the operand sits in the interior of a cached zero-bordered buffer for conv2d_cl_operand."""
    split = _fmt() if (for_gemm and splitting()) else None
    N, H, W, C = x_cl.shape
    buf = padded_operand_buffer(N, H, W, C, pad, x_cl.device) if (split and pad) else None
    return ops.groupnorm_cl(x_cl, gn.weight, gn.bias, gn.num_groups, gn.eps, lowres, relu,
                            want_f32 or not split, split, pad if split else 0, buf)


def conv2d_cl(x_cl, weight, bias=None, padding=0):
    """Convolution on a channel-last activation [N,H,W,Cin] -> [N,H,W,Cout] (storage channel-last).
    Splitting policy: a kxk "same" convolution is k*k shifted GEMMs over the zero-padded, split activation viewed
    as one 2-D token matrix (row = padded pixel), accumulated by the GEMM epilogues -- every tensor-core chain is one
    tap deep."""
    N, H, W, Cin = x_cl.shape
    if isinstance(padding, (tuple, list)):
        assert padding[0] == padding[1]
        padding = int(padding[0])
    if not splitting():
        y = F.conv2d(x_cl.permute(0, 3, 1, 2), weight, bias, padding=padding)
        return y.permute(0, 2, 3, 1)
    Cout, _, kh, kw = weight.shape
    assert kh == kw and padding == kh // 2, "only 'same' square convolutions are used on this path"
    p = padding
    xs = F.pad(ops.split_operand(x_cl.contiguous(), _fmt()), (0, 0, p, p, p, p))  # [N,Hp,Wp,2Cin], zeros split to zeros
    return _conv_taps(xs, H, W, weight, bias, _conv_weight_taps(weight))


def conv2d_cl_operand(xs_padded, H, W, weight, bias=None):
    """conv2d_cl on an activation that already is a spatially zero-padded split operand [N,H+2p,W+2p,width]
    (groupnorm_cl(..., for_gemm=True, pad=p)).  Splitting policies only."""
    assert splitting()
    Cout, _, kh, kw = weight.shape
    assert kh == kw and xs_padded.shape[1] == H + 2 * (kh // 2)
    return _conv_taps(xs_padded, H, W, weight, bias, _conv_weight_taps(weight))


def _conv_weight_taps(weight):
    Cout, _, kh, kw = weight.shape
    key = ("conv", weight.data_ptr(), tuple(weight.shape))
    sig = (weight.data_ptr(), weight._version)
    ent = _wcache.get(key)
    if ent is None or ent[0] != sig or not _same_owner(ent[2], weight):
        taps = [_split_weight(weight.detach()[:, :, dy, dx].contiguous(), cache=False) for dy in range(kh) for dx in range(kw)]
        ent = (sig, taps, _owner_ref(weight))
        _wcache[key] = ent
    return ent[1]


def _conv_weight_fused(weight):
    """k x k convolution weight [Cout,Cin,kh,kw] -> compact fp16 operand [Cout, 2*T*Cin] = [hi | lo*2^11] of W*2^s with the
    taps side by side (tap-major columns, T = kh*kw) and ONE scale for all taps, alpha = 2^-s: the weight of the fused
    shifted-row GEMM (ops.gemm_f16x3_tc with tap_rows).  Cached per parameter version."""
    key = ("convf", weight.data_ptr(), tuple(weight.shape))
    sig = (weight.data_ptr(), weight._version)
    ent = _wcache.get(key)
    if ent is None or ent[0] != sig or not _same_owner(ent[3], weight):
        Cout, Cin, kh, kw = weight.shape
        w2d = weight.detach().float().permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin).contiguous()
        amax = float(w2d.abs().max())
        s = 0 if amax == 0.0 else max(-24, min(24, int(torch.floor(torch.log2(torch.tensor(8192.0 / amax))))))
        ent = (sig, ops.split_operand(w2d * (2.0 ** s), "f16c"), 2.0 ** -s, _owner_ref(weight))
        _wcache[key] = ent
    return ent[1], ent[2]


def _conv_taps(xs, H, W, weight, bias, taps):
    """k*k shifted GEMMs over the padded split activation xs [N,Hp,Wp,width] viewed as one token matrix."""
    Cout, Cin, kh, kw = weight.shape
    N, Hp, Wp = xs.shape[0], xs.shape[1], xs.shape[2]
    x2 = xs.view(N * Hp * Wp, xs.shape[-1])
    R = N * Hp * Wp - ((kh - 1) * Wp + (kw - 1))                                  # rows every tap can address
    y = torch.empty((N * Hp * Wp, Cout), device=xs.device, dtype=torch.float32)
    yr = y[:R]
    f32 = torch.float32
    if gemm_tc() and _conv_fused and Cin % 64 == 0 and Cin <= 1536:
        # ONE accumulation per group of taps (K <= 1536 per tensor-core chain) instead of k*k GEMMs that re-read and re-write
        # the fp32 result: the taps are shifted row windows of the same token matrix
        wf, alpha = _conv_weight_fused(weight)
        T = kh * kw
        groups = -(-T * Cin // 1536)
        per = -(-T // groups)
        x_offs = (0, Cin) if x2.shape[-1] == 2 * Cin else (2 * Cin, 0)         # compact / K-chunk container
        for g0 in range(0, T, per):
            g1 = min(T, g0 + per)
            ops.gemm_f16x3_tc(x2, x_offs, wf, (g0 * Cin, T * Cin + g0 * Cin), Cin, alpha,
                              bias.float() if (bias is not None and g0 == 0) else None, yr if g0 else None, out=yr,
                              tap_rows=[(t // kw) * Wp + (t % kw) for t in range(g0, g1)], rows=R)
        return y.view(N, Hp, Wp, Cout)[:, :H, :W]
    for t, (wh, wlh) in enumerate(taps):
        off = (t // kw) * Wp + (t % kw)
        a = x2[off: off + R]
        if gemm_tc() and Cin % 8 == 0 and Cin <= 1536:
            x_offs = (0, Cin) if a.shape[-1] == 2 * Cin else (2 * Cin, 0)       # compact / K-chunk container
            ops.gemm_f16x3_tc(a, x_offs, wh, (2 * Cin, Cin), Cin, wlh, bias.float() if (bias is not None and t == 0) else None,
                              yr if t else None, out=yr)
            continue
        if _policy == "tf32x3":
            if t == 0:
                torch.addmm(bias if bias is not None else y.new_zeros(Cout), a[:, :Cin], wh.t(), out=yr)
            else:
                yr.addmm_(a[:, :Cin], wh.t())
            yr.addmm_(a, wlh.t())
        else:
            if t == 0:
                base = bias.float() if bias is not None else _zeros(Cout, y.device)
                try:        # write straight into the padded output buffer (rows beyond R are never read)
                    torch.addmm(base, a, wh.t(), beta=0 if bias is None else 1, alpha=wlh, out_dtype=f32, out=yr)
                except (RuntimeError, TypeError):
                    yr.copy_(torch.addmm(base, a, wh.t(), beta=0 if bias is None else 1, alpha=wlh, out_dtype=f32))
            else:
                res = _addmm16(yr, a, wh.t(), wlh)
                if res.data_ptr() != yr.data_ptr():
                    yr.copy_(res)
    return y.view(N, Hp, Wp, Cout)[:, :H, :W]


@contextlib.contextmanager
def ieee_fp32():
    """Small matmuls / convs that stay on plain torch ops (patch embedding, prompt pooling) under exact fp32."""
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    if _policy == "tf32":
        yield
        return
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = a
        torch.backends.cudnn.allow_tf32 = b
