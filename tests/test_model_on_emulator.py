"""End to end on the CPU: the product model's clip forward with EVERY opt-in path switched on, in which the kernels that
have never run on hardware execute -- as compiled from their sources -- on the CPU emulator (tests/emu), inside the real
host flow (strided conv outputs into GroupNorm, the cached zero-bordered operand buffers, operands carried between encoder
layers, attention-mask bits into the tcgen05 cross-attention, fp16 hi|lo mask features into the cluster einsum, pooled
masks), against the same model with pure oracle operators.

Backend of this test: `oracle_ops("tf32x3")` (exact CPU GEMMs over the split-operand plumbing) with the following
operators handed back to the real `univs_b200.ops` wrappers bound to the emulated library: MSDeformAttn (tiled + fused
biases), channel-last GroupNorm + FPN glue, PatchMerging gather-LayerNorm, frame ingest, multi-consumer LayerNorm, pooled
mask features + direct mask bits, tcgen05 window attention (window 12), tcgen05 cross- / self-attention core, cluster
einsum.  What stays on the oracle: the kernels validated on the B200 (row-wise LayerNorm / GELU / split -- the emulator
runs them in tests/test_kernels_cpu_emulation.py, here their thousands of tiny blocks would only cost time; the mma.sync /
cp.async kernels, out of the emulator's reach) and ProCA."""
import os
import shutil
import time

import pytest
import torch

from oracle import cpu_backend
from oracle.cpu_backend import oracle_ops
from tests import model_factory as mf
from tests.emu import build_emu
from tests.test_kernels_cpu_emulation import _Dev, _load, plain
from univs_b200 import _cabi, nn_ops, ops
from univs_b200.meta_arch import UniVS_Prompt
from univs_b200.modeling.head import MaskFormerHead
from univs_b200.registry import ShapeSpec

if shutil.which("g++") is None or not os.path.exists(os.path.join(build_emu.CUDA_INCLUDE, "cuda_runtime.h")):
    pytest.skip("needs g++ and the CUDA headers", allow_module_level=True)

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
EMULATED = ("ms_deform_attn_encoder", "groupnorm_cl", "layernorm_multi", "layernorm_merge2x2", "patchify_normalize",
            "mask_feature_pool", "attn_mask_bits_direct", "attn_mask_bits", "swin_window_attention",
            "swin_window_attention_operand", "mha_core", "mask_einsum", "prepare_mask_features")


def _as_dev(x):
    return torch.Tensor._make_subclass(_Dev, x if x.is_contiguous() or x.dim() == 4 else x.contiguous()) if torch.is_tensor(x) and not isinstance(x, _Dev) else x


def _model(T, Q, swin):
    parts = mf.build_product_model(swin, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(), enc_layers=2, dec_layers=3)
    mf.load_keyed(parts)
    E = swin["embed_dim"]
    shapes = {f"res{i + 2}": ShapeSpec(channels=E * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=parts[1], transformer_predictor=parts[2])
    return UniVS_Prompt(backbone=parts[0], sem_seg_head=head, pixel_mean=MEAN, pixel_std=STD)


def _rel(a, b):
    a, b = plain(a).float(), plain(b).float()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _cpu_fp16_gemm(monkeypatch):
    """torch.addmm(..., out_dtype=float32) on fp16 operands is a CUDA-only library entry; the fp16x3 policy needs it.
    CPU stand-in for this test: products of fp16 values are exact in fp32, accumulation in fp32 -- what the tensor core does
    up to the order of the sum."""
    real = torch.addmm

    def addmm(inp, a, b, *, beta=1, alpha=1, out_dtype=None, out=None):
        if out_dtype is None:
            return real(inp, a, b, beta=beta, alpha=alpha) if out is None else real(inp, a, b, beta=beta, alpha=alpha, out=out)
        r = alpha * (plain(a).float() @ plain(b).float())
        if beta != 0:
            r = r + beta * plain(inp).float()
        if out is not None:
            out.copy_(r)
            return out
        return r
    monkeypatch.setattr(torch, "addmm", addmm)


@pytest.mark.parametrize("policy", ["fp16x3"])        # the production policy ("tf32x3" works too: same plumbing, fp32 operands)
def test_clip_forward_with_all_switches_on_the_emulator(monkeypatch, tmp_path, policy):
    """fp16x3 is the production default: fp16 operand formats everywhere (window attention writes the projection GEMM's
    operand itself, GroupNorm writes the padded fp16 operand of the 3x3 taps, the MSDeformAttn kernel the output_proj one)"""
    if policy == "fp16x3":
        _cpu_fp16_gemm(monkeypatch)
        monkeypatch.setattr(nn_ops, "_inplace16", [None])
    lib_path = str(tmp_path / "libunivs_emu_model.so")
    shutil.copy(build_emu.build(), lib_path)
    monkeypatch.setenv("UNIVS_EMU_SMS", "4")
    T, Q = 2, 10
    swin = dict(embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], window_size=12)
    model = _model(T, Q, swin)
    g = torch.Generator().manual_seed(12)
    frames = (torch.rand(T, 3, 60, 90, generator=g) * 255).round()            # pads to 64 x 96: token grids 16x24 ... 2x3
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}]

    # reference: the default host path with oracle operators (what the reference-parity tests pin)
    with oracle_ops("fp32"):
        want = model.clip_forward(frames, tg())

    real = {k: getattr(ops, k) for k in EMULATED}                 # the real wrappers, before the oracle context patches them
    calls = {k: 0 for k in EMULATED}
    seconds = {}

    def bound(name):
        fn = real[name]

        def call(*a, **k):
            calls[name] += 1
            t0 = time.perf_counter()
            try:
                return fn(*[_as_dev(x) for x in a], **{kk: _as_dev(v) for kk, v in k.items()})
            finally:
                seconds[name] = seconds.get(name, 0.0) + time.perf_counter() - t0
        return call

    dec = model.sem_seg_head.predictor
    chk = ops._chk
    monkeypatch.setattr(_cabi, "_lib", _load(lib_path))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_chk", lambda t, name, dtype=torch.float32: chk(_as_dev(t), name, dtype))
    monkeypatch.setattr(ops, "_chk_t", lambda t, name, dtype=torch.float32: (chk(_as_dev(t), name, dtype), _as_dev(t))[1])
    monkeypatch.setattr(ops, "_ws_cache", {})
    monkeypatch.setattr(ops, "_gn_ws", {})
    monkeypatch.setattr(nn_ops, "_pad_cache", {})
    monkeypatch.setattr(ops, "_win_tc", 2)                # tcgen05 window attention (production version: compact operand)
    monkeypatch.setattr(ops, "_mha_tc", 1)                # tcgen05 attention core ...
    monkeypatch.setattr(ops, "MHA_TC_MIN_KEYS", 1)        # ... for the small key counts of this geometry too
    monkeypatch.setattr(ops, "_msda_tile", 8)             # tiled MSDeformAttn
    monkeypatch.setattr(ops, "_einsum_mc", 1)             # cluster einsum
    monkeypatch.setattr(ops, "_einsum_mode", "f16x3")
    monkeypatch.setattr(dec, "pooled_masks", True)        # intermediate heads from pooled mask features
    nn_ops.set_fused_glue(True)
    try:
        with oracle_ops(policy):
            for name in EMULATED:         # inside the context (which restores every operator on exit): not via monkeypatch,
                setattr(ops, name, bound(name))       # whose teardown would put the oracle functions back for good
            got = model.clip_forward(frames, tg())
    finally:
        nn_ops.set_fused_glue(False)

    print("emulated operators: calls", calls, "seconds", {k: round(v, 1) for k, v in seconds.items()})
    assert calls["swin_window_attention_operand" if policy == "fp16x3" else "swin_window_attention"] > 0
    for name in ("mha_core", "mask_einsum", "ms_deform_attn_encoder", "groupnorm_cl", "layernorm_multi",
                 "layernorm_merge2x2", "patchify_normalize", "mask_feature_pool", "attn_mask_bits_direct"):
        assert calls[name] > 0, f"{name} was not exercised"
    assert plain(got["pred_masks"]).shape == want["pred_masks"].shape == (1, Q, T, 16, 24)
    for k in ("pred_masks", "pred_logits", "pred_embds"):
        assert not torch.isnan(plain(got[k])).any()
        assert _rel(got[k], want[k]) < 1e-3, (k, _rel(got[k], want[k]))


DEFAULT_PATH = ("layernorm", "gelu", "relu", "split_operand", "split_tf32", "round_tf32", "ms_deform_attn_encoder", "swin_window_attention",
                "swin_window_attention_operand", "mha_core", "mask_einsum", "prepare_mask_features", "attn_mask_bits", "proca_core")


def test_default_path_clip_forward_on_the_emulator(monkeypatch, tmp_path):
    """The DEFAULT path (no switch on) under the production fp16x3 policy with EVERY operator running from its kernel source
    on the emulator -- the host code that changed after the last GPU run of round 1 (K/V operands prepared once per level,
    the MLP helper, encoder layers returning carries, scratch slots) driving the real wrappers and the real kernels.
    Category prompts (ProCA) included."""
    _cpu_fp16_gemm(monkeypatch)
    monkeypatch.setattr(nn_ops, "_inplace16", [None])
    monkeypatch.setattr(nn_ops, "_gemm_tc", False)      # the library-GEMM formulation: row-wise GELU / ReLU / split kernels in use
    monkeypatch.setattr(nn_ops, "_fused_glue", False)
    lib_path = str(tmp_path / "libunivs_emu_default.so")
    shutil.copy(build_emu.build(), lib_path)
    T, Q = 2, 6
    swin = dict(embed_dim=32, depths=[1, 1, 1, 1], num_heads=[1, 2, 4, 8], window_size=4)
    parts = mf.build_product_model(swin, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(), enc_layers=1, dec_layers=2)
    mf.load_keyed(parts)
    parts[2].pooled_masks = False           # this test binds the round-1 operator set only (pooled heads: the test below)
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=parts[1], transformer_predictor=parts[2])
    model = UniVS_Prompt(backbone=parts[0], sem_seg_head=head, pixel_mean=MEAN, pixel_std=STD)
    g = torch.Generator().manual_seed(13)
    frames = (torch.rand(T, 3, 32, 64, generator=g) * 255).round()
    tg = lambda: [{"task": "detection", "dataset_name": "bdd_track", "prompt_type": "text"}]        # 8 category prompts -> ProCA
    with oracle_ops("fp32"):
        want = model.clip_forward(frames, tg())
    real = {k: getattr(ops, k) for k in DEFAULT_PATH}
    calls = {k: 0 for k in DEFAULT_PATH}

    def bound(name, oracle_fn):
        fn = real[name]

        def call(*a, **k):
            if name == "split_operand" and a[0].numel() > 60000:
                # a big weight matrix (the 3938 x 640 class embeddings): one host thread per CUDA thread would make this one
                # split cost 600k thread creations -- the oracle splits it, the kernel keeps the activations
                return oracle_fn(*a, **k)
            calls[name] += 1
            return fn(*[_as_dev(x) for x in a], **{kk: _as_dev(v) for kk, v in k.items()})
        return call

    chk = ops._chk
    monkeypatch.setattr(_cabi, "_lib", _load(lib_path))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_chk", lambda t, name, dtype=torch.float32: chk(_as_dev(t), name, dtype))
    monkeypatch.setattr(ops, "_ws_cache", {})
    monkeypatch.setattr(ops, "_einsum_mode", "f16x3")
    monkeypatch.setattr(nn_ops, "_wcache", {})
    with oracle_ops("fp16x3"):
        for name in DEFAULT_PATH:
            setattr(ops, name, bound(name, getattr(ops, name)))
        got = model.clip_forward(frames, tg())
    for name in ("layernorm", "gelu", "split_operand", "ms_deform_attn_encoder", "swin_window_attention_operand", "mha_core",
                 "mask_einsum", "attn_mask_bits", "proca_core"):
        assert calls[name] > 0, f"{name} was not exercised"
    assert plain(got["pred_masks"]).shape == want["pred_masks"].shape == (1, Q + 8, T, 8, 16)
    for k in ("pred_masks", "pred_logits", "pred_embds"):
        assert not torch.isnan(plain(got[k])).any()
        assert _rel(got[k], want[k]) < 1e-3, (k, _rel(got[k], want[k]))
