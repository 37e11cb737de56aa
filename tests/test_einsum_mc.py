"""Cluster / TMA-multicast mask einsum (csrc/mask_einsum_mc.cu) against the validated tcgen05 kernel and the oracle.
Written without GPU time left in round 1: opt-in (UNIVS_GPU_EINSUM_MC=1) until it has run on a B200; the protocol is
simulated on the CPU in tests/test_einsum_mc_protocol.py."""
import os

import pytest
import torch

from oracle import ops_ref

_gate = pytest.mark.filterwarnings("default")      # validated on a B200 (round 2): no gate


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.fixture()
def cluster_ops():
    from univs_b200 import ops
    old = ops._einsum_mc
    yield ops
    ops._einsum_mc = old


@pytest.mark.gpu
@_gate
@pytest.mark.parametrize("T,Q,C,HW", [(1, 20, 256, 64 * 64), (2, 100, 256, 30 * 54), (3, 200, 256, 46 * 80),
                                        (1, 232, 256, 130), (2, 17, 64, 66), (1, 256, 64, 2), (1, 32, 64, 128),
                                        (5, 200, 256, 1000)])     # the last: more tiles than clusters x 2, frame changes
def test_cluster_einsum_equals_the_one_cta_kernel(cluster_ops, T, Q, C, HW):
    ops = cluster_ops
    torch.manual_seed(13)
    E, F = torch.randn(T, Q, C), torch.randn(T, HW, C)
    F16 = ops.prepare_mask_features(F.cuda(), "f16x3")
    ops._einsum_mc = 0
    base = ops.mask_einsum(E.cuda(), F16, mode="f16x3")
    ops._einsum_mc = 1
    got = ops.mask_einsum(E.cuda(), F16, mode="f16x3")
    torch.cuda.synchronize()
    # same operands, same MMA terms in the same order per output element: the results are bit-identical
    assert torch.equal(got, base)
    want = ops_ref.mask_einsum(E.double(), F.transpose(1, 2).double()).float()
    assert _rel(got, want) < 5e-6


@pytest.mark.gpu
@_gate
def test_cluster_einsum_full_size_checksum(cluster_ops):
    """north-star shape: equality with the validated kernel on every element + untouched guard rows around the output"""
    ops = cluster_ops
    T, Q, C, HW = 5, 200, 256, 184 * 320
    g = torch.Generator(device="cuda").manual_seed(5)
    E = torch.randn(T, Q, C, device="cuda", generator=g)
    F16 = ops.prepare_mask_features(torch.randn(T, HW, C, device="cuda", generator=g), "f16x3")
    guard = torch.full((Q + 2, T, HW), 7.0, device="cuda")
    ops._einsum_mc = 0
    base = ops.mask_einsum(E, F16, mode="f16x3")
    ops._einsum_mc = 1
    ops.mask_einsum(E, F16, out=guard[1:Q + 1], mode="f16x3")
    torch.cuda.synchronize()
    assert torch.equal(guard[1:Q + 1], base)
    assert bool((guard[0] == 7.0).all()) and bool((guard[Q + 1] == 7.0).all())


def test_cluster_entry_rejects_bad_arguments_without_a_gpu():
    """argument validation happens before any CUDA call"""
    from univs_b200._cabi import lib
    l = lib()
    f = l.univs_mask_einsum_f16x3_cluster
    assert f(None, 16, 16, 1, 16, 256, 128, 16) != 0            # Q <= 16: nothing to split
    assert b"16 queries" in l.univs_b200_last_error()
    assert f(None, 16, 16, 1, 200, 100, 128, 16) != 0           # C % 32
    assert f(None, 8, 16, 1, 200, 256, 128, 16) != 0            # misaligned operand
    assert f(None, None, None, 0, 200, 256, 128, None) == 0     # empty problem
