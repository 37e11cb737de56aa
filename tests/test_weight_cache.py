"""The split-weight cache of nn_ops must never serve another tensor's operands: the allocator reuses a freed model's
addresses for the next model of the same architecture (same shape, same _version, different weights).
(nn_ops.py:72)"""
import gc

import pytest
import torch
import torch.nn.functional as F

from oracle.cpu_backend import oracle_ops
from univs_b200 import nn_ops


@pytest.mark.parametrize("policy", ["tf32x3"])      # the fp16 GEMM has no CPU implementation; the cache code is shared
def test_cache_entry_dies_with_its_parameter(policy):
    old = nn_ops.policy()
    try:
        with oracle_ops(policy):
            x = torch.randn(5, 256)
            worst, reused, seen = 0.0, 0, set()
            for seed in range(8):       # freed storage is handed out again: same data_ptr, fresh weights
                torch.manual_seed(seed)
                lin = torch.nn.Linear(256, 192)
                reused += lin.weight.data_ptr() in seen
                seen.add(lin.weight.data_ptr())
                y = nn_ops.linear(x, lin.weight, lin.bias)
                worst = max(worst, float((y - F.linear(x, lin.weight, lin.bias)).detach().abs().max()))
                del lin, y
                gc.collect()
            if not reused:
                pytest.skip("the allocator never reused an address in this run")
            assert worst < 1e-3, worst
    finally:
        nn_ops.set_policy(old)


@pytest.mark.parametrize("policy", ["tf32x3"])      # the fp16 GEMM has no CPU implementation; the cache code is shared
def test_views_of_a_packed_parameter_hit_and_follow_updates(policy):
    old = nn_ops.policy()
    try:
        with oracle_ops(policy):
            p = torch.nn.Parameter(torch.randn(96, 32))
            x = torch.randn(3, 32)
            y0 = nn_ops.linear(x, p[:32])
            n = len(nn_ops._wcache)
            y1 = nn_ops.linear(x, p[:32])                 # a fresh view object of the same parameter: a hit
            assert len(nn_ops._wcache) == n and torch.equal(y0, y1)
            with torch.no_grad():
                p.mul_(2.0)                                # in-place update bumps _version
            y2 = nn_ops.linear(x, p[:32])
            assert float((y2 - 2 * y0).abs().max()) < 1e-4
            conv = torch.nn.Conv2d(8, 8, 3, padding=1)
            t0 = nn_ops._conv_weight_taps(conv.weight)
            assert nn_ops._conv_weight_taps(conv.weight) is t0
    finally:
        nn_ops.set_policy(old)
