"""Sine position encodings (host side, cached per shape).
2-D: mask2former/modeling/transformer_decoder/position_encoding.py:29-52 (normalize=True).
3-D ArbitraryT: univs/modeling/transformer_decoder/position_encoding.py:142-169 (z = frame_index/128*2pi)."""
from __future__ import annotations

import math

import torch

_cache = {}


def _dim_t(n, temperature, device):
    i = torch.arange(n, dtype=torch.float32, device=device)
    return temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / n)


def sine_2d(h, w, device, num_pos_feats=128, temperature=10000.0):
    """-> [h*w, 2*num_pos_feats] (token-major)"""
    key = ("2d", h, w, str(device), num_pos_feats)
    if key not in _cache:
        eps, scale = 1e-6, 2 * math.pi
        y = torch.arange(1, h + 1, dtype=torch.float32, device=device) / (h + eps) * scale
        x = torch.arange(1, w + 1, dtype=torch.float32, device=device) / (w + eps) * scale
        dt = _dim_t(num_pos_feats, temperature, device)
        py = y[:, None] / dt
        px = x[:, None] / dt
        py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), 2).flatten(1)    # [h, npf]
        px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), 2).flatten(1)    # [w, npf]
        pos = torch.cat((py[:, None, :].expand(h, w, -1), px[None, :, :].expand(h, w, -1)), 2)
        _cache[key] = pos.reshape(h * w, 2 * num_pos_feats).contiguous()
    return _cache[key]


def temporal_sine(frame_indices, device, num_pos_feats=128, temperature=10000.0, num_max_frames=128):
    """The z (frame) term of the 3-D encoding -> [T, 2*num_pos_feats].  It does not depend on the level: the decoder
    computes it once per clip and shares it between the three memory levels.  Frame indices that live on the host (or the
    default 0..T-1, passed as an int) key a cache, so a steady stream of clips launches nothing for it."""
    key = None
    if isinstance(frame_indices, int):
        key = ("z", tuple(range(frame_indices)), str(device), num_pos_feats)
    elif not frame_indices.is_cuda:
        key = ("z", tuple(int(v) for v in frame_indices.tolist()), str(device), num_pos_feats)
    if key is not None and key in _cache:
        return _cache[key]
    if isinstance(frame_indices, int):
        frame_indices = torch.arange(frame_indices)
    z = frame_indices.to(device=device, dtype=torch.float32) / num_max_frames * (2 * math.pi)
    dz = _dim_t(2 * num_pos_feats, temperature, device)
    pz = z[:, None] / dz
    pz = torch.stack((pz[:, 0::2].sin(), pz[:, 1::2].cos()), 2).flatten(1)        # [T, C]
    if key is not None:
        if len(_cache) > 256:           # sliding-window heads walk through many index tuples
            for k in [k for k in _cache if k[0] == "z"]:
                del _cache[k]
        _cache[key] = pz
    return pz


def sine_3d_arbitrary_t(frame_indices, h, w, device, num_pos_feats=128, temperature=10000.0, num_max_frames=128, pz=None):
    """-> [T, h*w, 2*num_pos_feats]; `pz` = temporal_sine(frame_indices, ...) when the caller already has it"""
    base = sine_2d(h, w, device, num_pos_feats, temperature)
    if pz is None:
        pz = temporal_sine(frame_indices, device, num_pos_feats, temperature, num_max_frames)
    return base[None] + pz[:, None, :]


def sine_3d_points(xy, t_indices, device, num_pos_feats=128, temperature=10000.0, num_max_frames=128):
    """position_encoding.py:191-236 (normalize=True): xy [n,2] normalised (x,y); t_indices [T] -> [T, n, C]"""
    scale = 2 * math.pi
    dt = _dim_t(num_pos_feats, temperature, device)
    dz = _dim_t(2 * num_pos_feats, temperature, device)
    px = (xy[:, 0] * scale)[:, None] / dt
    py = (xy[:, 1] * scale)[:, None] / dt
    pz = (t_indices.to(device=device, dtype=torch.float32) / num_max_frames * scale)[:, None] / dz
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), -1).flatten(-2)
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), -1).flatten(-2)
    pz = torch.stack((pz[:, 0::2].sin(), pz[:, 1::2].cos()), -1).flatten(-2)
    return torch.cat((py, px), -1)[None] + pz[:, None, :]
