"""COCO run-length encoding of result masks with the scan on the device (result formats).

The reference heads move every boolean mask to the host and call `pycocotools.mask.encode` there
(inference_video_vis.py:526-531, inference_video_entity.py:943-947) -- H*W bytes per mask over PCIe.  Here the run
boundaries are found on the device (a transposed view, one comparison, one compaction), so only the boundary positions
travel (a few hundred int32 per mask); the host turns them into the `counts` string.  Format: pycocotools
`common/maskApi.c` (rleEncode + rleToString): column-major runs starting with the 0-run; string = delta against
counts[i-2] for i > 2, 5 data bits + continuation bit per character, offset 48.
Plain torch ops: works on CPU tensors too (that is how the tests check it against oracle/rle_ref.py)."""
from __future__ import annotations

import numpy as np
import torch


@torch.no_grad()
def run_boundaries(masks):
    """masks [N, H, W] bool / uint8 (any device) -> (positions int64 [K] on the HOST, offsets int64 [N+1] on the host):
    positions[offsets[n]:offsets[n+1]] are the column-major indices j (1 <= j < H*W) where mask n changes value,
    preceded by a 0 if the mask starts with a one (so that the first run is the zero run, of length 0)."""
    N, H, W = masks.shape
    flat = masks.to(torch.bool).transpose(1, 2).reshape(N, H * W)            # column-major scan order
    change = torch.ones((N, H * W), dtype=torch.bool, device=masks.device)
    change[:, 1:] = flat[:, 1:] != flat[:, :-1]
    change[:, 0] = flat[:, 0]                                                 # a leading one opens with an empty zero run
    idx = change.nonzero()                                                    # [K, 2] sorted by (mask, position)
    per_mask = torch.bincount(idx[:, 0], minlength=N)
    offsets = torch.zeros(N + 1, dtype=torch.int64)
    offsets[1:] = per_mask.cumsum(0).cpu()
    return idx[:, 1].cpu(), offsets


def counts_to_string(cnts):
    """maskApi.c rleToString on a sequence of run lengths"""
    out = bytearray()
    cnts = [int(c) for c in cnts]
    for i, x in enumerate(cnts):
        if i > 2:
            x -= cnts[i - 2]
        more = True
        while more:
            c = x & 0x1F
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return out.decode("ascii")


def encode(masks):
    """masks [N, H, W] (device or host) -> list of {"size": [H, W], "counts": str}, equal to
    [pycocotools.mask.encode(np.asfortranarray(m[:, :, None]))[0] with counts decoded to str for m in masks]."""
    N, H, W = masks.shape
    pos, off = run_boundaries(masks)
    pos = pos.numpy()
    out = []
    for n in range(N):
        p = pos[off[n]:off[n + 1]]
        edges = np.concatenate([[0], p, [H * W]]) if (len(p) == 0 or p[0] != 0) else np.concatenate([p, [H * W]])
        cnts = np.diff(edges) if (len(p) == 0 or p[0] != 0) else np.concatenate([[0], np.diff(edges)])
        out.append({"size": [int(H), int(W)], "counts": counts_to_string(cnts)})
    return out
