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
    """maskApi.c rleToString on a sequence of run lengths: value i > 2 is coded as the difference to value i - 2; every value
    as little-endian groups of 5 data bits + a continuation bit, offset 48.  Vectorised over the runs (a noisy mask has
    hundreds of thousands of them; the per-run Python loop made the RLE output 7x slower than copying dense masks)."""
    c = np.asarray(cnts, dtype=np.int64).reshape(-1)
    if c.size == 0:
        return ""
    x = c.copy()
    if c.size > 3:
        x[3:] -= c[1:-2]
    chars = np.zeros((c.size, 13), dtype=np.uint8)          # 64-bit values need at most 13 groups
    used = np.zeros(c.size, dtype=np.int64)
    alive = np.ones(c.size, dtype=bool)
    for k in range(13):
        g = x & 0x1F
        x = x >> 5                                           # arithmetic shift, like the C code on a signed long
        more = np.where((g & 0x10) != 0, x != -1, x != 0)
        ch = (g | np.where(more, 0x20, 0)) + 48
        chars[alive, k] = ch[alive]
        used[alive] += 1
        alive &= more
        if not alive.any():
            break
    return chars[np.arange(13)[None, :] < used[:, None]].tobytes().decode("ascii")


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
