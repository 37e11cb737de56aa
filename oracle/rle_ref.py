"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the COCO run-length mask codec the reference's result writers call:
`pycocotools.mask.encode` (inference_video_vis.py:526-531, inference_video_entity.py:943-947, comm.py:119).

pycocotools (cocodataset/cocoapi, PythonAPI; the reference does not pin a version and does not vendor it) is not
installed in this image, so this file restates the published algorithm of its `common/maskApi.c`:
  rleEncode   : column-major scan of a binary mask, run lengths of alternating 0s / 1s starting with the 0-run
  rleToString : counts -> ASCII, delta against counts[i-2] for i > 2, 5 data bits + continuation bit per character,
                sign-extended, characters offset by 48
  rleFrString : the inverse.
PARITY UNPINNED for the compressed string form (no pycocotools here to generate vectors); the uncompressed counts are
pinned against the two examples of the COCO API documentation (mask.py docstring: M=[0 0 1 1 1 0 1] -> [2 3 1 1],
M=[1 1 1 1 1 1 0] -> [0 6 1]) in tests/test_rle.py, and encode/decode must round-trip.  Pure-Python loops: small cases only."""
import numpy as np


def rle_counts(mask):
    """mask [h, w] (any integer / bool dtype) -> list of run lengths (column-major order, first run counts zeros)"""
    t = np.asarray(mask).astype(np.uint8).flatten(order="F")
    cnts, c, p = [], 0, 0
    for v in t:
        if v != p:
            cnts.append(c)
            c, p = 0, v
        c += 1
    cnts.append(c)
    return cnts


def rle_to_string(cnts):
    out = []
    for i, x in enumerate(cnts):
        x = int(x)
        if i > 2:
            x -= int(cnts[i - 2])
        more = True
        while more:
            c = x & 0x1F
            x >>= 5                      # arithmetic shift: Python ints behave like C's signed long here
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return "".join(out)


def rle_from_string(s):
    cnts, p = [], 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x)
    return cnts


def encode(mask):
    """pycocotools.mask.encode for one [h, w] mask -> {"size": [h, w], "counts": str}"""
    h, w = np.asarray(mask).shape
    return {"size": [int(h), int(w)], "counts": rle_to_string(rle_counts(mask))}


def decode(rle):
    h, w = rle["size"]
    cnts = rle_from_string(rle["counts"])
    flat = np.zeros(h * w, np.uint8)
    pos, v = 0, 0
    for c in cnts:
        flat[pos:pos + c] = v
        pos += c
        v ^= 1
    return flat.reshape((h, w), order="F")
