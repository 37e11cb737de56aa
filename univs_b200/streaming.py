"""Sliding-window video inference over the hot path with per-frame feature reuse.

Every UniVS task head slides a T-frame clip over the video with stride 1 and calls, per clip,
`model.backbone(window)` (cached per NUM_FRAMES_WINDOW chunk) and `model.sem_seg_head(features)`
(inference_video_vis_fast.py:223-236) -- so each frame goes through the 6-layer MSDeformAttn pixel decoder T times.
Backbone and pixel decoder are frame-independent (frames are the batch dimension; LayerNorm / GroupNorm / window
attention / deformable sampling never mix frames -- the property frame sharding relies on, tests/test_sharding_gloo.py),
so `ClipStream` runs them ONCE per frame, keeps the per-frame {mask_features, three multi-scale maps} of the last T
frames in a ring, and runs only the decoder per clip.  Results equal the per-clip recomputation up to GEMM
batch-size rounding noise (~1e-6).  The decoder, the prompt memory pool in `targets` and the callers' tracking logic are
untouched: `clip(start, targets)` returns exactly what `sem_seg_head(features_of_clip, targets=targets)` returns.
"""
from __future__ import annotations

from collections import OrderedDict

import torch


class ClipStream:
    def __init__(self, model, num_frames: int, max_cached_frames: int | None = None):
        self.model = model
        self.T = num_frames
        self.capacity = max_cached_frames or (2 * num_frames)
        self._cache = OrderedDict()          # frame index -> (mask_features [1,C,h,w], [ms_1/32, ms_1/16, ms_1/8])
        self.image_size = None
        self.frames_encoded = 0
        self._floor = 0                      # first frame a future clip may still need (clips move forward)

    @torch.no_grad()
    def push(self, first_index: int, frames):
        """Encode frames [first_index, first_index + k) (k >= 1; any k -- a chunk is one backbone / pixel-decoder batch)."""
        from . import nn_ops
        if nn_ops.fused_glue() and hasattr(self.model, "backbone_from_frames"):      # ingest fused into the patch embedding
            if isinstance(frames, (list, tuple)):
                frames = torch.stack([f.to(self.model.device) for f in frames])
            self.image_size = tuple(frames.shape[-2:])
            self._store(first_index, self.model.backbone_from_frames(frames))
            return
        x, self.image_size = self.model.preprocess(frames)
        self.push_preprocessed(first_index, x)

    @torch.no_grad()
    def push_preprocessed(self, first_index: int, x):
        """Same for frames that are already normalised and padded ([k,3,Hp,Wp] float32 on the device) -- what the
        task heads hold after ImageList.from_tensors (inference_video_vis_fast.py:198-207)."""
        self._store(first_index, self.model.backbone(x))

    def _store(self, first_index: int, feats):
        mf, _bfe, _enc, ms = self.model.sem_seg_head.pixel_decoder.forward_features(feats)
        for j in range(mf.shape[0]):
            self._cache[first_index + j] = (mf[j:j + 1], [m[j:j + 1] for m in ms])
            self.frames_encoded += 1
        # `capacity` is a soft bound: a frame at or after the start of the oldest clip still to come is never evicted
        # (callers push NUM_FRAMES_WINDOW_TEST frames at a time, which may exceed 2*T: reference configs T=3 / W=5)
        while len(self._cache) > self.capacity and next(iter(self._cache)) < self._floor:
            self._cache.popitem(last=False)

    @torch.no_grad()
    def clip(self, start: int, targets, length: int | None = None):
        """Decoder output for frames [start, start + T) (or a shorter last clip of `length` frames); every frame must
        have been pushed."""
        n = self.T if length is None else length
        idx = range(start, start + n)
        missing = [i for i in idx if i not in self._cache]
        if missing:
            raise KeyError(f"frames {missing} are not cached (push them first / raise max_cached_frames)")
        mf = torch.cat([self._cache[i][0] for i in idx], 0)
        ms = [torch.cat([self._cache[i][1][l] for i in idx], 0) for l in range(3)]
        # keep the channel-last storage the decoder consumes without a copy
        mf = mf.contiguous(memory_format=torch.channels_last)
        self._floor = max(self._floor, start)
        while len(self._cache) > self.capacity and next(iter(self._cache)) < self._floor:
            self._cache.popitem(last=False)
        tg = targets[0]
        tg["frame_indices"] = torch.arange(start, start + n, device=mf.device)
        tg.setdefault("num_frames", self.T)
        return self.model.sem_seg_head.predictor(ms, mf, mf, None, targets)

    @torch.no_grad()
    def run(self, video_frames, make_targets, stride: int = 1, chunk: int | None = None):
        """Generator over (start, decoder_output) for a whole video tensor [V,3,H,W] (host or device).
        `make_targets(start)` returns the targets list for the clip (callers keep their own state in it)."""
        V = video_frames.shape[0]
        chunk = chunk or self.T
        pushed = 0
        for start in range(0, V - self.T + 1, stride):
            while pushed < start + self.T:
                k = min(chunk, V - pushed)
                self.push(pushed, video_frames[pushed:pushed + k])
                pushed += k
            yield start, self.clip(start, make_targets(start))
