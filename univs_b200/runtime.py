"""CUDA-graph execution of the per-clip forward.

The prompt-free detection clip (the north-star workload) has static shapes and static host control flow, so the
whole forward -- ~1000 kernel launches: library GEMMs, LayerNorms and the hand-written kernels -- is captured once
into a CUDA graph and replayed per clip; host launch overhead disappears from the critical path.  Paths with
data-dependent host logic (visual-prompt sampling) run eagerly."""
from __future__ import annotations

import torch


class GraphedClip:
    def __init__(self, model, frames_example, targets_factory, warmup=2):
        self.model = model
        self.targets_factory = targets_factory
        self.static_frames = torch.empty_like(frames_example, device=model.device)
        self.static_frames.copy_(frames_example)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                model.clip_forward(self.static_frames, targets_factory())
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._targets = targets_factory()
        with torch.cuda.graph(self.graph):
            self.out = model.clip_forward(self.static_frames, self._targets)

    def __call__(self, frames):
        """frames: host (pinned) or device tensor of the captured shape/dtype.  Returns the static output dict
        (overwritten by the next call)."""
        self.static_frames.copy_(frames, non_blocking=True)
        self.graph.replay()
        return self.out
