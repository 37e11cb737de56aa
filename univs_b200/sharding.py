"""Frame sharding of a clip across the GPUs of one node (new capability, not in the reference).

Backbone and pixel decoder treat frames as the batch dimension with no cross-frame operation, so rank r owns frames
{f : f mod n == r}.  The only exchange step is one all-gather of the per-frame {mask_features, 3 multi-scale maps}
before the decoder.  Ragged clips (T not a multiple of n) are padded to ceil(T/n) frames per rank for the collective;
padding frames are zeros and are dropped after the gather.  Works with NCCL (GPU) and gloo (CPU tests).

`TokenExchange` is the cheaper alternative of SURVEY.md 8e: cross-attention, mask einsum and attention-mask generation
are per-frame too, so the decoder can stay frame-sharded and only the [Q,256] query tokens of each frame travel (one
small all-gather per decoder layer for the Q*T self-attention, one for the T-mean of the class logits, one for the final
mask logits) instead of 80 MB of features per frame."""
from __future__ import annotations

import torch
import torch.distributed as dist


def frame_plan(num_frames: int, world_size: int):
    """-> list over ranks of the frame indices each rank owns (round robin)."""
    return [list(range(r, num_frames, world_size)) for r in range(world_size)]


class FrameSharder:
    def __init__(self, process_group=None):
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if process_group is not None else 1
        self.rank = dist.get_rank(process_group) if process_group is not None else 0
        self._order = {}

    def local_frames(self, x):
        if self.rank >= x.shape[0]:     # an idle rank still takes part in the collective with a padding frame
            return x[:0]
        # round robin = a strided slice: no index tensor, hence no host->device copy (the step is captured in a CUDA graph)
        return x[self.rank::self.world_size].contiguous()

    def all_gather_frames(self, tensors, num_frames):
        """tensors: list of [T_local, ...] (same trailing shapes on every rank).  Returns the list of [T, ...]
        tensors in global frame order.  One fused collective: all tensors are packed into one flat buffer."""
        n = self.world_size
        per_rank = (num_frames + n - 1) // n
        plan = frame_plan(num_frames, n)
        trailing = [t.shape[1:] for t in tensors]
        numels = [int(torch.Size(s).numel()) for s in trailing]
        frame_numel = sum(numels)
        dev, dt = tensors[0].device, tensors[0].dtype
        send = torch.zeros(per_rank * frame_numel, device=dev, dtype=dt)
        t_local = tensors[0].shape[0]
        if t_local:
            view = send[: t_local * frame_numel].view(t_local, frame_numel)
            off = 0
            for t, ne in zip(tensors, numels):
                view[:, off:off + ne] = t.reshape(t_local, ne)
                off += ne
        recv = torch.empty(n * per_rank * frame_numel, device=dev, dtype=dt)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        recv = recv.view(n, per_rank, frame_numel)
        key = (num_frames, n, str(dev))
        order = self._order.get(key)
        if order is None:                      # cached on the device: no host->device copy in the (graph-captured) forward
            idx = [0] * num_frames
            for r, frames in enumerate(plan):
                for j, f in enumerate(frames):
                    idx[f] = r * per_rank + j
            order = torch.tensor(idx, dtype=torch.long, device=dev)
            self._order[key] = order
        flat = recv.view(n * per_rank, frame_numel)[order]
        outs, off = [], 0
        for s, ne in zip(trailing, numels):
            outs.append(flat[:, off:off + ne].reshape(num_frames, *s))
            off += ne
        return outs


class TokenExchange:
    """Frame-sharded decoder plumbing.  Rank r keeps frames {f : f mod n == r} through the decoder; `gather` reassembles a
    per-frame tensor [T_local, ...] to [T, ...] in global frame order on every rank, `local` selects this rank's frames
    again.  With more ranks than frames a rank without a frame shadows frame (rank mod T): it computes like its owner
    (so every rank runs the same program and reaches every collective) but contributes nothing to the gathers."""

    def __init__(self, sharder: FrameSharder, num_frames: int):
        own = frame_plan(num_frames, sharder.world_size)[sharder.rank]
        self.sharder = sharder
        self.num_frames = num_frames
        self.shadow = not own
        self.frames = own if own else [sharder.rank % num_frames]
        self._idx = {}

    def frames_tensor(self, device):
        key = str(device)
        if key not in self._idx:
            self._idx[key] = torch.tensor(self.frames, dtype=torch.long, device=device)
        return self._idx[key]

    def gather(self, x):
        return self.sharder.all_gather_frames([x[:0] if self.shadow else x], self.num_frames)[0]

    def local(self, x):
        return x.index_select(0, self.frames_tensor(x.device))
