"""Frame-level data parallelism for the pillar encoder.

Frames are independent units: the batch index is the most significant part of the merged key
(``dynamic_pillar_vfe.py:208``), so no pillar spans two frames and a rank can encode its frames with no
exchange of points, keys or pillars.  Rank ``r`` of ``W`` takes frames ``r, r + W, r + 2W, ...`` of the global
batch and renumbers them ``0, 1, 2, ...`` locally (what ``DistributedSampler`` + ``collate_batch`` give each
rank in the reference, ``pcdet/datasets/__init__.py:79-86``).  The only collective of a training step is DDP's
gradient all-reduce of the PFN parameters -- never inside the encoder.

Host-side helpers only (numpy / torch on any device); used by bench.py and the multi-rank tests.

``GradientAllReduce`` is that one collective, issued without DistributedDataParallel's per-step machinery (bucket
rebuilds, autograd hooks, a buffer broadcast in front of every forward): the PFN parameter gradients of all wrapped
modules are packed into one flat buffer and averaged with ONE ``all_reduce`` per step (2 176 + 2 048 bytes for the
radar + LiDAR encoders).  Same result as DDP's gradient averaging; ``bench.py --dp ddp`` runs the torch wrapper instead.
"""
from __future__ import annotations

import numpy as np
import torch


def frames_of_rank(global_batch: int, rank: int, world: int) -> list:
    """Global frame ids encoded by ``rank``, in local order."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, global_batch, world))


def shard_points(points, global_batch: int, rank: int, world: int):
    """Rows of ``points`` (N, 1+C; column 0 = global frame id) that belong to ``rank``, with column 0 renumbered
    to the local frame id.  Row order inside a frame is preserved.  Works on numpy arrays and torch tensors."""
    b = points[:, 0]
    is_np = isinstance(points, np.ndarray)
    bi = b.astype(np.int64) if is_np else b.long()
    mask = (bi % world) == rank
    mask &= (bi >= 0) & (bi < global_batch)
    out = points[mask].copy() if is_np else points[mask].clone()
    out[:, 0] = (bi[mask] // world).astype(points.dtype) if is_np else (bi[mask] // world).to(points.dtype)
    local_batch = len(frames_of_rank(global_batch, rank, world))
    return out, local_batch


def unshard_coords(coords, rank: int, world: int):
    """Maps the local frame id in column 0 of a rank's pillar coords back to the global frame id."""
    out = coords.copy() if isinstance(coords, np.ndarray) else coords.clone()
    out[:, 0] = out[:, 0] * world + rank
    return out


def aggregate_throughput(rows_per_rank, seconds_per_rank):
    """Whole-job points/s: all rows processed divided by the slowest rank's time (device timed, max over ranks)."""
    return float(sum(rows_per_rank)) / float(max(seconds_per_rank))


class GradientAllReduce:
    """Averages the gradients of ``params`` over the ranks of ``group`` with one all-reduce of a flat buffer.

        reducer = GradientAllReduce(params)          # once
        loss.backward(); reducer.reduce()            # every step; p.grad then holds the average (views of the flat buffer)

    ``reduce`` is asynchronous with respect to the host: the collective runs on the backend's stream and the current
    stream waits for it (like DDP's finalize step), so the next kernels that read the gradients are ordered after it."""

    def __init__(self, params, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        ref = self.params[0] if self.params else None
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device) if ref is not None else None
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def reduce(self):
        if self.flat is None:
            return None
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)          # pack (one fused kernel)
        work = None
        if self.world > 1:
            if self.flat.is_cuda:
                work = self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.AVG, group=self.group, async_op=True)
                work.wait()                              # stream-level wait on CUDA: the host does not block
            else:                                        # gloo has no AVG
                self.dist.all_reduce(self.flat, group=self.group)
                self.flat /= self.world
        for p, v in zip(self.params, self.views):
            p.grad = v
        return work
