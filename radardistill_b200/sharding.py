"""Frame-level data parallelism for the pillar encoder.

Frames are independent units: the batch index is the most significant part of the merged key
(``dynamic_pillar_vfe.py:208``), so no pillar spans two frames and a rank can encode its frames with no
exchange of points, keys or pillars.  Rank ``r`` of ``W`` takes frames ``r, r + W, r + 2W, ...`` of the global
batch and renumbers them ``0, 1, 2, ...`` locally (what ``DistributedSampler`` + ``collate_batch`` give each
rank in the reference, ``pcdet/datasets/__init__.py:79-86``).  The only collective of a training step is DDP's
gradient all-reduce of the PFN parameters -- never inside the encoder.

Host-side helpers only (numpy / torch on any device); used by bench.py and the multi-rank tests.
"""
from __future__ import annotations

import numpy as np


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
