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
    """Averages the gradients of ``params`` over the ranks of ``group`` with one collective per step.

        reducer = GradientAllReduce(params)          # once
        loss.backward(); reducer.reduce()            # every step; p.grad then holds the average (views of the flat buffer)

    Two data paths, same result layout:
      * ``p2p`` (default on CUDA when torch's symmetric memory can map the ranks' buffers): ``rdp_allreduce_small`` -- ONE kernel
        on the current stream packs the gradients into a peer-mapped staging buffer, signals the peers over NVLink, waits for
        them and sums all ranks' buffers in rank order (~8 us on 8 B200s; bit-identical on every rank);
      * ``nccl``: ``_foreach_copy_`` + ``all_reduce(AVG)`` of the flat buffer (~33 us + a pack kernel; NCCL runs on its own
        stream).  Used when the mapping is unavailable, on CPU (gloo), or with ``RDP_ALLREDUCE=nccl``.
    ``reduce`` never blocks the host: the next kernels on the current stream are ordered after the collective."""

    def __init__(self, params, group=None, backend: str = "auto"):
        import os
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
        self.p2p = None
        backend = os.environ.get("RDP_ALLREDUCE", backend)
        if backend in ("auto", "p2p") and self.world > 1 and ref is not None and ref.is_cuda and ref.dtype == torch.float32 \
                and len(self.params) <= 16:
            try:
                self.p2p = _PeerAllReduce(n, ref.device, group if group is not None else dist.group.WORLD, dist)
            except Exception as e:   # no peer mapping on this box: NCCL does the same job
                if backend == "p2p":
                    raise
                self.p2p = None
                self.p2p_error = f"{type(e).__name__}: {e}"

    def reduce(self):
        if self.flat is None:
            return None
        if self.p2p is not None:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
            self.p2p(grads, self.flat)
            for p, v in zip(self.params, self.views):
                p.grad = v
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


class _PeerAllReduce:
    """Host side of ``rdp_allreduce_small``: a symmetric (peer-mapped) staging buffer per rank, the device table of the peers'
    mappings and the step counter.  torch's symmetric memory does the CUDA IPC exchange; the data path is librdp's kernel."""

    def __init__(self, n: int, device, group, dist):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.C, self.lib, self._lib = C, _lib.load(), _lib
        self.rank, self.world, self.n, self.device = dist.get_rank(group), dist.get_world_size(group), n, device
        nbytes = int(self.lib.rdp_allreduce_staging_bytes(n, self.world))
        self.staging = symm.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.staging.zero_()
        self.handle = symm.rendezvous(self.staging, group)
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)          # every rank's signal words are zero before anybody's first step
        self.step = 0
        self.scale = 1.0 / self.world

    def __call__(self, grads, out):
        C = self.C
        k = len(grads)
        for g in grads:
            if not g.is_contiguous() or g.dtype != torch.float32 or g.device != self.device:
                raise ValueError("gradients must be contiguous float32 tensors on the reducer's device")
        seg = (C.c_void_p * k)(*[g.data_ptr() for g in grads])
        cnt = (C.c_int32 * k)(*[g.numel() for g in grads])
        self.step += 1
        with torch.cuda.device(self.device):
            self._lib.check(self.lib.rdp_allreduce_small(seg, cnt, k, self.peers.data_ptr(), self.rank, self.world, self.step & 0xFFFFFFFF or 1,
                                                          self.scale, out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream),
                            "rdp_allreduce_small")
