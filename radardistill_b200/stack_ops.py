"""Operators for stacked PFN layers and the sibling encoders and for device-side input prep (8f-2).

Torch plumbing over ``include/rdp.h``: ``rdp_index_fwd_frames`` (2-D or 3-D key), ``rdp_decorate``,
``rdp_segment_max_fwd / _bwd``, ``rdp_voxel_mean``, ``rdp_prepare_points``.  The only computation that does not run in
librdp's kernels is the dense ``Linear`` + ``BatchNorm1d`` of a stacked layer, which is a plain library GEMM / cuDNN call
through torch (``vfe.PFNLayerV2.forward``).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, ops
from .ops import EncoderSpec, _plan, _raw_stream, _take_host, _host_pool


class IndexResult:
    """Outputs of the index pass (``dynamic_pillar_vfe.py:201-212,243-248``; voxel form ``dynamic_voxel_vfe.py:57-71,94-100``)
    plus the workspace that holds the pillar-grouped rows and the pillar table for the kernels that follow."""
    __slots__ = ("coords", "inverse", "counts", "n_kept", "n_pillars", "n_points", "spec", "batch_size", "buf", "plan", "device")

    def _ws(self):
        base = self.buf.data_ptr()
        return base, self.plan.ws_bytes, base + self.plan.off_counters


def index_forward(points: torch.Tensor, spec: EncoderSpec, batch_size: int,
                  frame_offsets: Optional[torch.Tensor] = None) -> IndexResult:
    lib = _lib.load()
    if not points.is_cuda:
        raise _lib.RdpError("the pillar encoder has no CPU path: `points` must be a CUDA tensor")
    in_cols = spec.cols - (1 if frame_offsets is not None else 0)
    if points.dim() != 2 or points.shape[1] != in_cols:
        raise ValueError(f"points must be (N, {in_cols}), got {tuple(points.shape)}")
    pts = points.detach()
    if pts.dtype != torch.float32:
        pts = pts.float()
    if not pts.is_contiguous() or pts.data_ptr() % 16:
        pts = pts.contiguous().clone() if pts.data_ptr() % 16 else pts.contiguous()
    dev, n0 = pts.device, int(pts.shape[0])
    pl = _plan(spec, int(batch_size), n0, False)
    with torch.cuda.device(dev):
        buf = torch.empty(pl.total_bytes, dtype=torch.uint8, device=dev)
        coords = torch.empty((pl.cap, spec.coord_cols), dtype=torch.int32, device=dev)
        st = _raw_stream(dev.index)
        host, host_np, event = _take_host(dev.index, st)
        base = buf.data_ptr()
        _lib.check(lib.rdp_index_fwd_frames(pts.data_ptr(), None if frame_offsets is None else frame_offsets.data_ptr(), n0,
                                            C.byref(pl.geom), spec.coord_cols, base, pl.ws_bytes, coords.data_ptr(),
                                            base + pl.off_inverse, base + pl.off_counts, base + pl.off_counters,
                                            host.data_ptr(), event.cuda_event, st), "rdp_index_fwd_frames")
        event.synchronize()
    n, p, err = int(host_np[_lib.CNT_N]), int(host_np[_lib.CNT_P]), int(host_np[_lib.CNT_ERRFLAGS])
    _host_pool.setdefault(dev.index, []).append((host, host_np, event))
    if err & 1:
        raise ValueError(f"points[:, 0] holds a batch index outside [0, {batch_size})")
    r = IndexResult()
    r.coords = coords[:p]
    r.inverse = buf[pl.off_inverse:pl.off_inverse + 4 * (pl.cap + 4)].view(torch.int32)[:n]
    r.counts = buf[pl.off_counts:pl.off_counts + 4 * (pl.cap + 4)].view(torch.int32)[:p]
    r.n_kept, r.n_pillars, r.n_points, r.spec, r.batch_size, r.buf, r.plan, r.device = n, p, n0, spec, int(batch_size), buf, pl, dev
    return r


def decorate(idx: IndexResult) -> torch.Tensor:
    """(N, c_in) decorated point features in kept-point order (what the reference feeds its first PFNLayerV2)."""
    spec = idx.spec
    out = torch.empty((max(idx.n_points, 1), spec.c_in), dtype=torch.float32, device=idx.device)
    ws, ws_bytes, cnt = idx._ws()
    with torch.cuda.device(idx.device):
        _lib.check(_lib.load().rdp_decorate(idx.n_points, C.byref(idx.plan.geom), C.byref(idx.plan.layout), ws, ws_bytes, cnt,
                                           out.data_ptr(), _raw_stream(idx.device.index)), "rdp_decorate")
    return out[:idx.n_kept]


class _SegmentMaxFn(torch.autograd.Function):
    """``torch_scatter.scatter_max(x, unq_inv, dim=0)`` (PFNLayerV2.forward :40) over the pillars of an IndexResult."""

    @staticmethod
    def forward(ctx, x, idx):
        x = x.contiguous().float()   # autocast (--use_amp) hands over half-precision activations: the pooling runs and returns fp32
        n, c = x.shape
        if n != idx.n_kept:
            raise ValueError(f"activations have {n} rows, the index pass kept {idx.n_kept} points")
        out = torch.empty((max(idx.n_pillars, 1), c), dtype=torch.float32, device=x.device)
        arg = torch.empty((max(idx.n_pillars, 1), c), dtype=torch.int32, device=x.device)
        ws, ws_bytes, cnt = idx._ws()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().rdp_segment_max_fwd(x.data_ptr(), c, idx.n_points, C.byref(idx.plan.geom), ws, ws_bytes, cnt,
                                                      out.data_ptr(), arg.data_ptr(), _raw_stream(x.device.index)),
                       "rdp_segment_max_fwd")
        out, arg = out[:idx.n_pillars], arg[:idx.n_pillars]
        ctx.save_for_backward(arg)
        ctx.n = n
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g_out, _g_arg):
        (arg,) = ctx.saved_tensors
        p, c = arg.shape
        g = g_out.contiguous().float()
        gx = torch.zeros((ctx.n, c), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().rdp_segment_max_bwd(g.data_ptr(), arg.data_ptr(), p, c, gx.data_ptr(),
                                                      _raw_stream(g.device.index)), "rdp_segment_max_bwd")
        return gx, None


def segment_max(x: torch.Tensor, idx: IndexResult):
    """(P, C) per-pillar maximum of x (N, C) and the kept index of the winning row (lowest index on ties)."""
    return _SegmentMaxFn.apply(x, idx)


def voxel_mean(idx: IndexResult) -> torch.Tensor:
    """(P, cols - 1) per-voxel mean of every point column (DynamicMeanVFE, dynamic_mean_vfe.py:63-65)."""
    c = idx.spec.cols - 1
    out = torch.empty((max(idx.n_points, 1), c), dtype=torch.float32, device=idx.device)
    ws, ws_bytes, cnt = idx._ws()
    with torch.cuda.device(idx.device):
        _lib.check(_lib.load().rdp_voxel_mean(idx.n_points, C.byref(idx.plan.geom), ws, ws_bytes, cnt, out.data_ptr(),
                                             _raw_stream(idx.device.index)), "rdp_voxel_mean")
    return out[:idx.n_pillars]


def prepare_points(points: torch.Tensor, point_cloud_range, x_col: int = 0, shuffle_seed: int = 0) -> torch.Tensor:
    """Device-side ``mask_points_and_boxes_outside_range`` + ``shuffle_points`` of the data processor
    (pcdet/datasets/processor/data_processor.py:80-86, :99-114) for one frame already on the GPU: rows with
    lo <= x, y <= hi (inclusive) are kept in order; ``shuffle_seed != 0`` writes them through a fixed pseudo-random
    permutation instead.  One 4-byte read-back (the kept count)."""
    if not points.is_cuda or points.dtype != torch.float32 or points.dim() != 2:
        raise ValueError("points must be a 2-D CUDA float32 tensor")
    pts = points.contiguous()
    n0, cols = int(pts.shape[0]), int(pts.shape[1])
    lib = _lib.load()
    dev = pts.device
    rng = (C.c_float * 4)(float(point_cloud_range[0]), float(point_cloud_range[1]), float(point_cloud_range[3]),
                          float(point_cloud_range[4]))
    out = torch.empty_like(pts)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = int(lib.rdp_prepare_scratch_bytes(n0))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rdp_prepare_points(pts.data_ptr(), n0, cols, int(x_col), rng, int(shuffle_seed) & 0xFFFFFFFFFFFFFFFF,
                                          scratch.data_ptr(), nbytes, out.data_ptr(), n_out.data_ptr(), _raw_stream(dev.index)),
                   "rdp_prepare_points")
    return out[:int(n_out.item())]
