"""PyTorch-facing operators over the C ABI (``include/rdp.h``): device memory, streams and autograd
plumbing only -- every computation runs in librdp.so's sm_100a kernels.

    encode(points, spec, params, training) -> EncodeResult

replaces the body of the reference ``forward`` between reading ``points`` and writing the
``batch_dict`` keys (``dynamic_pillar_vfe.py:200-249``).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import Geom, Layout, PfnParams


@dataclass(frozen=True)
class EncoderSpec:
    """Static description of one encoder (constructor arguments of the reference classes)."""
    cols: int                     # 1 + num_point_features
    layout: int                   # _lib.LAYOUT_*
    use_abs: bool
    use_cluster: bool
    use_relative: bool
    with_distance: bool
    c_in: int
    c_out: int
    coord_cols: int               # 3: [b,y,x]   4: [b,0,y,x]
    lo: tuple                     # point_cloud_range[0:3] (fp32 values)
    vsz: tuple                    # voxel size (fp32 values)
    off: tuple                    # voxel/2 + lo (double, rounded once to fp32)
    nx: int
    ny: int
    eps: float = 1e-3
    momentum: float = 0.01
    nz: int = 0                   # > 1: voxels (3-D key, z quantised and masked): DynamicVoxelVFE / DynamicMeanVFE

    def geom(self, batch_size: int) -> Geom:
        g = Geom()
        for k in range(3):
            g.lo[k], g.vsz[k], g.off[k] = self.lo[k], self.vsz[k], self.off[k]
        g.nx, g.ny, g.batch_size, g.cols, g.nz = self.nx, self.ny, int(batch_size), self.cols, int(self.nz)
        return g

    def layout_struct(self) -> Layout:
        return Layout(self.layout, int(self.use_abs), int(self.use_cluster), int(self.use_relative),
                      int(self.with_distance), self.c_in, self.c_out, self.coord_cols)


def make_spec(num_point_features: int, voxel_size, grid_size, point_cloud_range, layout: int, use_abs: bool,
              use_cluster: bool, use_relative: bool, with_distance: bool, c_out: int) -> EncoderSpec:
    c = int(num_point_features)
    nz = 0
    if layout == _lib.LAYOUT_SIMPLE2D:
        c_in = 3 + (c if use_abs else c - 3) + (3 if use_cluster else 0) + (3 if use_relative else 0)
        coord_cols = 3
    else:
        c_in = (c if use_abs else c - 3) + 6
        use_cluster, use_relative, coord_cols = True, False, 4
        if layout == _lib.LAYOUT_DYNVOXEL:
            nz = int(grid_size[2])
    c_in += 1 if with_distance else 0
    pcr = np.asarray(point_cloud_range, dtype=np.float32)
    vs = np.asarray([float(v) for v in voxel_size], dtype=np.float64)
    lo = tuple(float(pcr[k]) for k in range(3))
    vsz = tuple(float(np.float32(vs[k])) for k in range(3))
    off = tuple(float(np.float32(vs[k] / 2.0 + float(pcr[k]))) for k in range(3))   # (:180-182)
    return EncoderSpec(cols=c + 1, layout=layout, use_abs=bool(use_abs), use_cluster=bool(use_cluster),
                       use_relative=bool(use_relative), with_distance=bool(with_distance), c_in=c_in, c_out=int(c_out),
                       coord_cols=coord_cols, lo=lo, vsz=vsz, off=off, nx=int(grid_size[0]), ny=int(grid_size[1]), nz=nz)


class EncodeResult:
    """Outputs + saved state of one forward.  ``features`` / ``coords`` are what the modules publish; ``inverse``,
    ``counts``, ``bn_state``, ``counters`` and ``workspace`` are views into the one scratch allocation of the call and
    are only materialised when somebody asks for them (they are inspection / backward state, not hot-path outputs)."""
    __slots__ = ("features", "coords", "argpos", "n_kept", "n_pillars", "spec", "batch_size", "n_points", "train_bn",
                 "_buf", "_plan", "_argmax", "_views", "params_struct", "sync_group")

    def __init__(self, features, coords, argpos, n_kept, n_pillars, spec, batch_size, n_points, buf, plan, train_bn=False,
                 params_struct=None, sync_group=None):
        self.features, self.coords, self.argpos = features, coords, argpos
        self.n_kept, self.n_pillars, self.spec, self.batch_size, self.n_points = n_kept, n_pillars, spec, batch_size, n_points
        self._buf, self._plan, self.train_bn, self.params_struct = buf, plan, train_bn, params_struct
        self._argmax, self._views = None, {}
        self.sync_group = sync_group   # SyncBatchNorm: process group of the batch statistics (None: per-rank statistics)

    def without_outputs(self):
        """Copy for the autograd node: must not reference the node's own outputs (reference cycle => ~1 GB waits for the GC)."""
        r = EncodeResult(None, None, self.argpos, self.n_kept, self.n_pillars, self.spec, self.batch_size, self.n_points,
                         self._buf, self._plan, self.train_bn, self.params_struct, self.sync_group)
        return r

    def with_outputs(self, features, coords):
        r = EncodeResult(features, coords, self.argpos, self.n_kept, self.n_pillars, self.spec, self.batch_size,
                         self.n_points, self._buf, self._plan, self.train_bn, self.params_struct, self.sync_group)
        r._views = self._views
        return r

    def _view(self, name, off, nbytes, dtype, narrow):
        v = self._views.get(name)
        if v is None:
            v = self._buf[off:off + nbytes].view(dtype)
            if narrow is not None:
                v = v[:narrow]
            self._views[name] = v
        return v

    def _state_ptrs(self):
        """(workspace ptr, workspace bytes, counters ptr, bn_state ptr | None) for the backward call."""
        base = self._buf.data_ptr()
        return base, self._plan.ws_bytes, base + self._plan.off_counters, (base + self._plan.off_bn) if self.train_bn else None

    @property
    def workspace(self) -> torch.Tensor:
        return self._view("workspace", 0, self._plan.ws_bytes, torch.uint8, None)

    @property
    def counters(self) -> torch.Tensor:
        return self._view("counters", self._plan.off_counters, 4 * _lib.RDP_NUM_COUNTERS, torch.int32, None)

    @property
    def inverse(self) -> torch.Tensor:
        """(N,) int32: pillar of each KEPT point (torch.unique's inverse, :212)."""
        return self._view("inverse", self._plan.off_inverse, 4 * (self._plan.cap + 4), torch.int32, self.n_kept)

    @property
    def counts(self) -> torch.Tensor:
        return self._view("counts", self._plan.off_counts, 4 * (self._plan.cap + 4), torch.int32, self.n_pillars)

    @property
    def bn_state(self) -> Optional[torch.Tensor]:
        if not self.train_bn:
            return None
        return self._view("bn_state", self._plan.off_bn, 8 * self._plan.bn_doubles, torch.float64, None)

    def pillar_lookup(self) -> torch.Tensor:
        """Dense ``(batch_size, ny, nx)`` int32 map cell -> row of ``features`` / ``coords`` (-1 = empty): the table
        SparseEnc's first SubMConv2d needs for its rule book (spconv_backbone_2d.py:262-271), read off the occupancy
        bitmap the index kernels left in the workspace."""
        dev = self._buf.device
        out = torch.empty((self.batch_size, self.spec.ny, self.spec.nx), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().rdp_pillar_lookup(self.n_points, C.byref(self._plan.geom), self._buf.data_ptr(),
                                                     self._plan.ws_bytes, out.data_ptr(), _raw_stream(dev.index)),
                       "rdp_pillar_lookup")
        return out

    @property
    def argmax(self) -> Optional[torch.Tensor]:
        """scatter_max's argmax in the reference's numbering: (P, c_out) int32 index among the KEPT points."""
        if self.argpos is None:
            return None
        if self._argmax is None:
            lib = _lib.load()
            out = torch.empty((max(self.n_pillars, 1), self.spec.c_out), dtype=torch.int32, device=self.argpos.device)
            with torch.cuda.device(self.argpos.device):
                base = self._buf.data_ptr()
                _lib.check(lib.rdp_argmax_kept(self.n_points, C.byref(self._plan.geom), C.byref(self._plan.layout),
                                               C.c_void_p(base), self._plan.ws_bytes,
                                               C.c_void_p(base + self._plan.off_counters), _ptr(self.argpos), _ptr(out),
                                               _stream_ptr()), "rdp_argmax_kept")
            self._argmax = out[:self.n_pillars]
        return self._argmax


class ExternalState:
    """Backward state handed around as separate tensors (the torch.library ops return them individually)."""
    __slots__ = ("argpos", "train_bn", "params_struct", "_plan", "workspace", "counters", "bn_state")

    def __init__(self, spec, batch_size, n_points, argpos, workspace, counters, bn_state):
        self.argpos, self.workspace, self.counters, self.bn_state = argpos, workspace, counters, bn_state
        self.train_bn, self.params_struct = bn_state is not None, None
        self._plan = _plan(spec, int(batch_size), int(n_points), self.train_bn)

    def _state_ptrs(self):
        return (self.workspace.data_ptr(), self.workspace.numel(), self.counters.data_ptr(),
                None if self.bn_state is None else self.bn_state.data_ptr())


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _raw_stream(dev_index: int) -> int:
    try:
        return torch._C._cuda_getCurrentRawStream(dev_index)
    except AttributeError:  # pragma: no cover
        return torch.cuda.current_stream(dev_index).cuda_stream


def _stream_ptr():
    return C.c_void_p(_raw_stream(torch.cuda.current_device()))


class _Plan:
    """Everything about a (spec, batch_size, n_points, train_bn) call that does not change from step to step: the ctypes
    geometry / layout structs, the workspace size and the carving of the call's single scratch allocation
    [ librdp workspace | counters | bn_state | inverse | counts ]."""
    __slots__ = ("spec", "geom", "layout", "ws_bytes", "cap", "bn_doubles", "off_counters", "off_bn", "off_inverse",
                 "off_counts", "total_bytes", "_stats")

    def stats_buffers(self, n0: int):
        """(stats offset, stats doubles, bwd offset, bwd doubles) inside the workspace: the two fp64 vectors a SyncBatchNorm
        forward / backward all-reduces between its phases (rdp_stats_buffers)."""
        if self._stats is None:
            so, bo, sd, bd = C.c_size_t(0), C.c_size_t(0), C.c_int64(0), C.c_int64(0)
            _lib.check(_lib.load().rdp_stats_buffers(n0, C.byref(self.geom), C.byref(self.layout), C.byref(so), C.byref(sd),
                                                    C.byref(bo), C.byref(bd)), "rdp_stats_buffers")
            self._stats = (so.value, sd.value, bo.value, bd.value)
        return self._stats


_plan_cache = {}


def _plan(spec: "EncoderSpec", batch_size: int, n0: int, train_bn: bool) -> _Plan:
    key = (id(spec), batch_size, n0, train_bn)
    pl = _plan_cache.get(key)
    if pl is not None and pl.spec is spec:
        return pl
    lib = _lib.load()
    pl = _Plan()
    pl._stats = None
    pl.spec, pl.geom, pl.layout = spec, spec.geom(batch_size), spec.layout_struct()
    nbytes = C.c_size_t(0)
    _lib.check(lib.rdp_workspace_bytes(n0, C.byref(pl.geom), C.byref(pl.layout), C.byref(nbytes)), "rdp_workspace_bytes")
    up = lambda v: (v + 255) // 256 * 256
    pl.ws_bytes, pl.cap = up(nbytes.value), max(n0, 1)
    pl.bn_doubles = int(lib.rdp_bn_state_doubles(C.byref(pl.layout))) if train_bn else 0
    pl.off_counters = pl.ws_bytes
    pl.off_bn = pl.off_counters + 256
    pl.off_inverse = pl.off_bn + up(8 * pl.bn_doubles)
    pl.off_counts = pl.off_inverse + up(4 * (pl.cap + 4))
    pl.total_bytes = pl.off_counts + up(4 * (pl.cap + 4))
    if len(_plan_cache) > 512:
        _plan_cache.clear()
    _plan_cache[key] = pl
    return pl


_prm_cache = {}


def _params_struct(spec: "EncoderSpec", weight, bias, gamma, beta, running_mean, running_var, train_bn: bool,
                   num_batches_tracked=None) -> PfnParams:
    """ctypes parameter struct for these tensors, cached by their addresses (they rarely move between steps)."""
    key = (weight.data_ptr(), 0 if bias is None else bias.data_ptr(), 0 if gamma is None else gamma.data_ptr(),
           0 if beta is None else beta.data_ptr(), 0 if running_mean is None else running_mean.data_ptr(),
           0 if running_var is None else running_var.data_ptr(), bool(train_bn),
           0 if num_batches_tracked is None else num_batches_tracked.data_ptr(), spec.eps, spec.momentum)
    p = _prm_cache.get(key)
    if p is None:
        p = PfnParams()
        p.weight, p.bias, p.gamma, p.beta = key[0], key[1] or None, key[2] or None, key[3] or None
        p.running_mean, p.running_var = key[4] or None, key[5] or None
        p.eps, p.momentum, p.train_bn = spec.eps, spec.momentum, int(train_bn)
        p.num_batches_tracked = key[7] or None
        p.stats_phase, p.local_stats, p.global_bwd = 0, None, None
        if len(_prm_cache) > 512:
            _prm_cache.clear()
        _prm_cache[key] = p
    return p


def _phase_params(prm: PfnParams, phase: int, local_stats=None, global_bwd=None) -> PfnParams:
    """A copy of a (cached) parameter struct for one phase of a SyncBatchNorm forward / backward."""
    q = PfnParams()
    C.memmove(C.byref(q), C.byref(prm), C.sizeof(PfnParams))
    q.stats_phase = phase
    q.local_stats = None if local_stats is None else local_stats.data_ptr()
    q.global_bwd = None if global_bwd is None else global_bwd.data_ptr()
    return q


def _check_param(t: Optional[torch.Tensor], shape, name: str, device=None):
    if t is None:
        return None
    # no silent copies: the kernels keep raw addresses (the train-mode forward updates the running statistics in place and
    # the backward re-reads the parameters), so a temporary contiguous copy would be updated / read instead of the module's tensor
    if not t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA float32 tensor of shape {tuple(shape)}, got {t.dtype} "
                         f"{tuple(t.shape)} (contiguous={t.is_contiguous()}) on {t.device}")
    if device is not None and t.device != device:
        raise ValueError(f"{name} lives on {t.device} but the points are on {device}")
    return t


class PendingEncode:
    """Kernels of one forward have been enqueued on the current stream; `finish` waits for the early (N, P) publication."""
    __slots__ = ("spec", "batch_size", "n_points", "points", "coords", "features", "argpos", "buf", "plan", "host", "event",
                 "train_bn", "prm", "sync_group", "keep")


_host_pool = {}


def _take_host(dev_index: int, stream_handle: int):
    """A (pinned counters buffer, its numpy view, CUDA event) triple; the event exists (has been recorded once) so that
    librdp can record it by handle from inside the forward call."""
    pool = _host_pool.get(dev_index)
    if pool:
        return pool.pop()
    host = torch.empty(_lib.RDP_NUM_COUNTERS, dtype=torch.int32).pin_memory()
    event = torch.cuda.Event()
    event.record(torch.cuda.current_stream(dev_index))
    return host, host.numpy(), event


def encode_launch(points: torch.Tensor, spec: EncoderSpec, batch_size: int, weight, bias, gamma, beta, running_mean,
                  running_var, train_bn: bool, want_argmax: bool, num_batches_tracked=None,
                  frame_offsets: Optional[torch.Tensor] = None, sync_group=None) -> PendingEncode:
    """Enqueues index + PFN forward on the current stream (one call into librdp) and returns without synchronising.

    ``frame_offsets`` (int32 CUDA, ``batch_size + 1`` entries): ``points`` then holds the frames back to back WITHOUT the
    batch column, ``(N, cols - 1)``, frame b owning rows ``[offsets[b], offsets[b + 1])`` -- the batch-index padding of
    ``collate_batch`` (dataset_distill.py:237-244) and its 4 bytes per point of upload are skipped."""
    lib = _lib.load()
    if not points.is_cuda:
        raise _lib.RdpError("the pillar encoder has no CPU path: `points` must be a CUDA tensor")
    in_cols = spec.cols - (1 if frame_offsets is not None else 0)
    if points.dim() != 2 or points.shape[1] != in_cols:
        raise ValueError(f"points must be (N, {in_cols}), got {tuple(points.shape)}")
    if frame_offsets is not None:
        if (not frame_offsets.is_cuda or frame_offsets.dtype != torch.int32 or frame_offsets.dim() != 1
                or frame_offsets.shape[0] != int(batch_size) + 1 or not frame_offsets.is_contiguous()):
            raise ValueError(f"frame_offsets must be a contiguous CUDA int32 tensor with batch_size + 1 = {int(batch_size) + 1} entries")
    pts = points
    if pts.requires_grad:
        pts = pts.detach()
    if pts.dtype != torch.float32:
        pts = pts.float()
    if not pts.is_contiguous() or pts.data_ptr() % 16:
        pts = pts.contiguous().clone() if pts.data_ptr() % 16 else pts.contiguous()
    dev = pts.device
    n0 = pts.shape[0]
    use_norm = gamma is not None
    weight = _check_param(weight, (spec.c_out, spec.c_in), "linear.weight", dev)
    bias = _check_param(bias, (spec.c_out,), "linear.bias", dev)
    gamma = _check_param(gamma, (spec.c_out,), "norm.weight", dev)
    beta = _check_param(beta, (spec.c_out,), "norm.bias", dev)
    running_mean = _check_param(running_mean, (spec.c_out,), "norm.running_mean", dev)
    running_var = _check_param(running_var, (spec.c_out,), "norm.running_var", dev)
    train_bn = bool(train_bn and use_norm)
    batch_size = int(batch_size)
    pl = _plan(spec, batch_size, int(n0), train_bn)

    switch = torch.cuda.current_device() != dev.index
    if switch:
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev)
    try:
        buf = torch.empty(pl.total_bytes, dtype=torch.uint8, device=dev)
        coords = torch.empty((pl.cap, spec.coord_cols), dtype=torch.int32, device=dev)
        features = torch.empty((pl.cap, spec.c_out), dtype=torch.float32, device=dev)
        argpos = torch.empty((pl.cap, spec.c_out), dtype=torch.int32, device=dev) if want_argmax else None
        st = _raw_stream(dev.index)
        # (N, P) are published to pinned host memory (zero-copy kernel store: never queues behind bulk DMA) right after
        # the bitmap scan, and `event` is recorded there: encode_finish returns while the rest of the forward still runs
        host, host_np, event = _take_host(dev.index, st)
        prm = _params_struct(spec, weight, bias, gamma, beta, running_mean, running_var, train_bn,
                             num_batches_tracked if train_bn else None)
        base = buf.data_ptr()

        def call(q):
            _lib.check(lib.rdp_encode_fwd_frames(pts.data_ptr(), None if frame_offsets is None else frame_offsets.data_ptr(), n0,
                                                 C.byref(pl.geom), C.byref(pl.layout), C.byref(q), base, pl.ws_bytes,
                                                 coords.data_ptr(), base + pl.off_inverse, base + pl.off_counts,
                                                 base + pl.off_counters, features.data_ptr(),
                                                 None if argpos is None else argpos.data_ptr(),
                                                 (base + pl.off_bn) if train_bn else None, host.data_ptr(), event.cuda_event, st),
                       "rdp_encode_fwd_frames")

        keep = None
        if sync_group is not None and train_bn:
            # SyncBatchNorm (tools/train.py:34,144-145): batch statistics over the points of every rank of the group.
            # Phase 1 = index + feature moments; one all-reduce of the (<= 184)-double moment vector; phase 2 = the rest.
            import torch.distributed as dist
            if n0 == 0:
                raise NotImplementedError("SyncBatchNorm with an empty local batch (this rank would skip the collectives)")
            so, sd, _, _ = pl.stats_buffers(n0)
            call(_phase_params(prm, 1))
            stats = buf[so:so + 8 * sd].view(torch.float64)
            local = stats.clone()
            dist.all_reduce(stats, group=sync_group)
            keep = (local, _phase_params(prm, 2, local_stats=local))
            call(keep[1])
        else:
            sync_group = None
            call(prm)
    finally:
        if switch:
            torch.cuda.set_device(prev)
    p = PendingEncode()
    p.spec, p.batch_size, p.n_points, p.points, p.coords, p.features, p.argpos = spec, batch_size, int(n0), pts, coords, features, argpos
    p.buf, p.plan, p.host, p.event, p.train_bn, p.prm = buf, pl, (host, host_np), event, train_bn, prm
    p.sync_group, p.keep = sync_group, keep
    return p


def encode_finish(p: PendingEncode) -> EncodeResult:
    """Waits for the early (N, P) publication -- not for the forward's kernels, which may still be running on the
    launching stream -- and narrows the capacity-sized outputs."""
    p.event.synchronize()
    host, host_np = p.host
    n_kept, n_pillars, err = int(host_np[_lib.CNT_N]), int(host_np[_lib.CNT_P]), int(host_np[_lib.CNT_ERRFLAGS])
    _host_pool.setdefault(p.points.device.index, []).append((host, host_np, p.event))
    if err & 1:
        raise ValueError(f"points[:, 0] holds a batch index outside [0, {p.batch_size})")
    return EncodeResult(p.features[:n_pillars], p.coords[:n_pillars], None if p.argpos is None else p.argpos[:n_pillars],
                        n_kept, n_pillars, p.spec, p.batch_size, p.n_points, p.buf, p.plan, p.train_bn, p.prm, p.sync_group)


def encode_forward(points: torch.Tensor, spec: EncoderSpec, batch_size: int, weight, bias, gamma, beta, running_mean,
                   running_var, train_bn: bool, want_argmax: bool) -> EncodeResult:
    """index + PFN forward on the current stream; one 64-byte read-back to learn N and P."""
    return encode_finish(encode_launch(points, spec, batch_size, weight, bias, gamma, beta, running_mean, running_var,
                                       train_bn, want_argmax))


def encode_backward(points: torch.Tensor, spec: EncoderSpec, batch_size: int, res: EncodeResult, features, grad_features,
                    weight, bias, gamma, beta, running_mean, running_var, train_bn: bool):
    """Parameter gradients (d_weight, d_gamma | None, d_beta_or_bias)."""
    lib = _lib.load()
    dev = points.device
    use_norm = gamma is not None
    train_bn = bool(train_bn and use_norm)
    pl = res._plan
    switch = torch.cuda.current_device() != dev.index
    if switch:
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev)
    try:
        d_w = torch.empty((spec.c_out, spec.c_in), dtype=torch.float32, device=dev)
        d_g = torch.empty(spec.c_out, dtype=torch.float32, device=dev) if use_norm else None
        d_b = torch.empty(spec.c_out, dtype=torch.float32, device=dev)
        g = grad_features
        if g.dtype != torch.float32 or not g.is_contiguous():
            g = g.contiguous().float()
        if g.data_ptr() % 16:   # the backward stages gradient rows with 16-byte bulk copies (a view into a flat buffer may be offset)
            g = g.clone()
        prm = res.params_struct
        if prm is None or bool(prm.train_bn) != train_bn:
            prm = _params_struct(spec, weight, bias, gamma, beta, running_mean, running_var, train_bn)
        ws_ptr, ws_bytes, cnt_ptr, bn_ptr = res._state_ptrs()

        def call(q):
            _lib.check(lib.rdp_pfn_bwd(points.data_ptr(), points.shape[0], C.byref(pl.geom), C.byref(pl.layout), C.byref(q),
                                       ws_ptr, ws_bytes, cnt_ptr, g.data_ptr(), None, res.argpos.data_ptr(), bn_ptr, d_w.data_ptr(),
                                       None if d_g is None else d_g.data_ptr(), d_b.data_ptr(), _raw_stream(dev.index)),
                       "rdp_pfn_bwd")

        sync_group = getattr(res, "sync_group", None)
        if sync_group is not None and train_bn:
            # SyncBatchNorm backward: the per-channel sums (dbeta, dgamma enter every rank's correction terms) are all-reduced
            # between the tile kernel (phase 1) and the closed-form epilogue (phase 2)
            import torch.distributed as dist
            _, _, bo, bd = pl.stats_buffers(points.shape[0])
            call(_phase_params(prm, 1))
            glob = res._buf[bo:bo + 8 * bd].view(torch.float64).clone()
            dist.all_reduce(glob, group=sync_group)
            call(_phase_params(prm, 2, global_bwd=glob))
        else:
            call(prm)
    finally:
        if switch:
            torch.cuda.set_device(prev)
    return d_w, d_g, d_b


class _Holder:
    """Out-parameter of the autograd functions (a plain list would be copied by torch.amp.custom_fwd's argument casting)."""
    __slots__ = ("items",)

    def __init__(self):
        self.items = []


class _PillarEncodeFn(torch.autograd.Function):
    """Autograd node: saves (points, argpos, BN state, workspace) -- never the (N, C) activations.  The forward kernels
    were already enqueued (``pending``); this only waits for the (N, P) publication and wires up the backward."""

    # --use_amp (tools/train.py:52, train_utils.py:57-64): the node opts out of autocast by construction -- nothing in its
    # forward or backward is a torch op that autocast could re-type: the kernels read the fp32 master parameters and points
    # through raw pointers and write fp32 (a half-precision upstream gradient is widened in encode_backward).  The
    # torch.amp.custom_fwd / custom_bwd wrappers would add ~30 us of context-manager overhead per step for no effect.
    @staticmethod
    def forward(ctx, weight, bias, gamma, beta, running_mean, running_var, pending, holder):
        res = encode_finish(pending)
        holder.items.append(res)
        # the node must not reference its own outputs (reference cycle => the ~1 GB of state would wait for the GC)
        ctx.res = res.without_outputs()
        ctx.spec, ctx.batch_size, ctx.train_bn = pending.spec, pending.batch_size, pending.train_bn
        ctx.points = pending.points
        ctx.save_for_backward(weight, bias, gamma, beta)
        ctx.rm, ctx.rv = running_mean, running_var
        ctx.mark_non_differentiable(res.coords)
        return res.features, res.coords

    @staticmethod
    def backward(ctx, grad_features, _grad_coords):
        weight, bias, gamma, beta = ctx.saved_tensors
        res = ctx.res
        if res.argpos is None:
            raise RuntimeError("backward through a forward that ran without requires_grad parameters")
        d_w, d_g, d_b = encode_backward(ctx.points, ctx.spec, ctx.batch_size, res, None, grad_features, weight, bias,
                                        gamma, beta, ctx.rm, ctx.rv, ctx.train_bn)
        use_norm = gamma is not None
        return (d_w, (None if use_norm else d_b), d_g, (d_b if use_norm else None), None, None, None, None)


class _PairEncodeFn(torch.autograd.Function):
    """One autograd node for two independent encoders (``vfe.forward_pair``): `a` ran on the current stream, `b` on
    ``side``.  Halves the per-step autograd bookkeeping of the pair; the two backward launches still overlap on the two
    streams.  Same gradients as two ``_PillarEncodeFn`` nodes."""

    @staticmethod
    def forward(ctx, wa, ba, ga, bea, wb, bb, gb, beb, pma, pmb, side, holder):
        ra, rb = encode_finish(pma.pending), encode_finish(pmb.pending)
        holder.items.extend((ra, rb))
        ctx.ra, ctx.rb = ra.without_outputs(), rb.without_outputs()
        ctx.pa, ctx.pb, ctx.side = pma, pmb, side
        ctx.save_for_backward(wa, ba, ga, bea, wb, bb, gb, beb)
        ctx.mark_non_differentiable(ra.coords, rb.coords)
        ctx.set_materialize_grads(False)
        return ra.features, ra.coords, rb.features, rb.coords

    @staticmethod
    def backward(ctx, gfa, _gca, gfb, _gcb):
        wa, ba, ga, bea, wb, bb, gb, beb = ctx.saved_tensors

        def one(pm, res, gf, w, b, g, be):
            if gf is None:
                return (None, None, None, None)
            if res.argpos is None:
                raise RuntimeError("backward through a forward that ran without requires_grad parameters")
            p = pm.pending
            d_w, d_g, d_b = encode_backward(p.points, p.spec, p.batch_size, res, None, gf, w, b, g, be, pm.params[4], pm.params[5],
                                            p.train_bn)
            return (d_w, None, d_g, d_b) if g is not None else (d_w, d_b, None, None)

        main = torch.cuda.current_stream()
        side = ctx.side
        grads_b = (None, None, None, None)
        if gfb is not None:   # the short one goes out first, on the side stream
            side.wait_stream(main)
            gfb.record_stream(side)
            torch.cuda.set_stream(side)
            try:
                grads_b = one(ctx.pb, ctx.rb, gfb, wb, bb, gb, beb)
            finally:
                torch.cuda.set_stream(main)
        grads_a = one(ctx.pa, ctx.ra, gfa, wa, ba, ga, bea)
        if gfb is not None:
            main.wait_stream(side)
            for t in grads_b:
                if t is not None:
                    t.record_stream(main)
        return grads_a + grads_b + (None, None, None, None)


def encode_wait_pair(pma: "PendingModuleEncode", pmb: "PendingModuleEncode", side):
    """``encode_wait`` for two pending encodes that both need gradients: one autograd node for the pair."""
    holder = _Holder()
    fa, ca, fb, cb = _PairEncodeFn.apply(*pma.params[:4], *pmb.params[:4], pma, pmb, side, holder)
    return holder.items[0].with_outputs(fa, ca), holder.items[1].with_outputs(fb, cb)


class PendingModuleEncode:
    __slots__ = ("pending", "params", "needs_grad")

    def __init__(self, pending, params, needs_grad):
        self.pending, self.params, self.needs_grad = pending, params, needs_grad


def encode_async(points, spec: EncoderSpec, batch_size: int, weight, bias=None, gamma=None, beta=None, running_mean=None,
                 running_var=None, train_bn: bool = False, num_batches_tracked=None, frame_offsets=None,
                 sync_group=None) -> PendingModuleEncode:
    """Enqueues a (differentiable) encode on the current stream; pair with ``encode_wait``."""
    if points.requires_grad:
        raise NotImplementedError("gradients w.r.t. points are not produced (points are a leaf in the reference)")
    needs_grad = torch.is_grad_enabled() and (weight.requires_grad or (bias is not None and bias.requires_grad) or
                                              (gamma is not None and gamma.requires_grad) or
                                              (beta is not None and beta.requires_grad))
    pending = encode_launch(points, spec, batch_size, weight, bias, gamma, beta, running_mean, running_var, train_bn,
                            want_argmax=needs_grad, num_batches_tracked=num_batches_tracked, frame_offsets=frame_offsets,
                            sync_group=sync_group)
    return PendingModuleEncode(pending, (weight, bias, gamma, beta, running_mean, running_var), needs_grad)


def encode_wait(pm: PendingModuleEncode) -> EncodeResult:
    if pm.needs_grad:
        holder = _Holder()
        feats, coords = _PillarEncodeFn.apply(*pm.params, pm.pending, holder)
        return holder.items[0].with_outputs(feats, coords)
    return encode_finish(pm.pending)


def encode(points, spec: EncoderSpec, batch_size: int, weight, bias=None, gamma=None, beta=None, running_mean=None,
           running_var=None, train_bn: bool = False) -> EncodeResult:
    """Differentiable (w.r.t. the PFN parameters) pillar encoding.  Returns the full EncodeResult."""
    return encode_wait(encode_async(points, spec, batch_size, weight, bias, gamma, beta, running_mean, running_var, train_bn))
