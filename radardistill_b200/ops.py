"""PyTorch-facing operators over the C ABI (``include/rdp.h``): device memory, streams and autograd
plumbing only -- every computation runs in librdp.so's sm_100a kernels.

    encode(points, spec, params, training) -> EncodeResult

replaces the body of the reference ``forward`` between reading ``points`` and writing the
``batch_dict`` keys (``dynamic_pillar_vfe.py:200-249``).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
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

    def geom(self, batch_size: int) -> Geom:
        g = Geom()
        for k in range(3):
            g.lo[k], g.vsz[k], g.off[k] = self.lo[k], self.vsz[k], self.off[k]
        g.nx, g.ny, g.batch_size, g.cols = self.nx, self.ny, int(batch_size), self.cols
        return g

    def layout_struct(self) -> Layout:
        return Layout(self.layout, int(self.use_abs), int(self.use_cluster), int(self.use_relative),
                      int(self.with_distance), self.c_in, self.c_out, self.coord_cols)


def make_spec(num_point_features: int, voxel_size, grid_size, point_cloud_range, layout: int, use_abs: bool,
              use_cluster: bool, use_relative: bool, with_distance: bool, c_out: int) -> EncoderSpec:
    c = int(num_point_features)
    if layout == _lib.LAYOUT_SIMPLE2D:
        c_in = 3 + (c if use_abs else c - 3) + (3 if use_cluster else 0) + (3 if use_relative else 0)
        coord_cols = 3
    else:
        c_in = (c if use_abs else c - 3) + 6
        use_cluster, use_relative, coord_cols = True, False, 4
    c_in += 1 if with_distance else 0
    pcr = np.asarray(point_cloud_range, dtype=np.float32)
    vs = np.asarray([float(v) for v in voxel_size], dtype=np.float64)
    lo = tuple(float(pcr[k]) for k in range(3))
    vsz = tuple(float(np.float32(vs[k])) for k in range(3))
    off = tuple(float(np.float32(vs[k] / 2.0 + float(pcr[k]))) for k in range(3))   # (:180-182)
    return EncoderSpec(cols=c + 1, layout=layout, use_abs=bool(use_abs), use_cluster=bool(use_cluster),
                       use_relative=bool(use_relative), with_distance=bool(with_distance), c_in=c_in, c_out=int(c_out),
                       coord_cols=coord_cols, lo=lo, vsz=vsz, off=off, nx=int(grid_size[0]), ny=int(grid_size[1]))


@dataclass
class EncodeResult:
    features: torch.Tensor            # (P, c_out) fp32
    coords: torch.Tensor              # (P, coord_cols) int32
    inverse: torch.Tensor             # (N,) int32   pillar of each KEPT point (torch.unique's inverse)
    counts: torch.Tensor              # (P,) int32
    argpos: Optional[torch.Tensor]    # (P, c_out) int32 winning row as a position in the grouped order (internal)
    n_kept: int
    n_pillars: int
    # state for the backward / inspection
    pillar_mean: Optional[torch.Tensor] = None
    bn_state: Optional[torch.Tensor] = None
    workspace: Optional[torch.Tensor] = None
    counters: Optional[torch.Tensor] = None
    spec: Optional["EncoderSpec"] = None
    batch_size: int = 1
    n_points: int = 0
    _argmax: Optional[torch.Tensor] = None

    @property
    def argmax(self) -> Optional[torch.Tensor]:
        """scatter_max's argmax in the reference's numbering: (P, c_out) int32 index among the KEPT points."""
        if self.argpos is None:
            return None
        if self._argmax is None:
            lib = _lib.load()
            out = torch.empty((max(self.n_pillars, 1), self.spec.c_out), dtype=torch.int32, device=self.argpos.device)
            with torch.cuda.device(self.argpos.device):
                geom, layout = self.spec.geom(self.batch_size), self.spec.layout_struct()
                _lib.check(lib.rdp_argmax_kept(self.n_points, C.byref(geom), C.byref(layout), _ptr(self.workspace),
                                               self.workspace.numel(), _ptr(self.counters), _ptr(self.argpos), _ptr(out),
                                               _stream_ptr()), "rdp_argmax_kept")
            self._argmax = out[:self.n_pillars]
        return self._argmax


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_pinned = {}


def _pinned_counters(device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    t = _pinned.get(key)
    if t is None:
        t = torch.empty(_lib.RDP_NUM_COUNTERS, dtype=torch.int32).pin_memory()
        _pinned[key] = t
    return t


def _params_struct(spec: EncoderSpec, weight, bias, gamma, beta, running_mean, running_var, train_bn: bool) -> PfnParams:
    p = PfnParams()
    p.weight, p.bias = weight.data_ptr(), (bias.data_ptr() if bias is not None else None)
    p.gamma = gamma.data_ptr() if gamma is not None else None
    p.beta = beta.data_ptr() if beta is not None else None
    p.running_mean = running_mean.data_ptr() if running_mean is not None else None
    p.running_var = running_var.data_ptr() if running_var is not None else None
    p.eps, p.momentum, p.train_bn = spec.eps, spec.momentum, int(train_bn)
    return p


def _check_param(t: Optional[torch.Tensor], shape, name: str):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected a CUDA float32 tensor of shape {tuple(shape)}, got {t.dtype} {tuple(t.shape)} on {t.device}")
    return t.detach().contiguous() if not t.is_contiguous() else t.detach()


@dataclass
class PendingEncode:
    """Kernels of one forward have been enqueued on `stream`; `finish` waits for the 64-byte (N, P) publication."""
    spec: EncoderSpec
    batch_size: int
    n_points: int
    points: torch.Tensor
    coords: torch.Tensor
    inverse: torch.Tensor
    counts: torch.Tensor
    features: torch.Tensor
    argpos: Optional[torch.Tensor]
    bn_state: Optional[torch.Tensor]
    workspace: torch.Tensor
    counters: torch.Tensor
    host: torch.Tensor
    event: "torch.cuda.Event"
    stream: "torch.cuda.Stream"
    train_bn: bool = False


_host_pool = {}


def _take_host(device, stream):
    """A (pinned counters buffer, CUDA event) pair; the event exists (has been recorded once) so that librdp can record
    it by handle from inside rdp_index_fwd_publish."""
    pool = _host_pool.setdefault(device.index, [])
    if pool:
        return pool.pop()
    host = torch.empty(_lib.RDP_NUM_COUNTERS, dtype=torch.int32).pin_memory()
    event = torch.cuda.Event()
    event.record(stream)
    return host, event


_struct_cache = {}


def _structs(spec: EncoderSpec, batch_size: int, n0: int):
    """ctypes geometry / layout structs and the workspace size for (spec, batch_size, n0), cached."""
    key = (spec, batch_size, n0)
    hit = _struct_cache.get(key)
    if hit is None:
        lib = _lib.load()
        geom, layout = spec.geom(batch_size), spec.layout_struct()
        nbytes = C.c_size_t(0)
        _lib.check(lib.rdp_workspace_bytes(n0, C.byref(geom), C.byref(layout), C.byref(nbytes)), "rdp_workspace_bytes")
        if len(_struct_cache) > 256:
            _struct_cache.clear()
        hit = _struct_cache[key] = (geom, layout, nbytes.value)
    return hit


def encode_launch(points: torch.Tensor, spec: EncoderSpec, batch_size: int, weight, bias, gamma, beta, running_mean,
                  running_var, train_bn: bool, want_argmax: bool) -> PendingEncode:
    """Enqueues index + PFN forward on the current stream and returns without synchronising."""
    lib = _lib.load()
    if not points.is_cuda:
        raise _lib.RdpError("the pillar encoder has no CPU path: `points` must be a CUDA tensor")
    if points.dim() != 2 or points.shape[1] != spec.cols:
        raise ValueError(f"points must be (N, {spec.cols}), got {tuple(points.shape)}")
    pts = points.detach()
    if pts.dtype != torch.float32:
        pts = pts.float()
    if not pts.is_contiguous() or pts.data_ptr() % 16:
        pts = pts.contiguous().clone() if pts.data_ptr() % 16 else pts.contiguous()
    dev = pts.device
    n0 = pts.shape[0]
    use_norm = gamma is not None
    weight = _check_param(weight, (spec.c_out, spec.c_in), "linear.weight")
    bias = _check_param(bias, (spec.c_out,), "linear.bias")
    gamma = _check_param(gamma, (spec.c_out,), "norm.weight")
    beta = _check_param(beta, (spec.c_out,), "norm.bias")
    running_mean = _check_param(running_mean, (spec.c_out,), "norm.running_mean")
    running_var = _check_param(running_var, (spec.c_out,), "norm.running_var")
    train_bn = bool(train_bn and use_norm)

    with torch.cuda.device(dev):
        geom, layout, ws_bytes = _structs(spec, int(batch_size), int(n0))
        nbytes = C.c_size_t(ws_bytes)
        cap = max(n0, 1)
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        coords = torch.empty((cap, spec.coord_cols), dtype=torch.int32, device=dev)
        inverse = torch.empty(cap + 4, dtype=torch.int32, device=dev)
        counts = torch.empty(cap + 4, dtype=torch.int32, device=dev)
        counters = torch.empty(_lib.RDP_NUM_COUNTERS, dtype=torch.int32, device=dev)
        features = torch.empty((cap, spec.c_out), dtype=torch.float32, device=dev)
        argpos = torch.empty((cap, spec.c_out), dtype=torch.int32, device=dev) if want_argmax else None
        bn_state = None
        if train_bn:
            bn_state = torch.zeros(int(lib.rdp_bn_state_doubles(C.byref(layout))), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream()
        st = C.c_void_p(stream.cuda_stream)
        # (N, P) are published to pinned host memory (zero-copy kernel store: never queues behind bulk DMA) right after
        # the bitmap scan, and `event` is recorded there: encode_finish returns while the rest of the forward still runs
        host, event = _take_host(dev, stream)
        _lib.check(lib.rdp_index_fwd_publish(_ptr(pts), n0, C.byref(geom), spec.coord_cols, _ptr(ws), nbytes.value,
                                             _ptr(coords), _ptr(inverse), _ptr(counts), _ptr(counters),
                                             C.c_void_p(host.data_ptr()), C.c_void_p(event.cuda_event), st),
                   "rdp_index_fwd_publish")
        prm = _params_struct(spec, weight, bias, gamma, beta, running_mean, running_var, train_bn)
        _lib.check(lib.rdp_pfn_fwd(_ptr(pts), n0, C.byref(geom), C.byref(layout), C.byref(prm), _ptr(ws), nbytes.value,
                                   _ptr(counters), _ptr(features), _ptr(argpos), None, _ptr(bn_state), st), "rdp_pfn_fwd")
    return PendingEncode(spec=spec, batch_size=int(batch_size), n_points=int(n0), points=pts, coords=coords, inverse=inverse,
                         counts=counts, features=features, argpos=argpos, bn_state=bn_state, workspace=ws, counters=counters,
                         host=host, event=event, stream=stream, train_bn=train_bn)


def encode_finish(p: PendingEncode) -> EncodeResult:
    """Waits for the early (N, P) publication -- not for the forward's kernels, which may still be running on
    `p.stream` -- and narrows the capacity-sized outputs."""
    p.event.synchronize()
    hv = p.host.tolist()
    n_kept, n_pillars, err = hv[_lib.CNT_N], hv[_lib.CNT_P], hv[_lib.CNT_ERRFLAGS]
    _host_pool.setdefault(p.points.device.index, []).append((p.host, p.event))
    if err & 1:
        raise ValueError(f"points[:, 0] holds a batch index outside [0, {p.batch_size})")
    return EncodeResult(features=p.features[:n_pillars], coords=p.coords[:n_pillars], inverse=p.inverse[:n_kept],
                        counts=p.counts[:n_pillars], argpos=None if p.argpos is None else p.argpos[:n_pillars],
                        n_kept=n_kept, n_pillars=n_pillars, pillar_mean=None, bn_state=p.bn_state, workspace=p.workspace,
                        counters=p.counters, spec=p.spec, batch_size=p.batch_size, n_points=p.n_points)


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
    with torch.cuda.device(dev):
        geom, layout = spec.geom(batch_size), spec.layout_struct()
        d_w = torch.empty((spec.c_out, spec.c_in), dtype=torch.float32, device=dev)
        d_g = torch.empty(spec.c_out, dtype=torch.float32, device=dev) if use_norm else None
        d_b = torch.empty(spec.c_out, dtype=torch.float32, device=dev)
        g = grad_features.contiguous().float()
        prm = _params_struct(spec, weight.detach(), None if bias is None else bias.detach(),
                             None if gamma is None else gamma.detach(), None if beta is None else beta.detach(),
                             running_mean, running_var, bool(train_bn and use_norm))
        feats = features if features.is_contiguous() else features.contiguous()
        _lib.check(lib.rdp_pfn_bwd(_ptr(points), points.shape[0], C.byref(geom), C.byref(layout), C.byref(prm),
                                   _ptr(res.workspace), res.workspace.numel(), _ptr(res.counters), _ptr(g),
                                   _ptr(feats), _ptr(res.argpos), _ptr(res.bn_state), _ptr(d_w),
                                   _ptr(d_g), _ptr(d_b), _stream_ptr()), "rdp_pfn_bwd")
    return d_w, d_g, d_b


class _PillarEncodeFn(torch.autograd.Function):
    """Autograd node: saves (points, argpos, BN state, workspace) -- never the (N, C) activations.  The forward kernels
    were already enqueued (``pending``); this only waits for them and wires up the backward."""

    @staticmethod
    def forward(ctx, weight, bias, gamma, beta, running_mean, running_var, pending, holder):
        res = encode_finish(pending)
        holder.append(res)
        # the node must not reference its own outputs (reference cycle => the ~1 GB of state would wait for the GC)
        ctx.res = dataclasses.replace(res, features=None, coords=None)
        ctx.spec, ctx.batch_size, ctx.train_bn = pending.spec, pending.batch_size, pending.train_bn
        ctx.points = pending.points
        ctx.save_for_backward(weight, bias, gamma, beta, res.features)
        ctx.rm, ctx.rv = running_mean, running_var
        ctx.mark_non_differentiable(res.coords)
        return res.features, res.coords

    @staticmethod
    def backward(ctx, grad_features, _grad_coords):
        weight, bias, gamma, beta, features = ctx.saved_tensors
        res = ctx.res
        if res.argpos is None:
            raise RuntimeError("backward through a forward that ran without requires_grad parameters")
        d_w, d_g, d_b = encode_backward(ctx.points, ctx.spec, ctx.batch_size, res, features, grad_features, weight, bias,
                                        gamma, beta, ctx.rm, ctx.rv, ctx.train_bn)
        use_norm = gamma is not None
        return (d_w, (None if use_norm else d_b), d_g, (d_b if use_norm else None), None, None, None, None)


@dataclass
class PendingModuleEncode:
    pending: PendingEncode
    params: tuple
    needs_grad: bool


def encode_async(points, spec: EncoderSpec, batch_size: int, weight, bias=None, gamma=None, beta=None, running_mean=None,
                 running_var=None, train_bn: bool = False) -> PendingModuleEncode:
    """Enqueues a (differentiable) encode on the current stream; pair with ``encode_wait``."""
    if points.requires_grad:
        raise NotImplementedError("gradients w.r.t. points are not produced (points are a leaf in the reference)")
    needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (weight, bias, gamma, beta))
    pending = encode_launch(points, spec, batch_size, weight, bias, gamma, beta, running_mean, running_var, train_bn,
                            want_argmax=needs_grad)
    return PendingModuleEncode(pending, (weight, bias, gamma, beta, running_mean, running_var), needs_grad)


def encode_wait(pm: PendingModuleEncode) -> EncodeResult:
    if pm.needs_grad:
        holder = []
        feats, coords = _PillarEncodeFn.apply(*pm.params, pm.pending, holder)
        return dataclasses.replace(holder[0], features=feats, coords=coords)
    return encode_finish(pm.pending)


def encode(points, spec: EncoderSpec, batch_size: int, weight, bias=None, gamma=None, beta=None, running_mean=None,
           running_var=None, train_bn: bool = False) -> EncodeResult:
    """Differentiable (w.r.t. the PFN parameters) pillar encoding.  Returns the full EncodeResult."""
    return encode_wait(encode_async(points, spec, batch_size, weight, bias, gamma, beta, running_mean, running_var, train_bn))
