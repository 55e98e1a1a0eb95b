"""``torch.library`` registration of the encoder: ``torch.ops.rdp.pillar_encode`` / ``pillar_encode_backward``.

The kernels of ``librdp.so`` (C ABI: ``include/rdp.h``) exposed as PyTorch custom ops with schemas, a fake (meta)
implementation that gives the data-dependent row counts unbacked symbolic sizes, and an autograd formula.  The drop-in
modules (``vfe.py``) call the same C entry points through ``ops.py`` (which can also enqueue a forward without
synchronising); this file is the op-level surface for callers that want ``torch.ops`` / ``torch.export`` visibility.

    feats, coords, inverse, counts, argpos, bn_state, workspace, counters, new_rm, new_rv = torch.ops.rdp.pillar_encode(
        points, weight, bias, gamma, beta, running_mean, running_var, spec_ints, spec_floats, batch_size, train_bn, want_argmax)

The op is functional (autograd formulas cannot be attached to mutating ops): in train mode the updated BatchNorm running
statistics come back as ``new_rm`` / ``new_rv`` and the caller copies them into its buffers.

``spec_ints``  = [cols, layout, use_abs, use_cluster, use_relative, with_distance, c_in, c_out, coord_cols, nx, ny]
``spec_floats`` = [lo x,y,z, voxel x,y,z, offset x,y,z, eps, momentum]
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops
from .ops import EncoderSpec


def spec_to_lists(spec: EncoderSpec) -> Tuple[List[int], List[float]]:
    ints = [spec.cols, spec.layout, int(spec.use_abs), int(spec.use_cluster), int(spec.use_relative), int(spec.with_distance),
            spec.c_in, spec.c_out, spec.coord_cols, spec.nx, spec.ny]
    floats = [*spec.lo, *spec.vsz, *spec.off, spec.eps, spec.momentum]
    return ints, [float(v) for v in floats]


def spec_from_lists(ints: List[int], floats: List[float]) -> EncoderSpec:
    return EncoderSpec(cols=ints[0], layout=ints[1], use_abs=bool(ints[2]), use_cluster=bool(ints[3]), use_relative=bool(ints[4]),
                       with_distance=bool(ints[5]), c_in=ints[6], c_out=ints[7], coord_cols=ints[8], lo=tuple(floats[0:3]),
                       vsz=tuple(floats[3:6]), off=tuple(floats[6:9]), nx=ints[9], ny=ints[10], eps=floats[9], momentum=floats[10])


def _empty(dev, dtype=torch.float32):
    return torch.empty(0, dtype=dtype, device=dev)


@torch.library.custom_op("rdp::pillar_encode", mutates_args=(), device_types="cuda")
def pillar_encode(points: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], gamma: Optional[torch.Tensor],
                  beta: Optional[torch.Tensor], running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
                  spec_ints: List[int], spec_floats: List[float], batch_size: int, train_bn: bool,
                  want_argmax: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                              torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    spec = spec_from_lists(spec_ints, spec_floats)
    dev = points.device
    rm = running_mean.clone() if running_mean is not None else None   # the kernels update their copies in train mode
    rv = running_var.clone() if running_var is not None else None
    r = ops.encode_forward(points, spec, batch_size, weight, bias, gamma, beta, rm, rv, train_bn, want_argmax)
    # ops.encode_forward carves inverse / counts / counters / bn_state out of the call's one scratch allocation; custom-op
    # outputs must not alias each other, so the small ones are copied and the allocation itself is the `workspace` output
    return (r.features, r.coords, r.inverse.clone(), r.counts.clone(),
            r.argpos if r.argpos is not None else _empty(dev, torch.int32),
            r.bn_state.clone() if r.bn_state is not None else _empty(dev, torch.float64), r._buf, r.counters.clone(),
            rm if rm is not None else _empty(dev), rv if rv is not None else _empty(dev))


@pillar_encode.register_fake
def _(points, weight, bias, gamma, beta, running_mean, running_var, spec_ints, spec_floats, batch_size, train_bn, want_argmax):
    ctx = torch.library.get_ctx()
    n_kept, n_pillars = ctx.new_dynamic_size(), ctx.new_dynamic_size()   # data dependent: kept points, pillars
    c_in, c_out, kc = spec_ints[6], spec_ints[7], spec_ints[8]
    f32 = lambda *s: points.new_empty(s, dtype=torch.float32)
    i32 = lambda *s: points.new_empty(s, dtype=torch.int32)
    argpos = i32(n_pillars, c_out) if want_argmax else i32(0)
    bn = points.new_empty((4 * c_out + 1 + 14 + 14 * 14,) if (train_bn and gamma is not None) else (0,), dtype=torch.float64)
    ws = points.new_empty((ctx.new_dynamic_size(),), dtype=torch.uint8)
    stat = f32(c_out) if running_mean is not None else f32(0)
    return (f32(n_pillars, c_out), i32(n_pillars, kc), i32(n_kept), i32(n_pillars), argpos, bn, ws, i32(16), stat,
            torch.empty_like(stat))


@torch.library.custom_op("rdp::pillar_encode_backward", mutates_args=(), device_types="cuda")
def pillar_encode_backward(points: torch.Tensor, grad_features: torch.Tensor, features: torch.Tensor, argpos: torch.Tensor,
                           bn_state: torch.Tensor, workspace: torch.Tensor, counters: torch.Tensor, weight: torch.Tensor,
                           bias: Optional[torch.Tensor], gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
                           running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor], spec_ints: List[int],
                           spec_floats: List[float], batch_size: int, train_bn: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    spec = spec_from_lists(spec_ints, spec_floats)
    res = ops.ExternalState(spec, batch_size, int(points.shape[0]), argpos, workspace, counters,
                            bn_state if bn_state.numel() else None)
    d_w, d_g, d_b = ops.encode_backward(points, spec, batch_size, res, features, grad_features, weight, bias, gamma, beta,
                                        running_mean, running_var, train_bn)
    return d_w, (d_g if d_g is not None else _empty(points.device)), d_b


@pillar_encode_backward.register_fake
def _(points, grad_features, features, argpos, bn_state, workspace, counters, weight, bias, gamma, beta, running_mean, running_var,
      spec_ints, spec_floats, batch_size, train_bn):
    c_out = spec_ints[7]
    return (torch.empty_like(weight), weight.new_empty((c_out,) if gamma is not None else (0,)), weight.new_empty((c_out,)))


def _setup_context(ctx, inputs, output):
    (points, weight, bias, gamma, beta, running_mean, running_var, spec_ints, spec_floats, batch_size, train_bn, want_argmax) = inputs
    feats, _coords, _inv, _cnt, argpos, bn_state, workspace, counters, _rm, _rv = output
    ctx.save_for_backward(points, feats, argpos, bn_state, workspace, counters, weight, bias, gamma, beta)
    ctx.rm, ctx.rv = running_mean, running_var
    ctx.cfg = (spec_ints, spec_floats, batch_size, train_bn, want_argmax)
    ctx.set_materialize_grads(False)


def _backward(ctx, g_feats, *_unused):
    points, feats, argpos, bn_state, workspace, counters, weight, bias, gamma, beta = ctx.saved_tensors
    spec_ints, spec_floats, batch_size, train_bn, want_argmax = ctx.cfg
    if g_feats is None:
        return (None,) * 12
    if not want_argmax:
        raise RuntimeError("rdp::pillar_encode was called with want_argmax=False; no gradient is available")
    d_w, d_g, d_b = torch.ops.rdp.pillar_encode_backward(points, g_feats.contiguous(), feats, argpos, bn_state, workspace, counters,
                                                         weight, bias, gamma, beta, ctx.rm, ctx.rv, spec_ints, spec_floats,
                                                         batch_size, train_bn)
    use_norm = gamma is not None
    return (None, d_w, (None if use_norm else d_b), (d_g if use_norm else None), (d_b if use_norm else None), None, None,
            None, None, None, None, None)


pillar_encode.register_autograd(_backward, setup_context=_setup_context)
