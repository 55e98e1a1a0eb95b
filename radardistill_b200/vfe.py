"""Drop-in replacements for the reference's dynamic pillar encoders.

Same class names, constructor signature, ``forward(batch_dict) -> batch_dict`` contract,
``batch_dict`` keys and ``state_dict`` names as
``/root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py``:

* ``DynamicPillarVFE``                      (:49-142)   reads ``points``        -> ``pillar_features`` / ``voxel_features``, ``voxel_coords`` (P,4)
* ``DynamicPillarVFESimple2D``              (:146-252)  reads ``points``        -> ``pillar_features``, ``pillar_coords`` (P,3)
* ``Radar_DynamicPillarVFESimple2D``        (:255-313)  reads ``radar_points``  -> ``radar_pillar_features``, ``radar_pillar_coords``
* ``Radar_DynamicPillarVFESimple2D_Test``   (:315-373)  reads ``points``        -> ``radar_pillar_*``

``register(vfe_all)`` overwrites the four entries of the reference registry
(``pcdet/models/backbones_3d/vfe/__init__.py:9-21``).  The forward runs entirely in the sm_100a
kernels of librdp.so; there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _ext, _lib, ops, stack_ops


class VFETemplate(nn.Module):
    """Mirror of ``vfe_template.py:4-22``."""

    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg

    def get_output_feature_dim(self):
        raise NotImplementedError

    def forward(self, **kwargs):
        raise NotImplementedError


def _world_size() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _cfg_get(cfg, key, default=None):
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _cfg(cfg, key):
    try:
        return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)
    except (KeyError, AttributeError):
        raise AttributeError(f"model_cfg.{key} is required") from None


class PFNLayerV2(nn.Module):
    """``dynamic_pillar_vfe.py:14-46`` with the reference's parameter names (``linear``, ``norm``).

    A single (last) layer over pillars is fused into the encoder kernels and never calls ``forward``.  Stacked layers
    (``NUM_FILTERS`` longer than 1) and the voxel encoder run ``forward``: Linear + BatchNorm1d + ReLU as library calls,
    ``scatter_max`` by ``rdp_segment_max_fwd`` over the pillar-grouped rows of the index pass."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        self.last_vfe = last_layer
        self.use_norm = use_norm
        if not self.last_vfe:
            out_channels = out_channels // 2
        if self.use_norm:
            self.linear = nn.Linear(in_channels, out_channels, bias=False)
            self.norm = nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01)
        else:
            self.linear = nn.Linear(in_channels, out_channels, bias=True)
        self.relu = nn.ReLU()

    def forward(self, inputs, unq_inv):
        """``unq_inv``: the ``stack_ops.IndexResult`` of the index pass (it carries ``inverse`` and the grouped rows the
        pooling kernel walks).  There is no torch_scatter path."""
        if not isinstance(unq_inv, stack_ops.IndexResult):
            raise _lib.RdpError("PFNLayerV2.forward needs the IndexResult of the encoder's index pass; call the VFE module")
        x = self.linear(inputs)                                        # (:36)
        x = self.norm(x) if self.use_norm else x                       # (:37)
        x = self.relu(x)                                               # (:38)
        x_max = stack_ops.segment_max(x, unq_inv)[0]                   # (:40)
        if self.last_vfe:
            return x_max                                               # (:42-43)
        return torch.cat([x, x_max[unq_inv.inverse.long(), :]], dim=1)  # (:44-46)


class _FusedDynamicPillarVFE(VFETemplate):
    _layout = _lib.LAYOUT_SIMPLE2D
    _points_key = "points"
    _feature_keys = ("pillar_features",)
    _coords_key = "pillar_coords"

    def __init__(self, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, **kwargs):
        super().__init__(model_cfg=model_cfg)
        self.use_norm = _cfg(model_cfg, "USE_NORM")
        self.with_distance = _cfg(model_cfg, "WITH_DISTANCE")
        self.use_absolute_xyz = _cfg(model_cfg, "USE_ABSLOTE_XYZ")
        if self._layout == _lib.LAYOUT_SIMPLE2D:
            self.use_cluster_xyz = _cfg_get(model_cfg, "USE_CLUSTER_XYZ", True)
            self.use_relative_xyz = _cfg_get(model_cfg, "USE_RELATIVE_XYZ", True)
            self.double_flip = _cfg_get(model_cfg, "DOUBLE_FLIP", False)
        else:
            self.use_cluster_xyz, self.use_relative_xyz, self.double_flip = True, False, False
        self.num_filters = list(_cfg(model_cfg, "NUM_FILTERS"))
        assert len(self.num_filters) > 0
        first_out = self.num_filters[0] if len(self.num_filters) == 1 else self.num_filters[0] // 2
        self.spec = ops.make_spec(num_point_features, voxel_size, grid_size, point_cloud_range, self._layout,
                                  bool(self.use_absolute_xyz), bool(self.use_cluster_xyz), bool(self.use_relative_xyz),
                                  bool(self.with_distance), first_out)
        self.num_point_features = self.spec.c_in
        filters = [self.spec.c_in] + self.num_filters                   # (:63-72)
        self.pfn_layers = nn.ModuleList([PFNLayerV2(filters[i], filters[i + 1], self.use_norm, last_layer=(i >= len(filters) - 2))
                                         for i in range(len(filters) - 1)])
        # One PFN layer over pillars runs in the fused kernels when they are compiled for this row width / channel count;
        # stacked layers, voxel grids and any other shape take the layer-stack path (index + rdp_decorate + PFNLayerV2.forward).
        self.fused = (len(self.num_filters) == 1 and self.spec.nz <= 1 and
                      bool(_lib.load().rdp_config_supported(ops.C.byref(self.spec.geom(1)), ops.C.byref(self.spec.layout_struct()))))
        self.voxel_x, self.voxel_y, self.voxel_z = voxel_size[0], voxel_size[1], voxel_size[2]
        self.x_offset, self.y_offset, self.z_offset = self.spec.off
        self.scale_xy = int(grid_size[0]) * int(grid_size[1])
        self.scale_y = int(grid_size[1])
        # EncodeResult of the latest forward (inverse / counts / argmax for inspection); kept out of nn.Module.__setattr__,
        # which costs ~15 us per assignment on the hot path
        object.__setattr__(self, "last_result", None)

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    def _batch_size(self, batch_dict, points):
        bs = batch_dict.get("batch_size", None)
        if bs is not None:
            return int(bs)
        return int(points[:, 0].max().item()) + 1 if points.shape[0] else 1  # host sync; pass batch_size to avoid it

    def launch(self, batch_dict):
        """Enqueues this encoder's kernels on the current stream without synchronising; pair with ``finish``.
        Lets independent encoders (LiDAR teacher / radar student) overlap on two streams -- see ``forward_pair``."""
        if self.double_flip:
            # the reference path uses `np` without importing it (dynamic_pillar_vfe.py:196-199) and cannot run
            raise NotImplementedError("DOUBLE_FLIP is dead code in the reference (NameError) and is not provided")
        points = batch_dict[self._points_key]
        # optional device-side input prep: `<points key>_offsets` (int32, batch_size + 1) => `points` is (N, C) without the
        # batch column, frames back to back (what the dataset produces before collate_batch pads the frame index in)
        offsets = batch_dict.get(self._points_key + "_offsets", None)
        bs = (int(offsets.shape[0]) - 1) if offsets is not None else self._batch_size(batch_dict, points)
        if not self.fused:
            return ("stack", points, offsets, bs), False
        pfn = self.pfn_layers[0]
        norm = pfn.norm if self.use_norm else None
        train_bn = bool(self.use_norm and norm.training)
        sync_group = None
        if train_bn and isinstance(norm, nn.SyncBatchNorm) and _world_size() > 1:
            # tools/train.py --sync_bn (:34,144-145) converts the PFN's BatchNorm1d: batch statistics over every rank's points --
            # the fused forward / backward run in two phases around one small all-reduce each (ops.encode_launch)
            import torch.distributed as dist
            sync_group = norm.process_group if norm.process_group is not None else dist.group.WORLD
        pm = ops.encode_async(points, self.spec, bs, pfn.linear.weight,
                              bias=None if self.use_norm else pfn.linear.bias,
                              gamma=norm.weight if norm is not None else None, beta=norm.bias if norm is not None else None,
                              running_mean=norm.running_mean if norm is not None else None,
                              running_var=norm.running_var if norm is not None else None, train_bn=train_bn,
                              num_batches_tracked=norm.num_batches_tracked if train_bn else None, frame_offsets=offsets,
                              sync_group=sync_group)
        return pm, train_bn

    def _native_args(self, batch_dict, grad_enabled):
        """Descriptor of this encoder's call for the native host path (csrc/rdp_torch.cpp), or None when this call has to go
        through the Python glue (layer-stack path, SyncBatchNorm, DOUBLE_FLIP error path)."""
        if not self.fused or self.double_flip:
            return None
        pfn = self.pfn_layers[0]
        norm = pfn.norm if self.use_norm else None
        train_bn = bool(self.use_norm and norm.training)
        if train_bn and isinstance(norm, nn.SyncBatchNorm) and _world_size() > 1:
            return None
        points = batch_dict[self._points_key]
        if points.requires_grad:
            return None
        offsets = batch_dict.get(self._points_key + "_offsets", None)
        bs = (int(offsets.shape[0]) - 1) if offsets is not None else self._batch_size(batch_dict, points)
        spec = self.spec
        pl = ops._plan(spec, bs, int(points.shape[0]), train_bn)
        w = pfn.linear.weight
        want_grad = grad_enabled and (w.requires_grad or (norm is not None and (norm.weight.requires_grad or norm.bias.requires_grad)) or
                                      (norm is None and pfn.linear.bias.requires_grad))
        geom = spec.lo + spec.vsz + spec.off + (spec.nx, spec.ny, bs, spec.cols, spec.nz)
        lay = (spec.layout, int(spec.use_abs), int(spec.use_cluster), int(spec.use_relative), int(spec.with_distance), spec.c_in, spec.c_out,
               spec.coord_cols)
        plan = (pl.ws_bytes, pl.off_counters, pl.off_bn, pl.off_inverse, pl.off_counts, pl.total_bytes, pl.cap)
        if norm is not None:
            prm = (w, None, norm.weight, norm.bias, norm.running_mean, norm.running_var, norm.num_batches_tracked)
        else:
            prm = (w, pfn.linear.bias, None, None, None, None, None)
        return (geom, lay, plan, spec.eps, spec.momentum, points, offsets) + prm + (train_bn, bool(want_grad)), pl, bs, train_bn

    def _publish_native(self, batch_dict, out, pl, bs, train_bn, n_points):
        feats, coords, argpos, buf, n_kept, n_pillars = out
        res = ops.EncodeResult(feats, coords, argpos, n_kept, n_pillars, self.spec, bs, n_points, buf, pl, train_bn, None)
        return self._publish(batch_dict, res, train_bn)

    def finish(self, batch_dict, token):
        pm, train_bn = token
        if isinstance(pm, tuple) and pm[0] == "stack":
            return self._forward_stack(batch_dict, *pm[1:])
        return self._publish(batch_dict, ops.encode_wait(pm), train_bn)

    def _forward_stack(self, batch_dict, points, offsets, bs):
        """Layer-stack path (``dynamic_pillar_vfe.py:93-140``, ``dynamic_voxel_vfe.py:55-105``): index kernels, decorated
        per-point features, then the PFN layers as the reference chains them (:123-124)."""
        idx = stack_ops.index_forward(points, self.spec, bs, offsets)
        features = stack_ops.decorate(idx)
        for pfn in self.pfn_layers:
            features = pfn(features, idx)
        object.__setattr__(self, "last_result", idx)
        for k in self._feature_keys:
            batch_dict[k] = features
        batch_dict[self._coords_key] = idx.coords
        return batch_dict

    def _publish(self, batch_dict, res, train_bn):
        if train_bn:
            if res.n_kept == 1:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size "
                                 f"torch.Size([1, {self.spec.c_out}])")
            if res.n_points == 0:   # no kernel ran; otherwise bn_finalize_kernel has counted the batch (BatchNorm1d, :29)
                self.pfn_layers[0].norm.num_batches_tracked.add_(1)
        object.__setattr__(self, "last_result", res)
        for k in self._feature_keys:
            batch_dict[k] = res.features
        batch_dict[self._coords_key] = res.coords
        return batch_dict

    def forward(self, batch_dict, **kwargs):
        return self.finish(batch_dict, self.launch(batch_dict))


def forward_pair(first, second, batch_dict, side_stream=None, first_no_grad=False):
    """Runs two independent encoders (e.g. ``vfe`` and ``radar_vfe``, pillarnet.py:28-33) concurrently: ``second`` is
    enqueued on a side stream while ``first`` runs on the current one.  Results are identical to calling them in turn."""
    grad_on = torch.is_grad_enabled()
    ext = _ext.load() if side_stream is None else None
    if ext is not None and isinstance(first, _FusedDynamicPillarVFE) and isinstance(second, _FusedDynamicPillarVFE):
        # native host path: one call enqueues both encoders on two streams, waits for both (N, P) and wires one autograd node
        na = first._native_args(batch_dict, grad_on and not first_no_grad)
        nb = second._native_args(batch_dict, grad_on) if na is not None else None
        if na is not None and nb is not None:
            out = ext.pair_forward(na[0], nb[0])
            batch_dict = first._publish_native(batch_dict, out[:6], na[1], na[2], na[3], int(na[0][5].shape[0]))
            return second._publish_native(batch_dict, out[6:], nb[1], nb[2], nb[3], int(nb[0][5].shape[0]))
    main = torch.cuda.current_stream()
    side = side_stream if side_stream is not None else _side_stream(main.device)
    side.wait_stream(main)   # inputs of `second` were produced on the current stream
    try:
        if first_no_grad and grad_on:   # frozen teacher (FREEZE_PIPELINE)
            torch.set_grad_enabled(False)
        tok1 = first.launch(batch_dict)   # the long kernels go first: they keep the GPU busy while `second` is enqueued
        torch.set_grad_enabled(grad_on)
        torch.cuda.set_stream(side)       # plain stream switches: the `with torch.cuda.stream(...)` manager costs ~25 us a time
        tok2 = second.launch(batch_dict)
        torch.cuda.set_stream(main)
        fused_pair = not isinstance(tok1[0], tuple) and not isinstance(tok2[0], tuple)
        if fused_pair and tok1[0].needs_grad and tok2[0].needs_grad:   # one autograd node for the pair
            res1, res2 = ops.encode_wait_pair(tok1[0], tok2[0], side)
            batch_dict = first._publish(batch_dict, res1, tok1[1])
            batch_dict = second._publish(batch_dict, res2, tok2[1])
        else:
            if first_no_grad and grad_on:
                torch.set_grad_enabled(False)
            batch_dict = first.finish(batch_dict, tok1)
            torch.set_grad_enabled(grad_on)
            torch.cuda.set_stream(side)
            batch_dict = second.finish(batch_dict, tok2)
    finally:
        torch.cuda.set_stream(main)
        torch.set_grad_enabled(grad_on)
    main.wait_stream(side)
    for k in second._feature_keys + (second._coords_key,):
        batch_dict[k].record_stream(main)
    return batch_dict


_side_streams = {}


def _side_stream(device):
    s = _side_streams.get(device.index)
    if s is None:
        s = torch.cuda.Stream(device)
        _side_streams[device.index] = s
    return s


class DynamicPillarVFE(_FusedDynamicPillarVFE):
    _layout = _lib.LAYOUT_DYNPILLAR
    _feature_keys = ("voxel_features", "pillar_features")
    _coords_key = "voxel_coords"


class DynamicPillarVFESimple2D(_FusedDynamicPillarVFE):
    pass


class DynamicVoxelVFE(_FusedDynamicPillarVFE):
    """``dynamic_voxel_vfe.py:15-106``: 3-D voxel key (z quantised and masked too), voxel centre incl. z, features
    ``[points | f_cluster | f_center | dist]``, PFN layer stack, coords ``[b, z, y, x]``.  Layer-stack path."""
    _layout = _lib.LAYOUT_DYNVOXEL
    _feature_keys = ("pillar_features", "voxel_features")
    _coords_key = "voxel_coords"


class DynamicMeanVFE(VFETemplate):
    """``dynamic_mean_vfe.py:14-76``: per-voxel mean of every point column; no parameters, no gradient."""

    def __init__(self, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, **kwargs):
        super().__init__(model_cfg=model_cfg)
        self.num_point_features = num_point_features
        self.spec = ops.make_spec(num_point_features, voxel_size, grid_size, point_cloud_range, _lib.LAYOUT_DYNVOXEL,
                                  True, True, False, False, num_point_features)
        self.voxel_x, self.voxel_y, self.voxel_z = voxel_size[0], voxel_size[1], voxel_size[2]
        self.x_offset, self.y_offset, self.z_offset = self.spec.off
        object.__setattr__(self, "last_result", None)

    def get_output_feature_dim(self):
        return self.num_point_features

    @torch.no_grad()
    def forward(self, batch_dict, **kwargs):
        points = batch_dict["points"]
        idx = stack_ops.index_forward(points, self.spec, int(batch_dict["batch_size"]))
        object.__setattr__(self, "last_result", idx)
        batch_dict["voxel_features"] = stack_ops.voxel_mean(idx)
        batch_dict["voxel_coords"] = idx.coords
        return batch_dict


class Radar_DynamicPillarVFESimple2D(DynamicPillarVFESimple2D):
    _points_key = "radar_points"
    _feature_keys = ("radar_pillar_features",)
    _coords_key = "radar_pillar_coords"


class Radar_DynamicPillarVFESimple2D_Test(DynamicPillarVFESimple2D):
    _points_key = "points"
    _feature_keys = ("radar_pillar_features",)
    _coords_key = "radar_pillar_coords"


REGISTRY = {
    "VFETemplate": VFETemplate,
    "DynPillarVFE": DynamicPillarVFE,
    "DynamicPillarVFESimple2D": DynamicPillarVFESimple2D,
    "DynamicVoxelVFE": DynamicVoxelVFE,
    "DynMeanVFE": DynamicMeanVFE,
    "Radar_DynamicPillarVFESimple2D": Radar_DynamicPillarVFESimple2D,
    "Radar_DynamicPillarVFESimple2D_Test": Radar_DynamicPillarVFESimple2D_Test,
}


def register(vfe_all: dict) -> dict:
    """Overwrites the dynamic-pillar entries of ``pcdet.models.backbones_3d.vfe.__all__`` in place."""
    for name in ("DynPillarVFE", "DynamicPillarVFESimple2D", "DynamicVoxelVFE", "DynMeanVFE", "Radar_DynamicPillarVFESimple2D",
                 "Radar_DynamicPillarVFESimple2D_Test"):
        vfe_all[name] = REGISTRY[name]
    return vfe_all
