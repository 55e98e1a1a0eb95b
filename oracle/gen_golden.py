"""Generates tests/golden/*.npz by running the REAL reference encoder on CPU.

    python oracle/gen_golden.py            # needs /root/reference (build container only)

TEST INFRASTRUCTURE ONLY.  The reference (``pcdet/models/backbones_3d/vfe/
dynamic_pillar_vfe.py``) is imported from ``/root/reference`` through
``oracle/ref_loader.py`` (torch_scatter shim, see there); it cannot travel to the GPU
box, so the vectors it produces are committed.  Each file holds one case: inputs
(points, parameters, upstream gradient) and the reference's outputs (features, coords,
inverse, counts, argmax, parameter gradients, updated running statistics).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as rl  # noqa: E402
from radardistill_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

S2D = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])


def variants():
    """(tag, class name, C, points_key, model_cfg, voxel) -- the shipped configs first."""
    yield "lidar_s2d", "DynamicPillarVFESimple2D", 5, "points", dict(S2D), synth.VOXEL_SIZE
    yield "radar_s2d", "Radar_DynamicPillarVFESimple2D", 6, "radar_points", dict(S2D), synth.VOXEL_SIZE
    yield "radar_test", "Radar_DynamicPillarVFESimple2D_Test", 6, "points", dict(S2D), synth.VOXEL_SIZE
    yield "dynpillar64", "DynPillarVFE", 4, "points", dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True,
                                                          NUM_FILTERS=[64]), synth.VOXEL_SIZE
    yield "dynpillar_dist_noabs", "DynPillarVFE", 4, "points", dict(USE_NORM=True, WITH_DISTANCE=True,
                                                                   USE_ABSLOTE_XYZ=False, NUM_FILTERS=[32]), [0.2, 0.2, 8.0]
    yield "s2d_nonorm_dist", "DynamicPillarVFESimple2D", 5, "points", dict(S2D, USE_NORM=False, WITH_DISTANCE=True), \
        synth.VOXEL_SIZE
    yield "s2d_minimal", "DynamicPillarVFESimple2D", 5, "points", dict(S2D, USE_ABSLOTE_XYZ=False, USE_CLUSTER_XYZ=False,
                                                                       USE_RELATIVE_XYZ=False), [0.3, 0.3, 8.0]


def clouds(tag, C):
    """(cloud tag, (N, 1+C) float32) per variant."""
    if tag == "lidar_s2d":
        for kind in synth.EDGE_KINDS:
            yield kind, synth.edge_case_points(kind)
        yield "sweep2", synth.collate([synth.lidar_frame(11, sweeps=1, beams=8, azimuths=160),
                                       synth.lidar_frame(12, sweeps=2, beams=6, azimuths=96)])
    elif tag.startswith("radar"):
        yield "b2", synth.collate([synth.radar_frame(3, 500), synth.radar_frame(4, 420)])
        yield "dup", synth.collate([np.concatenate([synth.radar_frame(5, 150)] * 3)])
    else:
        base = synth.collate([synth.lidar_frame(21, sweeps=1, beams=8, azimuths=120),
                              synth.lidar_frame(22, sweeps=1, beams=8, azimuths=100)])
        yield "b2", np.ascontiguousarray(base[:, :1 + C])


def run_case(name, C, key, cfg, voxel, pts_np, training, seed):
    g = torch.Generator().manual_seed(seed)
    grid = synth.grid_size_of(synth.PC_RANGE, voxel)
    m = rl.build_reference(name, cfg, C, voxel, grid, synth.PC_RANGE)
    pfn = m.pfn_layers[0]
    with torch.no_grad():
        pfn.linear.weight.copy_(torch.randn(pfn.linear.weight.shape, generator=g) * 0.3)
        if cfg["USE_NORM"]:
            pfn.norm.weight.copy_(torch.rand(pfn.norm.weight.shape, generator=g) + 0.5)
            pfn.norm.weight[::5] *= -1.0  # negative gammas flip the arg ordering
            pfn.norm.bias.copy_(torch.randn(pfn.norm.bias.shape, generator=g) * 0.2)
            pfn.norm.running_mean.copy_(torch.randn(pfn.norm.running_mean.shape, generator=g))
            pfn.norm.running_var.copy_(torch.rand(pfn.norm.running_var.shape, generator=g) * 4 + 0.5)
        else:
            pfn.linear.bias.copy_(torch.randn(pfn.linear.bias.shape, generator=g) * 0.2)
    m.train(training)
    params = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    cap = {}
    pts = torch.from_numpy(pts_np.copy())
    out = rl.run_reference(m, pts, points_key=key, capture=cap)
    fkey = [k for k in out if k.endswith("pillar_features")][0]
    ckey = [k for k in out if k.endswith("_coords")][0]
    feats = out[fkey]
    rec = dict(points=pts_np, features=feats.detach().numpy(), coords=out[ckey].numpy(),
               feature_key=fkey, coords_key=ckey, training=training,
               inverse=cap.get("inverse", torch.zeros(0, dtype=torch.long)).numpy().astype(np.int32),
               counts=cap.get("counts", torch.zeros(0, dtype=torch.long)).numpy().astype(np.int32),
               argmax=cap.get("argmax", torch.zeros((0, feats.shape[1]), dtype=torch.long)).numpy().astype(np.int32))
    for k, v in params.items():
        rec["param." + k] = v
    if feats.requires_grad and feats.numel() > 0:
        gout = torch.randn(feats.shape, generator=g)
        feats.backward(gout)
        rec["grad_features"] = gout.numpy()
        rec["grad.linear.weight"] = pfn.linear.weight.grad.numpy()
        if cfg["USE_NORM"]:
            rec["grad.norm.weight"] = pfn.norm.weight.grad.numpy()
            rec["grad.norm.bias"] = pfn.norm.bias.grad.numpy()
        else:
            rec["grad.linear.bias"] = pfn.linear.bias.grad.numpy()
    if training and cfg["USE_NORM"]:
        rec["new.running_mean"] = pfn.norm.running_mean.numpy().copy()
        rec["new.running_var"] = pfn.norm.running_var.numpy().copy()
        rec["new.num_batches_tracked"] = pfn.norm.num_batches_tracked.numpy().copy()
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    seed, total = 0, 0
    for tag, name, C, key, cfg, voxel in variants():
        for ctag, pts in clouds(tag, C):
            for training in (False, True):
                seed += 1
                if training and cfg["USE_NORM"] and ctag in ("single",):
                    continue  # reference raises (BatchNorm needs > 1 value) -- covered by a raise-test instead
                rec = run_case(name, C, key, cfg, voxel, pts, training, seed)
                rec.update(class_name=name, num_point_features=C, voxel_size=np.asarray(voxel, np.float64),
                           model_cfg=repr(cfg))
                fn = os.path.join(OUT, f"{tag}__{ctag}__{'train' if training else 'eval'}.npz")
                np.savez_compressed(fn, **rec)
                total += os.path.getsize(fn)
                print(f"{os.path.basename(fn):58s} N0={len(pts):6d} P={rec['features'].shape[0]:6d}")
    print(f"total {total / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
