"""Generates tests/golden/stack__*.npz: the REAL reference on CPU for the layer-stack / voxel / mean encoders.

    python oracle/gen_golden_stack.py      # needs /root/reference (build container only)

TEST INFRASTRUCTURE ONLY (see oracle/gen_golden.py).  Cases: DynamicPillarVFE and DynamicPillarVFESimple2D with two PFN
layers (dynamic_pillar_vfe.py:24-25,42-46,123-124), DynamicVoxelVFE with one and two layers (dynamic_voxel_vfe.py) and
DynamicMeanVFE (dynamic_mean_vfe.py); eval and train mode; outputs, inverse / counts, and the gradients of every parameter.
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
VOX = [1.2, 1.2, 1.6]     # 90 x 90 x 5 voxels over the nuScenes range: several points per voxel


def variants():
    """(tag, class name, C, model_cfg, voxel)"""
    yield "dynpillar_2layer", "DynPillarVFE", 4, dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[32, 64]), synth.VOXEL_SIZE
    yield "s2d_2layer_dist", "DynamicPillarVFESimple2D", 5, dict(USE_NORM=True, WITH_DISTANCE=True, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True,
                                                                 NUM_FILTERS=[64, 32]), [0.9, 0.9, 8.0]
    yield "s2d_3layer_nonorm", "DynamicPillarVFESimple2D", 5, dict(USE_NORM=False, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True,
                                                                   NUM_FILTERS=[16, 32, 24]), [0.9, 0.9, 8.0]
    yield "voxel_1layer", "DynamicVoxelVFE", 4, dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[32]), VOX
    yield "voxel_2layer_dist_noabs", "DynamicVoxelVFE", 5, dict(USE_NORM=True, WITH_DISTANCE=True, USE_ABSLOTE_XYZ=False, NUM_FILTERS=[32, 48]), VOX
    yield "mean_vfe", "DynMeanVFE", 5, dict(), VOX


def cloud(C):
    base = synth.collate([synth.lidar_frame(31, sweeps=1, beams=8, azimuths=120), synth.lidar_frame(32, sweeps=2, beams=6, azimuths=90)])
    extra = np.array([[0, 1.0, 1.0, 3.0, 1, 0], [0, 1.0, 1.0, -5.0, 1, 0], [1, 2.0, 2.0, 2.99999, 1, 0], [1, 2.0, 2.0, -5.0001, 1, 0],
                      [1, 54.0, 0.0, 0.0, 1, 0], [0, np.nan, 0.0, 0.0, 1, 0]], np.float32)   # z on / beyond the bounds, xy bound, NaN
    pts = np.concatenate([base, extra], 0)
    dup = pts[:40].copy()                                                                      # exact duplicates -> argmax ties
    return np.ascontiguousarray(np.concatenate([pts, dup], 0)[:, :1 + C])


def run_case(name, C, cfg, voxel, pts_np, training, seed):
    g = torch.Generator().manual_seed(seed)
    grid = synth.grid_size_of(synth.PC_RANGE, voxel)
    m = rl.build_reference(name, cfg, C, voxel, grid, synth.PC_RANGE)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("linear.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif n.endswith("norm.weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
                p[::5] *= -1.0
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)
        for n, b in m.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) * 4 + 0.5)
    m.train(training)
    params = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    cap = {}
    batch = int(pts_np[:, 0].max()) + 1
    out = rl.run_reference(m, torch.from_numpy(pts_np.copy()), capture=cap, extra={"batch_size": batch})
    feats = out["voxel_features"] if "voxel_features" in out else out["pillar_features"]
    ckey = "voxel_coords" if "voxel_coords" in out else "pillar_coords"
    rec = dict(points=pts_np, features=feats.detach().numpy(), coords=out[ckey].numpy().astype(np.int32), coords_key=ckey, training=training,
               batch_size=batch, counts=cap.get("counts", torch.zeros(0, dtype=torch.long)).numpy().astype(np.int32))
    if "inverse" in cap:
        rec["inverse"] = cap["inverse"].numpy().astype(np.int32)
    for k, v in params.items():
        rec["param." + k] = v
    if feats.requires_grad:
        gout = torch.randn(feats.shape, generator=g)
        feats.backward(gout)
        rec["grad_features"] = gout.numpy()
        for n, p in m.named_parameters():
            rec["grad." + n] = p.grad.numpy()
    if training:
        for k, v in m.state_dict().items():
            if "running" in k or "num_batches" in k:
                rec["new." + k] = v.detach().clone().numpy()
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    seed = 500
    for tag, name, C, cfg, voxel in variants():
        pts = cloud(C)
        for training in ((False,) if name == "DynMeanVFE" else (False, True)):
            seed += 1
            rec = run_case(name, C, cfg, voxel, pts, training, seed)
            rec.update(class_name=name, num_point_features=C, voxel_size=np.asarray(voxel, np.float64), model_cfg=repr(cfg))
            fn = os.path.join(OUT, f"stack__{tag}__{'train' if training else 'eval'}.npz")
            np.savez_compressed(fn, **rec)
            print(f"{os.path.basename(fn):52s} N0={len(pts):6d} P={rec['features'].shape[0]:6d} C={rec['features'].shape[1]} {os.path.getsize(fn)/1e3:.0f} kB")


if __name__ == "__main__":
    main()
