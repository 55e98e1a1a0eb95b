"""GPU: the layer-stack path (multi-layer PFN, DynamicVoxelVFE, DynamicMeanVFE), device-side input prep and the AMP opt-out.

Bars: coords / inverse / counts bit-exact against the reference goldens (tests/golden/stack__*.npz) and the numpy oracle;
features within 1e-5 (norm-relative) of both; every parameter gradient within 1e-5 of the reference's autograd."""
import glob
import os

import numpy as np
import pytest
import torch

from radardistill_b200 import synth
from tests import helpers as H
from tests.test_stack_oracle_golden import FILES, load, oracle_forward

pytestmark = pytest.mark.gpu


def build(g):
    from oracle.ref_loader import Cfg
    from radardistill_b200.vfe import REGISTRY
    voxel = [float(v) for v in g["voxel_size"]]
    m = REGISTRY[g["class_name"]](model_cfg=Cfg(g["model_cfg"]), num_point_features=int(g["num_point_features"]), voxel_size=voxel,
                                  grid_size=synth.grid_size_of(synth.PC_RANGE, voxel), point_cloud_range=synth.PC_RANGE)
    sd = {k: torch.from_numpy(v) for k, v in g["params"].items()}
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.cuda()


@pytest.mark.parametrize("fn", FILES, ids=[os.path.basename(f)[7:-4] for f in FILES])
def test_stack_path_matches_reference_and_oracle(fn):
    g = load(fn)
    m = build(g).train(bool(g["training"]))
    out = m({"points": torch.from_numpy(g["points"]).cuda(), "batch_size": int(g["batch_size"])})
    feats = out["voxel_features"] if "voxel_features" in out else out["pillar_features"]
    coords = out[str(g["coords_key"])]
    r = oracle_forward(g)
    np.testing.assert_array_equal(coords.cpu().numpy(), g["coords"])
    idx = m.last_result
    np.testing.assert_array_equal(idx.counts.cpu().numpy(), g["counts"])
    if "inverse" in g:
        np.testing.assert_array_equal(idx.inverse.cpu().numpy(), g["inverse"])
    f = feats.detach().cpu().numpy()
    assert f.shape == g["features"].shape and feats.dtype == torch.float32
    assert H.norm_rel_err(f, g["features"]) <= H.RTOL_FEATURES
    assert H.norm_rel_err(f, r["features"]) <= H.RTOL_FEATURES
    if "grad_features" in g:
        feats.backward(torch.from_numpy(g["grad_features"]).cuda())
        for n, p in m.named_parameters():
            assert H.norm_rel_err(p.grad.cpu().numpy(), g["grad." + n]) <= H.RTOL_GRADS, n
    if g["training"]:
        for k, v in g.items():
            if k.startswith("new."):
                got = m.state_dict()[k[4:]].cpu().numpy()
                assert H.norm_rel_err(got, v) <= 1e-5, k


def test_unsupported_fused_shape_takes_the_stack_path():
    """A single PFN layer whose channel count the fused kernels are not compiled for (48) must still work (the reference accepts it)."""
    from oracle.ref_loader import Cfg
    from oracle import stack_oracle as so
    from radardistill_b200 import vfe
    cfg = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[48])
    voxel = [0.9, 0.9, 8.0]
    torch.manual_seed(3)
    m = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(cfg), num_point_features=5, voxel_size=voxel,
                                     grid_size=synth.grid_size_of(synth.PC_RANGE, voxel), point_cloud_range=synth.PC_RANGE).cuda().eval()
    assert not m.fused
    pts = synth.collate([synth.lidar_frame(41, sweeps=1, beams=8, azimuths=100)])
    out = m({"points": torch.from_numpy(pts).cuda(), "batch_size": 1})
    params = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    r = so.forward(pts, params, "DynamicPillarVFESimple2D", cfg, 5, voxel, synth.grid_size_of(synth.PC_RANGE, voxel), synth.PC_RANGE, False)
    np.testing.assert_array_equal(out["pillar_coords"].cpu().numpy(), r["coords"])
    assert H.norm_rel_err(out["pillar_features"].detach().cpu().numpy(), r["features"]) <= H.RTOL_FEATURES


def test_prepare_points_mask_and_shuffle():
    """Device-side mask_points_by_range (+ shuffle) against numpy (data_processor.py:80-86, :99-114)."""
    from radardistill_b200 import stack_ops
    rng = np.random.default_rng(5)
    pts = rng.uniform(-60, 60, (50_000, 5)).astype(np.float32)
    pts[:7, 0] = [54.0, -54.0, 54.00001, -54.00001, np.nan, np.inf, 0.0]
    pts[:7, 1] = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 54.0]
    pr = synth.PC_RANGE
    keep = (pts[:, 0] >= pr[0]) & (pts[:, 0] <= pr[3]) & (pts[:, 1] >= pr[1]) & (pts[:, 1] <= pr[4])
    d = torch.from_numpy(pts).cuda()
    out = stack_ops.prepare_points(d, pr).cpu().numpy()
    np.testing.assert_array_equal(out, pts[keep])                       # stable compaction
    sh = stack_ops.prepare_points(d, pr, shuffle_seed=12345).cpu().numpy()
    assert sh.shape == out.shape and not np.array_equal(sh, out)
    order = lambda a: a[np.lexsort(a.T[::-1])]
    np.testing.assert_array_equal(order(sh), order(out))                # a permutation of the kept rows
    sh2 = stack_ops.prepare_points(d, pr, shuffle_seed=12345).cpu().numpy()
    np.testing.assert_array_equal(sh, sh2)                              # deterministic per seed
    assert stack_ops.prepare_points(d[:0], pr).shape[0] == 0
    # batched rows [b, x, y, ...]: x in column 1
    bpts = np.concatenate([np.zeros((len(pts), 1), np.float32), pts], 1)
    outb = stack_ops.prepare_points(torch.from_numpy(bpts).cuda(), pr, x_col=1).cpu().numpy()
    np.testing.assert_array_equal(outb[:, 1:], pts[keep])


def test_autocast_leaves_the_encoder_in_fp32():
    """--use_amp (tools/train.py:52, train_utils.py:57-64): under torch.autocast the fused encoder still computes and returns fp32,
    bit-identical to the plain call, and its backward accepts a half-precision upstream gradient."""
    g = H.load_golden(os.path.join(H.GOLDEN_DIR, "radar_s2d__b2__train.npz"))
    pts = torch.from_numpy(g["points"]).cuda()
    m1, m2 = H.module_from_golden(g).train(), H.module_from_golden(g).train()
    ref = m1({"radar_points": pts})["radar_pillar_features"]
    with torch.autocast("cuda", dtype=torch.float16):
        got = m2({"radar_points": pts})["radar_pillar_features"]
        assert got.dtype == torch.float32
        loss = (got.half() * torch.from_numpy(g["grad_features"]).cuda().half()).sum()
    assert torch.equal(ref, got)
    loss.backward()
    ref.backward(torch.from_numpy(g["grad_features"]).cuda().half().float())
    w1, w2 = m1.pfn_layers[0].linear.weight.grad, m2.pfn_layers[0].linear.weight.grad
    assert H.norm_rel_err(w2.cpu().numpy(), w1.cpu().numpy()) <= 1e-3   # the upstream gradient went through fp16
    # the layer-stack path's pooling op opts out as well: fp16 activations in, fp32 maxima out
    g2 = load([f for f in FILES if "dynpillar_2layer__eval" in f][0])
    ms = build(g2).eval()
    with torch.autocast("cuda", dtype=torch.float16):
        o = ms({"points": torch.from_numpy(g2["points"]).cuda(), "batch_size": int(g2["batch_size"])})
    assert o["pillar_features"].dtype == torch.float32
    assert H.norm_rel_err(o["pillar_features"].detach().cpu().numpy(), g2["features"]) <= 5e-3   # the GEMMs ran in fp16, as the reference's would
