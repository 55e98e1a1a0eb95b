"""CPU, build container only: the C oracle against the REAL reference module run live from /root/reference.

The committed goldens (tests/golden, oracle/gen_golden.py) are outputs of the reference on fixed seeds; this test draws
fresh clouds and fresh weights every time it is parametrised, runs ``dynamic_pillar_vfe.py`` itself through
``oracle/ref_loader.py`` and holds the oracle to the same bars.  ``/root/reference`` does not exist on the GPU box, so the
whole module is skipped there (nothing under ``-m gpu`` may read the reference).
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import ref_loader
from radardistill_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present (GPU box)")

CASES = [  # (class name, C, model_cfg overrides, voxel size, training)
    ("DynamicPillarVFESimple2D", 5, {}, [0.075, 0.075, 0.2], False),
    ("DynamicPillarVFESimple2D", 5, {"USE_RELATIVE_XYZ": False}, [0.3, 0.3, 8.0], True),
    ("Radar_DynamicPillarVFESimple2D", 6, {}, [0.075, 0.075, 0.2], True),
    ("Radar_DynamicPillarVFESimple2D_Test", 6, {"USE_CLUSTER_XYZ": False, "WITH_DISTANCE": True}, [0.6, 0.6, 8.0], True),
    ("DynPillarVFE", 4, {"NUM_FILTERS": [64]}, [0.2, 0.2, 8.0], True),
    ("DynPillarVFE", 4, {"USE_NORM": False, "NUM_FILTERS": [32]}, [0.4, 0.4, 8.0], False),
]


@pytest.mark.parametrize("seed", [101, 202])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-C{c[1]}-{'train' if c[4] else 'eval'}")
def test_oracle_matches_live_reference(case, seed):
    name, C, over, voxel, training = case
    cfg = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])
    cfg.update(over)
    rng = np.random.default_rng(seed)
    # a few hundred to a few thousand points over 3 frames, some exactly on the range boundary, some duplicated
    frames = []
    for b in range(3):
        n = int(rng.integers(300, 1500))
        f = np.concatenate([rng.uniform(-54, 54, (n, 2)), rng.normal(0, 1.5, (n, 1)), rng.normal(0, 3, (n, C - 3))], 1).astype(np.float32)
        f[:5, 0] = 54.0
        f[5:25] = f[25:45]
        frames.append(f)
    pts = synth.collate(frames)
    grid = synth.grid_size_of(synth.PC_RANGE, voxel)
    torch.manual_seed(seed)
    ref = ref_loader.build_reference(name, cfg, C, voxel, grid, np.asarray(synth.PC_RANGE, np.float32))
    pfn = ref.pfn_layers[0]
    with torch.no_grad():
        if cfg["USE_NORM"]:
            pfn.norm.weight.uniform_(0.5, 1.5); pfn.norm.bias.normal_(0, 0.2)
            pfn.norm.running_mean.normal_(0, 1); pfn.norm.running_var.uniform_(0.5, 4)
    ref.train(training)
    cap = {}
    key = "radar_points" if name == "Radar_DynamicPillarVFESimple2D" else "points"
    out = ref_loader.run_reference(ref, torch.from_numpy(pts), key, capture=cap)
    fkey = [k for k in out if k.endswith("pillar_features")][0]
    ckey = [k for k in out if k.endswith("_coords")][0]
    feats, coords = out[fkey], out[ckey]

    ocfg = orc.config_for(name, C, voxel, grid, synth.PC_RANGE, cfg)
    cp = lambda t: t.detach().numpy().copy()
    kw = dict(weight=cp(pfn.linear.weight))
    if cfg["USE_NORM"]:
        # module state BEFORE the forward for the running statistics (train mode updates them in place)
        kw.update(gamma=cp(pfn.norm.weight), beta=cp(pfn.norm.bias))
    else:
        kw.update(bias=cp(pfn.linear.bias))
    if cfg["USE_NORM"] and not training:
        kw.update(running_mean=cp(pfn.norm.running_mean), running_var=cp(pfn.norm.running_var))
    o = orc.PillarOracle(ocfg, **kw)
    r = o.forward(pts, training=training)

    np.testing.assert_array_equal(r["coords"], coords.numpy())
    np.testing.assert_array_equal(r["inverse"], cap["inverse"].numpy())
    np.testing.assert_array_equal(r["counts"], cap["counts"].numpy())
    assert H.norm_rel_err(r["features"], feats.detach().numpy()) <= H.RTOL_FEATURES
    scale = max(float(np.abs(feats.detach().numpy()).max()), 1e-30)
    post = np.maximum(r["x"] * r["scale"] + r["shift"], 0.0)
    assert H.argmax_mismatch_is_near_tie(r["argmax"], cap["argmax"].numpy(), post, r["inverse"], H.RTOL_FEATURES * scale)
    if training:
        g = torch.from_numpy(rng.standard_normal(tuple(feats.shape)).astype(np.float32))
        feats.backward(g)
        b = o.backward(r, g.numpy())
        assert H.norm_rel_err(b["d_weight"], pfn.linear.weight.grad.numpy()) <= H.RTOL_GRADS_REF
        if cfg["USE_NORM"]:
            assert H.norm_rel_err(b["d_gamma"], pfn.norm.weight.grad.numpy()) <= H.RTOL_GRADS_REF
            assert H.norm_rel_err(b["d_beta"], pfn.norm.bias.grad.numpy()) <= H.RTOL_GRADS_REF
