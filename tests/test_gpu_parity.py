"""GPU: the CUDA encoder (through the drop-in modules -> C ABI) against the oracle and the goldens.

Bars (tests/helpers.py):  coords / inverse / counts / argmax bit-exact vs the oracle; forward features
bit-exact vs the oracle's canonical arithmetic and within RTOL_FEATURES (norm-relative 1e-5) of the
reference's own output; parameter gradients within RTOL_GRADS of oracle and reference.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu
FILES = H.golden_files()


def run_module(m, g, pts_np, key, batch_size=None):
    pts = torch.from_numpy(pts_np).cuda()
    bd = {key: pts}
    if batch_size is not None:
        bd["batch_size"] = batch_size
    out = m(bd)
    fkey = [k for k in out if k.endswith("pillar_features")][0]
    ckey = [k for k in out if k.endswith("_coords")][0]
    return out, out[fkey], out[ckey], m.last_result


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_matches_oracle_and_reference(path):
    g = H.load_golden(path)
    o = H.oracle_from_golden(g, orc.FOLDED)
    m = H.module_from_golden(g)
    m.train(g["training"])
    r = o.forward(g["points"], training=g["training"])
    out, feats, coords, res = run_module(m, g, g["points"], H.points_key_of(g))
    assert str(g["feature_key"]) in out and str(g["coords_key"]) in out
    # ---- integer outputs: bit-exact vs oracle AND vs the reference's golden
    assert res.n_kept == r["n"] and res.n_pillars == r["p"]
    assert coords.dtype == torch.int32 and feats.dtype == torch.float32
    assert tuple(coords.shape) == tuple(g["coords"].shape) and tuple(feats.shape) == tuple(g["features"].shape)
    np.testing.assert_array_equal(coords.cpu().numpy(), r["coords"])
    np.testing.assert_array_equal(coords.cpu().numpy(), g["coords"])
    np.testing.assert_array_equal(res.inverse.cpu().numpy(), g["inverse"])
    np.testing.assert_array_equal(res.counts.cpu().numpy(), g["counts"])
    if r["p"] == 0:
        return
    # ---- features: bit-exact vs the oracle's canonical arithmetic, tolerance vs the reference
    f = feats.detach().cpu().numpy()
    if not g["training"]:
        np.testing.assert_array_equal(f, r["features"])
    else:  # batch statistics are fp64 sums in a different order: allow the last-ulp effect of that
        assert H.norm_rel_err(f, r["features"]) <= 1e-6
    assert H.norm_rel_err(f, g["features"]) <= H.RTOL_FEATURES
    # ---- backward: argmax bit-exact, gradients in tolerance
    if "grad_features" in g:
        np.testing.assert_array_equal(res.argmax.cpu().numpy(), r["argmax"])
        gout = torch.from_numpy(g["grad_features"]).cuda()
        feats.backward(gout)
        b = o.backward(r, g["grad_features"])
        pfn = m.pfn_layers[0]
        dW = pfn.linear.weight.grad.cpu().numpy()
        assert H.norm_rel_err(dW, b["d_weight"]) <= H.RTOL_GRADS
        assert H.norm_rel_err(dW, g["grad.linear.weight"]) <= H.RTOL_GRADS
        if o.cfg.use_norm:
            assert H.norm_rel_err(pfn.norm.weight.grad.cpu().numpy(), b["d_gamma"]) <= H.RTOL_GRADS
            assert H.norm_rel_err(pfn.norm.bias.grad.cpu().numpy(), b["d_beta"]) <= H.RTOL_GRADS
            assert H.norm_rel_err(pfn.norm.weight.grad.cpu().numpy(), g["grad.norm.weight"]) <= H.RTOL_GRADS
            assert H.norm_rel_err(pfn.norm.bias.grad.cpu().numpy(), g["grad.norm.bias"]) <= H.RTOL_GRADS
        else:
            assert H.norm_rel_err(pfn.linear.bias.grad.cpu().numpy(), g["grad.linear.bias"]) <= H.RTOL_GRADS
    if "new.running_mean" in g:
        pfn = m.pfn_layers[0]
        assert H.norm_rel_err(pfn.norm.running_mean.cpu().numpy(), g["new.running_mean"]) <= H.RTOL_FEATURES
        assert H.norm_rel_err(pfn.norm.running_var.cpu().numpy(), g["new.running_var"]) <= H.RTOL_FEATURES
        assert int(pfn.norm.num_batches_tracked) == int(g["new.num_batches_tracked"])


def _shipped_module(kind, train):
    from oracle.ref_loader import Cfg
    from radardistill_b200 import synth, vfe
    cfg = Cfg(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])
    cls, c = (vfe.DynamicPillarVFESimple2D, 5) if kind == "lidar" else (vfe.Radar_DynamicPillarVFESimple2D, 6)
    torch.manual_seed(7)
    m = cls(model_cfg=cfg, num_point_features=c, voxel_size=synth.VOXEL_SIZE, grid_size=synth.grid_size_of(),
            point_cloud_range=synth.PC_RANGE).cuda()
    n = m.pfn_layers[0].norm
    with torch.no_grad():
        n.weight.uniform_(0.5, 1.5); n.bias.normal_(0, 0.2); n.running_mean.normal_(0, 1); n.running_var.uniform_(0.5, 4)
    m.train(train)
    return m


def _oracle_of(m, c):
    from radardistill_b200 import synth
    pfn = m.pfn_layers[0]
    cfg = orc.OracleConfig(num_point_features=c, voxel_size=tuple(synth.VOXEL_SIZE), grid_size=tuple(synth.grid_size_of()),
                           point_cloud_range=tuple(synth.PC_RANGE))
    cp = lambda t: t.detach().cpu().numpy().copy()
    return orc.PillarOracle(cfg, cp(pfn.linear.weight), cp(pfn.norm.weight), cp(pfn.norm.bias), cp(pfn.norm.running_mean),
                            cp(pfn.norm.running_var))


def test_lidar_two_frames_bit_exact_vs_oracle():
    """Config-2-sized case (2 x ~340 k points): every output bit-exact against the oracle."""
    from radardistill_b200 import synth
    pts = synth.lidar_batch(2)
    m = _shipped_module("lidar", train=False)
    o = _oracle_of(m, 5)
    orc.set_threads(8)
    r = o.forward(pts, training=False, keep_intermediates=False)
    with torch.no_grad():
        out = m({"points": torch.from_numpy(pts).cuda(), "batch_size": 2})
    res = m.last_result
    assert res.n_pillars == r["p"] and res.n_kept == r["n"]
    np.testing.assert_array_equal(out["pillar_coords"].cpu().numpy(), r["coords"])
    np.testing.assert_array_equal(res.inverse.cpu().numpy(), r["inverse"])
    np.testing.assert_array_equal(res.counts.cpu().numpy(), r["counts"])
    np.testing.assert_array_equal(out["pillar_features"].cpu().numpy(), r["features"])


def test_radar_train_step_vs_oracle():
    """Config-1 radar branch: train-mode BN forward + backward on a batch of 4 frames."""
    from radardistill_b200 import synth
    pts = synth.radar_batch(4)
    m = _shipped_module("radar", train=True)
    o = _oracle_of(m, 6)
    r = o.forward(pts, training=True)
    out = m({"radar_points": torch.from_numpy(pts).cuda(), "batch_size": 4})
    res = m.last_result
    np.testing.assert_array_equal(out["radar_pillar_coords"].cpu().numpy(), r["coords"])
    np.testing.assert_array_equal(res.argmax.cpu().numpy(), r["argmax"])
    f = out["radar_pillar_features"]
    assert H.norm_rel_err(f.detach().cpu().numpy(), r["features"]) <= 1e-6
    gout = torch.randn(f.shape, generator=torch.Generator().manual_seed(3)).cuda()
    f.backward(gout)
    b = o.backward(r, gout.cpu().numpy())
    pfn = m.pfn_layers[0]
    assert H.norm_rel_err(pfn.linear.weight.grad.cpu().numpy(), b["d_weight"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(pfn.norm.weight.grad.cpu().numpy(), b["d_gamma"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(pfn.norm.bias.grad.cpu().numpy(), b["d_beta"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(pfn.norm.running_mean.cpu().numpy(), r["new_running_mean"]) <= 1e-6


def test_deterministic_and_idempotent():
    """Two runs on the same input give bit-identical outputs (atomics only touch integers)."""
    from radardistill_b200 import synth
    pts = torch.from_numpy(synth.lidar_batch(1)).cuda()
    m = _shipped_module("lidar", train=False)
    with torch.no_grad():
        a = m({"points": pts, "batch_size": 1})
        fa, ca = a["pillar_features"].clone(), a["pillar_coords"].clone()
        b = m({"points": pts, "batch_size": 1})
    assert torch.equal(fa, b["pillar_features"]) and torch.equal(ca, b["pillar_coords"])


def test_single_point_train_raises_and_missing_batch_size():
    g = H.load_golden([f for f in FILES if "single__eval" in f][0])
    m = H.module_from_golden(g).train()
    with pytest.raises(ValueError):
        m({"points": torch.from_numpy(g["points"]).cuda()})
    with pytest.raises(Exception):
        m.eval()({"points": torch.from_numpy(g["points"])})  # CPU tensor: no fallback


def test_bad_batch_index_is_reported():
    g = H.load_golden([f for f in FILES if "lidar_s2d__sweep2__eval" in f][0])
    m = H.module_from_golden(g).eval()
    with pytest.raises(ValueError):
        m({"points": torch.from_numpy(g["points"]).cuda(), "batch_size": 1})


def test_stress_frame_bit_exact_vs_oracle():
    """Config-5 stress cloud (1 M points, 0.05 m pillars, 2160 x 2160 grid, hot pillars with > 1000 points, 20 % exact
    duplicates): eval forward bit-exact, and a train step whose argmax (lowest index on the many exact ties) is bit-exact."""
    from oracle.ref_loader import Cfg
    from radardistill_b200 import synth, vfe
    grid = synth.grid_size_of(synth.PC_RANGE, synth.STRESS_VOXEL_SIZE)
    pts = synth.stress_batch(2, n_points=500_000)
    cfg = Cfg(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])
    torch.manual_seed(11)
    m = vfe.DynamicPillarVFESimple2D(model_cfg=cfg, num_point_features=5, voxel_size=synth.STRESS_VOXEL_SIZE, grid_size=grid,
                                     point_cloud_range=synth.PC_RANGE).cuda()
    n = m.pfn_layers[0].norm
    with torch.no_grad():
        n.weight.uniform_(0.5, 1.5); n.bias.normal_(0, 0.2); n.running_mean.normal_(0, 1); n.running_var.uniform_(0.5, 4)
    pfn = m.pfn_layers[0]
    ocfg = orc.OracleConfig(num_point_features=5, voxel_size=tuple(synth.STRESS_VOXEL_SIZE), grid_size=tuple(grid),
                            point_cloud_range=tuple(synth.PC_RANGE))
    cp = lambda t: t.detach().cpu().numpy().copy()
    o = orc.PillarOracle(ocfg, cp(pfn.linear.weight), cp(n.weight), cp(n.bias), cp(n.running_mean), cp(n.running_var))
    orc.set_threads(8)
    dev = torch.from_numpy(pts).cuda()
    # eval
    m.eval()
    r = o.forward(pts, training=False, keep_intermediates=False)
    assert r["counts"].max() > 1000
    with torch.no_grad():
        out = m({"points": dev, "batch_size": 2})
    np.testing.assert_array_equal(out["pillar_coords"].cpu().numpy(), r["coords"])
    np.testing.assert_array_equal(m.last_result.counts.cpu().numpy(), r["counts"])
    np.testing.assert_array_equal(m.last_result.inverse.cpu().numpy(), r["inverse"])
    np.testing.assert_array_equal(out["pillar_features"].cpu().numpy(), r["features"])
    # train: argmax with ties, gradients
    m.train()
    r = o.forward(pts, training=True)
    out = m({"points": dev, "batch_size": 2})
    np.testing.assert_array_equal(m.last_result.argmax.cpu().numpy(), r["argmax"])
    f = out["pillar_features"]
    assert H.norm_rel_err(f.detach().cpu().numpy(), r["features"]) <= 1e-6
    gout = torch.randn(f.shape, generator=torch.Generator().manual_seed(5)).cuda()
    f.backward(gout)
    b = o.backward(r, gout.cpu().numpy())
    assert H.norm_rel_err(pfn.linear.weight.grad.cpu().numpy(), b["d_weight"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(n.weight.grad.cpu().numpy(), b["d_gamma"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(n.bias.grad.cpu().numpy(), b["d_beta"]) <= H.RTOL_GRADS


def test_host_buffer_c_abi_entry_point():
    """rdp_encode_host: host pointers in, host pointers out (what a non-torch caller binds)."""
    import ctypes as C
    from radardistill_b200 import _lib, ops, synth
    lib = _lib.load()
    pts = synth.lidar_batch(1, sweeps=2)
    spec = ops.make_spec(5, synth.VOXEL_SIZE, synth.grid_size_of(), synth.PC_RANGE, _lib.LAYOUT_SIMPLE2D, True, True, True, False, 32)
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((32, 14)) * 0.2).astype(np.float32)
    gamma, beta = rng.uniform(0.5, 1.5, 32).astype(np.float32), rng.normal(0, 0.2, 32).astype(np.float32)
    rm, rv = rng.normal(0, 1, 32).astype(np.float32), rng.uniform(0.5, 4, 32).astype(np.float32)
    n = len(pts)
    feats = np.empty((n, 32), np.float32); coords = np.empty((n, 3), np.int32)
    inv = np.empty(n + 4, np.int32); cnt = np.empty(n + 4, np.int32)
    prm = _lib.PfnParams()
    vp = lambda a: a.ctypes.data_as(C.c_void_p).value
    prm.weight, prm.gamma, prm.beta, prm.running_mean, prm.running_var = vp(w), vp(gamma), vp(beta), vp(rm), vp(rv)
    prm.eps, prm.momentum, prm.train_bn = 1e-3, 0.01, 0
    geom, layout = spec.geom(1), spec.layout_struct()
    nk, npil = C.c_int64(0), C.c_int64(0)
    rc = lib.rdp_encode_host(vp(pts), n, C.byref(geom), C.byref(layout), C.byref(prm), vp(feats), vp(coords), vp(inv), vp(cnt),
                             C.byref(nk), C.byref(npil))
    assert rc == 0, lib.rdp_status_string(rc)
    cfg = orc.OracleConfig(num_point_features=5, voxel_size=tuple(synth.VOXEL_SIZE), grid_size=tuple(synth.grid_size_of()),
                           point_cloud_range=tuple(synth.PC_RANGE))
    r = orc.PillarOracle(cfg, w, gamma, beta, rm, rv).forward(pts)
    assert (nk.value, npil.value) == (r["n"], r["p"])
    np.testing.assert_array_equal(coords[:r["p"]], r["coords"])
    np.testing.assert_array_equal(inv[:r["n"]], r["inverse"])
    np.testing.assert_array_equal(feats[:r["p"]], r["features"])


def test_forward_pair_two_streams_matches_sequential():
    """forward_pair (radar on a side stream, LiDAR on the current one) gives bit-identical outputs and gradients."""
    from radardistill_b200 import synth
    from radardistill_b200.vfe import forward_pair
    lp, rp = torch.from_numpy(synth.lidar_batch(2, sweeps=3)).cuda(), torch.from_numpy(synth.radar_batch(2)).cuda()
    outs = []
    for paired in (False, True):
        lid, rad = _shipped_module("lidar", train=False), _shipped_module("radar", train=True)
        bd = {"points": lp, "radar_points": rp, "batch_size": 2}
        if paired:
            bd = forward_pair(lid, rad, bd, first_no_grad=True)
        else:
            with torch.no_grad():
                bd = lid(bd)
            bd = rad(bd)
        assert not bd["pillar_features"].requires_grad and bd["radar_pillar_features"].requires_grad
        bd["radar_pillar_features"].square().sum().backward()
        torch.cuda.synchronize()
        outs.append((bd["pillar_features"].clone(), bd["pillar_coords"].clone(), bd["radar_pillar_features"].detach().clone(),
                     rad.pfn_layers[0].linear.weight.grad.clone(), rad.pfn_layers[0].norm.running_var.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_host_pipeline_matches_direct_call():
    from radardistill_b200 import synth
    from radardistill_b200.pipeline import HostPipeline
    lid = _shipped_module("lidar", train=False)
    frames = [synth.lidar_batch(1, seed0=s, sweeps=2) for s in range(4)]
    pipe = HostPipeline(lambda d: lid({"points": d["points"], "batch_size": 1}), torch.device("cuda", 0),
                        ("pillar_features", "pillar_coords"))
    hosts = [{"points": torch.from_numpy(f).pin_memory()} for f in frames]
    with torch.no_grad():
        for i, h in enumerate(hosts):
            pipe.submit(h, hosts[i + 1] if i + 1 < len(hosts) else None)
            if i >= 2:  # slots are recycled after `depth` steps: read step i-2 before it is overwritten
                pass
        pipe.finish()
        for i in (1, 2, 3):  # depth 3: the last three steps are still resident
            got = pipe.result(i)
            ref = lid({"points": torch.from_numpy(frames[i]).cuda(), "batch_size": 1})
            assert torch.equal(got["pillar_features"], ref["pillar_features"].cpu())
            assert torch.equal(got["pillar_coords"], ref["pillar_coords"].cpu())


def test_torch_library_op_matches_module():
    """torch.ops.rdp.pillar_encode (+ its registered autograd) == the drop-in module, bit for bit."""
    from radardistill_b200 import synth, torch_ops
    pts = torch.from_numpy(synth.radar_batch(3)).cuda()
    m = _shipped_module("radar", train=True)
    pfn, n = m.pfn_layers[0], m.pfn_layers[0].norm
    w = pfn.linear.weight.detach().clone().requires_grad_(True)
    gamma, beta = n.weight.detach().clone().requires_grad_(True), n.bias.detach().clone().requires_grad_(True)
    ints, floats = torch_ops.spec_to_lists(m.spec)
    out = torch.ops.rdp.pillar_encode(pts, w, None, gamma, beta, n.running_mean.clone(), n.running_var.clone(), ints, floats, 3,
                                      True, True)
    feats, coords, new_rm = out[0], out[1], out[8]
    ref = m({"radar_points": pts, "batch_size": 3})
    assert torch.equal(coords, ref["radar_pillar_coords"]) and torch.equal(feats, ref["radar_pillar_features"])
    assert torch.equal(new_rm, n.running_mean)
    g = torch.randn(feats.shape, generator=torch.Generator().manual_seed(1)).cuda()
    feats.backward(g)
    ref["radar_pillar_features"].backward(g)
    assert torch.equal(w.grad, pfn.linear.weight.grad) and torch.equal(gamma.grad, n.weight.grad) and torch.equal(beta.grad, n.bias.grad)


@pytest.mark.parametrize("kind,train", [("lidar", False), ("radar", True)])
def test_frames_without_batch_column_match_padded_rows(kind, train):
    """Device-side input prep (rdp_index_fwd_frames): frames back to back + offsets give the same bits as collate_batch's
    padded rows -- coords, inverse, counts, features, argmax and parameter gradients."""
    from radardistill_b200 import synth
    frames = [synth.lidar_frame(s, sweeps=2, beams=16, azimuths=256) if kind == "lidar" else synth.radar_frame(s, n_points=700 + 37 * s)
              for s in range(5)]
    frames[3] = frames[3][:0]                       # an empty frame in the middle
    padded = synth.collate(frames)
    raw, offs = synth.collate_frames(frames)
    key = "points" if kind == "lidar" else "radar_points"
    outs = []
    for use_offsets in (False, True):
        m = _shipped_module(kind, train)
        if use_offsets:
            bd = {key: torch.from_numpy(raw).cuda(), key + "_offsets": torch.from_numpy(offs).cuda()}
        else:
            bd = {key: torch.from_numpy(padded).cuda(), "batch_size": len(frames)}
        out = m(bd)
        fk = [k for k in out if k.endswith("pillar_features")][0]
        ck = [k for k in out if k.endswith("_coords")][0]
        res = m.last_result
        rec = dict(f=out[fk].detach().cpu().numpy(), c=out[ck].cpu().numpy(), inv=res.inverse.cpu().numpy(), cnt=res.counts.cpu().numpy())
        if train:
            rec["arg"] = res.argmax.cpu().numpy()
            out[fk].backward(torch.ones_like(out[fk]))
            rec["dw"] = m.pfn_layers[0].linear.weight.grad.cpu().numpy()
        outs.append(rec)
    a, b = outs
    for k in a:
        if k in ("f", "dw") and train:     # train-mode statistics / gradient sums are fp64 sums in run-dependent order
            assert H.norm_rel_err(b[k], a[k]) <= 1e-6
        else:
            np.testing.assert_array_equal(b[k], a[k])


def test_pillar_lookup_is_the_inverse_of_coords():
    """rdp_pillar_lookup: lookup[b, y, x] == row p  <=>  coords[p] == [b, y, x]; every other cell is -1."""
    from radardistill_b200 import synth
    pts = synth.lidar_batch(3, sweeps=2, beams=16, azimuths=256)
    m = _shipped_module("lidar", train=False)
    with torch.no_grad():
        out = m({"points": torch.from_numpy(pts).cuda(), "batch_size": 3})
    coords = out["pillar_coords"].cpu().numpy().astype(np.int64)
    lut = m.last_result.pillar_lookup().cpu().numpy()
    assert lut.shape == (3, m.spec.ny, m.spec.nx) and lut.dtype == np.int32
    ref = np.full(lut.shape, -1, np.int32)
    ref[coords[:, 0], coords[:, 1], coords[:, 2]] = np.arange(len(coords), dtype=np.int32)
    np.testing.assert_array_equal(lut, ref)


@pytest.mark.parametrize("use_both", [True, False])
def test_forward_pair_single_autograd_node_matches_sequential(use_both):
    """Both encoders trainable: forward_pair wires ONE autograd node for the pair; outputs and gradients must equal two
    separate calls bit for bit (train-mode sums: to 1e-6), also when only one of the two outputs reaches the loss."""
    from radardistill_b200 import synth
    from radardistill_b200.vfe import forward_pair
    lp, rp = torch.from_numpy(synth.lidar_batch(2, sweeps=3)).cuda(), torch.from_numpy(synth.radar_batch(2)).cuda()
    outs = []
    for paired in (False, True):
        lid, rad = _shipped_module("lidar", train=True), _shipped_module("radar", train=True)
        bd = {"points": lp, "radar_points": rp, "batch_size": 2}
        bd = forward_pair(lid, rad, bd) if paired else rad(lid(bd))
        assert bd["pillar_features"].requires_grad and bd["radar_pillar_features"].requires_grad
        loss = bd["radar_pillar_features"].square().sum()
        if use_both:
            loss = loss + (bd["pillar_features"] * 0.5).sum()
        loss.backward()
        torch.cuda.synchronize()
        gl = lid.pfn_layers[0].linear.weight.grad
        assert (gl is not None) == use_both
        outs.append((bd["pillar_features"].detach().clone(), bd["pillar_coords"].clone(), bd["radar_pillar_features"].detach().clone(),
                     rad.pfn_layers[0].linear.weight.grad.clone(), rad.pfn_layers[0].norm.weight.grad.clone(),
                     gl.clone() if gl is not None else torch.zeros(1), lid.pfn_layers[0].norm.running_var.clone(),
                     lid.pfn_layers[0].norm.num_batches_tracked.clone()))
    for a, b in zip(*outs):
        if a.dtype.is_floating_point:
            assert H.norm_rel_err(b.cpu().numpy(), a.cpu().numpy()) <= 1e-6
        else:
            assert torch.equal(a, b)


def test_full_size_properties_batch_of_eight():
    """BASELINE configs[2] at full size (8 LiDAR frames, 2.7 M rows): size-independent properties instead of the oracle.
    Sortedness of the merged keys, count / inverse consistency against a numpy re-quantisation, permutation invariance of
    the pillar set and of the max-pooled features (bit-exact in eval mode)."""
    from radardistill_b200 import synth
    pts = synth.lidar_batch(8)
    m = _shipped_module("lidar", train=False)
    lo, vs, nx, ny = np.float32(synth.PC_RANGE[:2]), np.float32(synth.VOXEL_SIZE[:2]), m.spec.nx, m.spec.ny

    def run(p):
        with torch.no_grad():
            out = m({"points": torch.from_numpy(p).cuda(), "batch_size": 8})
        r = m.last_result
        return (out["pillar_features"].cpu().numpy(), out["pillar_coords"].cpu().numpy().astype(np.int64), r.inverse.cpu().numpy(),
                r.counts.cpu().numpy(), r.n_kept)

    f, c, inv, cnt, n = run(pts)
    # sortedness: rows are in ascending merged-key order, no duplicates (torch.unique, :212)
    key = (c[:, 0] * nx + c[:, 2]) * ny + c[:, 1]
    assert np.all(np.diff(key) > 0)
    # counts / inverse against an independent numpy quantisation (:201-210)
    q = np.floor((pts[:, 1:3] - lo) / vs).astype(np.int64)
    keep = (q[:, 0] >= 0) & (q[:, 0] < nx) & (q[:, 1] >= 0) & (q[:, 1] < ny)
    assert n == int(keep.sum()) and cnt.sum() == n and cnt.min() >= 1
    pk = (pts[keep, 0].astype(np.int64) * nx + q[keep, 0]) * ny + q[keep, 1]
    np.testing.assert_array_equal(key[inv], pk)
    np.testing.assert_array_equal(np.bincount(inv, minlength=len(cnt)), cnt)
    # permutation invariance: another row order gives the same pillars and bit-identical features
    perm = np.random.default_rng(3).permutation(len(pts))
    f2, c2, inv2, cnt2, n2 = run(np.ascontiguousarray(pts[perm]))
    np.testing.assert_array_equal(c2, c)
    np.testing.assert_array_equal(cnt2, cnt)
    np.testing.assert_array_equal(f2, f)
    kept_perm = keep[perm]   # the inverse follows the rows: point j of the permuted input still lands in its own pillar
    np.testing.assert_array_equal(key[inv2], ((pts[perm][kept_perm, 0].astype(np.int64) * nx + q[perm][kept_perm, 0]) * ny + q[perm][kept_perm, 1]))
    # features are a max over each pillar's rows of a ReLU: non-negative and finite
    assert np.isfinite(f).all() and f.min() >= 0.0


def _train_step_against_oracle(kind, pts, batch, grad_seed):
    """Train-mode forward + backward of a shipped encoder on `pts`, every output compared with the C oracle."""
    import os
    c, key = (5, "points") if kind == "lidar" else (6, "radar_points")
    m = _shipped_module(kind, train=True)
    o = _oracle_of(m, c)
    orc.set_threads(os.cpu_count() or 8)
    r = o.forward(pts, training=True)
    out = m({key: torch.from_numpy(pts).cuda(), "batch_size": batch})
    res = m.last_result
    fk = [k for k in out if k.endswith("pillar_features")][0]
    ck = [k for k in out if k.endswith("_coords")][0]
    assert res.n_kept == r["n"] and res.n_pillars == r["p"]
    np.testing.assert_array_equal(out[ck].cpu().numpy(), r["coords"])
    np.testing.assert_array_equal(res.inverse.cpu().numpy(), r["inverse"])
    np.testing.assert_array_equal(res.counts.cpu().numpy(), r["counts"])
    np.testing.assert_array_equal(res.argmax.cpu().numpy(), r["argmax"])
    f = out[fk]
    assert H.norm_rel_err(f.detach().cpu().numpy(), r["features"]) <= 1e-6
    gout = torch.randn(f.shape, generator=torch.Generator().manual_seed(grad_seed)).cuda()
    f.backward(gout)
    b = o.backward(r, gout.cpu().numpy())
    pfn = m.pfn_layers[0]
    assert H.norm_rel_err(pfn.linear.weight.grad.cpu().numpy(), b["d_weight"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(pfn.norm.weight.grad.cpu().numpy(), b["d_gamma"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(pfn.norm.bias.grad.cpu().numpy(), b["d_beta"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(pfn.norm.running_mean.cpu().numpy(), r["new_running_mean"]) <= 1e-6
    assert H.norm_rel_err(pfn.norm.running_var.cpu().numpy(), r["new_running_var"]) <= 1e-6
    return r


def test_full_size_train_step_vs_oracle_batch_of_eight():
    """BASELINE configs[2] exactly as bench.py's mode B times it: 8 LiDAR frames (2.7 M rows) and 8 radar frames, both
    encoders train-mode BN, forward + backward, against the C oracle: coords / inverse / counts / argmax bit-exact,
    features <= 1e-6, dW / dgamma / dbeta <= 1e-5 (dynamic_pillar_vfe.py:195-252)."""
    from radardistill_b200 import synth
    r = _train_step_against_oracle("lidar", synth.lidar_batch(8), 8, 21)
    assert r["n"] > 2_500_000 and r["p"] > 1_000_000
    _train_step_against_oracle("radar", synth.radar_batch(8), 8, 22)


def test_stress_spec_size_one_million_points_per_frame():
    """BASELINE configs[4] at the specified size: 1 M points per frame (2 frames = one GPU's share of batch 16 over 8
    GPUs), 0.05 m pillars: eval forward bit-exact and a train step against the oracle."""
    import os
    from oracle.ref_loader import Cfg
    from radardistill_b200 import synth, vfe
    grid = synth.grid_size_of(synth.PC_RANGE, synth.STRESS_VOXEL_SIZE)
    pts = synth.stress_batch(2, n_points=1_000_000)
    cfg = Cfg(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])
    torch.manual_seed(12)
    m = vfe.DynamicPillarVFESimple2D(model_cfg=cfg, num_point_features=5, voxel_size=synth.STRESS_VOXEL_SIZE, grid_size=grid,
                                     point_cloud_range=synth.PC_RANGE).cuda()
    n = m.pfn_layers[0].norm
    with torch.no_grad():
        n.weight.uniform_(0.5, 1.5); n.bias.normal_(0, 0.2); n.running_mean.normal_(0, 1); n.running_var.uniform_(0.5, 4)
    pfn = m.pfn_layers[0]
    ocfg = orc.OracleConfig(num_point_features=5, voxel_size=tuple(synth.STRESS_VOXEL_SIZE), grid_size=tuple(grid),
                            point_cloud_range=tuple(synth.PC_RANGE))
    cp = lambda t: t.detach().cpu().numpy().copy()
    o = orc.PillarOracle(ocfg, cp(pfn.linear.weight), cp(n.weight), cp(n.bias), cp(n.running_mean), cp(n.running_var))
    orc.set_threads(os.cpu_count() or 8)
    dev = torch.from_numpy(pts).cuda()
    m.eval()
    r = o.forward(pts, training=False, keep_intermediates=False)
    assert r["n0"] == 2_000_000 and r["counts"].max() > 1000
    with torch.no_grad():
        out = m({"points": dev, "batch_size": 2})
    np.testing.assert_array_equal(out["pillar_coords"].cpu().numpy(), r["coords"])
    np.testing.assert_array_equal(m.last_result.counts.cpu().numpy(), r["counts"])
    np.testing.assert_array_equal(m.last_result.inverse.cpu().numpy(), r["inverse"])
    np.testing.assert_array_equal(out["pillar_features"].cpu().numpy(), r["features"])
    m.train()
    r = o.forward(pts, training=True)
    out = m({"points": dev, "batch_size": 2})
    np.testing.assert_array_equal(m.last_result.argmax.cpu().numpy(), r["argmax"])
    f = out["pillar_features"]
    assert H.norm_rel_err(f.detach().cpu().numpy(), r["features"]) <= 1e-6
    gout = torch.randn(f.shape, generator=torch.Generator().manual_seed(6)).cuda()
    f.backward(gout)
    b = o.backward(r, gout.cpu().numpy())
    assert H.norm_rel_err(pfn.linear.weight.grad.cpu().numpy(), b["d_weight"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(n.weight.grad.cpu().numpy(), b["d_gamma"]) <= H.RTOL_GRADS
    assert H.norm_rel_err(n.bias.grad.cpu().numpy(), b["d_beta"]) <= H.RTOL_GRADS


def test_backward_accepts_a_misaligned_upstream_gradient():
    """a contiguous upstream gradient whose storage offset is not a multiple of 16 bytes (a view into a flat
    buffer) must give the same parameter gradients as an aligned copy (the kernels stage gradient rows with 16-byte bulk copies)."""
    from radardistill_b200 import synth
    pts = torch.from_numpy(synth.radar_batch(2)).cuda()
    grads = []
    for misaligned in (False, True):
        m = _shipped_module("radar", train=True)
        f = m({"radar_points": pts, "batch_size": 2})["radar_pillar_features"]
        g = torch.randn(f.numel() + 1, generator=torch.Generator().manual_seed(4)).cuda()
        gv = g[1:].view_as(f) if misaligned else g[1:].clone().view_as(f)
        assert (gv.data_ptr() % 16 != 0) == misaligned
        f.backward(gv)
        grads.append(m.pfn_layers[0].linear.weight.grad.clone())
    assert H.norm_rel_err(grads[1].cpu().numpy(), grads[0].cpu().numpy()) <= 1e-6
