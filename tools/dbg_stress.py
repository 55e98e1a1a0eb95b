"""Debug: stress train step, compare every backward quantity with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as orc
from oracle.ref_loader import Cfg
from radardistill_b200 import synth, vfe
from tests import helpers as H

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
grid = synth.grid_size_of(synth.PC_RANGE, synth.STRESS_VOXEL_SIZE)
pts = synth.stress_batch(2, n_points=npts)
cfg = Cfg(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])
torch.manual_seed(12)
m = vfe.DynamicPillarVFESimple2D(model_cfg=cfg, num_point_features=5, voxel_size=synth.STRESS_VOXEL_SIZE, grid_size=grid,
                                 point_cloud_range=synth.PC_RANGE).cuda()
n = m.pfn_layers[0].norm
with torch.no_grad():
    n.weight.uniform_(0.5, 1.5); n.bias.normal_(0, 0.2); n.running_mean.normal_(0, 1); n.running_var.uniform_(0.5, 4)
pfn = m.pfn_layers[0]
ocfg = orc.OracleConfig(num_point_features=5, voxel_size=tuple(synth.STRESS_VOXEL_SIZE), grid_size=tuple(grid), point_cloud_range=tuple(synth.PC_RANGE))
cp = lambda t: t.detach().cpu().numpy().copy()
o = orc.PillarOracle(ocfg, cp(pfn.linear.weight), cp(n.weight), cp(n.bias), cp(n.running_mean), cp(n.running_var))
orc.set_threads(os.cpu_count() or 8)
dev = torch.from_numpy(pts).cuda()
m.train()
r = o.forward(pts, training=True)
out = m({"points": dev, "batch_size": 2})
res = m.last_result
print("N", r["n"], "P", r["p"], "max count", r["counts"].max(), "argmax equal", np.array_equal(res.argmax.cpu().numpy(), r["argmax"]))
f = out["pillar_features"]
print("feat err", H.norm_rel_err(f.detach().cpu().numpy(), r["features"]))
bn = res.bn_state.cpu().numpy()
print("mean err", np.abs(bn[:32] - r["batch_mean"]).max(), "var relerr", np.abs(bn[32:64] / r["batch_var"] - 1).max())
for mode in ("ones", "randn"):
    gout = torch.ones_like(f) if mode == "ones" else torch.randn(f.shape, generator=torch.Generator().manual_seed(6)).cuda()
    for p in m.parameters(): p.grad = None
    f.backward(gout, retain_graph=True)
    b = o.backward(r, gout.cpu().numpy())
    dW, dg, db = pfn.linear.weight.grad.cpu().numpy(), n.weight.grad.cpu().numpy(), n.bias.grad.cpu().numpy()
    print(mode, "dW", H.norm_rel_err(dW, b["d_weight"]), "dg", H.norm_rel_err(dg, b["d_gamma"]), "db", H.norm_rel_err(db, b["d_beta"]))
    print("  db gpu", db[:4], "oracle", b["d_beta"][:4])
    print("  dg gpu", dg[:4], "oracle", b["d_gamma"][:4])
    print("  dW col err", np.abs(dW - b["d_weight"]).max(0))
    # which pillars: per-pillar contribution check for channel 0 via dbeta restricted to big pillars
    cnt = r["counts"]
    big = cnt > 192
    arg = r["argmax"]; outf = r["features"]
    gy = np.where(outf > 0, gout.cpu().numpy(), 0.0)
    print("  dbeta from big pillars (oracle)", gy[big].sum(0)[:4], " #big", big.sum())
