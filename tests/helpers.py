"""Shared helpers for the parity tests: golden loading and the stated tolerances."""
from __future__ import annotations

import ast
import glob
import os

import numpy as np

from oracle import oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# ---- tolerances: norm-relative, max|a-b| <= RTOL * max|b|  (+ tiny absolute floor)
RTOL_FEATURES = 1e-5     # forward features vs the reference (north_star: rel 1e-5)
RTOL_GRADS = 1e-5        # parameter gradients vs the reference
RTOL_GRADS_REF = 1e-5    # oracle (fp64 reductions) vs the reference's fp32 autograd; measured max 8.0e-6


def golden_files():
    """The single-layer pillar goldens (oracle/gen_golden.py); the layer-stack / voxel / mean ones are `stack__*.npz`."""
    return sorted(f for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(f).startswith("stack__"))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["model_cfg"] = ast.literal_eval(str(d["model_cfg"]))
    d["class_name"] = str(d["class_name"])
    d["training"] = bool(d["training"])
    d["name"] = os.path.basename(path)[:-4]
    return d


def grid_of(g):
    from radardistill_b200 import synth
    return synth.grid_size_of(synth.PC_RANGE, g["voxel_size"])


def oracle_from_golden(g, mean_mode=orc.FOLDED):
    from radardistill_b200 import synth
    cfg = orc.config_for(g["class_name"], int(g["num_point_features"]), g["voxel_size"], grid_of(g),
                         synth.PC_RANGE, g["model_cfg"])
    pre = "param.pfn_layers.0."
    kw = dict(weight=g[pre + "linear.weight"])
    if cfg.use_norm:
        kw.update(gamma=g[pre + "norm.weight"], beta=g[pre + "norm.bias"],
                  running_mean=g[pre + "norm.running_mean"], running_var=g[pre + "norm.running_var"])
    else:
        kw.update(bias=g[pre + "linear.bias"])
    return orc.PillarOracle(cfg, mean_mode=mean_mode, **kw)


def norm_rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def argmax_mismatch_is_near_tie(arg_a, arg_b, x_post, inverse, tol):
    """Every (pillar, channel) where the two argmax tensors differ must be a near-tie:
    the post-activation values of the two candidate rows differ by <= tol."""
    bad = np.argwhere(arg_a != arg_b)
    for p, c in bad:
        ia, ib = int(arg_a[p, c]), int(arg_b[p, c])
        if inverse[ia] != p or inverse[ib] != p:
            return False
        if abs(float(x_post[ia, c]) - float(x_post[ib, c])) > tol:
            return False
    return True


# ------------------------------------------------------------------ CUDA-side helpers (gpu tests, smoke)
def module_from_golden(g, device="cuda"):
    """The drop-in module (radardistill_b200.vfe) configured like the golden's reference module."""
    import torch
    from oracle.ref_loader import Cfg
    from radardistill_b200 import synth
    from radardistill_b200.vfe import REGISTRY
    name = "DynPillarVFE" if g["class_name"] in ("DynPillarVFE", "DynamicPillarVFE") else g["class_name"]
    m = REGISTRY[name](model_cfg=Cfg(g["model_cfg"]), num_point_features=int(g["num_point_features"]),
                       voxel_size=[float(v) for v in g["voxel_size"]], grid_size=grid_of(g),
                       point_cloud_range=synth.PC_RANGE, depth_downsample_factor=None)
    sd = {k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.to(device)


def points_key_of(g):
    return "radar_points" if g["class_name"] == "Radar_DynamicPillarVFESimple2D" else "points"
