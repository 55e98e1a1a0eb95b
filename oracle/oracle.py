"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the C oracle (oracle/pillar_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product (``radardistill_b200``)
never does.  It restates the reference encoder
(``pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py``) on the CPU, and is pinned to
the reference's own outputs through ``tests/golden`` (see ``oracle/gen_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpillar_oracle.so")

MEAN_SEQ_F32 = 0   # torch_scatter CPU semantics (pinning against the reference)
MEAN_F64 = 1       # round-1 canonical form: fp64 mean, k-ascending fmaf chain over the reference's feature order
FOLDED = 2         # round-2 canonical form, what the CUDA kernels implement: fp64 mean + affine-folded linear layer
MAX_G = 14
LAYOUT_SIMPLE2D = 0
LAYOUT_DYNPILLAR = 1


class _Cfg(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("vsz", C.c_float * 3), ("off", C.c_float * 3),
                ("nx", C.c_int32), ("ny", C.c_int32), ("cols", C.c_int32),
                ("layout", C.c_int32), ("use_abs", C.c_int32), ("use_cluster", C.c_int32),
                ("use_relative", C.c_int32), ("with_distance", C.c_int32),
                ("c_in", C.c_int32), ("c_out", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pillar_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libpillar_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_abi_version.restype = C.c_int
        assert _lib.orc_abi_version() == 2
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def _p(a, ty=C.c_void_p):
    return a.ctypes.data_as(ty)


@dataclass
class OracleConfig:
    """Mirror of the reference constructor arguments (dynamic_pillar_vfe.py:50-85,147-190)."""
    num_point_features: int                      # C (raw features, without batch column)
    voxel_size: tuple
    grid_size: tuple
    point_cloud_range: tuple
    layout: int = LAYOUT_SIMPLE2D
    use_norm: bool = True
    with_distance: bool = False
    use_absolute_xyz: bool = True
    use_cluster_xyz: bool = True
    use_relative_xyz: bool = True
    c_out: int = 32
    eps: float = 1e-3
    momentum: float = 0.01
    c_in: int = field(init=False)
    coord_cols: int = field(init=False)

    def __post_init__(self):
        c = self.num_point_features
        if self.layout == LAYOUT_SIMPLE2D:
            cin = 3 + (c if self.use_absolute_xyz else c - 3)
            cin += 3 if self.use_cluster_xyz else 0
            cin += 3 if self.use_relative_xyz else 0
            self.coord_cols = 3
        else:
            cin = (c if self.use_absolute_xyz else c - 3) + 6
            self.use_cluster_xyz, self.use_relative_xyz = True, False
            self.coord_cols = 4
        cin += 1 if self.with_distance else 0
        self.c_in = cin

    def c_struct(self) -> _Cfg:
        s = _Cfg()
        pcr = np.asarray(self.point_cloud_range, dtype=np.float32)
        vs = np.asarray(self.voxel_size, dtype=np.float64)
        for k in range(3):
            s.lo[k] = float(pcr[k])
            s.vsz[k] = float(np.float32(vs[k]))
            # voxel/2 + range_lo evaluated in double, rounded once to fp32 (:180-182)
            s.off[k] = float(np.float32(vs[k] / 2.0 + float(pcr[k])))
        s.nx, s.ny = int(self.grid_size[0]), int(self.grid_size[1])
        s.cols = self.num_point_features + 1
        s.layout = self.layout
        s.use_abs, s.use_cluster = int(self.use_absolute_xyz), int(self.use_cluster_xyz)
        s.use_relative, s.with_distance = int(self.use_relative_xyz), int(self.with_distance)
        s.c_in, s.c_out = self.c_in, self.c_out
        return s


class PillarOracle:
    """CPU oracle with the module's state: W (Cout,Cin), gamma, beta, running stats."""

    def __init__(self, cfg: OracleConfig, weight, gamma=None, beta=None, running_mean=None,
                 running_var=None, bias=None, mean_mode: int = FOLDED):
        self.cfg, self.mean_mode = cfg, mean_mode
        f32 = lambda a, d: np.ascontiguousarray(d if a is None else a, dtype=np.float32)
        co = cfg.c_out
        self.weight = f32(weight, None).reshape(co, cfg.c_in)
        self.gamma = f32(gamma, np.ones(co))
        self.beta = f32(beta, np.zeros(co))
        self.running_mean = f32(running_mean, np.zeros(co))
        self.running_var = f32(running_var, np.ones(co))
        self.bias = f32(bias, np.zeros(co))
        self._cs = cfg.c_struct()

    # -- a1..a4, a10
    def index(self, points: np.ndarray) -> dict:
        L, cfg = lib(), self.cfg
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, cfg.num_point_features + 1)
        n0 = pts.shape[0]
        m = max(n0, 1)
        keep = np.empty(m, np.int32); pcoord = np.empty((m, 2), np.int32); inv = np.empty(m, np.int32)
        unq = np.empty(m, np.int32); cnt = np.empty(m, np.int32); coords = np.empty((m, cfg.coord_cols), np.int32)
        n, p = C.c_int64(0), C.c_int64(0)
        rc = L.orc_index(_p(pts), C.c_int64(n0), C.byref(self._cs), C.c_int(cfg.coord_cols), _p(keep), _p(pcoord),
                         _p(inv), _p(unq), _p(cnt), _p(coords), C.byref(n), C.byref(p))
        if rc != 0:
            raise ValueError("oracle: negative merged key (batch index < 0)")
        n, p = n.value, p.value
        return dict(points=pts, n0=n0, n=n, p=p, keep=keep[:n], pcoord=pcoord[:n], inverse=inv[:n],
                    unq=unq[:p], counts=cnt[:p], coords=coords[:p])

    # -- a5..a9
    def forward(self, points: np.ndarray, training: bool = False, keep_intermediates: bool = True) -> dict:
        L, cfg = lib(), self.cfg
        r = self.index(points)
        n, p, cin, co = r["n"], r["p"], cfg.c_in, cfg.c_out
        if training and cfg.use_norm and n == 1:
            raise ValueError("Expected more than 1 value per channel when training")  # torch BN behaviour
        mean = np.zeros((max(p, 1), 3), np.float32)
        L.orc_pillar_mean(_p(r["points"]), C.byref(self._cs), _p(r["keep"]), _p(r["inverse"]), _p(r["counts"]),
                          C.c_int64(n), C.c_int64(p), C.c_int(min(self.mean_mode, MEAN_F64)), _p(mean))
        if self.mean_mode == FOLDED:
            return self._forward_folded(r, mean, training, keep_intermediates)
        f = np.empty((max(n, 1), cin), np.float32)
        L.orc_features(_p(r["points"]), C.byref(self._cs), _p(r["keep"]), _p(r["pcoord"]), _p(r["inverse"]), _p(mean),
                       C.c_int64(n), _p(f))
        x = np.empty((max(n, 1), co), np.float32)
        L.orc_linear(_p(f), C.c_int64(n), C.c_int(cin), _p(self.weight), C.c_int(co), _p(x))
        bmean, bvar = np.zeros(co, np.float64), np.ones(co, np.float64)
        scale, shift = np.ones(co, np.float32), self.bias.copy()
        if cfg.use_norm:
            if training and self.mean_mode == MEAN_F64:   # canonical: from the fp64 feature moments
                L.orc_bn_batch_stats_moments(_p(f), C.c_int64(n), C.c_int(cin), _p(self.weight), C.c_int(co),
                                             _p(bmean), _p(bvar))
            elif training:                                # reference semantics: moments of the fp32 x values
                L.orc_bn_batch_stats(_p(x), C.c_int64(n), C.c_int(co), _p(bmean), _p(bvar))
            else:
                bmean, bvar = self.running_mean.astype(np.float64), self.running_var.astype(np.float64)
            L.orc_bn_fold(_p(self.gamma), _p(self.beta), _p(bmean), _p(bvar), C.c_double(cfg.eps), C.c_int(co),
                          _p(scale), _p(shift))
        out = np.empty((max(p, 1), co), np.float32); arg = np.empty((max(p, 1), co), np.int32)
        L.orc_act_max(_p(x), C.c_int64(n), C.c_int(co), _p(scale), _p(shift), _p(r["inverse"]), C.c_int64(p),
                      _p(out), _p(arg))
        r.update(features=out[:p], argmax=arg[:p], pillar_mean=mean[:p], scale=scale, shift=shift,
                 batch_mean=bmean, batch_var=bvar, training=training)
        if keep_intermediates:
            r.update(f=f[:n], x=x[:n])
        if training and cfg.use_norm and n > 0:
            m = cfg.momentum
            unbiased = bvar * (n / (n - 1.0)) if n > 1 else bvar
            r["new_running_mean"] = ((1 - m) * self.running_mean + m * bmean).astype(np.float32)
            r["new_running_var"] = ((1 - m) * self.running_var + m * unbiased).astype(np.float32)
        return r

    def _forward_folded(self, r, mean, training, keep_intermediates):
        """ORC_FOLDED: the canonical arithmetic of the CUDA kernels (see the section header in pillar_oracle.c)."""
        L, cfg = lib(), self.cfg
        n, p, cin, co = r["n"], r["p"], cfg.c_in, cfg.c_out
        T = np.zeros((max(cin, 1), MAX_G + 1), np.float32)
        kin, g = C.c_int32(0), C.c_int32(0)
        if L.orc_build_T(C.byref(self._cs), _p(T), C.byref(kin), C.byref(g)) != cin:
            raise ValueError("oracle: layout does not fit the folded basis")
        kin, g = kin.value, g.value
        wg, cst = np.empty((co, g), np.float32), np.empty(co, np.float32)
        L.orc_fold_weights(_p(self.weight), None if cfg.use_norm else _p(self.bias), _p(T), C.c_int(cin), C.c_int(co), C.c_int(g),
                           _p(wg), _p(cst))
        q = np.zeros((max(p, 1), 5), np.float32)
        L.orc_pillar_consts(_p(r["points"]), C.byref(self._cs), _p(r["keep"]), _p(r["inverse"]), _p(r["unq"]), _p(r["counts"]),
                            C.c_int64(n), C.c_int64(p), _p(q))
        rin = np.empty((max(n, 1), kin), np.float32)
        v, x = np.empty((max(n, 1), co), np.float32), np.empty((max(n, 1), co), np.float32)
        L.orc_forward_folded(_p(r["points"]), C.byref(self._cs), _p(r["keep"]), _p(r["inverse"]), _p(q), _p(wg), _p(cst),
                             C.c_int(kin), C.c_int(g), C.c_int64(n), _p(rin), _p(v), _p(x))
        bmean, bvar = np.zeros(co, np.float64), np.ones(co, np.float64)
        scale, shift = np.ones(co, np.float32), np.zeros(co, np.float32)
        if cfg.use_norm:
            if training:
                S1, S2 = np.zeros(g, np.float64), np.zeros((g, g), np.float64)
                L.orc_moments_folded(_p(rin), _p(q), _p(r["inverse"]), C.c_int64(n), C.c_int(kin), C.c_int(g), _p(S1), _p(S2))
                L.orc_bn_stats_from_moments(_p(S1), _p(S2), C.c_int64(n), C.c_int(g), _p(wg), _p(cst), C.c_int(co), _p(bmean), _p(bvar))
            else:
                bmean, bvar = self.running_mean.astype(np.float64), self.running_var.astype(np.float64)
            L.orc_bn_fold(_p(self.gamma), _p(self.beta), _p(bmean), _p(bvar), C.c_double(cfg.eps), C.c_int(co), _p(scale), _p(shift))
        out = np.empty((max(p, 1), co), np.float32); arg = np.empty((max(p, 1), co), np.int32)
        L.orc_act_max_folded(_p(x), _p(v), C.c_int64(n), C.c_int(co), _p(scale), _p(shift), _p(r["inverse"]), C.c_int64(p),
                             _p(out), _p(arg))
        r.update(features=out[:p], argmax=arg[:p], pillar_mean=mean[:p], scale=scale, shift=shift, batch_mean=bmean,
                 batch_var=bvar, training=training)
        if keep_intermediates:
            # the backward oracle works on the reference's own decorated features (orc_features) and the folded x
            f = np.empty((max(n, 1), cin), np.float32)
            L.orc_features(_p(r["points"]), C.byref(self._cs), _p(r["keep"]), _p(r["pcoord"]), _p(r["inverse"]), _p(mean),
                           C.c_int64(n), _p(f))
            r.update(f=f[:n], x=x[:n])
        if training and cfg.use_norm and n > 0:
            m = cfg.momentum
            unbiased = bvar * (n / (n - 1.0)) if n > 1 else bvar
            r["new_running_mean"] = ((1 - m) * self.running_mean + m * bmean).astype(np.float32)
            r["new_running_var"] = ((1 - m) * self.running_var + m * unbiased).astype(np.float32)
        return r

    # -- a12
    def backward(self, fwd: dict, grad_features: np.ndarray) -> dict:
        L, cfg = lib(), self.cfg
        n, p, cin, co = fwd["n"], fwd["p"], cfg.c_in, cfg.c_out
        g = np.ascontiguousarray(grad_features, np.float32).reshape(max(p, 0), co)
        dW = np.zeros((co, cin), np.float32); dg = np.zeros(co, np.float32); db = np.zeros(co, np.float32)
        if n > 0:
            L.orc_backward(_p(g), _p(fwd["f"]), _p(fwd["x"]), _p(fwd["features"]), _p(fwd["argmax"]), C.c_int64(n),
                           C.c_int64(p), C.c_int(cin), C.c_int(co), _p(self.gamma), _p(fwd["batch_mean"]),
                           _p(fwd["batch_var"]), C.c_double(cfg.eps), C.c_int(int(fwd["training"])),
                           C.c_int(int(cfg.use_norm)), _p(dW), _p(dg), _p(db))
        return dict(d_weight=dW, d_gamma=dg, d_beta=db)


def config_for(name: str, num_point_features: int, voxel_size, grid_size, point_cloud_range, model_cfg: dict) -> OracleConfig:
    """OracleConfig for a reference class name + its ``model_cfg`` (same keys as the YAML)."""
    layout = LAYOUT_DYNPILLAR if name in ("DynPillarVFE", "DynamicPillarVFE") else LAYOUT_SIMPLE2D
    nf = list(model_cfg["NUM_FILTERS"])
    assert len(nf) == 1, "oracle covers the single-PFN-layer configs the reference ships"
    return OracleConfig(num_point_features=num_point_features, voxel_size=tuple(voxel_size),
                        grid_size=tuple(int(v) for v in grid_size), point_cloud_range=tuple(point_cloud_range),
                        layout=layout, use_norm=bool(model_cfg["USE_NORM"]),
                        with_distance=bool(model_cfg["WITH_DISTANCE"]),
                        use_absolute_xyz=bool(model_cfg["USE_ABSLOTE_XYZ"]),
                        use_cluster_xyz=bool(model_cfg.get("USE_CLUSTER_XYZ", True)),
                        use_relative_xyz=bool(model_cfg.get("USE_RELATIVE_XYZ", True)), c_out=nf[-1])
