"""CPU: the C oracle (oracle/pillar_oracle.c) against the reference's own outputs (tests/golden).

Integer outputs (coords, inverse, counts) must be bit-exact.  Features / gradients use the
norm-relative tolerance of tests/helpers.py.  The argmax is compared exactly except at
near-ties: the reference's sgemm and the oracle's fmaf chain round differently, so two
rows whose activations differ by less than the feature tolerance may swap.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H

FILES = H.golden_files()


def test_goldens_present():
    assert len(FILES) >= 30


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
@pytest.mark.parametrize("mean_mode", [orc.MEAN_SEQ_F32, orc.MEAN_F64, orc.FOLDED], ids=["seqf32", "f64", "folded"])
def test_oracle_matches_reference(path, mean_mode):
    g = H.load_golden(path)
    o = H.oracle_from_golden(g, mean_mode)
    r = o.forward(g["points"], training=g["training"])
    # --- integer outputs: bit-exact
    assert r["p"] == g["features"].shape[0]
    np.testing.assert_array_equal(r["coords"], g["coords"])
    assert r["coords"].dtype == np.int32 and g["coords"].dtype == np.int32
    np.testing.assert_array_equal(r["inverse"], g["inverse"])
    np.testing.assert_array_equal(r["counts"], g["counts"])
    if r["p"] == 0:
        assert g["features"].size == 0
        return
    # --- features
    err = H.norm_rel_err(r["features"], g["features"])
    assert err <= H.RTOL_FEATURES, f"features norm-rel err {err:.3e}"
    # --- argmax: exact up to near-ties
    scale = max(float(np.abs(g["features"]).max()), 1e-30)
    post = np.maximum(r["x"] * r["scale"] + r["shift"], 0.0)
    assert H.argmax_mismatch_is_near_tie(r["argmax"], g["argmax"], post, r["inverse"], H.RTOL_FEATURES * scale)
    frac = float(np.mean(r["argmax"] != g["argmax"]))
    assert frac < 2e-3, f"too many argmax differences: {frac}"
    # --- running statistics
    if "new.running_mean" in g:
        assert H.norm_rel_err(r["new_running_mean"], g["new.running_mean"]) <= H.RTOL_FEATURES
        assert H.norm_rel_err(r["new_running_var"], g["new.running_var"]) <= H.RTOL_FEATURES
    # --- parameter gradients
    if "grad_features" in g:
        b = o.backward(r, g["grad_features"])
        e_w = H.norm_rel_err(b["d_weight"], g["grad.linear.weight"])
        assert e_w <= H.RTOL_GRADS_REF, f"dW err {e_w:.3e}"
        if o.cfg.use_norm:
            assert H.norm_rel_err(b["d_gamma"], g["grad.norm.weight"]) <= H.RTOL_GRADS_REF
            assert H.norm_rel_err(b["d_beta"], g["grad.norm.bias"]) <= H.RTOL_GRADS_REF
        else:
            assert H.norm_rel_err(b["d_beta"], g["grad.linear.bias"]) <= H.RTOL_GRADS_REF


def test_mean_modes_agree_within_one_ulp():
    g = H.load_golden([f for f in FILES if "dense_cell__eval" in f][0])
    a = H.oracle_from_golden(g, orc.MEAN_SEQ_F32).forward(g["points"])
    b = H.oracle_from_golden(g, orc.MEAN_F64).forward(g["points"])
    c = H.oracle_from_golden(g, orc.FOLDED).forward(g["points"])
    d = np.abs(a["pillar_mean"] - b["pillar_mean"])
    assert d.max() <= 64 * np.spacing(np.float32(54.0))  # 700-point pillar: sequential fp32 drift
    assert H.norm_rel_err(a["features"], b["features"]) <= H.RTOL_FEATURES
    np.testing.assert_array_equal(b["pillar_mean"], c["pillar_mean"])      # the folded form keeps the fp64 mean
    assert H.norm_rel_err(c["features"], b["features"]) <= H.RTOL_FEATURES


def test_single_point_train_raises():
    g = H.load_golden([f for f in FILES if "single__eval" in f][0])
    o = H.oracle_from_golden(g)
    with pytest.raises(ValueError):
        o.forward(g["points"], training=True)
