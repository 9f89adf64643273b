"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU: the oracle and the host build of the product headers reproduce them.  GPU: the CUDA path, through
the C ABI, reproduces them.  Bars: fp32 stage and H rows bit-exact; reductions 1e-12; state 1e-9.
"""
import glob
import os

import numpy as np
import pytest

import shim_binding as S
from conftest import oracle_params

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def _load(path, lv, O):
    z = np.load(path)
    prm = lv.params_from_yaml(os.path.join(lv.CONFIG_DIR, str(z["yaml"])), max_map_points=1 << 17, max_points=1 << 14)
    return z, prm, oracle_params(O, prm)


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_and_host_build_reproduce_golden(path, lv, O):
    z, prm, oprm = _load(path, lv, O)
    om = O.Map(O.KNN_KDTREE)
    om.build(z["map"])
    m = om.match_all(z["x_prop"], oprm, z["sweep"])
    for k in ("valid", "nn_sqd", "plane", "dist", "g"):
        assert (m[k] == z[k]).all(), k
    st, x, P, logs = om.update_iterated(z["x_prop"], z["P0"], oprm, z["sweep"])
    assert [l["n_matches"] for l in logs] == z["n_matches"].tolist()
    assert np.abs(x - z["x_final"]).max() < 1e-12 and np.abs(P - z["P_final"]).max() < 1e-12
    sm = S.ShimMap(z["map"], prm.voxel_size, prm.MAX_DIST_PLANE)
    sp = S.make_params(oprm, prm.voxel_size)
    g = sm.match_all(z["x_prop"], sp, z["sweep"])
    for k in ("valid", "plane", "dist", "g"):
        assert (g[k] == z[k]).all(), k
    rows = g["rows"][g["valid"] == 1]
    assert (rows[:, :12] == z["h_x"]).all() and (rows[:, 12] == z["h"]).all()
    ss, xs, Ps, ls = sm.update(z["x_prop"], z["P0"], sp, z["sweep"])
    assert [l["n_matches"] for l in ls] == z["n_matches"].tolist()
    assert np.abs(np.stack([l["dx"] for l in ls]) - z["dx"]).max() < 1e-9
    assert np.abs(xs - z["x_final"]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_reproduces_golden(path, lv, O):
    z, prm, oprm = _load(path, lv, O)
    loc = lv.Localizer(prm)
    loc.map_build(z["map"])
    g = loc.match_all(z["x_prop"], z["sweep"])
    for k in ("valid", "plane", "dist", "g"):
        assert (g[k] == z[k]).all(), k
    inside = np.isfinite(g["nn_sqd"][:, 4])
    assert (g["nn_sqd"][inside] == z["nn_sqd"][inside]).all()
    st, hx, h = loc.measure(z["x_prop"], z["sweep"])
    assert (hx == z["h_x"]).all() and (h == z["h"]).all()
    loc.set_state(z["x_prop"], z["P0"])
    st, x, P, logs = loc.correct(z["sweep"])
    assert st == 0 and [l["n_matches"] for l in logs] == z["n_matches"].tolist()
    assert [l["converged"] for l in logs] == z["converged"].tolist()
    assert np.abs(np.stack([l["HTH"] for l in logs]) - z["HTH"]).max() <= 1e-12 * np.abs(z["HTH"]).max()
    assert np.abs(np.stack([l["dx"] for l in logs]) - z["dx"]).max() < 1e-9
    assert np.abs(x - z["x_final"]).max() < 1e-9
    assert np.abs(P - z["P_final"]).max() < 1e-8 * np.abs(z["P_final"]).max()
    loc.close()
