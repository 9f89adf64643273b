"""Pins the oracle's restatement of Eigen's arithmetic against Eigen itself — where Eigen exists.

oracle/build_with_eigen.sh builds oracle/_ref/libeigen_sites.so from oracle/ref_eigen/eigen_sites.cpp (the reference's own
statements at Utils.cpp:34-54, esekfom.hpp:1722-1729 and :1736).  This image has no Eigen, so the tests skip here; they are
the recipe a reviewer with Eigen runs."""
import ctypes as C
import os

import numpy as np
import pytest

import shim_binding as S

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libeigen_sites.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libeigen_sites.so not built (needs Eigen3: oracle/build_with_eigen.sh)")


def test_plane_fit_bitwise_against_eigen(O, scene_xaloc):
    E = C.CDLL(LIB)
    sc = scene_xaloc
    om = O.Map(O.KNN_KDTREE)
    om.build(sc.map)
    ref = om.match_all(sc.x_prop, sc.oprm, sc.sweep)
    ok = ref["valid"].astype(bool)
    sets = sc.map[ref["nn_idx"][ok]]
    fp = C.POINTER(C.c_float)
    differ = 0
    for i in range(len(sets)):
        pts = np.ascontiguousarray(sets[i], np.float32)
        out = np.zeros(4, np.float32)
        E.eig_estimate_plane(pts.ctypes.data_as(fp), 5, out.ctypes.data_as(fp))
        differ += int((out != ref["plane"][ok][i]).any())
        assert np.abs(out - ref["plane"][ok][i]).max() <= 8 * np.finfo(np.float32).eps * max(1.0, abs(float(out[3])))
    print("\nplanes differing from Eigen in any bit: %d of %d" % (differ, len(sets)))
    assert differ == 0, "Eigen's packet-order summation differs from the restatement: state the count, widen the bar to ulp"


def test_gain_against_eigen_inverses(scene_xaloc):
    E = C.CDLL(LIB)
    rng = np.random.default_rng(1)
    H = rng.normal(size=(4000, 12))
    HTH, HTh = H.T @ H, H.T @ rng.normal(size=4000)
    P = np.diag(np.r_[np.ones(6), 1e-5 * np.ones(6), np.ones(3), 1e-4 * np.ones(3), 1e-3 * np.ones(3), 1e-5 * np.ones(2)])
    Kh, Kx = np.zeros(23), np.zeros((23, 12))
    dp = C.POINTER(C.c_double)
    E.eig_gain(P.ctypes.data_as(dp), C.c_double(0.001), np.ascontiguousarray(HTH).ctypes.data_as(dp), HTh.ctypes.data_as(dp),
               Kh.ctypes.data_as(dp), Kx.ctypes.data_as(dp))
    S11 = P[:, :12] / 0.001
    Y = np.linalg.solve(np.eye(12) + HTH @ S11[:12], np.c_[HTH, HTh])          # the product's 12x12 formulation (lv_ieskf.h)
    assert np.abs(S11 @ Y[:, 12] - Kh).max() <= 1e-9 * np.abs(Kh).max()
    assert np.abs(S11 @ Y[:, :12] - Kx).max() <= 1e-9 * np.abs(Kx).max()
