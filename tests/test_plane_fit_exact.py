"""What "bit-exact vs the oracle" rests on for the 5x3 plane fit.

The oracle's QR (oracle/lv_oracle.cpp) and the device's (csrc/lv_point_math.h) are two transcriptions of Eigen's
colPivHouseholderQr by the same author; agreeing with each other says nothing about agreeing with the mathematics.  Here the
least-squares problem A n = -1 of R3Math::estimate_plane (src/Utils/Utils.cpp:32-57) is solved in float64 (numpy lstsq, ~1e-16)
for every neighbour set of a scene, and the fp32 results of BOTH implementations are held against it:
  * the normalised plane (A, B, C, D) lies within cond(A) x a few fp32 ulp of the exact solution (backward stability);
  * the accept bit of is_plane (Utils.cpp:59-66) is decided by the exact residuals wherever the exact margin exceeds the
    fp32 uncertainty of the residuals; the sets closer to the threshold than that are COUNTED and reported (they are the
    only points where an Eigen build could legitimately differ from either transcription)."""
import numpy as np
import pytest

import shim_binding as S

EPS32 = float(np.finfo(np.float32).eps)


def _neighbour_sets(O, sc):
    om = O.Map(O.KNN_KDTREE)
    om.build(sc.map)
    ref = om.match_all(sc.x_prop, sc.oprm, sc.sweep)
    ok = (ref["nn_idx"][:, 4] >= 0) & (ref["nn_sqd"][:, 4].astype(np.float64) < sc.prm.MAX_DIST_PLANE ** 2)
    return sc.map[ref["nn_idx"][ok]], ref["plane"][ok], ref["valid"][ok].astype(bool)


@pytest.mark.parametrize("name", ["scene_xaloc", "scene_kitti", "scene_ouster"])
def test_plane_fit_against_float64_least_squares(O, name, request):
    sc = request.getfixturevalue(name)
    sets, planes, valid = _neighbour_sets(O, sc)
    thr = float(sc.prm.PLANES_THRESHOLD)
    A = sets.astype(np.float64)                                     # (m, 5, 3)
    m = len(A)
    n_exact = np.empty((m, 3))
    cond = np.empty(m)
    for i in range(m):
        sol, _, rank, sv = np.linalg.lstsq(A[i], -np.ones(5), rcond=None)
        n_exact[i] = sol
        cond[i] = sv[0] / sv[-1]
    norm = np.linalg.norm(n_exact, axis=1)
    exact = np.concatenate([n_exact / norm[:, None], (1.0 / norm)[:, None]], axis=1)      # Utils.cpp:50-54
    res = np.abs(np.einsum("mjk,mk->mj", A, exact[:, :3]) + exact[:, 3:4])              # exact residuals, Utils.cpp:60-63
    margin = thr - res.max(axis=1)
    # the product's fp32 fit for every set (the oracle's output carries the plane only where it was accepted)
    fit = [S.plane_fit(sets[i], thr) for i in range(m)]
    dev = np.array([f[0] for f in fit], dtype=np.float64)
    ok_dev = np.array([f[1] for f in fit])
    assert (ok_dev == valid).all()
    assert (dev[valid].astype(np.float32) == planes[valid]).all()                        # the two transcriptions agree bit for bit
    # accuracy of the coefficients: within cond(A) x a few fp32 ulp of the exact solution (backward stability of QR)
    err = np.abs(dev - exact)
    coef_ulp = (err[:, :3].max(axis=1) / (EPS32 * cond))
    assert coef_ulp.max() <= 64, coef_ulp.max()
    assert (err[:, 3] <= 64 * EPS32 * cond * np.maximum(1.0, np.abs(exact[:, 3]))).all()
    # what decides the accept bit is the residual at the five points, and that is far better conditioned than the coefficients:
    # the computed plane is the exact plane of data perturbed by ~eps |q|, so its residuals move by ~eps |q| (|q| up to 100 m)
    scale = np.abs(sets).max(axis=(1, 2)).astype(np.float64)
    res32 = np.abs(np.einsum("mjk,mk->mj", A, dev[:, :3]) + dev[:, 3:4])
    res_err = np.abs(res32 - res).max(axis=1)
    unc = 32 * EPS32 * (scale + 1.0)
    assert (res_err <= unc).all(), (res_err / (EPS32 * (scale + 1.0))).max()
    decided = np.abs(margin) > 2 * unc
    assert (valid[decided] == (margin[decided] > 0)).all()
    sensitive = int((~decided).sum())
    print("\n%s: %d neighbour sets, %d accepted; coefficients within %.1f x cond x ulp32 of the float64 solution; residual error <= %.1f x "
          "ulp32 x |q|; accept bits closer to the threshold than that: %d (%.4f %%)"
          % (name, m, int(valid.sum()), coef_ulp.max(), (res_err / (EPS32 * (scale + 1.0))).max(), sensitive, 100.0 * sensitive / m))
    assert sensitive <= 2e-3 * m
