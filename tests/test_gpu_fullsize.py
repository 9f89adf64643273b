"""BASELINE.json sizes on the GPU: 65 536-point Velodyne-64 sweep against a 1 M-point map (configs[1]),
plus the size-independent properties the domain offers (determinism, permutation invariance of the
normal equations, streaming predict -> correct -> map update against the oracle)."""
import numpy as np
import pytest

import bench

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full(lv, O):
    prm = lv.params_from_yaml(lv.CONFIG_DIR + "/xaloc.yaml", max_map_points=bench.MAP_POINTS + 4 * 65536,
                              max_points=65536)
    world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=3, prm=prm)
    x0, P0 = lv.init_state_host(prm)
    return dict(prm=prm, world=world, map=mp, sweeps=sweeps, x_props=x_props, truths=truths, P0=P0,
                oprm=bench.oracle_params(O, prm))


def boundary_ties(om, g):
    """The queries (world points g, n x 3) whose 5th and 6th nearest map points are EXACTLY equidistant (fp32 squared
    distance, the search's arithmetic).  WHICH of the two joins the five is then decided by the order a search meets them:
    the reference's max-heap keeps the first one its tree walk visits (ikd_Tree.cpp:1087-1093, strict <), the voxel
    search has other walks, and inside an update a query may keep the five an earlier evaluation found.  About one query
    in 10^6 is such a tie, i.e. one update in ten has one.  Found with the oracle's own 6-NN."""
    out = []
    for i in range(len(g)):
        found, idx, sqd, nn = om.knn(g[i], 6)
        if found == 6 and sqd[4] == sqd[5]:
            out.append(i)
    return np.array(out, dtype=np.int64)


def check_every_evaluation(O, om, oprm, x_prior, P_prior, sweep, logs, loc=None):
    """The tight form of "the update matches the oracle": evaluation by evaluation, AT THE ITERATE THE GPU MEASURED.
    For evaluation k the GPU's iterate is x_prior (k = 0) or the previous log's x_after.  At that very iterate the oracle's
    measurement (kNN = the reference ikd-Tree where available, plane fit, rows) must give the same Nm and the same normal
    equations up to summation order (1e-12): there is no room for a flipped gate, both sides look at the same fp32 world points.
    The oracle's 23-DoF step from those normal equations must then reproduce the GPU's dx_ and next iterate to 1e-9
    (SURVEY 8c asks 1e-6 / 1e-5 m).  Nothing is inferred: a per-point disagreement shows up as Nm or HTH.

    The one disagreement the domain itself leaves open is a query whose 5th and 6th nearest map points are EXACTLY equidistant
    (boundary_ties).  With `loc` given, an evaluation that differs is taken apart:
      * the tied queries T are found by brute force (oracle 6-NN at the same fp32 world points); there must be 1..4 of them;
      * with T removed from the sweep, a fresh GPU measurement and the oracle must again agree to 1e-12 (every other query
        is bit-for-bit the same on both sides);
      * the normal equations the update logged may differ from a fresh GPU measurement of the full sweep only by the rows of
        T: the difference matrix must have rank <= 2 |T| (one row taken out, one put in, per tied query)."""
    x_k = np.array(x_prior, dtype=np.float64)
    ties = 0
    for k, lg in enumerate(logs):
        st, HTH, HTh, nm = om.measure_reduced(x_k, oprm, sweep)
        assert st == 0
        same = nm == lg["n_matches"] and np.abs(HTH - lg["HTH"]).max() <= 1e-12 * np.abs(HTH).max()
        if not same:
            assert loc is not None, (k, nm, lg["n_matches"], np.abs(HTH - lg["HTH"]).max())
            g = loc.match_all(x_k, sweep)["g"]
            T = boundary_ties(om, g)
            assert 0 < len(T) <= 4, (k, T)
            keep = np.ones(len(sweep), bool)
            keep[T] = False
            st_g, HTH_g, HTh_g, nm_g = loc.measure_reduced(x_k, sweep[keep])
            st_o, HTH_o, HTh_o, nm_o = om.measure_reduced(x_k, oprm, sweep[keep])
            assert st_g == st_o == 0 and nm_g == nm_o
            assert np.abs(HTH_g - HTH_o).max() <= 1e-12 * np.abs(HTH_o).max()
            st_f, HTH_f, HTh_f, nm_f = loc.measure_reduced(x_k, sweep)
            assert abs(nm_f - lg["n_matches"]) <= len(T) and abs(nm - lg["n_matches"]) <= len(T)
            sv = np.linalg.svd(lg["HTH"] - HTH_f, compute_uv=False)
            assert sv[min(2 * len(T), 11):].max() <= 1e-11 * np.abs(HTH_f).max(), (k, T, sv)
            HTH, HTh = lg["HTH"], lg["HTh"]        # the step is checked on the normal equations the GPU summed
            ties += len(T)
        assert np.abs(HTh - lg["HTh"]).max() <= 1e-12 * max(1.0, np.abs(HTh).max())
        dx, x_new, P_now, Kx, conv = O.update_step(x_prior, P_prior, x_k, oprm, HTH, HTh)
        assert np.abs(dx - lg["dx"]).max() < 1e-9, (k, np.abs(dx - lg["dx"]).max())
        assert np.abs(x_new - lg["x_after"]).max() < 1e-9, (k, np.abs(x_new - lg["x_after"]).max())
        x_k = lg["x_after"]
    return ties


def test_full_size_update_matches_oracle(lv, O, full):
    loc = lv.Localizer(full["prm"])
    loc.map_build(full["map"])
    om = O.Map(O.KNN_REF_IKDTREE if O.ref_available() else O.KNN_KDTREE)
    om.build(full["map"])
    sweep, x_prop = full["sweeps"][0], full["x_props"][0]
    loc.set_state(x_prop, full["P0"])
    st, x, P, logs = loc.correct(sweep)
    st_o, x_o, P_o, logs_o = om.update_iterated(x_prop, full["P0"], full["oprm"], sweep)
    assert st == st_o == 0 and len(logs) == len(logs_o)
    for k, (a, b) in enumerate(zip(logs, logs_o)):
        assert a["n_matches"] == b["n_matches"] > 40000
        # first evaluation: identical rows, only the summation order differs; later ones inherit the
        # ~1e-12 difference of the iterate (12x12 gain solve vs two 23x23 inverses) times 100 m lever arms
        assert np.abs(a["HTH"] - b["HTH"]).max() <= (1e-12 if k == 0 else 1e-9) * np.abs(b["HTH"]).max()
        # 65 536 points with hard fp32 gates: by the third evaluation the two implementations' iterates differ
        # by ~1e-10 (different but equally valid fp64 evaluation orders); SURVEY 8c states 1e-6 for dx_
        assert np.abs(a["dx"] - b["dx"]).max() < (1e-9 if k == 0 else 1e-8), (k, np.abs(a["dx"] - b["dx"]).max())
    assert np.abs(x - x_o).max() < 1e-8
    check_every_evaluation(O, om, full["oprm"], x_prop, full["P0"], sweep, logs, loc)
    err = np.abs(O.boxminus(x, full["truths"][0]))
    assert err[:3].max() < 5e-3 and err[3:6].max() < 5e-4          # centimetre-level localisation
    # per-point parity at full size
    got = loc.match_all(x_prop, sweep)
    ref = om.match_all(x_prop, full["oprm"], sweep)
    inside = np.isfinite(got["nn_sqd"][:, 4])
    assert (got["nn_sqd"][inside] == ref["nn_sqd"][inside]).all()
    assert (got["valid"] == ref["valid"]).all() and (got["plane"] == ref["plane"]).all()
    loc.close()


def test_determinism_and_permutation_invariance(lv, full):
    loc = lv.Localizer(full["prm"])
    loc.map_build(full["map"])
    sweep, x_prop = full["sweeps"][1], full["x_props"][1]
    a = loc.measure_reduced(x_prop, sweep)
    b = loc.measure_reduced(x_prop, sweep)
    assert (a[1] == b[1]).all() and (a[2] == b[2]).all() and a[3] == b[3]       # bit-identical reruns
    perm = np.random.default_rng(0).permutation(len(sweep))
    c = loc.measure_reduced(x_prop, sweep[perm])
    assert c[3] == a[3]
    assert np.abs(c[1] - a[1]).max() <= 1e-12 * np.abs(a[1]).max()              # sums are order-insensitive
    # splitting the sweep: the normal equations are additive over points
    d1 = loc.measure_reduced(x_prop, sweep[:30000])
    d2 = loc.measure_reduced(x_prop, sweep[30000:])
    assert d1[3] + d2[3] == a[3]
    assert np.abs(d1[1] + d2[1] - a[1]).max() <= 1e-12 * np.abs(a[1]).max()
    loc.close()


def test_neighbour_reuse_and_graph_replay_change_nothing(lv, full, monkeypatch):
    """evaluations after the first reuse the stored neighbours where lv_reuse_kernel can vouch for them, an update is
    replayed as a CUDA graph, its kernels overlap their launches (programmatic dependent launch): none of it may
    change a single bit of the update"""
    def run(env):
        for k in ("LV_NO_REUSE", "LV_NO_GRAPH", "LV_NO_PDL", "LV_SEARCH_GROUP"):
            monkeypatch.delenv(k, raising=False)
        for k in env:
            monkeypatch.setenv(*(k.split("=") if "=" in k else (k, "1")))
        loc = lv.Localizer(full["prm"])                 # the switches are read by lv_create
        loc.map_build(full["map"])
        out = []
        for k in range(2):
            loc.set_state(full["x_props"][k], full["P0"])
            st, x, P, logs = loc.correct(full["sweeps"][k])
            out.append((st, x, P, logs))
        loc.close()
        return out
    base = run(["LV_NO_REUSE", "LV_NO_GRAPH", "LV_NO_PDL"])
    # ... nor may the shape of the level-0 search kernel: 8 lanes per query (default), one lane per query, or the
    # query-per-lane prologue with 8-lane scans (lv_search_coop_kernel)
    for env in ([], ["LV_NO_GRAPH"], ["LV_NO_REUSE"], ["LV_NO_PDL"], ["LV_NO_GRAPH", "LV_NO_PDL"],
                ["LV_SEARCH_GROUP=1"], ["LV_SEARCH_GROUP=32"], ["LV_SEARCH_GROUP=32", "LV_NO_GRAPH", "LV_NO_REUSE"]):
        got = run(env)
        for (st0, x0, P0, l0), (st1, x1, P1, l1) in zip(base, got):
            assert st0 == st1 == 0 and len(l0) == len(l1)
            assert (x0 == x1).all() and (P0 == P1).all()
            for a, b in zip(l0, l1):
                assert a["n_matches"] == b["n_matches"]
                assert (a["HTH"] == b["HTH"]).all() and (a["HTh"] == b["HTh"]).all() and (a["dx"] == b["dx"]).all()


def test_streaming_predict_correct_map_update(lv, O, full):
    """three sweeps of the sequence: IMU propagation, iterated update, Mapper::add with the 0.2 m rule.

    Three comparisons: every evaluation at the GPU's own iterate (check_every_evaluation: same Nm, HTH to 1e-12, dx to 1e-9 —
    the tight one); `step`, an oracle chain restarted every update from the GPU's prior (x, P); `free`, one that never sees the
    GPU's state.  The free chain's
    prior differs from the GPU's by ~1e-8 after the first update, which moves ~0.1 % of the fp32 world points by
    one ulp and flips a handful of hard gates (Mapper.cpp:81, Plane::is_plane); one flipped match of 58 000
    moves the pose by ~1e-5 m, so its bar is 2e-4 m (SURVEY 8c's 1e-5 m bar is for identical priors)."""
    prm = full["prm"]
    loc = lv.Localizer(prm)
    loc.map_build(full["map"])
    om = O.Map(O.KNN_KDTREE)
    om.build(full["map"])
    x, P = full["x_props"][0].copy(), full["P0"].copy()
    xo, Po = x.copy(), P.copy()
    loc.set_state(x, P)
    cov = (prm.covariance_gyroscope, prm.covariance_acceleration, prm.covariance_bias_gyroscope,
           prm.covariance_bias_acceleration)
    for k in range(3):
        sweep = full["sweeps"][k]
        x_prior, P_prior = loc.get_state()
        st, x, P, logs = loc.correct(sweep, time=0.1 * k)
        used = loc.last_neighbours(len(sweep))                      # before any inspection call searches again
        st_s, xs, Ps, logs_s = om.update_iterated(x_prior, P_prior, full["oprm"], sweep)       # step chain
        st_o, xo, Po, logs_o = om.update_iterated(xo, Po, full["oprm"], sweep)                 # free chain
        assert st == st_s == st_o == 0 and len(logs) == len(logs_s) == len(logs_o)
        # every evaluation, at the GPU's own iterate: same matches, same normal equations, same step (1e-9) — no inference
        check_every_evaluation(O, om, full["oprm"], x_prior, P_prior, sweep, logs, loc)
        # what the LAST evaluation really fitted its planes to (stored neighbours where lv_reuse_kernel vouched for them, searched
        # ones elsewhere) against a fresh search at the same iterate: the same five map points for every query, except where the
        # 5th and 6th are exactly equidistant
        fresh = loc.match_all(logs[-2]["x_after"], sweep)
        inside = np.isfinite(fresh["nn_sqd"][:, 4])
        differ = np.nonzero(inside & (np.sort(used, 1) != np.sort(fresh["nn_idx"], 1)).any(1))[0]
        assert len(differ) <= 4
        assert len(boundary_ties(om, fresh["g"][differ])) == len(differ)
        loc.set_state(x, P)                                        # the inspection calls moved the device state
        # the two free-running oracle chains are weaker statements (their iterates drift by ~1e-9, which re-rounds ~0.1 % of the fp32
        # world points and now and then flips one hard gate of 58 000): unconditional bars, no flip detection
        for a, b, c in zip(logs, logs_s, logs_o):
            assert abs(a["n_matches"] - b["n_matches"]) <= 4
            assert abs(a["n_matches"] - c["n_matches"]) <= 1e-3 * c["n_matches"]
            assert np.abs(a["dx"] - b["dx"]).max() < 2e-5, np.abs(a["dx"] - b["dx"]).max()
        assert np.abs(x - xs).max() < 2e-5, np.abs(x - xs)
        assert np.abs(P - Ps).max() <= 1e-5 * np.abs(Ps).max()
        assert np.abs(x[:7] - xo[:7]).max() < (1e-7 if k == 0 else 2e-4), np.abs(x - xo)
        assert np.abs(x - xo).max() < (1e-7 if k == 0 else 2e-3), np.abs(x - xo)
        assert loc.L.lv_last_time_updated(loc.h) == pytest.approx(0.1 * k)
        g = bench.world_points(sweep, x)                           # main.cpp:101: map.add(global points, t2, true)
        loc.map_add(g, downsample=True)
        om.add(g, downsample=True)
        assert loc.map_size() == om.size()
        for _ in range(4):                                         # Localizator::propagate_to: IMU samples to the next sweep
            acc, gyr = -x[23:26] + np.array([0.05, 0.0, 0.0]), np.array([0.0, 0.0, 0.01])   # at rest the accelerometer reads -grav
            x_b, P_b = loc.get_state()
            loc.predict(acc, gyr, 0.025)
            xo, Po = O.predict(xo, Po, acc, gyr, 0.025, *cov)
            x_a, P_a = loc.get_state()
            x_r, P_r = O.predict(x_b, P_b, acc, gyr, 0.025, *cov)
            assert np.abs(x_a - x_r).max() < 1e-12 and np.abs(P_a - P_r).max() <= 1e-12 * np.abs(P_r).max()
        # re-anchor on the next ground truth so that the synthetic sweeps (made at the true poses) stay consistent
        if k < 2:
            x, P = loc.get_state()
            x[:7] = full["x_props"][k + 1][:7]
            xo[:7] = full["x_props"][k + 1][:7]
            loc.set_state(x, P)
    assert sorted(map(tuple, loc.map_points().tolist())) == sorted(map(tuple, om.points().tolist()))
    loc.close()
