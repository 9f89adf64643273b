"""Diagnosis: the neighbours the update path used in its last evaluation against a fresh search at the same iterate."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import bench
lv = G.load_package(); O = G.load_oracle()
prm = lv.params_from_yaml(lv.CONFIG_DIR + "/xaloc.yaml", max_map_points=bench.MAP_POINTS + 4 * 65536, max_points=65536)
world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=3, prm=prm)
x0, P0 = lv.init_state_host(prm)
oprm = bench.oracle_params(O, prm)
cov = (prm.covariance_gyroscope, prm.covariance_acceleration, prm.covariance_bias_gyroscope, prm.covariance_bias_acceleration)
loc = lv.Localizer(prm); loc.map_build(mp)
om = O.Map(O.KNN_KDTREE); om.build(mp)
x, P = x_props[0].copy(), P0.copy()
loc.set_state(x, P)
for k in range(3):
    sweep = sweeps[k]
    x_prior, P_prior = loc.get_state()
    st, x, P, logs = loc.correct(sweep, time=0.1 * k)
    used = loc.last_neighbours(len(sweep))
    x_k = logs[-2]["x_after"] if len(logs) > 1 else x_prior
    fresh = loc.match_all(x_k, sweep)
    ref = om.match_all(x_k, oprm, sweep)
    ok = np.isfinite(fresh["nn_sqd"][:, 4])
    diff = np.nonzero(ok & (np.sort(used, 1) != np.sort(fresh["nn_idx"], 1)).any(1))[0]
    order = np.nonzero(ok & (used != fresh["nn_idx"]).any(1))[0]
    print(f"sweep {k}: {len(logs)} evaluations; queries whose last-evaluation neighbour SET differs from a fresh search: {diff[:10]} ({len(diff)}); order differs: {len(order)}", flush=True)
    pts = om.points().astype(np.float32)
    for i in list(diff[:4]) + [j for j in order[:3] if j not in diff]:
        g = fresh["g"][i]
        d = pts - g
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        o = np.argsort(d2, kind="stable")[:7]
        print("  query", i, "g", g, "\n    brute-force 7 nearest d2:", d2[o], "\n    used ids ", used[i], "\n    fresh ids", fresh["nn_idx"][i], "\n    fresh d2 ", fresh["nn_sqd"][i], "\n    oracle d2", ref["nn_sqd"][i],
              "\n    plane fresh", fresh["plane"][i], "oracle", ref["plane"][i])
    g = bench.world_points(sweep, x)
    loc.map_add(g, downsample=True); om.add(g, downsample=True)
    for _ in range(4):
        acc, gyr = -x[23:26] + np.array([0.05, 0.0, 0.0]), np.array([0.0, 0.0, 0.01])
        loc.predict(acc, gyr, 0.025)
    if k < 2:
        x, P = loc.get_state(); x[:7] = x_props[k + 1][:7]; loc.set_state(x, P)
loc.close()
