"""CPU repro attempt: oracle chain over the streaming sequence; at every evaluation compare the CPU build of the product's
per-point code (tests/cpu_shim) with the oracle, per point."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import bench
import shim_binding as S
lv = G.load_package(); O = G.load_oracle()
prm = lv.params_from_yaml(lv.CONFIG_DIR + "/xaloc.yaml", max_map_points=bench.MAP_POINTS + 4 * 65536, max_points=65536)
world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=3, prm=prm)
x0, P0 = lv.init_state_host(prm)
oprm = bench.oracle_params(O, prm)
sprm = S.make_params(oprm, voxel_size=prm.voxel_size)
om = O.Map(O.KNN_KDTREE); om.build(mp)
sm = S.ShimMap(mp, cell=prm.voxel_size, max_dist=prm.MAX_DIST_PLANE)
x, P = x_props[0].copy(), P0.copy()
for k in range(2):
    sweep = sweeps[k]
    st, xn, Pn, logs = om.update_iterated(x, P, oprm, sweep)
    x_k = x.copy()
    for e, lg in enumerate(logs):
        ref = om.match_all(x_k, oprm, sweep)
        got = sm.match_all(x_k, sprm, sweep)
        bad = np.nonzero((got["valid"] != ref["valid"]) | (got["plane"] != ref["plane"]).any(1))[0]
        print(f"sweep {k} eval {e}: per-point plane/valid mismatches: {bad[:10]}", flush=True)
        for i in bad[:3]:
            print("   point", i, "g", ref["g"][i], "\n     shim plane", got["plane"][i], "\n     ref  plane", ref["plane"][i], "\n     d", got["nn_sqd"][i], ref["nn_sqd"][i], "idx", got["nn_idx"][i], ref["nn_idx"][i])
            np.save(f"/tmp/bad_{k}_{e}_{i}.npy", np.concatenate([x_k]))
        x_k = lg["x_after"]
    g = bench.world_points(sweep, xn)
    om.add(g, downsample=True); sm.add(g, downsample=True)
    x, P = xn.copy(), Pn.copy()
    if k < 2: x[:7] = x_props[k + 1][:7]
