"""What compute-sanitizer runs (tools/gpu_sanitize.sh): every kernel of the path once or twice on a small scene, checked against
the oracle so that a sanitizer-induced slowdown cannot hide a wrong result.  Both search paths, the map update, the device tick."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G
lv = G.load_package(); O = G.load_oracle()
VARIANTS = ((0, None), (0, "4"), (0, "32"), (0, "1"), (1, None))                    # every shape of the level-0 search
if os.environ.get("LV_SANITIZE_QUICK"):
    VARIANTS = ((0, "4"), (1, None))
for sort_queries, group in VARIANTS:
    os.environ.pop("LV_SEARCH_GROUP", None)
    if group:
        os.environ["LV_SEARCH_GROUP"] = group
    prm = lv.params_from_yaml(os.path.join(lv.CONFIG_DIR, "xaloc.yaml"), max_map_points=1 << 17, max_points=1 << 13, sort_queries=sort_queries)
    world = lv.SynthWorld(20260924, 40000)
    truth = world.pose(10.0, prm)
    sweep = world.sweep(truth, rings=16, azimuths=256, seed=1)
    far = sweep[:512] + np.float32([0.0, 0.0, 0.9])                 # queries off the surfaces: ring search, blind queries
    sweep = np.ascontiguousarray(np.concatenate([sweep, far]))
    loc = lv.Localizer(prm)
    loc.map_build(world.map())
    loc.init_state()
    x0, P0 = loc.get_state()
    d = np.zeros(23); d[0:3] = [0.04, -0.03, 0.02]; d[3:6] = [0.004, -0.003, 0.005]
    x_prop = O.boxplus(truth, d)
    om = O.Map(O.KNN_KDTREE); om.build(world.map())
    oprm = O.make_params(max_num_iters=prm.MAX_NUM_ITERS, estimate_extrinsics=prm.estimate_extrinsics, max_dist_plane=prm.MAX_DIST_PLANE,
                         planes_threshold=prm.PLANES_THRESHOLD, lidar_noise=prm.LiDAR_noise, degeneracy_threshold=prm.degeneracy_threshold)
    got, ref = loc.match_all(x_prop, sweep), om.match_all(x_prop, oprm, sweep)
    inside = np.isfinite(got["nn_sqd"][:, 4])
    assert (got["nn_sqd"][inside] == ref["nn_sqd"][inside]).all() and (got["valid"] == ref["valid"]).all() and (got["plane"] == ref["plane"]).all()
    for k in range(2):
        loc.set_state(x_prop, P0)
        st, x, P, logs = loc.correct(sweep)
        st_o, x_o, P_o, logs_o = om.update_iterated(x_prop, P0, oprm, sweep)
        assert st == st_o == 0 and len(logs) == len(logs_o) and np.abs(x - x_o).max() < 1e-7
        assert loc.last_neighbours(len(sweep)).shape == (len(sweep), 5)
        loc.map_add_last_sweep(True)
        om.add(om.match_all(x, oprm, sweep)["g"], downsample=True)
        loc.map_status()
        assert abs(loc.map_size() - len(om.points())) <= 2
        loc.propagate_device(np.tile([0.0, 0.0, 9.8], (4, 1)), np.zeros((4, 3)), np.full(4, 1e-3))
    loc.close()
    print("sanitize target: sort_queries=%d search group %s ok" % (sort_queries, group or "default"), flush=True)
