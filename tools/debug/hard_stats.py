"""How many queries of a cfg1 sweep level 0 cannot settle (CPU build of the product's search code)."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import bench
import shim_binding as S
lv = G.load_package()
prm = lv.params_from_yaml(lv.CONFIG_DIR + "/xaloc.yaml", max_map_points=bench.MAP_POINTS + 4 * 65536, max_points=65536)
world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=2, prm=prm)
for vox in (0.4, 0.6):
    sm = S.ShimMap(mp, cell=vox, max_dist=prm.MAX_DIST_PLANE)
    n = len(sweeps[0])
    lvl = np.zeros(n, np.int32); sc = np.zeros(n, np.int32)
    x = np.ascontiguousarray(x_props[0], np.float64); xyz = np.ascontiguousarray(sweeps[0], np.float32)
    sm.L.shim_query_stats(C.c_void_p(sm.h), x.ctypes.data_as(C.POINTER(C.c_double)), xyz.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(n), C.c_double(prm.MAX_DIST_PLANE),
                          lvl.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p))
    print("voxel", vox, "settled", (lvl == 0).sum(), "bucket-uncertified", (lvl == 1).sum(), "no slot", (lvl == 2).sum(), "mean bucket", sc[lvl < 2].mean())

# reuse fraction between evaluations (oracle chain for the iterates)
O = G.load_oracle()
oprm = bench.oracle_params(O, prm)
om = O.Map(O.KNN_KDTREE); om.build(mp)
x0, P0 = lv.init_state_host(prm)
st, xn, Pn, logs = om.update_iterated(x_props[0], P0, oprm, sweeps[0])
sm = S.ShimMap(mp, cell=0.4, max_dist=prm.MAX_DIST_PLANE)
xs = [x_props[0]] + [lg["x_after"] for lg in logs]
for e in range(1, len(logs)):
    reused, same = sm.reuse_check(xs[0], xs[e], sweeps[0], max_dist=prm.MAX_DIST_PLANE)
    print("eval", e, "reusable from the eval-0 reference:", int(reused.sum()), "of", len(reused), " |dx| from eval 0:", np.linalg.norm(xs[e][:3] - xs[0][:3]))
    reused, same = sm.reuse_check(xs[e - 1], xs[e], sweeps[0], max_dist=prm.MAX_DIST_PLANE)
    print("        reusable from the previous iterate:", int(reused.sum()))
