import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import bench
import shim_binding as S
lv = G.load_package(); O = G.load_oracle()
prm = lv.params_from_yaml(lv.CONFIG_DIR + "/xaloc.yaml", max_map_points=bench.MAP_POINTS + 4 * 65536, max_points=65536)
world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=3, prm=prm)
oprm = bench.oracle_params(O, prm)
sprm = S.make_params(oprm, voxel_size=prm.voxel_size)
D = os.path.join(ROOT, "gpurun_out", "r2l")
xa0 = np.load(D + "/xafter_0.npy"); xk = np.load(D + "/xk_1_2.npy")
om = O.Map(O.KNN_KDTREE); om.build(mp)
sm = S.ShimMap(mp, cell=prm.voxel_size, max_dist=prm.MAX_DIST_PLANE)
g = bench.world_points(sweeps[0], xa0)
om.add(g, downsample=True); sm.add(g, downsample=True)
i = 53270
ref = om.match_all(xk, oprm, sweeps[1]); got = sm.match_all(xk, sprm, sweeps[1])
print("oracle plane", ref["plane"][i].tolist(), "shim plane", got["plane"][i].tolist())
print("d", ref["nn_sqd"][i], got["nn_sqd"][i])
allp = np.concatenate([mp, g]).astype(np.float32)
idx = ref["nn_idx"][i]
print("idx", idx, got["nn_idx"][i])
pts = om.points()
# find the five points by distance
gq = ref["g"][i]
d2 = ((pts.astype(np.float32) - gq) ** 2).sum(1)
o = np.argsort(d2)[:6]
print("five nearest (brute force):"); print(pts[o], d2[o])
np.save("/tmp/five.npy", pts[o[:5]])
A = pts[o[:5]].astype(np.float64)
print("singular values of A:", np.linalg.svd(A, compute_uv=False))
print("centered sv:", np.linalg.svd(A - A.mean(0), compute_uv=False))
n64 = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]; print("f64 LS normal", n64 / np.linalg.norm(n64), 1 / np.linalg.norm(n64))
