"""Diagnosis: cfg2, first sub-sweep: per-point difference between the GPU's fresh search and the oracle, with the 6-NN of the differing queries."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import bench
lv = G.load_package(); O = G.load_oracle()
cfg = "cfg2"
prm = bench.config_params(lv, cfg)
world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=10, prm=prm, cfg=cfg)
loc = lv.Localizer(prm); loc.map_build(mp)
om = O.Map(O.KNN_REF_IKDTREE if O.ref_available() else O.KNN_KDTREE); om.build(mp)
okd = O.Map(O.KNN_KDTREE); okd.build(mp)
oprm = bench.oracle_params(O, prm)
for k in range(len(sweeps)):
    got = loc.match_all(x_props[k], sweeps[k]); ref = om.match_all(x_props[k], oprm, sweeps[k])
    bad = np.nonzero((got["valid"] != ref["valid"]) | (got["plane"] != ref["plane"]).any(1))[0]
    print("sub-sweep", k, "differing queries:", bad, flush=True)
    for i in bad[:4]:
        found, idx, sqd, nn = okd.knn(ref["g"][i], 7)
        print("   query", i, "7-NN d2 (kd-tree):", sqd, "\n     gpu d2", got["nn_sqd"][i], "ref d2", ref["nn_sqd"][i], "valid", got["valid"][i], ref["valid"][i])
loc.close()
