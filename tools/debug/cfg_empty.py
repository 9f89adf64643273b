"""How many queries of a config's update have fewer than five map points within MAX_DIST_PLANE (they can never be reused today)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import bench
lv = G.load_package(); O = G.load_oracle()
for cfg in sys.argv[1:]:
    prm = bench.config_params(lv, cfg)
    world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=2, prm=prm, cfg=cfg)
    om = O.Map(O.KNN_KDTREE); om.build(mp)
    oprm = bench.oracle_params(O, prm)
    ref = om.match_all(x_props[0], oprm, sweeps[0])
    d4 = ref["nn_sqd"][:, 4]
    gate = prm.MAX_DIST_PLANE ** 2
    print(cfg, "queries", len(d4), " fewer than 5 points at all / 5th beyond the gate:", int((~np.isfinite(d4)).sum()), int((np.isfinite(d4) & (d4 >= gate)).sum()), " valid", int(ref["valid"].sum()), flush=True)
