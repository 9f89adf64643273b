#!/usr/bin/env python
"""Generate the committed golden fixtures (tests/golden/*.npz) with the CPU oracle.

The reference ships no golden vectors for this path, so these pin the ORACLE: its
kNN is itself pinned to the reference's own ikd-Tree (tests/test_oracle.py), the rest is the Eigen-free
restatement.  Run in the authoring container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
from conftest import Scene  # noqa: E402

CASES = {
    "xaloc_small": dict(yaml="xaloc.yaml", seed=101, m=20000, rings=16, azimuths=64),
    "kitti_small": dict(yaml="kitti.yaml", seed=202, m=20000, rings=16, azimuths=64),
    "ouster_small": dict(yaml="ouster.yaml", seed=303, m=20000, rings=16, azimuths=64, elev=(-22.5, 22.5)),
}


def main():
    lv = G.load_package()
    O = G.load_oracle()
    for name, kw in CASES.items():
        sc = Scene(lv, O, **kw)
        backend = O.KNN_REF_IKDTREE if O.ref_available() else O.KNN_KDTREE
        om = O.Map(backend)
        om.build(sc.map)
        m = om.match_all(sc.x_prop, sc.oprm, sc.sweep)
        st, hx, h = om.measure(sc.x_prop, sc.oprm, sc.sweep)
        st, x, P, logs = om.update_iterated(sc.x_prop, sc.P0, sc.oprm, sc.sweep)
        assert st == 0
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), yaml=kw["yaml"], map=sc.map, sweep=sc.sweep, x_prop=sc.x_prop, P0=sc.P0,
            truth=sc.truth, valid=m["valid"], nn_sqd=m["nn_sqd"], plane=m["plane"], dist=m["dist"], g=m["g"],
            h_x=hx, h=h, n_matches=np.array([l["n_matches"] for l in logs]),
            converged=np.array([l["converged"] for l in logs]), HTH=np.stack([l["HTH"] for l in logs]),
            HTh=np.stack([l["HTh"] for l in logs]), dx=np.stack([l["dx"] for l in logs]),
            x_after=np.stack([l["x_after"] for l in logs]), x_final=x, P_final=P,
            knn_backend="reference ikd-Tree (oracle/_ref)" if backend == O.KNN_REF_IKDTREE else "oracle kd-tree")
        print(name, "evals", len(logs), "Nm", [l["n_matches"] for l in logs], "kNN:", "ref" if backend == 2 else "port")


if __name__ == "__main__":
    main()
