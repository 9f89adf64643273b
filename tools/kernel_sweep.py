#!/usr/bin/env python
"""Micro-benchmark of the fused measure kernel for tuning (not the judged bench): device time of one
h-evaluation (CUDA events inside the library) for several voxel sizes / query orders."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", default="0.25,0.5,1.0")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--map", type=int, default=bench.MAP_POINTS)
    args = ap.parse_args()
    lv = G.load_package()
    O = G.load_oracle()
    base = lv.params_from_yaml(os.path.join(lv.CONFIG_DIR, "xaloc.yaml"))
    bench.MAP_POINTS = args.map
    world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=2, prm=base)
    sweep = sweeps[0]
    # variants of the query set
    rng = np.random.default_rng(0)
    g = bench.world_points(sweep, x_props[0])
    has_nb = np.ones(len(sweep), bool)
    variants = {"firing_order": sweep, "shuffled": sweep[rng.permutation(len(sweep))]}
    out = []
    for vs in [float(v) for v in args.voxels.split(",")]:
        prm = lv.params_from_yaml(os.path.join(lv.CONFIG_DIR, "xaloc.yaml"), voxel_size=vs,
                                  max_map_points=args.map + 2 * len(sweep), max_points=len(sweep))
        loc = lv.Localizer(prm)
        loc.profile_enable(True)
        loc.map_build(mp)
        pb = loc.profile(reset=True)
        if vs == 0.5:
            m = loc.match_all(x_props[0], sweep)
            inside = np.isfinite(m["nn_sqd"][:, 4])
            variants["only_points_with_5nn_in_radius"] = sweep[inside]
            variants["sorted_by_voxel"] = sweep[np.lexsort((np.floor(g[:, 0] / vs), np.floor(g[:, 1] / vs), np.floor(g[:, 2] / vs)))]
            print("points without 5 neighbours in radius:", int((~inside).sum()), "accepted:", int(m["valid"].sum()), flush=True)
        for name, q in variants.items():
            for _ in range(3):
                loc.measure_reduced(x_props[0], q)
            loc.profile(reset=True)
            for _ in range(args.reps):
                loc.flush_l2()
                loc.measure_reduced(x_props[0], q)
            p = loc.profile(reset=True)
            ms = p["measure_ms"] / max(1, p["measure_launches"])
            row = {"voxel": vs, "queries": name, "n": len(q), "measure_us": 1e3 * ms, "build_ms": pb["build_ms"],
                   "Gpts_per_s": len(q) / ms / 1e6}
            out.append(row)
            print(json.dumps(row), flush=True)
        loc.close()


if __name__ == "__main__":
    main()
