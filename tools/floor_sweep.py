"""Tuning aid: per-kernel device time of one h-evaluation as a function of the number of queries (latency floors)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G, bench
lv = G.load_package()
prm = lv.params_from_yaml(lv.CONFIG_DIR + "/xaloc.yaml", max_map_points=bench.MAP_POINTS + 4 * 65536, max_points=65536)
world, mp, sweeps, x_props, truths = bench.make_scene(lv, 0, n_sweeps=1, prm=prm)
loc = lv.Localizer(prm); loc.map_build(mp); loc.profile_enable(True)
for flush in (False, True):
    for n in (256, 2048, 8192, 32768, 65536):
        q = sweeps[0][:: 65536 // n][:n]
        for _ in range(3): loc.measure_reduced(x_props[0], q)
        loc.profile(reset=True)
        reps = 30
        for _ in range(reps):
            if flush: loc.flush_l2()
            loc.measure_reduced(x_props[0], q)
        p = loc.profile(reset=True)
        print(f"flush={int(flush)} n={n:6d}  measure group {1e3*p['measure_ms']/reps:7.2f} us  (search {1e3*p['search_ms']/reps:6.2f} upper {1e3*p['search_upper_ms']/reps:6.2f} fit {1e3*p['fit_ms']/reps:6.2f})")
