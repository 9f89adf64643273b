"""Diagnosis: the streaming sequence of tests/test_gpu_fullsize.py, evaluation by evaluation, with the per-point
comparison at the first evaluation whose normal equations differ from the oracle's."""
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

def run(env):
    for k in ("LV_NO_REUSE", "LV_NO_GRAPH", "LV_NO_PDL"): os.environ.pop(k, None)
    for k in env: os.environ[k] = "1"
    print("=== env", env, flush=True)
    loc = lv.Localizer(prm); loc.map_build(mp)
    om = O.Map(O.KNN_KDTREE); om.build(mp)
    loc.set_state(x_props[0], P0)
    for k in range(3):
        sweep = sweeps[k]
        x_prior, P_prior = loc.get_state()
        st, x, P, logs = loc.correct(sweep, time=0.1 * k)
        x_k = np.array(x_prior)
        for e, lg in enumerate(logs):
            st_o, HTH, HTh, nm = om.measure_reduced(x_k, oprm, sweep)
            rel = np.abs(HTH - lg["HTH"]).max() / np.abs(HTH).max()
            print(f"sweep {k} eval {e}: nm gpu {lg['n_matches']} oracle {nm}  rel HTH diff {rel:.3e}", flush=True)
            if rel > 1e-12 or nm != lg["n_matches"]:
                np.save(os.path.join(ROOT, "gpurun_out", "r2l", f"xk_{k}_{e}.npy"), x_k)
                got = loc.match_all(x_k, sweep); ref = om.match_all(x_k, oprm, sweep)
                bad = np.nonzero((got["valid"] != ref["valid"]) | (got["plane"] != ref["plane"]).any(1))[0]
                fin = np.isfinite(got["nn_sqd"][:, 4]) & np.isfinite(ref["nn_sqd"][:, 4])
                badd = np.nonzero(fin & (got["nn_sqd"] != ref["nn_sqd"]).any(1))[0]
                print("   fresh search at this iterate: points with different valid/plane:", bad[:10], " different distances:", badd[:10])
                st_f, HTH_f, _, nm_f = loc.measure_reduced(x_k, sweep)
                print("   fresh lv_measure_reduced at this iterate: nm", nm_f, "rel diff", np.abs(HTH - HTH_f).max() / np.abs(HTH).max())
                for i in list(bad[:3]) + list(badd[:3]):
                    print("   planes", got["plane"][i].tolist(), ref["plane"][i].tolist(), "idx", got["nn_idx"][i], ref["nn_idx"][i])
                    print("   point", i, "g", got["g"][i], "gpu d", got["nn_sqd"][i], "ref d", ref["nn_sqd"][i], "valid", got["valid"][i], ref["valid"][i])
            x_k = lg["x_after"]
        np.save(os.path.join(ROOT, "gpurun_out", "r2l", f"xafter_{k}.npy"), x)
        g = bench.world_points(sweep, x)
        loc.map_add(g, downsample=True); om.add(g, downsample=True)
        if k < 2:
            x, P = loc.get_state(); x[:7] = x_props[k + 1][:7]; loc.set_state(x, P)
    loc.close()

run([])
