"""Shared fixtures.  `-m "not gpu"` runs here (no GPU); `-m gpu` runs on the B200 box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lv():
    """the product package (ctypes over liblimovelo_b200.so)"""
    if not os.path.exists(os.path.join(ROOT, "limo-velo_b200", "liblimovelo_b200.so")):
        G.build()
    return G.load_package()


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (test infrastructure)"""
    return G.load_oracle()


def oracle_params(O, prm):
    return O.make_params(max_num_iters=prm.MAX_NUM_ITERS, estimate_extrinsics=prm.estimate_extrinsics,
                         max_dist_plane=prm.MAX_DIST_PLANE, planes_threshold=prm.PLANES_THRESHOLD,
                         lidar_noise=prm.LiDAR_noise, degeneracy_threshold=prm.degeneracy_threshold,
                         limits=list(prm.LIMITS))


class Scene:
    """seeded world + one sweep + predicted state, for a named YAML"""

    def __init__(self, lv, O, yaml="xaloc.yaml", seed=20260924, m=100000, rings=32, azimuths=256, s=8.0,
                 elev=(-24.8, 2.0), perturb=(0.05, 0.5), **over):
        self.prm = lv.params_from_yaml(os.path.join(lv.CONFIG_DIR, yaml), **over)
        self.world = lv.SynthWorld(seed, m)
        self.map = self.world.map()
        self.truth = self.world.pose(s, self.prm)
        self.sweep = self.world.sweep(self.truth, rings=rings, azimuths=azimuths, elev=elev, seed=seed + 7)
        rng = np.random.default_rng(seed)
        d = np.zeros(23)
        d[0:3] = rng.uniform(-perturb[0], perturb[0], 3)
        d[3:6] = rng.uniform(-perturb[1], perturb[1], 3) * np.pi / 180
        self.x_prop = O.boxplus(self.truth, d)
        self.x0, self.P0 = O.init_state(initial_gravity=self.prm.initial_gravity[:],
                                        I_Rotation_L=self.prm.I_Rotation_L[:],
                                        I_Translation_L=self.prm.I_Translation_L[:])
        self.oprm = oracle_params(O, self.prm)


@pytest.fixture(scope="session")
def scene_xaloc(lv, O):
    return Scene(lv, O, "xaloc.yaml", max_map_points=1 << 19, max_points=1 << 16)


@pytest.fixture(scope="session")
def scene_kitti(lv, O):
    return Scene(lv, O, "kitti.yaml", seed=20260926, m=80000, max_map_points=1 << 19, max_points=1 << 16)


@pytest.fixture(scope="session")
def scene_ouster(lv, O):
    return Scene(lv, O, "ouster.yaml", seed=20260927, m=80000, rings=32, azimuths=256, elev=(-22.5, 22.5),
                 max_map_points=1 << 19, max_points=1 << 16)


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
