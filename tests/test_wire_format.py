"""Wire format of the LiDAR message (last item): PointCloudProcessor::msg2points restated on raw bytes
(host C++) against a numpy restatement — the four point structs of Common.hpp:109-221, the time semantics of the YAML
(stamp_beginning / offset_beginning / full_rotation_time) and the int arithmetic of Conversions::*2Sec.  No GPU."""
import numpy as np
import pytest

DTYPES = {
    # unaligned on purpose: drivers pack differently, the reader goes by the declared offsets
    "velodyne": np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                          "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"], "offsets": [0, 4, 8, 16, 20, 22], "itemsize": 32}),
    "hesai": np.dtype({"names": ["x", "y", "z", "intensity", "timestamp", "ring"],
                       "formats": ["<f4", "<f4", "<f4", "u1", "<f8", "<u2"], "offsets": [0, 4, 8, 16, 24, 32], "itemsize": 48}),
    "ouster": np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring", "range"],
                        "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "u1", "<u4"], "offsets": [0, 4, 8, 16, 20, 24, 26, 28], "itemsize": 48}),
    "custom": np.dtype({"names": ["x", "y", "z", "rgb", "intensity", "range", "timestamp", "ring"],
                        "formats": ["<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<f8", "<u2"], "offsets": [0, 4, 8, 16, 32, 36, 40, 48], "itemsize": 64}),
}


def make_cloud(lidar, n, rng, t0=1_695_000_000.25):
    pts = np.zeros(n, DTYPES[lidar])
    xyz = rng.uniform(-80, 80, (n, 3)).astype(np.float32)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rel = np.sort(rng.uniform(0.0, 0.1, n))
    if lidar == "velodyne":
        pts["intensity"] = rng.uniform(0, 255, n).astype(np.float32)
        pts["time"] = rel.astype(np.float32)
    elif lidar == "hesai":
        pts["intensity"] = rng.integers(0, 256, n)
        pts["timestamp"] = t0 + rel
    elif lidar == "ouster":
        pts["t"] = (rel * 1e9).astype(np.uint32)
        pts["reflectivity"] = rng.integers(0, 65536, n)
        pts["range"] = (np.linalg.norm(xyz, axis=1) * 1000).astype(np.uint32)
    else:
        pts["intensity"] = rng.uniform(0, 1, n).astype(np.float32)
        pts["timestamp"] = t0 + rel
    return pts


@pytest.mark.parametrize("lidar", ["velodyne", "hesai", "ouster", "custom"])
@pytest.mark.parametrize("stamp_beginning,offset_beginning", [(False, True), (True, True), (False, False), (True, False)])
def test_msg2points_matches_numpy_restatement(lv, O, lidar, stamp_beginning, offset_beginning):
    rng = np.random.default_rng(5)
    pts = make_cloud(lidar, 5000, rng)
    stamp_us = 1_695_000_000_350_000
    got = lv.pointcloud2_to_points(lidar, pts, stamp_us, stamp_beginning, offset_beginning, 0.1)
    ref = O.pointcloud2_to_points(lidar, pts, stamp_us, stamp_beginning, offset_beginning, 0.1)
    for a, b in zip(got, ref):
        assert a.dtype == b.dtype and (a == b).all()
    if lidar == "velodyne" and not stamp_beginning and offset_beginning:
        # get_begin_time: stamp + front.time - back.time, so the LAST point lands on stamp + front.time (xaloc.yaml:29-30)
        assert got[1][-1] == pytest.approx(1_695_000_000.35 + float(pts["time"][0]), abs=1e-6)
    if lidar == "hesai":
        assert (got[1] == pts["timestamp"]).all()                             # absolute stamps pass through
    # sort_points: stable order by time
    shuffled = rng.permutation(len(pts))
    idx = lv.time_sort_indices(got[1][shuffled])
    assert (np.diff(got[1][shuffled][idx]) >= 0).all()
    assert (idx == np.argsort(got[1][shuffled], kind="stable")).all()


def test_msg2points_edge_cases(lv, O):
    rng = np.random.default_rng(1)
    for lidar in DTYPES:
        empty = lv.pointcloud2_to_points(lidar, make_cloud(lidar, 0, rng), 0, False, True, 0.1)
        assert all(len(a) == 0 for a in empty)
        one = make_cloud(lidar, 1, rng)
        a = lv.pointcloud2_to_points(lidar, one, 1_000_000, False, True, 0.1)
        b = O.pointcloud2_to_points(lidar, one, 1_000_000, False, True, 0.1)
        assert all((u == v).all() for u, v in zip(a, b))
    # the whole Accumulator::process chain stays consistent: decimate by index, then sort
    pts = make_cloud("velodyne", 4000, rng)
    xyz, t, inten, rngv = lv.pointcloud2_to_points("velodyne", pts, 2_000_000, False, True, 0.1)
    keep = O.temporal_downsample(xyz, 4, 4.0)
    order = lv.time_sort_indices(t[keep])
    assert (np.diff(t[keep][order]) >= 0).all()
