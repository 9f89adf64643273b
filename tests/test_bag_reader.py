"""lv_bag_*: the rosbag (format 2.0) reader that stands in for roscpp in front of Accumulator::receive_lidar / receive_imu.
A bag written by tests/bagwriter.py (the layout `rosbag record` produces: bag header, chunk with connection + message
records, index, connection records, chunk info) is read back: message order, topics, types, record times; the
sensor_msgs/PointCloud2 payloads go through lv_bag_parse_pointcloud2 + lv_pointcloud2_to_points_checked and must equal the
numpy restatement of PointCloudProcessor::msg2points; sensor_msgs/Imu payloads through lv_bag_parse_imu."""
import ctypes as C

import numpy as np
import pytest

import bagwriter as W

VELO = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                 "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"], "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
OUSTER = np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring", "ambient", "range"],
                   "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "<u1", "<u2", "<u4"],
                   "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48})


class BagMessage(C.Structure):
    _fields_ = [("conn", C.c_int32), ("topic", C.c_char_p), ("type", C.c_char_p), ("sec", C.c_uint32), ("nsec", C.c_uint32),
                ("data", C.POINTER(C.c_uint8)), ("size", C.c_int64)]


class CloudLayout(C.Structure):
    _fields_ = [("point_step", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("off_time", C.c_int32), ("off_range", C.c_int32)]


class PC2View(C.Structure):
    _fields_ = [("stamp_sec", C.c_uint32), ("stamp_nsec", C.c_uint32), ("height", C.c_uint32), ("width", C.c_uint32),
                ("point_step", C.c_uint32), ("row_step", C.c_uint32), ("is_bigendian", C.c_int32), ("is_dense", C.c_int32),
                ("n_points", C.c_int64), ("data", C.POINTER(C.c_uint8)), ("data_bytes", C.c_int64), ("layout", CloudLayout)]


class ImuSample(C.Structure):
    _fields_ = [("stamp_sec", C.c_uint32), ("stamp_nsec", C.c_uint32), ("orientation", C.c_double * 4),
                ("angular_velocity", C.c_double * 3), ("linear_acceleration", C.c_double * 3)]


def _cloud(dtype, n, rng, lidar):
    p = np.zeros(n, dtype)
    for k in "xyz":
        p[k] = rng.uniform(-50, 50, n).astype(np.float32)
    if lidar == "velodyne":
        p["intensity"] = rng.uniform(0, 255, n).astype(np.float32)
        p["time"] = np.sort(rng.uniform(0, 0.1, n)).astype(np.float32)
    else:
        p["reflectivity"] = rng.integers(0, 60000, n)
        p["t"] = np.sort(rng.integers(0, 100_000_000, n)).astype(np.uint32)
        p["range"] = rng.integers(1000, 90000, n)
    return p


def test_bag_round_trip(lv, tmp_path):
    L = lv.lib()
    rng = np.random.default_rng(4)
    clouds = [_cloud(VELO, 500, rng, "velodyne"), _cloud(VELO, 321, rng, "velodyne"), _cloud(OUSTER, 400, rng, "ouster")]
    msgs = [(0, 100, 5, W.imu(100, 5, (0, 0, 0, 1), (0.1, 0.2, 0.3), (0.0, 0.1, 9.8))),
            (1, 100, 100000, W.pointcloud2(clouds[0], 100, 50_000_000)),
            (0, 100, 200000, W.imu(100, 200000, (0, 0, 0.1, 0.99), (-0.1, 0.0, 0.05), (0.3, 0.0, 9.7))),
            (2, 100, 300000, W.pointcloud2(clouds[2], 100, 150_000_000)),
            (1, 100, 400000, W.pointcloud2(clouds[1], 100, 250_000_000))]
    conns = {0: ("/imu/data", "sensor_msgs/Imu"), 1: ("/velodyne_points", "sensor_msgs/PointCloud2"),
             2: ("/os_cloud_node/points", "sensor_msgs/PointCloud2")}
    path = str(tmp_path / "t.bag")
    W.write_bag(path, conns, msgs)
    bag = C.c_void_p()
    assert L.lv_bag_open(path.encode(), C.byref(bag)) == lv.OK
    L.lv_bag_connection_count.argtypes = [C.c_void_p]
    L.lv_bag_next.argtypes = [C.c_void_p, C.POINTER(BagMessage)]
    L.lv_bag_rewind.argtypes = [C.c_void_p]
    L.lv_bag_close.argtypes = [C.c_void_p]
    assert L.lv_bag_connection_count(bag) == 3
    for rewind in range(2):
        got = []
        m = BagMessage()
        while L.lv_bag_next(bag, C.byref(m)) == 1:
            got.append((m.conn, m.topic.decode(), m.type.decode(), m.sec, m.nsec, bytes(C.string_at(m.data, m.size))))
        assert [(g[0], g[3], g[4], g[5]) for g in got] == msgs
        assert [g[1:3] for g in got] == [conns[c] for c, *_ in msgs]
        L.lv_bag_rewind(bag)
    # payloads
    L.lv_bag_parse_pointcloud2.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.POINTER(PC2View)]
    L.lv_bag_parse_imu.argtypes = [C.c_char_p, C.c_int64, C.POINTER(ImuSample)]
    L.lv_pointcloud2_to_points_checked.argtypes = [C.c_int, C.POINTER(CloudLayout), C.POINTER(C.c_uint8), C.c_int64, C.c_int64, C.c_uint64,
                                                   C.c_int, C.c_int, C.c_double, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                                   C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for (cid, sec, nsec, payload), cloud, lidar in ((msgs[1], clouds[0], "velodyne"), (msgs[4], clouds[1], "velodyne"), (msgs[3], clouds[2], "ouster")):
        v = PC2View()
        assert L.lv_bag_parse_pointcloud2(payload, len(payload), lv.LIDAR_TYPES[lidar], C.byref(v)) == lv.OK
        assert v.n_points == len(cloud) and v.point_step == cloud.dtype.itemsize and v.data_bytes == cloud.nbytes
        assert (v.layout.off_x, v.layout.off_y, v.layout.off_z) == (0, 4, 8)
        n = v.n_points
        xyz, t = np.zeros((n, 3), np.float32), np.zeros(n)
        inten, rngv = np.zeros(n, np.float32), np.zeros(n, np.float32)
        stamp_us = v.stamp_sec * 1_000_000 + v.stamp_nsec // 1000
        fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
        assert L.lv_pointcloud2_to_points_checked(lv.LIDAR_TYPES[lidar], C.byref(v.layout), v.data, v.data_bytes, n, stamp_us, 1, 1, 0.1,
                                                  xyz.ctypes.data_as(fp), t.ctypes.data_as(dp), inten.ctypes.data_as(fp),
                                                  rngv.ctypes.data_as(fp)) == lv.OK
        ref = lv.pointcloud2_to_points(lidar, cloud, stamp_us, True, True, 0.1)       # the same parser fed from the numpy array
        assert (xyz == ref[0]).all() and (t == ref[1]).all() and (inten == ref[2]).all() and (rngv == ref[3]).all()
        assert (xyz == np.stack([cloud["x"], cloud["y"], cloud["z"]], 1)).all()
        # bounds: a count the data cannot hold, a field outside the point
        assert L.lv_pointcloud2_to_points_checked(lv.LIDAR_TYPES[lidar], C.byref(v.layout), v.data, v.data_bytes, n + 1, stamp_us, 1, 1, 0.1,
                                                  xyz.ctypes.data_as(fp), t.ctypes.data_as(dp), None, None) == lv.ERR_ARG
        bad = CloudLayout.from_buffer_copy(v.layout)
        bad.off_time = v.layout.point_step - 2
        assert L.lv_pointcloud2_to_points_checked(lv.LIDAR_TYPES[lidar], C.byref(bad), v.data, v.data_bytes, n, stamp_us, 1, 1, 0.1,
                                                  xyz.ctypes.data_as(fp), t.ctypes.data_as(dp), None, None) == lv.ERR_ARG
        # the wrong LiDAR type does not find its fields
        other = "ouster" if lidar == "velodyne" else "velodyne"
        assert L.lv_bag_parse_pointcloud2(payload, len(payload), lv.LIDAR_TYPES[other], C.byref(v)) == lv.ERR_IO
        assert L.lv_bag_parse_pointcloud2(payload, len(payload) - 7, lv.LIDAR_TYPES[lidar], C.byref(v)) == lv.ERR_IO   # truncated
    s = ImuSample()
    assert L.lv_bag_parse_imu(msgs[2][3], len(msgs[2][3]), C.byref(s)) == lv.OK
    assert (s.stamp_sec, s.stamp_nsec) == (100, 200000)
    assert list(s.orientation) == [0, 0, 0.1, 0.99] and list(s.angular_velocity) == [-0.1, 0.0, 0.05]
    assert list(s.linear_acceleration) == [0.3, 0.0, 9.7]
    assert L.lv_bag_parse_imu(msgs[2][3], 100, C.byref(s)) == lv.ERR_IO
    L.lv_bag_close(bag)


def test_bag_errors(lv, tmp_path):
    L = lv.lib()
    bag = C.c_void_p()
    assert L.lv_bag_open(b"/nonexistent.bag", C.byref(bag)) == lv.ERR_IO
    p = tmp_path / "junk.bag"
    p.write_bytes(b"not a bag at all, just bytes")
    assert L.lv_bag_open(str(p).encode(), C.byref(bag)) == lv.ERR_IO
    # compressed chunks are refused, not misread
    path = str(tmp_path / "c.bag")
    W.write_bag(path, {0: ("/imu", "sensor_msgs/Imu")}, [(0, 1, 2, W.imu(1, 2, (0, 0, 0, 1), (0, 0, 0), (0, 0, 9.8)))], compression="lz4")
    assert L.lv_bag_open(path.encode(), C.byref(bag)) == lv.ERR_IO
    # truncated file
    good = str(tmp_path / "g.bag")
    W.write_bag(good, {0: ("/imu", "sensor_msgs/Imu")}, [(0, 1, 2, W.imu(1, 2, (0, 0, 0, 1), (0, 0, 0), (0, 0, 9.8)))])
    raw = open(good, "rb").read()
    (tmp_path / "trunc.bag").write_bytes(raw[:len(raw) - 50])
    assert L.lv_bag_open(str(tmp_path / "trunc.bag").encode(), C.byref(bag)) == lv.ERR_IO
