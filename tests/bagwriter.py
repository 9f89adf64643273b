"""A minimal rosbag (format 2.0) WRITER for the tests of lv_bag_*: bag header, ONE chunk (uncompressed or pretend-compressed)
with connection + message records, then the connection records and a chunk-info record as `rosbag record` leaves them.
Serialises sensor_msgs/PointCloud2 and sensor_msgs/Imu by hand (little endian, u32-prefixed strings and arrays)."""
import struct

import numpy as np


def _field(name, value):
    b = name.encode() + b"=" + value
    return struct.pack("<I", len(b)) + b


def _record(header_fields, data):
    h = b"".join(_field(k, v) for k, v in header_fields)
    return struct.pack("<I", len(h)) + h + struct.pack("<I", len(data)) + data


def _string(s):
    s = s.encode()
    return struct.pack("<I", len(s)) + s


def _header(seq, sec, nsec, frame):
    return struct.pack("<III", seq, sec, nsec) + _string(frame)


_DT = {np.dtype("uint8"): 2, np.dtype("uint16"): 4, np.dtype("uint32"): 6, np.dtype("float32"): 7, np.dtype("float64"): 8}


def pointcloud2(points, sec, nsec, frame="lidar", seq=0):
    """numpy structured array -> serialised sensor_msgs/PointCloud2"""
    pts = np.ascontiguousarray(points)
    out = _header(seq, sec, nsec, frame) + struct.pack("<II", 1, len(pts))
    names = [n for n in pts.dtype.names if not n.startswith("_pad")]
    out += struct.pack("<I", len(names))
    for n in names:
        dt, off = pts.dtype.fields[n][:2]
        out += _string(n) + struct.pack("<IBI", off, _DT[dt], 1)
    out += struct.pack("<BII", 0, pts.dtype.itemsize, pts.dtype.itemsize * len(pts))
    raw = pts.tobytes()
    out += struct.pack("<I", len(raw)) + raw + struct.pack("<B", 1)
    return out


def imu(sec, nsec, quat, gyro, acc, frame="imu", seq=0):
    out = _header(seq, sec, nsec, frame)
    out += struct.pack("<4d", *quat) + struct.pack("<9d", *([0.0] * 9))
    out += struct.pack("<3d", *gyro) + struct.pack("<9d", *([0.0] * 9))
    out += struct.pack("<3d", *acc) + struct.pack("<9d", *([0.0] * 9))
    return out


def write_bag(path, connections, messages, compression="none"):
    """connections: {conn_id: (topic, type)}; messages: [(conn_id, sec, nsec, payload bytes)] in file order"""
    conn_recs = {}
    for cid, (topic, typ) in connections.items():
        ch = _field("topic", topic.encode()) + _field("type", typ.encode()) + _field("md5sum", b"0" * 32) + \
            _field("message_definition", b"# test")
        conn_recs[cid] = _record([("op", b"\x07"), ("conn", struct.pack("<I", cid)), ("topic", topic.encode())], ch)
    chunk = b""
    seen = set()
    for cid, sec, nsec, payload in messages:
        if cid not in seen:                       # rosbag writes the connection record before its first message in a chunk
            chunk += conn_recs[cid]
            seen.add(cid)
        chunk += _record([("op", b"\x02"), ("conn", struct.pack("<I", cid)), ("time", struct.pack("<II", sec, nsec))], payload)
    chunk_rec = _record([("op", b"\x05"), ("compression", compression.encode()), ("size", struct.pack("<I", len(chunk)))], chunk)
    index = _record([("op", b"\x04"), ("ver", struct.pack("<I", 1)), ("conn", struct.pack("<I", 0)), ("count", struct.pack("<I", 0))], b"")
    tail = b"".join(conn_recs.values())
    tail += _record([("op", b"\x06"), ("ver", struct.pack("<I", 1)), ("chunk_pos", struct.pack("<Q", 4096 + 13)),
                     ("start_time", struct.pack("<II", 0, 0)), ("end_time", struct.pack("<II", 0, 0)), ("count", struct.pack("<I", 0))], b"")
    bh_fields = [("op", b"\x03"), ("index_pos", struct.pack("<Q", 13 + 4096 + len(chunk_rec) + len(index))),
                 ("conn_count", struct.pack("<I", len(connections))), ("chunk_count", struct.pack("<I", 1))]
    h = b"".join(_field(k, v) for k, v in bh_fields)
    pad = 4096 - 4 - len(h) - 4
    bag_header = struct.pack("<I", len(h)) + h + struct.pack("<I", pad) + b" " * pad
    with open(path, "wb") as f:
        f.write(b"#ROSBAG V2.0\n" + bag_header + chunk_rec + index + tail)
