"""ctypes binding of the CPU oracle (oracle/liblv_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
STATE_LEN = 26
DOF = 23

KNN_BRUTE, KNN_KDTREE, KNN_REF_IKDTREE = 0, 1, 2
OK, EMPTY_MAP, TOO_FEW_MATCHES, BAD_ARG = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [
        ("max_num_iters", C.c_int32),
        ("estimate_extrinsics", C.c_int32),
        ("max_dist_plane", C.c_double),
        ("planes_threshold", C.c_float),
        ("pad_", C.c_float),
        ("lidar_noise", C.c_double),
        ("degeneracy_threshold", C.c_double),
        ("limits", C.c_double * DOF),
    ]


class IterLog(C.Structure):
    _fields_ = [
        ("n_matches", C.c_int64),
        ("converged", C.c_int32),
        ("pad_", C.c_int32),
        ("HTH", C.c_double * 144),
        ("HTh", C.c_double * 12),
        ("dx", C.c_double * DOF),
        ("x_after", C.c_double * STATE_LEN),
    ]


def make_params(max_num_iters=3, estimate_extrinsics=True, max_dist_plane=2.0, planes_threshold=0.05,
                lidar_noise=0.001, degeneracy_threshold=5.0, limits=None):
    """Defaults are config/xaloc.yaml:13,31,45-50 and main.cpp:145."""
    p = Params()
    p.max_num_iters = int(max_num_iters)
    p.estimate_extrinsics = int(bool(estimate_extrinsics))
    p.max_dist_plane = float(max_dist_plane)
    p.planes_threshold = float(planes_threshold)
    p.lidar_noise = float(lidar_noise)
    p.degeneracy_threshold = float(degeneracy_threshold)
    lim = [0.001] * DOF if limits is None else list(limits)
    for i in range(DOF):
        p.limits[i] = float(lim[i])
    return p


def build(force=False):
    """Compile liblv_oracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liblv_oracle.so")
    src = os.path.join(_HERE, "lv_oracle.cpp")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src)
    if force or stale or not os.path.exists(os.path.join(_HERE, "_ref", "libikdtree_ref.so")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = build()
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    vp = C.c_void_p
    L.lvo_map_create.restype = vp
    L.lvo_map_create.argtypes = [C.c_int]
    L.lvo_map_destroy.argtypes = [vp]
    L.lvo_map_build.argtypes = [vp, fp, C.c_int64]
    L.lvo_map_add.argtypes = [vp, fp, C.c_int64, C.c_int]
    L.lvo_map_size.restype = C.c_int64
    L.lvo_map_size.argtypes = [vp]
    L.lvo_map_points.restype = C.c_int64
    L.lvo_map_points.argtypes = [vp, fp, C.c_int64]
    L.lvo_knn.argtypes = [vp, fp, C.c_int, C.POINTER(C.c_int32), fp, fp]
    L.lvo_match_all.argtypes = [vp, dp, C.POINTER(Params), fp, C.c_int64, C.POINTER(C.c_uint8),
                                C.POINTER(C.c_int32), fp, fp, fp, fp]
    L.lvo_measure.argtypes = [vp, dp, C.POINTER(Params), fp, C.c_int64, dp, dp, C.POINTER(C.c_int64)]
    L.lvo_measure_reduced.argtypes = [vp, dp, C.POINTER(Params), fp, C.c_int64, dp, dp, C.POINTER(C.c_int64)]
    L.lvo_update_iterated.argtypes = [vp, dp, dp, C.POINTER(Params), fp, C.c_int64, C.POINTER(IterLog),
                                      C.POINTER(C.c_int32)]
    L.lvo_update_step.argtypes = [dp, dp, dp, C.POINTER(Params), dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int32)]
    L.lvo_update_finish.argtypes = [dp, dp, dp, dp, dp, dp]
    L.lvo_predict.argtypes = [dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    L.lvo_init_state.argtypes = [dp, dp, fp, fp, fp, fp]
    L.lvo_boxplus.argtypes = [dp, dp]
    L.lvo_boxminus.argtypes = [dp, dp, dp]
    L.lvo_quat_to_rot.argtypes = [dp, dp]
    L.lvo_plane_fit.argtypes = [fp, C.c_float, fp, C.POINTER(C.c_int)]
    L.lvo_inverse.argtypes = [dp, C.c_int, dp]
    L.lvo_sym_eig6.argtypes = [dp, dp, dp]
    L.lvo_last_timing.argtypes = [dp, dp]
    _lib = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libikdtree_ref.so"))


class Map:
    """Mapper (src/Modules/Mapper.cpp) restated; backend selects the exact-kNN implementation."""

    def __init__(self, backend=KNN_KDTREE):
        self.L = lib()
        self.h = self.L.lvo_map_create(backend)
        if not self.h:
            raise RuntimeError("oracle map backend %d unavailable" % backend)
        self.backend = backend

    def __del__(self):
        try:
            if self.h:
                self.L.lvo_map_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def build(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        return self.L.lvo_map_build(self.h, _f(xyz), xyz.shape[0])

    def add(self, xyz, downsample=True):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        return self.L.lvo_map_add(self.h, _f(xyz), xyz.shape[0], int(downsample))

    def size(self):
        return int(self.L.lvo_map_size(self.h))

    def points(self):
        n = self.size()
        out = np.zeros((max(n, 1), 3), dtype=np.float32)
        k = self.L.lvo_map_points(self.h, _f(out), n)
        return out[:k]

    def knn(self, g, k=5):
        g = np.ascontiguousarray(g, dtype=np.float32).reshape(3)
        idx = np.full(k, -1, dtype=np.int32)
        sqd = np.full(k, np.inf, dtype=np.float32)
        nn = np.zeros((k, 3), dtype=np.float32)
        found = self.L.lvo_knn(self.h, _f(g), k, idx.ctypes.data_as(C.POINTER(C.c_int32)), _f(sqd), _f(nn))
        return found, idx, sqd, nn

    def match_all(self, x, prm, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = xyz.shape[0]
        out = dict(valid=np.zeros(n, np.uint8), nn_idx=np.zeros((n, 5), np.int32),
                   nn_sqd=np.zeros((n, 5), np.float32), plane=np.zeros((n, 4), np.float32),
                   dist=np.zeros(n, np.float32), g=np.zeros((n, 3), np.float32))
        st = self.L.lvo_match_all(self.h, _d(x), C.byref(prm), _f(xyz), n,
                                  out["valid"].ctypes.data_as(C.POINTER(C.c_uint8)),
                                  out["nn_idx"].ctypes.data_as(C.POINTER(C.c_int32)), _f(out["nn_sqd"]),
                                  _f(out["plane"]), _f(out["dist"]), _f(out["g"]))
        out["status"] = st
        return out

    def measure(self, x, prm, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = xyz.shape[0]
        hx = np.zeros(12 * n, np.float64)
        h = np.zeros(n, np.float64)
        nm = C.c_int64(0)
        st = self.L.lvo_measure(self.h, _d(x), C.byref(prm), _f(xyz), n, _d(hx), _d(h), C.byref(nm))
        k = nm.value
        return st, hx[:12 * k].reshape(12, k).T.copy(), h[:k].copy()

    def measure_reduced(self, x, prm, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, dtype=np.float64)
        HTH = np.zeros((12, 12))
        HTh = np.zeros(12)
        nm = C.c_int64(0)
        st = self.L.lvo_measure_reduced(self.h, _d(x), C.byref(prm), _f(xyz), xyz.shape[0], _d(HTH), _d(HTh),
                                        C.byref(nm))
        return st, HTH, HTh, nm.value

    def update_iterated(self, x, P, prm, xyz):
        """Localizator::correct body.  Returns (status, x, P, logs)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        x = np.array(x, dtype=np.float64).copy()
        P = np.array(P, dtype=np.float64).reshape(DOF, DOF).copy()
        cap = prm.max_num_iters + 1
        logs = (IterLog * cap)()
        ne = C.c_int32(0)
        st = self.L.lvo_update_iterated(self.h, _d(x), _d(P), C.byref(prm), _f(xyz), xyz.shape[0], logs,
                                        C.byref(ne))
        out = []
        for i in range(ne.value):
            lg = logs[i]
            out.append(dict(n_matches=int(lg.n_matches), converged=int(lg.converged),
                            HTH=np.array(lg.HTH[:]).reshape(12, 12), HTh=np.array(lg.HTh[:]),
                            dx=np.array(lg.dx[:]), x_after=np.array(lg.x_after[:])))
        return st, x, P, out


def update_step(x_prop, P_prop, x_cur, prm, HTH, HTh):
    L = lib()
    a = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    x_prop, P_prop, x_cur, HTH, HTh = a(x_prop), a(P_prop), a(x_cur), a(HTH), a(HTh)
    dx = np.zeros(DOF)
    x_new = np.zeros(STATE_LEN)
    P_now = np.zeros((DOF, DOF))
    Kx = np.zeros((DOF, 12))
    conv = C.c_int32(0)
    L.lvo_update_step(_d(x_prop), _d(P_prop), _d(x_cur), C.byref(prm), _d(HTH), _d(HTh), _d(dx), _d(x_new),
                      _d(P_now), _d(Kx), C.byref(conv))
    return dx, x_new, P_now, Kx, conv.value


def update_finish(x_prop, x_new, dx, P_now, Kx):
    L = lib()
    a = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    P_out = np.zeros((DOF, DOF))
    L.lvo_update_finish(_d(a(x_prop)), _d(a(x_new)), _d(a(dx)), _d(a(P_now)), _d(a(Kx)), _d(P_out))
    return P_out


def predict(x, P, acc, gyro, dt, cov_gyro, cov_acc, cov_bias_gyro, cov_bias_acc):
    L = lib()
    x = np.array(x, dtype=np.float64).copy()
    P = np.array(P, dtype=np.float64).reshape(DOF, DOF).copy()
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    gyro = np.ascontiguousarray(gyro, dtype=np.float64)
    L.lvo_predict(_d(x), _d(P), _d(acc), _d(gyro), dt, cov_gyro, cov_acc, cov_bias_gyro, cov_bias_acc)
    return x, P


def init_state(q_imu=(0, 0, 0, 1), initial_gravity=(0, 0, -9.807), I_Rotation_L=(1, 0, 0, 0, 1, 0, 0, 0, 1),
               I_Translation_L=(0, 0, 0)):
    L = lib()
    x = np.zeros(STATE_LEN)
    P = np.zeros((DOF, DOF))
    f = lambda v: np.ascontiguousarray(v, dtype=np.float32)
    q, g, R, t = f(q_imu), f(initial_gravity), f(I_Rotation_L), f(I_Translation_L)
    L.lvo_init_state(_d(x), _d(P), _f(q), _f(g), _f(R), _f(t))
    return x, P


def boxplus(x, d):
    x = np.array(x, dtype=np.float64).copy()
    d = np.ascontiguousarray(d, dtype=np.float64)
    lib().lvo_boxplus(_d(x), _d(d))
    return x


def boxminus(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    d = np.zeros(DOF)
    lib().lvo_boxminus(_d(x), _d(y), _d(d))
    return d


def quat_to_rot(q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    R = np.zeros((3, 3))
    lib().lvo_quat_to_rot(_d(q), _d(R))
    return R


def plane_fit(pts5, threshold):
    pts5 = np.ascontiguousarray(pts5, dtype=np.float32).reshape(5, 3)
    abcd = np.zeros(4, np.float32)
    ok = C.c_int(0)
    lib().lvo_plane_fit(_f(pts5), C.c_float(threshold), _f(abcd), C.byref(ok))
    return abcd, bool(ok.value)


def inverse(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    out = np.zeros((n, n))
    lib().lvo_inverse(_d(A), n, _d(out))
    return out


def sym_eig6(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    ev = np.zeros(6)
    V = np.zeros((6, 6))
    lib().lvo_sym_eig6(_d(A), _d(ev), _d(V))
    return ev, V


def set_threads(n):
    lib().lvo_set_threads(int(n))


def last_timing():
    a = C.c_double(0)
    b = C.c_double(0)
    lib().lvo_last_timing(C.byref(a), C.byref(b))
    return a.value, b.value


# ---- deskew (Compensator) ---------------------------------------------------------------------
class State32(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("pos", C.c_float * 3), ("vel", C.c_float * 3), ("bw", C.c_float * 3),
                ("ba", C.c_float * 3), ("g", C.c_float * 3), ("RLI", C.c_float * 9), ("tLI", C.c_float * 3),
                ("a", C.c_float * 3), ("w", C.c_float * 3), ("time", C.c_double)]

    def as_tuple(self):
        return tuple(np.array(getattr(self, k)[:], dtype=np.float32).tobytes() for k, _ in self._fields_[:-1]) + (self.time,)


def _f3(v):
    return (C.c_float * 3)(*[float(t) for t in v])


def state_from_ikfom(x, time, a, w, initial_gravity):
    out = State32()
    lib().lvo_state_from_ikfom(_d(np.ascontiguousarray(x, np.float64)), C.c_double(time), _f3(a), _f3(w),
                               _f3(initial_gravity), C.byref(out))
    return out


def state_add_imu(state, a, w, time):
    out = State32.from_buffer_copy(state)
    lib().lvo_state_add_imu(C.byref(out), _f3(a), _f3(w), C.c_double(time))
    return out


def upsample(states, imu_a, imu_w, imu_t):
    ns, ni = len(states), len(imu_t)
    arr = (State32 * ns)(*states)
    a = np.ascontiguousarray(imu_a, np.float32).reshape(ni, 3)
    w = np.ascontiguousarray(imu_w, np.float32).reshape(ni, 3)
    t = np.ascontiguousarray(imu_t, np.float64)
    cap = ns + ni + 8
    out = (State32 * cap)()
    n = lib().lvo_upsample(arr, ns, _f(a), _f(w), _d(t), ni, out, cap)
    assert n <= cap
    return [State32.from_buffer_copy(out[i]) for i in range(n)]


def get_t2(path, t2):
    arr = (State32 * len(path))(*path)
    out = State32()
    lib().lvo_get_t2(arr, len(path), C.c_double(t2), C.byref(out))
    return out


def compensate(path, Xt2, xyz, t):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(t, np.float64)
    out = np.zeros_like(xyz)
    arr = (State32 * len(path))(*path)
    lib().lvo_compensate.restype = C.c_int64
    n = lib().lvo_compensate(arr, len(path), C.byref(Xt2), _f(xyz), _d(t), C.c_int64(xyz.shape[0]), _f(out))
    return out[:n]


# ---- downsampling ---------------------------------------------------------------------------------
def temporal_downsample(xyz, rate, min_dist):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    idx = np.zeros(xyz.shape[0], np.int32)
    lib().lvo_temporal_downsample.restype = C.c_int64
    n = lib().lvo_temporal_downsample(_f(xyz), C.c_int64(xyz.shape[0]), C.c_int(rate), C.c_double(min_dist),
                                      idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return idx[:n]


def voxelgrid_downsample(xyz, leaf):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    out = np.zeros_like(xyz)
    lib().lvo_voxelgrid_downsample.restype = C.c_int64
    n = lib().lvo_voxelgrid_downsample(_f(xyz), C.c_int64(xyz.shape[0]), C.c_float(leaf), _f(out), C.c_int64(xyz.shape[0]))
    if n < 0:
        raise ValueError("leaf size too small for the extent of the cloud (PCL refuses)")
    return out[:n]


# ---- wire format (numpy restatement of PointCloudProcessor::msg2points, no C involved) ------------------
def _us2s(t):      # Conversions::microsec2Sec (Utils.cpp:18-23): int quotient + int remainder * 1e-6
    t = int(t)
    return float(int(t // 1000000) + int(t % 1000000) * 1e-6)


def _ns2s(t):      # Conversions::nanosec2Sec (Utils.cpp:25-30)
    t = np.asarray(t, dtype=np.int64)
    return (t // 1000000000).astype(np.float64) + (t % 1000000000).astype(np.float64) * 1e-9


def pointcloud2_to_points(lidar, pts, header_stamp_us, stamp_beginning, offset_beginning, full_rotation_time):
    """pts: numpy structured array with the fields of the reference's point struct (Common.hpp:109-221):
    velodyne x y z intensity time | hesai x y z intensity(u1) timestamp | ouster x y z t(u4) reflectivity(u2) range(u4) |
    custom x y z intensity timestamp.  Returns xyz (n,3) f32, time (n,) f64, intensity f32, range f32."""
    xyz = np.stack([pts["x"], pts["y"], pts["z"]], axis=1).astype(np.float32)
    nrm = np.sqrt((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]) + xyz[:, 2] * xyz[:, 2]).astype(np.float32)
    if lidar == "velodyne":
        raw = pts["time"].astype(np.float64)
    elif lidar == "ouster":
        raw = _ns2s(pts["t"])
    else:
        raw = pts["timestamp"].astype(np.float64)
    begin = 0.0
    if lidar in ("velodyne", "ouster") and len(pts):
        begin = _us2s(header_stamp_us) + raw[0] if stamp_beginning else _us2s(header_stamp_us) + raw[0] - raw[-1]
        t = raw if offset_beginning else full_rotation_time + raw
    else:
        t = raw
    t = t + begin
    if lidar == "ouster":
        inten, rng = pts["reflectivity"].astype(np.float32), pts["range"].astype(np.float32)
    else:
        inten, rng = pts["intensity"].astype(np.float32), nrm
    return xyz, t, inten, rng

