"""limovelo_b200 — Python (ctypes) access to liblimovelo_b200.so for tests and benchmarks.

The product is the C-ABI shared library (include/limovelo_b200.h) built from csrc/ (CUDA, sm_100a)
and host/ (C++); this module only marshals numpy arrays into it.  There is no Python or CPU
implementation behind it: if the library is missing, loading raises, and without a CUDA device
every compute call returns LV_ERR_CUDA (raised as RuntimeError here).

The directory name contains a hyphen, so import it through `__graft_entry__.load_package()`
(registers it as module `limovelo_b200`).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LV_LIB_PATH") or os.path.join(_HERE, "liblimovelo_b200.so")   # LV_LIB_PATH: variant builds of tools/
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "limovelo_b200.h")

STATE_LEN, DOF, MAX_EVALS = 26, 23, 8
OK, EMPTY_MAP, TOO_FEW_MATCHES, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_IO = range(7)
STATUS_NAMES = ["LV_OK", "LV_EMPTY_MAP", "LV_TOO_FEW_MATCHES", "LV_ERR_ARG", "LV_ERR_CUDA", "LV_ERR_CAPACITY",
                "LV_ERR_IO"]


class Params(C.Structure):
    _fields_ = [
        ("MAX_NUM_ITERS", C.c_int32), ("NUM_MATCH_POINTS", C.c_int32), ("estimate_extrinsics", C.c_int32),
        ("print_degeneracy_values", C.c_int32), ("MAX_DIST_PLANE", C.c_double), ("PLANES_THRESHOLD", C.c_float),
        ("pad0_", C.c_float), ("LiDAR_noise", C.c_double), ("degeneracy_threshold", C.c_double),
        ("LIMITS", C.c_double * DOF),
        ("covariance_gyroscope", C.c_double), ("covariance_acceleration", C.c_double),
        ("covariance_bias_gyroscope", C.c_double), ("covariance_bias_acceleration", C.c_double),
        ("initial_gravity", C.c_float * 3), ("I_Translation_L", C.c_float * 3), ("I_Rotation_L", C.c_float * 9),
        ("map_downsample_size", C.c_float), ("voxel_size", C.c_float), ("device", C.c_int32),
        ("sort_queries", C.c_int32), ("max_map_points", C.c_int64), ("max_points", C.c_int64),
        ("stream", C.c_void_p),
    ]


class IterLog(C.Structure):
    _fields_ = [("n_matches", C.c_int64), ("converged", C.c_int32), ("degenerate", C.c_int32),
                ("HTH", C.c_double * 144), ("HTh", C.c_double * 12), ("dx", C.c_double * DOF),
                ("x_after", C.c_double * STATE_LEN)]


class Profile(C.Structure):
    _fields_ = [("measure_ms", C.c_double), ("solve_ms", C.c_double), ("build_ms", C.c_double),
                ("measure_launches", C.c_int64), ("solve_launches", C.c_int64), ("build_launches", C.c_int64),
                ("total_launches", C.c_int64), ("idle_ms", C.c_double), ("idle_launches", C.c_int64),
                ("search_ms", C.c_double), ("search_upper_ms", C.c_double), ("fit_ms", C.c_double),
                ("reuse_ms", C.c_double), ("search_first_ms", C.c_double), ("search_first_launches", C.c_int64)]


def build(verbose=False):
    """Compile liblimovelo_b200.so in-tree with nvcc for sm_100a (works without a GPU)."""
    cmd = ["make", "-C", _HERE, "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liblimovelo_b200.so is not built (run __graft_entry__.build()); there is no fallback")
    L = C.CDLL(LIB_PATH)
    dp, fp, vp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p
    i64, i32p = C.c_int64, C.POINTER(C.c_int32)
    L.lv_default_params.argtypes = [C.POINTER(Params)]
    L.lv_default_params.restype = None
    L.lv_params_from_yaml.argtypes = [C.c_char_p, C.POINTER(Params)]
    L.lv_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.lv_destroy.argtypes = [vp]
    L.lv_destroy.restype = None
    L.lv_last_error.restype = C.c_char_p
    L.lv_version.restype = C.c_char_p
    L.lv_result_bytes.restype = C.c_int64
    L.lv_map_build.argtypes = [vp, fp, i64]
    L.lv_map_add.argtypes = [vp, fp, i64, C.c_int]
    L.lv_map_size.argtypes = [vp]
    L.lv_map_size.restype = i64
    L.lv_map_exists.argtypes = [vp]
    L.lv_map_points.argtypes = [vp, fp, i64]
    L.lv_map_points.restype = i64
    L.lv_map_build_device.argtypes = [vp, vp, i64]
    L.lv_map_add_device.argtypes = [vp, vp, i64, C.c_int]
    L.lv_map_status.argtypes = [vp]
    L.lv_map_add_sweep_device.argtypes = [vp, vp, i64, C.c_int]
    L.lv_map_add_last_sweep.argtypes = [vp, C.c_int]
    L.lv_measure.argtypes = [vp, dp, fp, i64, dp, dp, C.POINTER(i64)]
    L.lv_measure_reduced.argtypes = [vp, dp, fp, i64, dp, dp, C.POINTER(i64)]
    L.lv_match_all.argtypes = [vp, dp, fp, i64, C.POINTER(C.c_uint8), i32p, fp, fp, fp, fp]
    L.lv_last_neighbours.argtypes = [vp, i64, i32p]
    L.lv_set_state.argtypes = [vp, dp, dp]
    L.lv_get_state.argtypes = [vp, dp, dp]
    L.lv_init_state.argtypes = [vp, fp]
    L.lv_predict.argtypes = [vp, dp, dp, C.c_double]
    L.lv_propagate_device.argtypes = [vp, dp, dp, dp, C.c_int32]
    L.lv_init_state_host.argtypes = [C.POINTER(Params), fp, dp, dp]
    L.lv_predict_host.argtypes = [C.POINTER(Params), dp, dp, C.c_double, dp, dp]
    L.lv_correct.argtypes = [vp, fp, i64, C.c_double, C.POINTER(IterLog), i32p, dp, dp]
    L.lv_correct_device.argtypes = [vp, vp, i64, C.c_double]
    L.lv_last_logs.argtypes = [vp, C.POINTER(IterLog), i32p]
    L.lv_last_time_updated.argtypes = [vp]
    L.lv_last_time_updated.restype = C.c_double
    L.lv_host_alloc.argtypes = [i64]
    L.lv_host_alloc.restype = vp
    L.lv_host_free.argtypes = [vp]
    L.lv_host_free.restype = None
    L.lv_device_alloc.argtypes = [vp, i64]
    L.lv_device_alloc.restype = vp
    L.lv_device_free.argtypes = [vp, vp]
    L.lv_device_free.restype = None
    L.lv_memcpy_h2d.argtypes = [vp, vp, vp, i64]
    L.lv_synchronize.argtypes = [vp]
    L.lv_profile_enable.argtypes = [vp, C.c_int]
    L.lv_profile_get.argtypes = [vp, C.POINTER(Profile), C.c_int]
    L.lv_flush_l2.argtypes = [vp]
    _lib = L
    return L


SYNTH_LIB_PATH = os.path.join(_HERE, "liblv_synth.so")
_synth = None


def synth_lib():
    """liblv_synth.so (include/lv_synth.h): the synthetic reader + the YAML reader, plain C++ — a process that only
    needs inputs (bench.py's CPU reference arm) never loads the CUDA library."""
    global _synth
    if _synth is not None:
        return _synth
    if not os.path.exists(SYNTH_LIB_PATH):
        raise RuntimeError("liblv_synth.so is not built (run __graft_entry__.build())")
    L = C.CDLL(SYNTH_LIB_PATH)
    dp, fp, vp, i64 = C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p, C.c_int64
    L.lv_default_params.argtypes = [C.POINTER(Params)]
    L.lv_default_params.restype = None
    L.lv_params_from_yaml.argtypes = [C.c_char_p, C.POINTER(Params)]
    L.lv_init_state_host.argtypes = [C.POINTER(Params), fp, dp, dp]
    L.lv_synth_world_create.argtypes = [C.c_uint64, i64]
    L.lv_synth_world_create.restype = vp
    L.lv_synth_world_destroy.argtypes = [vp]
    L.lv_synth_world_destroy.restype = None
    L.lv_synth_world_map.argtypes = [vp, fp, i64]
    L.lv_synth_world_map.restype = i64
    L.lv_synth_world_extent.argtypes = [vp]
    L.lv_synth_world_extent.restype = C.c_double
    L.lv_synth_pose.argtypes = [vp, C.c_double, C.POINTER(Params), dp]
    L.lv_synth_pose.restype = None
    L.lv_synth_sweep.argtypes = [vp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.c_uint64, fp]
    L.lv_synth_sweep.restype = i64
    _synth = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_params(_L=None, **over):
    p = Params()
    (_L or lib()).lv_default_params(C.byref(p))
    for k, v in over.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(p, k, v)
    return p


def params_from_yaml(path, _L=None, **over):
    """lv_params from a LIMO-Velo YAML (+ field overrides).  _L: the library providing the reader (default: the product's;
    liblv_synth.so carries the same reader for processes that must not load the CUDA library)"""
    p = default_params(_L)
    st = (_L or lib()).lv_params_from_yaml(str(path).encode(), C.byref(p))
    if st != OK:
        raise RuntimeError("lv_params_from_yaml(%s) -> %s" % (path, STATUS_NAMES[st]))
    for k, v in over.items():
        setattr(p, k, v)
    return p


CONFIG_DIR = os.path.join(_HERE, "config")


def _check(st, allow=()):
    if st != OK and st not in allow:
        raise RuntimeError("%s: %s" % (STATUS_NAMES[st], lib().lv_last_error().decode()))
    return st


def _logs_to_py(logs, n):
    out = []
    for i in range(n):
        lg = logs[i]
        out.append(dict(n_matches=int(lg.n_matches), converged=int(lg.converged), degenerate=int(lg.degenerate),
                        HTH=np.array(lg.HTH[:]).reshape(12, 12), HTh=np.array(lg.HTh[:]),
                        dx=np.array(lg.dx[:]), x_after=np.array(lg.x_after[:])))
    return out


class State32(C.Structure):
    """lv_state32: `State` of the reference (Objects.hpp:97-120), single precision"""
    _fields_ = [("R", C.c_float * 9), ("pos", C.c_float * 3), ("vel", C.c_float * 3), ("bw", C.c_float * 3),
                ("ba", C.c_float * 3), ("g", C.c_float * 3), ("RLI", C.c_float * 9), ("tLI", C.c_float * 3),
                ("a", C.c_float * 3), ("w", C.c_float * 3), ("time", C.c_double)]

    def as_tuple(self):
        return tuple(np.array(getattr(self, k)[:], dtype=np.float32).tobytes() for k, _ in self._fields_[:-1]) + (self.time,)


def _f3(v):
    return (C.c_float * 3)(*[float(t) for t in v])


def state_from_ikfom(params, x, time, a, w):
    """State(const state_ikfom&, double) (State.cpp:40-62)"""
    out = State32()
    lib().lv_state_from_ikfom(C.byref(params), _d(np.ascontiguousarray(x, np.float64)), C.c_double(time), _f3(a), _f3(w),
                              C.byref(out))
    return out


def state_add_imu(state, a, w, time):
    """State::operator+=(IMU) on a copy"""
    out = State32.from_buffer_copy(state)
    lib().lv_state_add_imu(C.byref(out), _f3(a), _f3(w), C.c_double(time))
    return out


def compensator_upsample(states, imu_a, imu_w, imu_t):
    """Compensator::upsample: list of State32 + IMU arrays -> the integrated path (list of State32)"""
    ns, ni = len(states), len(imu_t)
    arr = (State32 * ns)(*states)
    a = np.ascontiguousarray(imu_a, np.float32).reshape(ni, 3)
    w = np.ascontiguousarray(imu_w, np.float32).reshape(ni, 3)
    t = np.ascontiguousarray(imu_t, np.float64)
    cap = ns + ni + 8
    out = (State32 * cap)()
    lib().lv_compensator_upsample.restype = C.c_int32
    n = lib().lv_compensator_upsample(arr, ns, _f(a), _f(w), _d(t), ni, out, cap)
    assert n <= cap
    return [State32.from_buffer_copy(out[i]) for i in range(n)]


def compensator_get_t2(path, t2):
    arr = (State32 * len(path))(*path)
    out = State32()
    lib().lv_compensator_get_t2(arr, len(path), C.c_double(t2), C.byref(out))
    return out


class CloudLayout(C.Structure):
    _fields_ = [("point_step", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("off_time", C.c_int32), ("off_range", C.c_int32)]


LIDAR_TYPES = {"velodyne": 0, "hesai": 1, "ouster": 2, "custom": 3}


def pointcloud2_to_points(lidar, pts, header_stamp_us, stamp_beginning, offset_beginning, full_rotation_time):
    """PointCloudProcessor::msg2points on a numpy structured array (its buffer is the message's `data`)"""
    pts = np.ascontiguousarray(pts)
    f = pts.dtype.fields
    tname = {"velodyne": "time", "ouster": "t"}.get(lidar, "timestamp")
    iname = "reflectivity" if lidar == "ouster" else "intensity"
    lay = CloudLayout(pts.dtype.itemsize, f["x"][1], f["y"][1], f["z"][1], f[iname][1], f[tname][1],
                      f["range"][1] if "range" in f else 0)
    n = len(pts)
    xyz, t = np.zeros((n, 3), np.float32), np.zeros(n, np.float64)
    inten, rng = np.zeros(n, np.float32), np.zeros(n, np.float32)
    _check(lib().lv_pointcloud2_to_points(LIDAR_TYPES[lidar], C.byref(lay), pts.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(n),
                                         C.c_uint64(header_stamp_us), int(stamp_beginning), int(offset_beginning),
                                         C.c_double(full_rotation_time), _f(xyz), _d(t), _f(inten), _f(rng)))
    return xyz, t, inten, rng


def time_sort_indices(t):
    t = np.ascontiguousarray(t, np.float64)
    idx = np.zeros(len(t), np.int32)
    _check(lib().lv_time_sort_indices(_d(t), C.c_int64(len(t)), idx.ctypes.data_as(C.POINTER(C.c_int32))))
    return idx


def init_state_host(params, q_imu=(0, 0, 0, 1)):
    """Localizator::init_IKFoM_state on the host (no GPU needed)."""
    x, P = np.zeros(STATE_LEN), np.zeros((DOF, DOF))
    q = np.ascontiguousarray(q_imu, dtype=np.float32)
    _check(lib().lv_init_state_host(C.byref(params), _f(q), _d(x), _d(P)))
    return x, P


def predict_host(params, x, P, acc, gyro, dt):
    """Localizator::propagate -> esekf::predict on the host (no GPU needed)."""
    x = np.array(x, dtype=np.float64).copy()
    P = np.array(P, dtype=np.float64).reshape(DOF, DOF).copy()
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    gyro = np.ascontiguousarray(gyro, dtype=np.float64)
    _check(lib().lv_predict_host(C.byref(params), _d(acc), _d(gyro), float(dt), _d(x), _d(P)))
    return x, P


class Localizer:
    """One lv_handle: the Localizator + Mapper pair of one sequence on one GPU."""

    def __init__(self, params):
        self.L = lib()
        self.params = params
        self.h = C.c_void_p()
        _check(self.L.lv_create(C.byref(params), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.lv_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # Mapper
    def map_build(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        return _check(self.L.lv_map_build(self.h, _f(xyz), xyz.shape[0]))

    def map_add(self, xyz, downsample=True):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        return _check(self.L.lv_map_add(self.h, _f(xyz), xyz.shape[0], int(downsample)))

    def map_add_device(self, d_xyz, n, downsample=True):
        """Mapper::add from a device buffer; asynchronous (no host round trip)"""
        return _check(self.L.lv_map_add_device(self.h, C.c_void_p(d_xyz), int(n), int(downsample)))

    def propagate_device(self, acc, gyro, dt):
        """Localizator::propagate_to on the device: k IMU samples (k x 3, k x 3, k), one launch, the state stays in HBM"""
        acc = np.ascontiguousarray(acc, np.float64).reshape(-1, 3)
        gyro = np.ascontiguousarray(gyro, np.float64).reshape(-1, 3)
        dt = np.ascontiguousarray(dt, np.float64).reshape(-1)
        assert len(acc) == len(gyro) == len(dt)
        return _check(self.L.lv_propagate_device(self.h, _d(acc), _d(gyro), _d(dt), len(dt)))

    def map_add_sweep_device(self, d_xyz_lidar, n, downsample=True):
        """main.cpp:99-105 on the device: the LiDAR-frame sweep, transformed by the filter's current state, joins the map"""
        return _check(self.L.lv_map_add_sweep_device(self.h, C.c_void_p(d_xyz_lidar), int(n), int(downsample)))

    def map_add_last_sweep(self, downsample=True):
        return _check(self.L.lv_map_add_last_sweep(self.h, int(downsample)))

    def map_status(self):
        """synchronises; raises on LV_ERR_CAPACITY (device map out of table / arena space)"""
        return _check(self.L.lv_map_status(self.h))

    def map_size(self):
        return int(self.L.lv_map_size(self.h))

    def map_points(self):
        n = self.map_size()
        out = np.zeros((max(n, 1), 3), np.float32)
        k = self.L.lv_map_points(self.h, _f(out), n)
        return out[:k]

    # operator boundary
    def measure_reduced(self, x, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, dtype=np.float64)
        HTH, HTh, nm = np.zeros((12, 12)), np.zeros(12), C.c_int64(0)
        st = _check(self.L.lv_measure_reduced(self.h, _d(x), _f(xyz), xyz.shape[0], _d(HTH), _d(HTh), C.byref(nm)),
                    allow=(EMPTY_MAP,))
        return st, HTH, HTh, nm.value

    def measure(self, x, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = xyz.shape[0]
        hx, h, nm = np.zeros(12 * n), np.zeros(n), C.c_int64(0)
        st = _check(self.L.lv_measure(self.h, _d(x), _f(xyz), n, _d(hx), _d(h), C.byref(nm)), allow=(EMPTY_MAP,))
        k = nm.value
        return st, hx[:12 * k].reshape(12, k).T.copy(), h[:k].copy()

    def match_all(self, x, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = xyz.shape[0]
        out = dict(valid=np.zeros(n, np.uint8), nn_idx=np.zeros((n, 5), np.int32), nn_sqd=np.zeros((n, 5), np.float32),
                   plane=np.zeros((n, 4), np.float32), dist=np.zeros(n, np.float32), g=np.zeros((n, 3), np.float32))
        st = _check(self.L.lv_match_all(self.h, _d(x), _f(xyz), n, out["valid"].ctypes.data_as(C.POINTER(C.c_uint8)),
                                        out["nn_idx"].ctypes.data_as(C.POINTER(C.c_int32)), _f(out["nn_sqd"]),
                                        _f(out["plane"]), _f(out["dist"]), _f(out["g"])), allow=(EMPTY_MAP,))
        out["status"] = st
        return out

    def last_neighbours(self, n):
        """map point ids (n x 5) of the neighbours the last evaluation handed to the plane fit"""
        out = np.zeros((n, 5), np.int32)
        _check(self.L.lv_last_neighbours(self.h, C.c_int64(n), out.ctypes.data_as(C.POINTER(C.c_int32))), allow=(EMPTY_MAP,))
        return out

    # Localizator
    def set_state(self, x, P):
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64)
        _check(self.L.lv_set_state(self.h, _d(x), _d(P)))

    def get_state(self):
        x, P = np.zeros(STATE_LEN), np.zeros((DOF, DOF))
        _check(self.L.lv_get_state(self.h, _d(x), _d(P)))
        return x, P

    def init_state(self, q_imu=(0, 0, 0, 1)):
        q = np.ascontiguousarray(q_imu, dtype=np.float32)
        _check(self.L.lv_init_state(self.h, _f(q)))

    def predict(self, acc, gyro, dt):
        acc = np.ascontiguousarray(acc, dtype=np.float64)
        gyro = np.ascontiguousarray(gyro, dtype=np.float64)
        _check(self.L.lv_predict(self.h, _d(acc), _d(gyro), float(dt)))

    def correct(self, xyz, time=0.0, raw_ptr=None, n=None):
        """Localizator::correct.  xyz: numpy (host) array, or raw_ptr/n for a pinned host buffer."""
        if raw_ptr is None:
            xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
            ptr, n = _f(xyz), xyz.shape[0]
        else:
            ptr = C.cast(raw_ptr, C.POINTER(C.c_float))
        logs = (IterLog * MAX_EVALS)()
        ne = C.c_int32(0)
        x, P = np.zeros(STATE_LEN), np.zeros((DOF, DOF))
        st = _check(self.L.lv_correct(self.h, ptr, n, float(time), logs, C.byref(ne), _d(x), _d(P)),
                    allow=(EMPTY_MAP, TOO_FEW_MATCHES))
        return st, x, P, _logs_to_py(logs, ne.value)

    def compensate(self, path, Xt2, xyz, t):
        """Compensator::compensate (deskew) on host buffers: returns the points in the LiDAR frame at t2"""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(t, np.float64)
        out = np.empty_like(xyz)
        arr = (State32 * len(path))(*path)
        _check(self.L.lv_compensate(self.h, arr, len(path), C.byref(Xt2), _f(xyz), _d(t), C.c_int64(xyz.shape[0]), _f(out)))
        return out

    def compensate_device(self, path, Xt2, d_xyz, d_t, n, d_out):
        arr = (State32 * len(path))(*path)
        return _check(self.L.lv_compensate_device(self.h, arr, len(path), C.byref(Xt2), C.c_void_p(d_xyz), C.c_void_p(d_t),
                                                  C.c_int64(n), C.c_void_p(d_out)))

    def temporal_downsample(self, xyz, rate, min_dist):
        """PointCloudProcessor::temporal_downsample: returns (kept points, their indices)"""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        out, idx, n = np.empty_like(xyz), np.empty(xyz.shape[0], np.int32), C.c_int64(0)
        _check(self.L.lv_temporal_downsample(self.h, _f(xyz), C.c_int64(xyz.shape[0]), C.c_int32(rate), C.c_double(min_dist),
                                             _f(out), idx.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)))
        return out[:n.value].copy(), idx[:n.value].copy()

    def voxelgrid_downsample(self, xyz, leaf):
        """Compensator::voxelgrid_downsample (pcl::VoxelGrid centroids)"""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        out, n = np.empty_like(xyz), C.c_int64(0)
        _check(self.L.lv_voxelgrid_downsample(self.h, _f(xyz), C.c_int64(xyz.shape[0]), C.c_float(leaf), _f(out), C.byref(n)))
        return out[:n.value].copy()

    def voxelgrid_downsample_device(self, d_xyz, n, leaf, d_out):
        m = C.c_int64(0)
        _check(self.L.lv_voxelgrid_downsample_device(self.h, C.c_void_p(d_xyz), C.c_int64(n), C.c_float(leaf), C.c_void_p(d_out),
                                                     C.byref(m)))
        return m.value

    def correct_buffers(self):
        """preallocated outputs for correct_raw (benchmarks: keeps numpy / ctypes construction out of the timed call)"""
        return dict(logs=(IterLog * MAX_EVALS)(), ne=C.c_int32(0), x=np.zeros(STATE_LEN), P=np.zeros((DOF, DOF)))

    def correct_raw(self, raw_ptr, n, buf, time=0.0):
        """lv_correct on a host pointer, nothing else: returns the status; unpack `buf` with correct_unpack"""
        return self.L.lv_correct(self.h, C.cast(raw_ptr, C.POINTER(C.c_float)), n, float(time), buf["logs"],
                                 C.byref(buf["ne"]), _d(buf["x"]), _d(buf["P"]))

    @staticmethod
    def correct_unpack(buf):
        return buf["x"].copy(), buf["P"].copy(), _logs_to_py(buf["logs"], buf["ne"].value)

    def correct_device(self, d_ptr, n, time=0.0):
        return _check(self.L.lv_correct_device(self.h, d_ptr, n, float(time)), allow=(EMPTY_MAP,))

    def result_bytes(self):
        return int(self.L.lv_result_bytes())

    def last_logs(self):
        logs = (IterLog * MAX_EVALS)()
        ne = C.c_int32(0)
        st = _check(self.L.lv_last_logs(self.h, logs, C.byref(ne)), allow=(TOO_FEW_MATCHES,))
        return st, _logs_to_py(logs, ne.value)

    # utilities
    def device_alloc(self, nbytes):
        p = self.L.lv_device_alloc(self.h, nbytes)
        if not p:
            raise RuntimeError("lv_device_alloc failed")
        return p

    def device_free(self, p):
        self.L.lv_device_free(self.h, p)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.device_alloc(arr.nbytes)
        _check(self.L.lv_memcpy_h2d(self.h, p, arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        return p

    def synchronize(self):
        _check(self.L.lv_synchronize(self.h))

    def profile_enable(self, on=True):
        _check(self.L.lv_profile_enable(self.h, int(on)))

    def profile(self, reset=True):
        p = Profile()
        _check(self.L.lv_profile_get(self.h, C.byref(p), int(reset)))
        return {k: getattr(p, k) for k, _ in Profile._fields_}

    def flush_l2(self):
        _check(self.L.lv_flush_l2(self.h))


class PinnedBuffer:
    """Pinned host memory (lv_host_alloc) viewed as a numpy float32 array."""

    def __init__(self, shape):
        self.L = lib()
        n = int(np.prod(shape))
        self.ptr = self.L.lv_host_alloc(n * 4)
        if not self.ptr:
            raise RuntimeError("lv_host_alloc failed (no CUDA device?)")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_float)), shape=(n,)).reshape(shape)

    def free(self):
        if self.ptr:
            self.L.lv_host_free(self.ptr)
            self.ptr = None


class SynthWorld:
    """Seeded synthetic world + LiDAR ray caster (liblv_synth.so, limo-velo_b200/host/lv_synth.cpp)."""

    def __init__(self, seed, m):
        self.L = synth_lib()
        self.w = self.L.lv_synth_world_create(int(seed), int(m))
        if not self.w:
            raise RuntimeError("lv_synth_world_create failed")
        self.m = int(m)

    def __del__(self):
        try:
            if self.w:
                self.L.lv_synth_world_destroy(self.w)
                self.w = None
        except Exception:
            pass

    def map(self):
        out = np.zeros((self.m, 3), np.float32)
        n = self.L.lv_synth_world_map(self.w, _f(out), self.m)
        return out[:n]

    def extent(self):
        return float(self.L.lv_synth_world_extent(self.w))

    def pose(self, s, params):
        x = np.zeros(STATE_LEN)
        self.L.lv_synth_pose(self.w, float(s), C.byref(params), _d(x))
        return x

    def sweep(self, x, rings=64, azimuths=1024, elev=(-24.8, 2.0), min_dist=4.0, range_sigma=0.02, seed=0, out=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if out is None:
            out = np.zeros((rings * azimuths, 3), np.float32)
        n = self.L.lv_synth_sweep(self.w, _d(x), rings, azimuths, elev[0], elev[1], min_dist, range_sigma, int(seed),
                                  _f(out))
        assert n == rings * azimuths
        return out
