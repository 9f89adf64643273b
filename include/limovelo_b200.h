/*
 * limovelo_b200.h — C ABI of liblimovelo_b200.so
 *
 * B200-native (sm_100a) implementation of LIMO-Velo's per-sweep localization hot path:
 * Localizator::correct -> IKFoM update_iterated_dyn_share_modified -> per iteration
 * Mapper::match (exact 5-NN, plane fit, gates) -> Localizator::calculate_H -> HtH / Hth ->
 * 23-DoF state update.  Plain pointers and sizes only; no torch / Eigen / ROS types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * LIMO-Velo source tree; esekfom.hpp = include/IKFoM/IKFoM_toolkit/esekfom/esekfom.hpp).
 * INTEGRATION.md shows the reference-side glue (the h_share_model / Localizator / Mapper
 * call sites that would bind these).
 *
 * Threading: like the reference (single main thread, main.cpp:52-130) a handle is not
 * re-entrant.  Handles are independent of each other: one handle = one sequence = one
 * CUDA device + stream.
 *
 * There is NO CPU fallback: every compute entry point returns LV_ERR_CUDA when no CUDA
 * device is usable.
 */
#ifndef LIMOVELO_B200_H_
#define LIMOVELO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LV_STATE_LEN 26   /* flat state_ikfom, see lv_state_layout below                    */
#define LV_DOF 23         /* state_ikfom::DOF (use-ikfom.hpp:12-21)                         */
#define LV_MEAS_COLS 12   /* non-zero Jacobian columns (Localizator.cpp:31, esekfom.hpp:1647) */
#define LV_MAX_EVALS 8    /* capacity of the per-update log: MAX_NUM_ITERS + 1 <= 8          */

typedef struct lv_context* lv_handle;

typedef enum lv_status {
    LV_OK = 0,
    LV_EMPTY_MAP = 1,          /* Localizator.cpp:24 / Mapper.cpp:42 silent returns           */
    LV_TOO_FEW_MATCHES = 2,    /* Nm < 23: esekfom.hpp:1701-1709 branch (undefined in the ref) */
    LV_ERR_ARG = 3,
    LV_ERR_CUDA = 4,           /* no device / CUDA failure (see lv_last_error)                */
    LV_ERR_CAPACITY = 5,       /* more points than the handle was created for                 */
    LV_ERR_IO = 6              /* YAML file unreadable                                        */
} lv_status;

/*
 * Flat state layout (26 doubles), field order of state_ikfom (use-ikfom.hpp:12-21),
 * quaternions in Eigen coeffs() order (x, y, z, w):
 *   [0:3) pos   [3:7) rot   [7:11) offset_R_L_I   [11:14) offset_T_L_I
 *   [14:17) vel [17:20) bg  [20:23) ba            [23:26) grav (S2, |grav| = 9.809)
 * Covariance P is 23 x 23 row-major in DOF order pos0 rot3 offR6 offT9 vel12 bg15 ba18 grav21.
 */

/* Configuration.  Upper-case / ROS-style names are the YAML keys read by fill_config
 * (src/main.cpp:135-176) with the same defaults; the last block is device tuning that has
 * no counterpart in the reference.                                                         */
typedef struct lv_params {
    /* Localizator */
    int32_t MAX_NUM_ITERS;            /* main.cpp:144, default 3                           */
    int32_t NUM_MATCH_POINTS;         /* main.cpp:146, must be 5                           */
    int32_t estimate_extrinsics;      /* main.cpp:139                                      */
    int32_t print_degeneracy_values;  /* main.cpp:156 (eigenvalues are returned in the log) */
    double MAX_DIST_PLANE;            /* main.cpp:148, default 2.0                         */
    float PLANES_THRESHOLD;           /* main.cpp:149, default 0.1                         */
    float pad0_;
    double LiDAR_noise;               /* main.cpp:152 -> R                                 */
    double degeneracy_threshold;      /* main.cpp:155 -> D                                 */
    double LIMITS[LV_DOF];            /* main.cpp:145, default 23 x 0.001                  */
    /* IMU process noise (Localizator.cpp:164-168) */
    double covariance_gyroscope, covariance_acceleration;
    double covariance_bias_gyroscope, covariance_bias_acceleration;
    /* extrinsics / gravity (Localizator.cpp:135-153) */
    float initial_gravity[3];
    float I_Translation_L[3];
    float I_Rotation_L[9];            /* row-major YAML list                               */
    float map_downsample_size;        /* ikd-Tree box_length, Mapper.cpp:65 = 0.2          */
    /* ---- device tuning (not in the reference) ---- */
    float voxel_size;                 /* edge of the hashed search voxels [m], default 0.4; rounded to a whole number
                                       * (1..3) of map_downsample_size cells: a voxel is k x k x k downsample cells */
    int32_t device;                   /* CUDA device ordinal                               */
    int32_t sort_queries;             /* 1: bin the sweep by home voxel once per update and search from shared memory
                                       * (lv_search_staged_kernel); 0: per-query search from global memory       */
    int64_t max_map_points;           /* capacity of the device map                        */
    int64_t max_points;               /* capacity of one sweep                             */
    void* stream;                     /* cudaStream_t to run on; NULL = own stream         */
} lv_params;

/* One record per h-evaluation of update_iterated_dyn_share_modified (esekfom.hpp:1634-1822) */
typedef struct lv_iter_log {
    int64_t n_matches;        /* Nm = rows of h_x (esekfom.hpp:1650)                       */
    int32_t converged;        /* dyn_share.converge after esekfom.hpp:1748-1762            */
    int32_t degenerate;       /* 1 if any eigenvalue of HTH[0:6,0:6] < D (esekfom.hpp:1740) */
    double HTH[144];          /* row-major 12 x 12 (esekfom.hpp:1723)                      */
    double HTh[12];           /* h_x^T h (esekfom.hpp:1727)                                */
    double dx[LV_DOF];        /* dx_ before the degeneracy mask (esekfom.hpp:1733)         */
    double x_after[LV_STATE_LEN];
} lv_iter_log;

typedef struct lv_profile {
    double measure_ms;        /* sum of device time of the measurement kernels (search + fit) */
    double solve_ms;          /* sum of device time of the IESKF step kernel               */
    double build_ms;          /* sum of device time of map (re)builds                      */
    int64_t measure_launches, solve_launches, build_launches;
    int64_t total_launches;   /* every kernel launched by this handle since the last reset */
    double idle_ms;           /* launches that found the update already finished (early exit) */
    int64_t idle_launches;
    /* measure_ms split by kernel (each launched once per h-evaluation) */
    double search_ms;         /* lv_search_kernel: exact 5-NN at level 0                   */
    double search_upper_ms;   /* lv_search_rings_kernel: the queries level 0 cannot certify */
    double fit_ms;            /* lv_fit_kernel: plane fit, Jacobian rows, normal equations  */
    double reuse_ms;          /* lv_reuse_kernel: neighbours carried over from the previous evaluation */
    double search_first_ms;   /* the part of search_ms spent in FIRST evaluations: those launches search every query of the
                               * sweep (later ones only the queries the reuse test handed back)                  */
    int64_t search_first_launches;
} lv_profile;

/* ------------------------------------------------------------------------------------------ */
/* configuration                                                                              */
void lv_default_params(lv_params* p);                         /* defaults of main.cpp:137-175 */
lv_status lv_params_from_yaml(const char* path, lv_params* p); /* config/<name>.yaml keys           */

/* lifetime: replaces the Localizator / Mapper singletons (Localizator.cpp:100-103,
 * Mapper.hpp:35-38): one handle owns one filter and one map.                                */
lv_status lv_create(const lv_params* p, lv_handle* out);
void lv_destroy(lv_handle h);
const char* lv_last_error(void);
const char* lv_version(void);
/* bytes one lv_correct() reads back from the device (state, covariance, logs): the D2H side of its e2e cost */
int64_t lv_result_bytes(void);

/* ---- Mapper boundary (include/Headers/Mapper.hpp:17-23) ---------------------------------- */
/* Mapper::add on an empty map == KD_TREE::Build, no downsampling (Mapper.cpp:26,68-71).      */
lv_status lv_map_build(lv_handle h, const float* xyz, int64_t m);
/* Mapper::add on an existing map == KD_TREE::Add_Points (Mapper.cpp:73-76,
 * ikd_Tree.cpp:478-573): 0.2 m voxel rule when downsample != 0.                              */
lv_status lv_map_add(lv_handle h, const float* xyz, int64_t n, int downsample);
/* Both are ASYNCHRONOUS on the handle's stream (no host round trip: every count lives on the device; the host
 * buffer may be reused when the call returns only if it is pinned — pageable memory is staged by the runtime).
 * Running out of table / arena space is flagged on the device and reported as LV_ERR_CAPACITY by the next call
 * that synchronises anyway (lv_map_size, lv_map_status, lv_correct, lv_last_logs).                              */
int64_t lv_map_size(lv_handle h);                             /* Mapper::size  (Mapper.cpp:32); synchronises */
int lv_map_exists(lv_handle h);                               /* Mapper::exists (Mapper.cpp:36) */
lv_status lv_map_status(lv_handle h);                         /* synchronises; LV_OK or LV_ERR_CAPACITY       */
/* KD_TREE::flatten: the map's points in insertion order; returns their number; synchronises  */
int64_t lv_map_points(lv_handle h, float* xyz_out, int64_t cap);
/* same as lv_map_build / lv_map_add but xyz is a DEVICE pointer (data already resident in HBM, e.g. the deskewed
 * sweep transformed by the update's own result): correct -> add never leaves the GPU.  The buffer must stay valid
 * until the stream has passed the call.                                                                        */
lv_status lv_map_build_device(lv_handle h, const float* d_xyz, int64_t m);
lv_status lv_map_add_device(lv_handle h, const float* d_xyz, int64_t n, int downsample);
/* main.cpp:99-105 on the device: `map.add(Xt2 * Xt2.I_Rt_L() * ds_compensated, t2, true)`.  The sweep is given in the
 * LiDAR frame (device memory) and is transformed by the filter's current state, i.e. the result of the lv_correct /
 * lv_correct_device that precedes the call, without that state ever visiting the host.  lv_map_add_last_sweep reuses
 * the sweep of that update (it is still resident), so one tick is lv_correct(...); lv_map_add_last_sweep(h, 1).
 * On an empty map both build it un-downsampled (Mapper.cpp:26).  Asynchronous.                                  */
lv_status lv_map_add_sweep_device(lv_handle h, const float* d_xyz_lidar, int64_t n, int downsample);
lv_status lv_map_add_last_sweep(lv_handle h, int downsample);

/* ---- operator boundary: the measurement model IKFoM calls (esekfom.hpp:128,1637) --------- */
/* compat mode: fills h_x (Nm x 12, COLUMN-major like Eigen::MatrixXd) and h (Nm) exactly as
 * IKFoM::h_share_model does (use-ikfom.cpp:16-37); rows in input-point order.                */
lv_status lv_measure(lv_handle h, const double* x /*LV_STATE_LEN*/, const float* xyz_lidar, int64_t n,
                     double* h_x, double* h_vec, int64_t* nm);
/* fast mode: what update_iterated_dyn_share_modified consumes when Nm >= 23
 * (esekfom.hpp:1722-1729): HTH = h_x^T h_x (row-major 12x12), HTh = h_x^T h.                 */
lv_status lv_measure_reduced(lv_handle h, const double* x, const float* xyz_lidar, int64_t n,
                             double* HTH, double* HTh, int64_t* nm);
/* Mapper::match without the compaction (Mapper.cpp:40-56): per input point the world point,
 * the 5 neighbours (indices into the current map order of lv_map_points, squared distances
 * ascending), the fitted plane (A,B,C,D), the point-to-plane distance and is_chosen().
 * Any output pointer may be NULL.  Neighbour data are defined for points whose 5th
 * neighbour is closer than MAX_DIST_PLANE (others report idx -1).                            */
lv_status lv_match_all(lv_handle h, const double* x, const float* xyz_lidar, int64_t n,
                       uint8_t* valid, int32_t* nn_idx, float* nn_sqd, float* plane, float* dist,
                       float* g_world);
/* The neighbours the LAST evaluation of the last lv_correct (or operator call) handed to the plane fit: nn_idx (n x 5,
 * indices as in lv_match_all, -1 where a query had fewer than five).  lv_match_all always searches afresh; inside an
 * update, evaluations after the first keep the stored five wherever the exact search provably returns them again
 * (the reference searches every time, Mapper.cpp:40-56) — this is the window on what the update really used.  */
lv_status lv_last_neighbours(lv_handle h, int64_t n, int32_t* nn_idx);

/* ---- module boundary: Localizator (include/Headers/Localizator.hpp:24-33) ---------------- */
lv_status lv_set_state(lv_handle h, const double* x, const double* P);   /* change_x / change_P */
lv_status lv_get_state(lv_handle h, double* x, double* P);               /* get_x / get_P       */
/* Localizator::init_IKFoM_state (Localizator.cpp:135-153); q_imu = (x,y,z,w)                 */
lv_status lv_init_state(lv_handle h, const float q_imu[4]);
/* Localizator::propagate -> esekf::predict (Localizator.cpp:159-173, esekfom.hpp:279-384)    */
lv_status lv_predict(lv_handle h, const double acc[3], const double gyro[3], double dt);
/* Localizator::propagate_to (Localizator.cpp:59-75) on the DEVICE: the k IMU samples between two sweeps (acc, gyro: k x 3,
 * dt: k) run through esekf::predict in one launch; the filter state stays in HBM, so update -> propagate -> update needs no
 * host round trip (lv_get_state fetches it when the caller wants to publish it).  Same arithmetic as lv_predict.        */
lv_status lv_propagate_device(lv_handle h, const double* acc, const double* gyro, const double* dt, int32_t k);
/* the same two host-side steps on caller-owned (x, P), no handle and no GPU involved           */
lv_status lv_init_state_host(const lv_params* p, const float q_imu[4], double* x, double* P);
lv_status lv_predict_host(const lv_params* p, const double acc[3], const double gyro[3], double dt,
                          double* x, double* P);
/* Localizator::correct(points, time) (Localizator.cpp:23-27): the whole iterated update.
 * xyz_lidar: HOST buffer of n deskewed points in the LiDAR frame.  logs (capacity
 * LV_MAX_EVALS) / n_evals / x_out / P_out may be NULL.                                       */
lv_status lv_correct(lv_handle h, const float* xyz_lidar, int64_t n, double time,
                     lv_iter_log* logs, int32_t* n_evals, double* x_out, double* P_out);
/* same, but the points are already resident in HBM (DEVICE pointer) and nothing is copied
 * back except through lv_get_state / lv_last_logs.                                           */
lv_status lv_correct_device(lv_handle h, const float* d_xyz_lidar, int64_t n, double time);
lv_status lv_last_logs(lv_handle h, lv_iter_log* logs, int32_t* n_evals);
double lv_last_time_updated(lv_handle h);                     /* Localizator::last_time_updated */

/* ---- utilities --------------------------------------------------------------------------- */
void* lv_host_alloc(int64_t bytes);                           /* pinned host memory            */
void lv_host_free(void* p);
void* lv_device_alloc(lv_handle h, int64_t bytes);
void lv_device_free(lv_handle h, void* p);
lv_status lv_memcpy_h2d(lv_handle h, void* dst, const void* src, int64_t bytes);
lv_status lv_synchronize(lv_handle h);
lv_status lv_profile_enable(lv_handle h, int on);             /* CUDA-event timing per kernel  */
lv_status lv_profile_get(lv_handle h, lv_profile* out, int reset);
lv_status lv_flush_l2(lv_handle h);                           /* writes a 256 MiB scratch      */

/* ---- Compensator boundary: deskew (include/Headers/Compensator.hpp, src/Modules/Compensator.cpp) ---- */
/* `State` of include/Headers/Objects.hpp:97-120 (single precision like the reference), row-major matrices */
typedef struct lv_state32 {
    float R[9];
    float pos[3], vel[3], bw[3], ba[3], g[3];
    float RLI[9], tLI[3];
    float a[3], w[3];          /* last controls (State.cpp:129-131) */
    double time;
} lv_state32;
/* State(const state_ikfom&, double) (State.cpp:40-62): x = flat state (layout above); a, w = the IMU sample that
 * follows `time` (Accumulator::get_next_imu); g = p->initial_gravity (State.cpp:21)                    */
void lv_state_from_ikfom(const lv_params* p, const double* x, double time, const float a[3], const float w[3],
                         lv_state32* out);
/* State::operator+=(const IMU&) (State.cpp:73-75 -> update :122-132 -> propagate_f :103-120), host */
void lv_state_add_imu(lv_state32* s, const float a[3], const float w[3], double time);
/* Compensator::upsample (Compensator.cpp:73-113): states (before t1 .. t2) + IMU samples (a, w: ni x 3, t: ni) ->
 * the integrated path; returns the number of states the path has (writes at most cap)                   */
int32_t lv_compensator_upsample(const lv_state32* states, int32_t ns, const float* imu_a, const float* imu_w,
                                const double* imu_t, int32_t ni, lv_state32* out, int32_t cap);
/* Compensator::get_t2 (Compensator.cpp:55-63) */
void lv_compensator_get_t2(const lv_state32* path, int32_t ns, double t2, lv_state32* out);
/* Compensator::compensate(states, Xt2, points) (Compensator.cpp:123-146): every point is carried from the pose at
 * its own timestamp to the LiDAR frame at t2.  xyz (n x 3) and t (n, ascending, inside [path[0].time,
 * path[ns-1].time]) are HOST buffers; xyz_out (n x 3, host) receives the deskewed points in input order, which is
 * what Localizator::correct takes.  LV_ERR_ARG if a timestamp lies outside the path (the reference asserts).   */
lv_status lv_compensate(lv_handle h, const lv_state32* path, int32_t ns, const lv_state32* Xt2, const float* xyz,
                        const double* t, int64_t n, float* xyz_out);
/* same on device-resident buffers (lv_device_alloc); d_xyz_out may alias d_xyz                              */
lv_status lv_compensate_device(lv_handle h, const lv_state32* path, int32_t ns, const lv_state32* Xt2,
                               const float* d_xyz, const double* d_t, int64_t n, float* d_xyz_out);

/* ---- downsamplers in front of the path -------------------------------------------- */
/* PointCloudProcessor::downsample -> temporal_downsample (src/Utils/PointCloudProcessor.cpp:18-21,101-112): keeps
 * point i iff (downsample_rate <= 1 or (i + 1) % downsample_rate == 0) and min_dist < |p|, order preserved.
 * HOST buffers; xyz_out (n x 3) and idx_out (n, may be NULL: indices of the kept points, for the caller's other
 * per-point fields) receive *n_out entries.                                                             */
lv_status lv_temporal_downsample(lv_handle h, const float* xyz, int64_t n, int32_t downsample_rate, double min_dist,
                                 float* xyz_out, int32_t* idx_out, int64_t* n_out);
/* Compensator::downsample -> voxelgrid_downsample (src/Modules/Compensator.cpp:115-118,148-163): pcl::VoxelGrid with
 * leaf downsample_prec: one centroid per occupied leaf, leaves in ascending PCL cell index.  xyz_out needs room for
 * n points.  When the leaf is too small for the extent (cell index overflow) PCL warns and returns the input: so does
 * this (n_out = n, LV_OK, the warning in lv_last_error()).                                                          */
lv_status lv_voxelgrid_downsample(lv_handle h, const float* xyz, int64_t n, float downsample_prec, float* xyz_out,
                                  int64_t* n_out);
/* the same on device-resident buffers (d_xyz_out must not alias d_xyz)                                    */
lv_status lv_voxelgrid_downsample_device(lv_handle h, const float* d_xyz, int64_t n, float downsample_prec,
                                         float* d_xyz_out, int64_t* n_out);

/* ---- wire format of the LiDAR message (last item) --------------------------------------
 * PointCloudProcessor::msg2points (src/Utils/PointCloudProcessor.cpp:24-99) on the raw `data` of a
 * sensor_msgs/PointCloud2: the point structs of include/Headers/Common.hpp:109-221 are read field by field at the
 * byte offsets the message declares (pcl::fromROSMsg maps fields by name), so no ROS / PCL type is needed here. */
typedef enum lv_lidar_type { LV_LIDAR_VELODYNE = 0, LV_LIDAR_HESAI = 1, LV_LIDAR_OUSTER = 2, LV_LIDAR_CUSTOM = 3 } lv_lidar_type;
typedef struct lv_cloud_layout {
    int32_t point_step;        /* bytes per point */
    int32_t off_x, off_y, off_z;           /* float32 */
    int32_t off_intensity;     /* velodyne / custom: float32 `intensity`; hesai: uint8 `intensity`; ouster: uint16 `reflectivity` */
    int32_t off_time;          /* velodyne: float32 `time`; hesai / custom: float64 `timestamp`; ouster: uint32 `t` (ns) */
    int32_t off_range;         /* ouster: uint32 `range`; others: unused (range = |p|, Point.cpp:166-169) */
} lv_cloud_layout;
/* the time semantics of the YAML (Common.hpp:56-107): stamp_beginning, offset_beginning, full_rotation_time.
 * header_stamp_us = pcl header stamp (microseconds).  Outputs (n entries each; intensity / range may be NULL):
 * xyz, absolute time per point (Point.cpp:37-111 + get_begin_time, PointCloudProcessor.cpp:42-88), intensity, range. */
lv_status lv_pointcloud2_to_points(lv_lidar_type type, const lv_cloud_layout* layout, const uint8_t* data, int64_t n,
                                   uint64_t header_stamp_us, int stamp_beginning, int offset_beginning,
                                   double full_rotation_time, float* xyz, double* time, float* intensity, float* range);
/* PointCloudProcessor::sort_points (PointCloudProcessor.cpp:114-123): order of the points by time (stable; the
 * reference's std::sort leaves equal stamps in unspecified order)                                          */
lv_status lv_time_sort_indices(const double* time, int64_t n, int32_t* idx_out);

/* lv_pointcloud2_to_points with the bounds pcl::fromROSMsg takes from the message: every field used lies inside point_step
 * and n points fit data_bytes (LV_ERR_ARG otherwise)                                                                    */
lv_status lv_pointcloud2_to_points_checked(lv_lidar_type type, const lv_cloud_layout* layout, const uint8_t* data, int64_t data_bytes,
                                           int64_t n, uint64_t header_stamp_us, int stamp_beginning, int offset_beginning,
                                           double full_rotation_time, float* xyz, double* time, float* intensity, float* range);

/* ---- rosbag reader (format 2.0; replaces roscpp / `rosbag play` in front of Accumulator::receive_lidar / receive_imu,
 * src/main.cpp:27-39, src/Modules/Accumulator.cpp:39-60).  Sequential read of uncompressed bags (bz2 / lz4 chunks are
 * refused with LV_ERR_IO: `rosbag decompress` first).  Host code.                                                       */
typedef struct lv_bag lv_bag;
typedef struct lv_bag_message {
    int32_t conn;              /* connection id                                                   */
    const char* topic;         /* e.g. "/velodyne_points"; valid while the bag is open            */
    const char* type;          /* e.g. "sensor_msgs/PointCloud2"                                  */
    uint32_t sec, nsec;        /* receive time of the record                                      */
    const uint8_t* data;       /* the serialised message (points into the bag's buffer)           */
    int64_t size;
} lv_bag_message;
lv_status lv_bag_open(const char* path, lv_bag** out);
void lv_bag_close(lv_bag* bag);
void lv_bag_rewind(lv_bag* bag);
int32_t lv_bag_connection_count(const lv_bag* bag);
/* next message in file order: 1 = *msg filled, 0 = end of the bag, -1 = malformed */
int lv_bag_next(lv_bag* bag, lv_bag_message* msg);
/* sensor_msgs/PointCloud2 -> a view of the point bytes + the offsets of the fields `type`'s point struct uses
 * (include/Headers/Common.hpp:109-221), ready for lv_pointcloud2_to_points_checked                                      */
typedef struct lv_pointcloud2_view {
    uint32_t stamp_sec, stamp_nsec;      /* header.stamp                                          */
    uint32_t height, width, point_step, row_step;
    int32_t is_bigendian, is_dense;
    int64_t n_points;                    /* height x width                                        */
    const uint8_t* data;
    int64_t data_bytes;
    lv_cloud_layout layout;
} lv_pointcloud2_view;
lv_status lv_bag_parse_pointcloud2(const uint8_t* msg, int64_t size, lv_lidar_type type, lv_pointcloud2_view* out);
/* sensor_msgs/Imu (what Accumulator::receive_imu reads: IMU.cpp:18-40) */
typedef struct lv_imu_sample {
    uint32_t stamp_sec, stamp_nsec;
    double orientation[4];               /* x y z w                                                */
    double angular_velocity[3];
    double linear_acceleration[3];
} lv_imu_sample;
lv_status lv_bag_parse_imu(const uint8_t* msg, int64_t size, lv_imu_sample* out);

#ifdef __cplusplus
}
#endif
#endif /* LIMOVELO_B200_H_ */
