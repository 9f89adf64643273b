/*
 * lv_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of LIMO-Velo's per-sweep
 * localization hot path (Localizator::correct and everything below it).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product library (liblimovelo_b200.so) never links,
 * loads or calls anything in oracle/.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or KATs for this
 * path and its Eigen/Boost/PCL/ROS dependencies are absent
 * from this image, so only the ikd-Tree part of the reference can be compiled
 * here (oracle/_ref, see oracle/Makefile).  The kNN of this oracle IS pinned
 * against that verbatim ikd-Tree build (tests/test_oracle_knn.py).  The plane
 * fit (Eigen colPivHouseholderQr), the 23x23 inverses (Eigen PartialPivLU) and
 * the 6x6 EigenSolver are restated from Eigen 3.3's published algorithms with
 * sequential summation order and are therefore "parity unpinned" at the level
 * of fp rounding (agreement to a few ulp is expected, bit equality is not
 * provable here).
 *
 * All reference citations are relative to /root/reference.
 */
#ifndef LV_ORACLE_H_
#define LV_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flat state layout, 26 doubles (use-ikfom.hpp:12-21; quaternions in Eigen
 * coeffs() order x,y,z,w):
 *   [0:3) pos  [3:7) rot  [7:11) offset_R_L_I  [11:14) offset_T_L_I
 *   [14:17) vel  [17:20) bg  [20:23) ba  [23:26) grav                       */
#define LVO_STATE_LEN 26
#define LVO_DOF 23

typedef struct lvo_params {
    int32_t max_num_iters;        /* MAX_NUM_ITERS (main.cpp:144)            */
    int32_t estimate_extrinsics;  /* main.cpp:139                            */
    double max_dist_plane;        /* MAX_DIST_PLANE (main.cpp:148)           */
    float planes_threshold;       /* PLANES_THRESHOLD (main.cpp:149)         */
    float pad_;
    double lidar_noise;           /* LiDAR_noise -> R (Localizator.cpp:132)  */
    double degeneracy_threshold;  /* D (Localizator.cpp:132)                 */
    double limits[LVO_DOF];       /* LIMITS (main.cpp:145)                   */
} lvo_params;

/* per h-evaluation record of update_iterated_dyn_share_modified */
typedef struct lvo_iter_log {
    int64_t n_matches;            /* Nm                                      */
    int32_t converged;            /* dyn_share.converge after the test       */
    int32_t pad_;
    double HTH[144];              /* row-major 12x12                         */
    double HTh[12];
    double dx[LVO_DOF];           /* dx_ (pre degeneracy mask)               */
    double x_after[LVO_STATE_LEN];
} lvo_iter_log;

enum { LVO_OK = 0, LVO_EMPTY_MAP = 1, LVO_TOO_FEW_MATCHES = 2, LVO_BAD_ARG = 3 };
enum { LVO_KNN_BRUTE = 0, LVO_KNN_KDTREE = 1, LVO_KNN_REF_IKDTREE = 2 };

typedef struct lvo_map lvo_map;

/* Mapper (Mapper.cpp:22-30, ikd_Tree.cpp:409-423,478-573) */
lvo_map* lvo_map_create(int backend);
void lvo_map_destroy(lvo_map*);
int lvo_map_build(lvo_map*, const float* xyz, int64_t m);            /* Build, no downsample */
int lvo_map_add(lvo_map*, const float* xyz, int64_t n, int downsample); /* Add_Points rule   */
int64_t lvo_map_size(const lvo_map*);
int64_t lvo_map_points(const lvo_map*, float* xyz_out, int64_t cap);  /* flatten            */
/* exact k-NN (ikd_Tree.cpp:426-461): ascending; returns number found */
int lvo_knn(const lvo_map*, const float g[3], int k, int32_t* idx, float* sqd, float* nn_xyz);

/* Mapper::match for every point, without the compaction (Mapper.cpp:40-56).
 * Outputs may be NULL.  valid[i] = Match::is_chosen().                    */
int lvo_match_all(const lvo_map*, const double* x, const lvo_params*, const float* xyz_lidar,
                  int64_t n, uint8_t* valid, int32_t* nn_idx /*n*5*/, float* nn_sqd /*n*5*/,
                  float* plane /*n*4*/, float* dist /*n*/, float* g_world /*n*3*/);

/* h_share_model (use-ikfom.cpp:16-37): h_x is Nm x 12 COLUMN-major (Eigen MatrixXd), h is Nm */
int lvo_measure(const lvo_map*, const double* x, const lvo_params*, const float* xyz_lidar,
                int64_t n, double* h_x, double* h, int64_t* nm);
/* the reduction IKFoM consumes (esekfom.hpp:1723,1727) */
int lvo_measure_reduced(const lvo_map*, const double* x, const lvo_params*, const float* xyz_lidar,
                        int64_t n, double* HTH /*144*/, double* HTh /*12*/, int64_t* nm);

/* update_iterated_dyn_share_modified (esekfom.hpp:1620-1823).  x, P in/out. */
int lvo_update_iterated(const lvo_map*, double* x, double* P /*23x23 row-major*/,
                        const lvo_params*, const float* xyz_lidar, int64_t n,
                        lvo_iter_log* logs /*cap max_num_iters+1*/, int32_t* n_evals);
/* same loop, but the measurement (HTH,HTh,Nm per evaluation) is supplied by the caller:
 * used to test the 23x23 algebra in isolation.                              */
int lvo_update_step(const double* x_prop, const double* P_prop, const double* x_cur,
                    const lvo_params*, const double* HTH, const double* HTh,
                    double* dx_out /*23 pre-mask*/, double* x_new, double* P_now /*P_ after the J blocks*/,
                    double* Kx_out /*23x12 row-major*/, int32_t* converged);
int lvo_update_finish(const double* x_prop, const double* x_new, const double* dx,
                      const double* P_now, const double* Kx, double* P_out);

/* esekf::predict with Localizator::propagate's Q (esekfom.hpp:279-384, Localizator.cpp:159-173) */
int lvo_predict(double* x, double* P, const double acc[3], const double gyro[3], double dt,
                double cov_gyro, double cov_acc, double cov_bias_gyro, double cov_bias_acc);

/* init (Localizator.cpp:135-153): q_imu is (x,y,z,w) */
int lvo_init_state(double* x, double* P, const float q_imu[4], const float initial_gravity[3],
                   const float I_Rotation_L[9], const float I_Translation_L[3]);

/* manifold primitives exported for tests */
void lvo_boxplus(double* x, const double* d23);
void lvo_boxminus(const double* x, const double* y, double* d23); /* x [-] y */
void lvo_quat_to_rot(const double q[4], double R[9]);
void lvo_plane_fit(const float* pts5 /*5x3*/, float threshold, float abcd[4], int* is_plane);
void lvo_inverse(const double* A, int n, double* Ainv);
void lvo_sym_eig6(const double* A, double* evals, double* evecs /*row-major, columns = vectors*/);

/* timing helper for bench.py: seconds spent inside kNN / rest during the last lvo_update_iterated */
void lvo_last_timing(double* knn_s, double* total_s);
/* OpenMP team size of the match loop (MP_PROC_NUM, CMakeLists.txt:19-36); default 1 */
void lvo_set_threads(int n);

/* ---- deskew: Compensator::compensate and its State arithmetic -------------------------
 * `State` of include/Headers/Objects.hpp:97-120, single precision like the reference.                        */
typedef struct lvo_state32 {
    float R[9];            /* row-major */
    float pos[3], vel[3], bw[3], ba[3], g[3];
    float RLI[9], tLI[3];
    float a[3], w[3];      /* last controls */
    double time;
} lvo_state32;
/* State(const state_ikfom&, double) (State.cpp:51-62); a, w = the IMU sample following `time` (State.cpp:45-48) */
void lvo_state_from_ikfom(const double* x26, double time, const float a[3], const float w[3],
                          const float initial_gravity[3], lvo_state32* out);
/* State::operator+=(IMU) = update (State.cpp:122-132) -> propagate_f (State.cpp:103-120) */
void lvo_state_add_imu(lvo_state32* s, const float a[3], const float w[3], double time);
/* Compensator::upsample (Compensator.cpp:73-113): returns the number of states written (<= cap) */
int lvo_upsample(const lvo_state32* states, int ns, const float* imu_a, const float* imu_w, const double* imu_t, int ni,
                 lvo_state32* out, int cap);
/* Compensator::get_t2 (Compensator.cpp:55-63) */
void lvo_get_t2(const lvo_state32* states, int ns, double t2, lvo_state32* out);
/* Compensator::compensate (Compensator.cpp:123-146): time-sorted points; returns the number of points written */
int64_t lvo_compensate(const lvo_state32* states, int ns, const lvo_state32* Xt2, const float* xyz, const double* t,
                       int64_t n, float* xyz_out);

/* ---- downsampling ------------------------------------------------------------------------
 * PointCloudProcessor::temporal_downsample (src/Utils/PointCloudProcessor.cpp:101-112): keeps point i iff
 * (rate <= 1 or (i + 1) % rate == 0) and min_dist < |p|; returns the number kept, their indices in idx_out.    */
int64_t lvo_temporal_downsample(const float* xyz, int64_t n, int rate, double min_dist, int32_t* idx_out);
/* Compensator::voxelgrid_downsample (Compensator.cpp:148-163) = pcl::VoxelGrid with leaf `leaf`: one centroid per
 * occupied leaf, leaves in ascending PCL cell index, fp32 sums.  PCL sorts (cell, point) pairs with an unstable
 * sort, so the summation order inside a leaf is unspecified upstream; this restatement uses input order.
 * Returns the number of leaves (writes at most cap centroids); -1 if PCL would refuse (cell index overflow).   */
int64_t lvo_voxelgrid_downsample(const float* xyz, int64_t n, float leaf, float* xyz_out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
