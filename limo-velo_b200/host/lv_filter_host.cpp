/*
 * lv_filter_host.cpp — the parts of the Localizator that stay on the host, as in the reference:
 *   Localizator::init_IKFoM_state   src/Modules/Localizator.cpp:135-153
 *   Localizator::propagate          src/Modules/Localizator.cpp:159-173
 *     -> esekf::predict             include/IKFoM/IKFoM_toolkit/esekfom/esekfom.hpp:279-384
 *     with f / df_dx / df_dw of     include/IKFoM/use-ikfom.cpp:49-90
 * 23x23 fp64 work per IMU sample: negligible next to the measurement path, so it is not
 * accelerated; it only has to leave (x, P) for lv_correct to upload.
 *
 * Formulation: instead of materialising the 24x23 / 24x12 Jacobians of the flattened state and
 * re-indexing them (esekfom.hpp:289-368), the three row-blocks that are not plain copies are
 * written directly: rows of an SO3 block get A(-f dt) applied, the S2 block gets
 * -Nx * hat(grav_before) * A(f dt)^T, and the "exp" factors of F_x1 are identities because the
 * reference evaluates scalar(1/2) as an integer division.
 */
#include <string.h>

#include "../csrc/lv_host.h"
#include "../csrc/lv_manifold.h"
#include "../csrc/lv_predict.h"

using namespace lv;

namespace {
const int N = kDof;    /* 23 */
const int W = 12;      /* process-noise dof */

}  // namespace

void lvh_init_state(const lv_params& prm, const float q_imu[4], double* x, double* P) {
    for (int i = 0; i < kStateLen; ++i) x[i] = 0.0;
    for (int i = 0; i < 4; ++i) x[kRot + i] = (double)q_imu[i];                       /* :137 */
    /* S2(-initial_gravity): normalised, then scaled to the manifold length (S2.hpp:124-127) */
    double g[3] = {-(double)prm.initial_gravity[0], -(double)prm.initial_gravity[1], -(double)prm.initial_gravity[2]};
    const double gn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    for (int i = 0; i < 3; ++i) x[kGrav + i] = g[i] / gn * LV_S2_LEN;                 /* :138 */
    /* Eigen::Map<Matrix3f>(I_Rotation_L.data(),3,3) reads the row-major YAML list COLUMN-major and
     * is not transposed here, unlike State.cpp:23 */
    Mat3d R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[i * 3 + j] = (double)prm.I_Rotation_L[j * 3 + i];
    store_quat(x + kOffR, rot_to_quat(R));                                            /* :140 */
    for (int i = 0; i < 3; ++i) x[kOffT + i] = (double)prm.I_Translation_L[i];        /* :141 */
    for (int i = 0; i < N * N; ++i) P[i] = 0.0;                                       /* :144-152 */
    for (int i = 0; i < N; ++i) {
        double d = 1.0;
        if (i >= 6 && i < 12) d = 0.00001;
        else if (i >= 15 && i < 18) d = 0.0001;
        else if (i >= 18 && i < 21) d = 0.001;
        else if (i >= 21) d = 0.00001;
        P[i * N + i] = d;
    }
}

void lvh_predict(const lv_params& prm, const double acc[3], const double gyro[3], double dt, double* x, double* P) {
    static thread_local PredictWork w;                      /* the same code as the device's lv_predict_kernel (lv_predict.h) */
    memcpy(w.x, x, sizeof(w.x));
    memcpy(w.P, P, sizeof(w.P));
    PredictNoise q = {prm.covariance_gyroscope, prm.covariance_acceleration, prm.covariance_bias_gyroscope, prm.covariance_bias_acceleration};
    ExecSerial ex;
    predict_step(ex, q, acc, gyro, dt, &w);
    memcpy(x, w.x, sizeof(w.x));
    memcpy(P, w.P, sizeof(w.P));
}

extern "C" lv_status lv_init_state_host(const lv_params* p, const float q_imu[4], double* x, double* P) {
    if (!p || !q_imu || !x || !P) return LV_ERR_ARG;
    lvh_init_state(*p, q_imu, x, P);
    return LV_OK;
}
extern "C" lv_status lv_predict_host(const lv_params* p, const double acc[3], const double gyro[3], double dt, double* x,
                                     double* P) {
    if (!p || !acc || !gyro || !x || !P) return LV_ERR_ARG;
    lvh_predict(*p, acc, gyro, dt, x, P);
    return LV_OK;
}
