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

using namespace lv;

namespace {
const int N = kDof;    /* 23 */
const int W = 12;      /* process-noise dof */

struct Dense {         /* small row-major matrix with runtime dims <= 24 x 24 */
    int r, c;
    double v[24 * 24];
    Dense(int r_, int c_) : r(r_), c(c_) { memset(v, 0, sizeof(v)); }
    double& operator()(int i, int j) { return v[i * c + j]; }
    double operator()(int i, int j) const { return v[i * c + j]; }
};
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
    const Mat3d R = quat_to_rot(load_quat(x + kRot));
    const Vec3d grav0 = load_vec3(x + kGrav);
    Vec3d omega, a_b;
    omega.x = gyro[0] - x[kBg]; omega.y = gyro[1] - x[kBg + 1]; omega.z = gyro[2] - x[kBg + 2];
    a_b.x = acc[0] - x[kBa]; a_b.y = acc[1] - x[kBa + 1]; a_b.z = acc[2] - x[kBa + 2];
    const Vec3d a_in = mat3_apply(R, a_b);
    /* f (use-ikfom.cpp:49-61): d(pos) = vel, d(rot) = omega, d(vel) = R (a - ba) + grav; others 0 */
    const double f_pos[3] = {x[kVel], x[kVel + 1], x[kVel + 2]};
    const double f_vel[3] = {a_in.x + grav0.x, a_in.y + grav0.y, a_in.z + grav0.z};

    /* continuous-time Jacobian rows in DOF indexing (df_dx, use-ikfom.cpp:63-79) */
    Dense Fc(N, N);     /* f_x_final */
    Dense G(N, W);      /* f_w_final */
    for (int i = 0; i < 3; ++i) Fc(i, 12 + i) = 1.0;                                  /* pos <- vel */
    {   /* vel rows: -R hat(a - ba) wrt rot, -R wrt ba, grav block wrt the 2 S2 dof */
        const Mat3d RH = mat3_mul(R, hat(a_b));
        double gm[6];
        s2_Mx(grav0, 0.0, 0.0, gm);
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) { Fc(12 + i, 3 + j) = -RH.m[i * 3 + j]; Fc(12 + i, 18 + j) = -R.m[i * 3 + j]; }
            for (int j = 0; j < 2; ++j) Fc(12 + i, 21 + j) = gm[i * 2 + j];
            for (int j = 0; j < 3; ++j) G(12 + i, 3 + j) = -R.m[i * 3 + j];            /* df_dw, :82-90 */
        }
    }
    {   /* rot rows: raw rows are -I wrt bg (and -I wrt ng); SO3 blocks get A(-f dt) applied (esekfom.hpp:327-348) */
        Vec3d seg; seg.x = -omega.x * dt; seg.y = -omega.y * dt; seg.z = -omega.z * dt;
        const Mat3d A = A_matrix(seg);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { Fc(3 + i, 15 + j) = -A.m[i * 3 + j]; G(3 + i, j) = -A.m[i * 3 + j]; }
        /* offset_R_L_I rows: f is zero there, the raw rows are zero, A(0) = I -> stay zero */
    }
    for (int i = 0; i < 3; ++i) { G(15 + i, 6 + i) = 1.0; G(18 + i, 9 + i) = 1.0; }    /* bias random walks */

    /* x <- x (+) f dt (build_manifold.hpp:195-197): vect += f dt, SO3 *= exp(f dt) */
    for (int i = 0; i < 3; ++i) x[kPos + i] += dt * f_pos[i];
    store_quat(x + kRot, quat_mul(load_quat(x + kRot), so3_exp(omega, dt / 2)));
    for (int i = 0; i < 3; ++i) x[kVel + i] += dt * f_vel[i];
    /* offset_R_L_I, offset_T_L_I, bg, ba, grav: f = 0 (exp(0) = identity) */

    /* S2 rows of F (esekfom.hpp:350-377).  f over the grav DIM block is zero, so seg = 0,
     * A(seg) = I and the S2 rows of f_x_final are -Nx hat(grav) applied to zero rows = 0;
     * what remains is the 2x2 block of F_x1: Nx(grav_after) * Mx(grav_before, 0). */
    Dense F(N, N);
    for (int i = 0; i < N; ++i) F(i, i) = 1.0;
    {
        double J2[4];
        s2_J(load_vec3(x + kGrav), grav0, 0.0, 0.0, J2);
        F(21, 21) = J2[0]; F(21, 22) = J2[1]; F(22, 21) = J2[2]; F(22, 22) = J2[3];
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) F(i, j) += Fc(i, j) * dt;                           /* :379 */

    /* P <- F P F^T + (dt G) Q (dt G)^T with the diagonal Q of Localizator.cpp:164-168 */
    double Q[W];
    for (int i = 0; i < 3; ++i) {
        Q[i] = prm.covariance_gyroscope; Q[3 + i] = prm.covariance_acceleration;
        Q[6 + i] = prm.covariance_bias_gyroscope; Q[9 + i] = prm.covariance_bias_acceleration;
    }
    Dense FP(N, N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += F(i, k) * P[k * N + j];
            FP(i, j) = s;
        }
    double Pn[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += FP(i, k) * F(j, k);
            double q = 0;
            for (int k = 0; k < W; ++k) q += (dt * G(i, k)) * Q[k] * (dt * G(j, k));
            Pn[i * N + j] = s + q;
        }
    memcpy(P, Pn, sizeof(Pn));
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
