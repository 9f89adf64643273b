/*
 * lv_predict.h — esekf::predict for the LIMO-Velo state, host+device.
 *
 * Replaces Localizator::propagate -> esekf::predict (src/Modules/Localizator.cpp:159-173,
 * include/IKFoM/IKFoM_toolkit/esekfom/esekfom.hpp:279-384) with f / df_dx / df_dw of include/IKFoM/use-ikfom.cpp:49-90.
 * 23x23 fp64 work per IMU sample.  The same code runs on the host (lv_predict, ExecSerial) and as ONE thread block on the
 * device (lv_propagate_device: all IMU samples between two sweeps in a single launch, the state never leaves HBM between
 * an update, the propagation to the next sweep and the next update).
 *
 * Formulation: instead of materialising the 24x23 / 24x12 Jacobians of the flattened state and re-indexing them
 * (esekfom.hpp:289-368), the three row-blocks that are not plain copies are written directly: rows of an SO3 block get
 * A(-f dt) applied, the S2 block gets -Nx * hat(grav_before) * A(f dt)^T, and the "exp" factors of F_x1 are identities
 * because the reference evaluates scalar(1/2) as an integer division.
 */
#ifndef LV_PREDICT_H_
#define LV_PREDICT_H_

#include "lv_ieskf.h"

namespace lv {

struct PredictNoise {          /* Localizator.cpp:164-168 */
    double gyr, acc, bias_gyr, bias_acc;
};
struct PredictWork {
    double x[kStateLen];
    double P[kN * kN];         /* in / out                                  */
    double F[kN * kN];         /* I + f_x_final dt (with the S2 block)      */
    double G[kN * 12];         /* dt * f_w_final                            */
    double FP[kN * kN];
};

/* one IMU sample: (w->x, w->P) <- predict(dt, Q, {acc, gyro}) */
template <class Ex>
LV_HD_NOINLINE void predict_step(Ex& ex, const PredictNoise& noise, const double* acc, const double* gyro, double dt, PredictWork* w) {
    const int N = kN, W = 12;
    LV_PAR(i, N * N) w->F[i] = (i % (N + 1) == 0) ? 1.0 : 0.0;
    LV_PAR(i, N * W) w->G[i] = 0.0;
    ex.sync();
    if (ex.tid == 0) {
        double* x = w->x;
        const Mat3d R = quat_to_rot(load_quat(x + kRot));
        const Vec3d grav0 = load_vec3(x + kGrav);
        Vec3d omega, a_b;
        omega.x = gyro[0] - x[kBg]; omega.y = gyro[1] - x[kBg + 1]; omega.z = gyro[2] - x[kBg + 2];
        a_b.x = acc[0] - x[kBa]; a_b.y = acc[1] - x[kBa + 1]; a_b.z = acc[2] - x[kBa + 2];
        const Vec3d a_in = mat3_apply(R, a_b);
        /* f (use-ikfom.cpp:49-61): d(pos) = vel, d(rot) = omega, d(vel) = R (a - ba) + grav; others 0 */
        const double f_pos[3] = {x[kVel], x[kVel + 1], x[kVel + 2]};
        const double f_vel[3] = {a_in.x + grav0.x, a_in.y + grav0.y, a_in.z + grav0.z};
        /* continuous-time Jacobian rows in DOF indexing (df_dx, use-ikfom.cpp:63-79), scaled by dt on the fly (:379) */
#define LV_F(i, j) w->F[(i) * N + (j)]
#define LV_G(i, j) w->G[(i) * W + (j)]
        for (int i = 0; i < 3; ++i) LV_F(i, 12 + i) += 1.0 * dt;                          /* pos <- vel */
        {   /* vel rows: -R hat(a - ba) wrt rot, -R wrt ba, grav block wrt the 2 S2 dof */
            const Mat3d RH = mat3_mul(R, hat(a_b));
            double gm[6];
            s2_Mx(grav0, 0.0, 0.0, gm);
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) { LV_F(12 + i, 3 + j) += -RH.m[i * 3 + j] * dt; LV_F(12 + i, 18 + j) += -R.m[i * 3 + j] * dt; }
                for (int j = 0; j < 2; ++j) LV_F(12 + i, 21 + j) += gm[i * 2 + j] * dt;
                for (int j = 0; j < 3; ++j) LV_G(12 + i, 3 + j) = dt * -R.m[i * 3 + j];      /* df_dw, :82-90 */
            }
        }
        {   /* rot rows: raw rows are -I wrt bg (and -I wrt ng); SO3 blocks get A(-f dt) applied (esekfom.hpp:327-348) */
            Vec3d seg; seg.x = -omega.x * dt; seg.y = -omega.y * dt; seg.z = -omega.z * dt;
            const Mat3d A = A_matrix(seg);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) { LV_F(3 + i, 15 + j) += -A.m[i * 3 + j] * dt; LV_G(3 + i, j) = dt * -A.m[i * 3 + j]; }
            /* offset_R_L_I rows: f is zero there, the raw rows are zero, A(0) = I -> stay zero */
        }
        for (int i = 0; i < 3; ++i) { LV_G(15 + i, 6 + i) = dt * 1.0; LV_G(18 + i, 9 + i) = dt * 1.0; }   /* bias random walks */
        /* x <- x (+) f dt (build_manifold.hpp:195-197): vect += f dt, SO3 *= exp(f dt) */
        for (int i = 0; i < 3; ++i) x[kPos + i] += dt * f_pos[i];
        store_quat(x + kRot, quat_mul(load_quat(x + kRot), so3_exp(omega, dt / 2)));
        for (int i = 0; i < 3; ++i) x[kVel + i] += dt * f_vel[i];
        /* offset_R_L_I, offset_T_L_I, bg, ba, grav: f = 0 (exp(0) = identity) */
        /* S2 rows of F (esekfom.hpp:350-377).  f over the grav DIM block is zero, so seg = 0, A(seg) = I and the S2 rows
         * of f_x_final are -Nx hat(grav) applied to zero rows = 0; what remains is the 2x2 block of F_x1:
         * Nx(grav_after) * Mx(grav_before, 0). */
        double J2[4];
        s2_J(load_vec3(x + kGrav), grav0, 0.0, 0.0, J2);
        LV_F(21, 21) = J2[0]; LV_F(21, 22) = J2[1]; LV_F(22, 21) = J2[2]; LV_F(22, 22) = J2[3];
#undef LV_F
#undef LV_G
    }
    ex.sync();
    /* P <- F P F^T + (dt G) Q (dt G)^T with the diagonal Q of Localizator.cpp:164-168 */
    LV_PAR(it, N * N) {
        const int i = it / N, j = it - i * N;
        double s = 0;
        for (int k = 0; k < N; ++k) s += w->F[i * N + k] * w->P[k * N + j];
        w->FP[it] = s;
    }
    ex.sync();
    LV_PAR(it, N * N) {
        const int i = it / N, j = it - i * N;
        double s = 0;
        for (int k = 0; k < N; ++k) s += w->FP[i * N + k] * w->F[j * N + k];
        double q = 0;
        for (int k = 0; k < W; ++k) {
            const double Qk = k < 3 ? noise.gyr : (k < 6 ? noise.acc : (k < 9 ? noise.bias_gyr : noise.bias_acc));
            q += w->G[i * W + k] * Qk * w->G[j * W + k];
        }
        w->P[it] = s + q;
    }
    ex.sync();
}

}  // namespace lv
#endif
