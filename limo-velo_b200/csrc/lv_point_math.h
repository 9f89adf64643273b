/*
 * lv_point_math.h — per-point arithmetic of the measurement model, host+device.
 *
 *   Frame            the rigid transforms one h-evaluation needs, derived from the state the way
 *                    State(const state_ikfom&, double) does (src/Objects/State.cpp:51-62: fp64 ->
 *                    fp32 casts) and composed like RotTransl (src/Objects/RotTransl.cpp:29-48)
 *   plane_fit        R3Math::estimate_plane + is_plane (src/Utils/Utils.cpp:32-66): 5x3 least
 *                    squares by column-pivoted Householder QR (Eigen 3.3 algorithm), fp32
 *   jacobian_row     one row of Localizator::calculate_H (src/Modules/Localizator.cpp:36-56)
 *
 * fp32 arithmetic is the reference's unfused arithmetic (see lv_hd.h).
 */
#ifndef LV_POINT_MATH_H_
#define LV_POINT_MATH_H_

#include "lv_manifold.h"

#if defined(__CUDACC__)
#define LV_UNROLL _Pragma("unroll")
#else
#define LV_UNROLL
#endif

namespace lv {

struct Rt32 {
    float R[9];
    float t[3];
};

LV_HD void rt_apply(const Rt32& a, float px, float py, float pz, float* g) {   /* RotTransl.cpp:43-48 */
    g[0] = fadd(dot3f(a.R[0], a.R[1], a.R[2], px, py, pz), a.t[0]);
    g[1] = fadd(dot3f(a.R[3], a.R[4], a.R[5], px, py, pz), a.t[1]);
    g[2] = fadd(dot3f(a.R[6], a.R[7], a.R[8], px, py, pz), a.t[2]);
}
LV_HD Rt32 rt_mul(const Rt32& a, const Rt32& b) {   /* RotTransl.cpp:36-41 */
    Rt32 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.R[i * 3 + j] = dot3f(a.R[i * 3], a.R[i * 3 + 1], a.R[i * 3 + 2], b.R[j], b.R[3 + j], b.R[6 + j]);
    for (int i = 0; i < 3; ++i)
        r.t[i] = fadd(dot3f(a.R[i * 3], a.R[i * 3 + 1], a.R[i * 3 + 2], b.t[0], b.t[1], b.t[2]), a.t[i]);
    return r;
}
LV_HD Rt32 rt_inv(const Rt32& a) {   /* RotTransl.cpp:29-34: (R^T, -R^T t) */
    Rt32 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.R[i * 3 + j] = a.R[j * 3 + i];
    for (int i = 0; i < 3; ++i)
        r.t[i] = dot3f(-r.R[i * 3], -r.R[i * 3 + 1], -r.R[i * 3 + 2], a.t[0], a.t[1], a.t[2]);
    return r;
}

/* Everything one h-evaluation reads about the state.  Written by the IESKF step kernel (or the
 * host for the stand-alone operator calls), read by every thread of the measure kernel.      */
struct Frame {
    Rt32 lidar_to_world;     /* X * X.I_Rt_L()                    (Mapper.cpp:51)            */
    Rt32 world_to_lidar;     /* S.I_Rt_L().inv() * S.inv()        (Localizator.cpp:38)       */
    Rt32 lidar_to_imu;       /* S.I_Rt_L()                        (Localizator.cpp:39)       */
    double R_inv[9];         /* s.rot.conjugate().toRotationMatrix()          (:43)          */
    double RLI_inv[9];       /* s.offset_R_L_I.conjugate().toRotationMatrix() (:44)          */
};

LV_HD void make_frame(const double* x, Frame* f) {
    Rt32 X, IL;
    Mat3d R = quat_to_rot(load_quat(x + kRot));
    Mat3d RL = quat_to_rot(load_quat(x + kOffR));
    for (int i = 0; i < 9; ++i) { X.R[i] = (float)R.m[i]; IL.R[i] = (float)RL.m[i]; }
    for (int i = 0; i < 3; ++i) { X.t[i] = (float)x[kPos + i]; IL.t[i] = (float)x[kOffT + i]; }
    f->lidar_to_world = rt_mul(X, IL);
    f->world_to_lidar = rt_mul(rt_inv(IL), rt_inv(X));
    f->lidar_to_imu = IL;
    Mat3d Ri = quat_to_rot(quat_conj(load_quat(x + kRot)));
    Mat3d RLi = quat_to_rot(quat_conj(load_quat(x + kOffR)));
    for (int i = 0; i < 9; ++i) { f->R_inv[i] = Ri.m[i]; f->RLI_inv[i] = RLi.m[i]; }
}

/* squared distance as KD_TREE::calc_dist evaluates it (ikd_Tree.cpp:1682-1687) */
LV_HD float sq_dist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = fsub(ax, bx), dy = fsub(ay, by), dz = fsub(az, bz);
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

/* ---- 5x3 least squares  A n = -1  by column-pivoted Householder QR ------------------------
 * Follows Eigen 3.3 ColPivHouseholderQR::computeInPlace / _solve_impl (the reference calls
 * A.colPivHouseholderQr().solve(b), Utils.cpp:47) with sequential summation.  q holds the five
 * neighbours row-wise.  Returns the un-normalised normal in n[3].                             */
LV_HD void qr_solve_5x3(const float (*q)[3], float* n) {
    float a[5][3];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = q[i][j];
    float tau[3];
    float nu[3], nd[3];        /* m_colNormsUpdated / m_colNormsDirect */
    int perm[3] = {0, 1, 2};   /* column permutation accumulated from the transpositions */
    const float eps = 1.1920928955078125e-07f;
    for (int k = 0; k < 3; ++k) {
        float s = 0.f;
        for (int i = 0; i < 5; ++i) s = fadd(s, fmul(a[i][k], a[i][k]));
        nd[k] = fsqrt(s);
        nu[k] = nd[k];
    }
    float mx = nu[0] > nu[1] ? nu[0] : nu[1];
    mx = mx > nu[2] ? mx : nu[2];
    const float me = fmul(mx, eps);
    const float threshold_helper = fdiv(fmul(me, me), 5.0f);
    const float downdate_threshold = 3.4526698300124393e-04f;   /* sqrt(eps) */
    int nonzero = 3;
    LV_UNROLL
    for (int k = 0; k < 3; ++k) {
        int big = k;
        float bigv = nu[k];
        LV_UNROLL
        for (int j = k + 1; j < 3; ++j)
            if (nu[j] > bigv) { big = j; bigv = nu[j]; }
        if (nonzero == 3 && fmul(bigv, bigv) < fmul(threshold_helper, (float)(5 - k))) nonzero = k;
        LV_UNROLL
        for (int j = k + 1; j < 3; ++j)
            if (big == j) {   /* static column indices after unrolling: no local-memory arrays */
                LV_UNROLL
                for (int i = 0; i < 5; ++i) { float t = a[i][k]; a[i][k] = a[i][j]; a[i][j] = t; }
                float t = nu[k]; nu[k] = nu[j]; nu[j] = t;
                t = nd[k]; nd[k] = nd[j]; nd[j] = t;
                int ti = perm[k]; perm[k] = perm[j]; perm[j] = ti;
            }
        /* makeHouseholderInPlace on a[k..4][k] */
        float tail = 0.f;
        for (int i = k + 1; i < 5; ++i) tail = fadd(tail, fmul(a[i][k], a[i][k]));
        const float c0 = a[k][k];
        float beta;
        if (tail <= 1.17549435e-38f) {
            tau[k] = 0.f;
            beta = c0;
            for (int i = k + 1; i < 5; ++i) a[i][k] = 0.f;
        } else {
            beta = fsqrt(fadd(fmul(c0, c0), tail));
            if (c0 >= 0.f) beta = -beta;
            const float den = fsub(c0, beta);
            for (int i = k + 1; i < 5; ++i) a[i][k] = fdiv(a[i][k], den);
            tau[k] = fdiv(fsub(beta, c0), beta);
        }
        a[k][k] = beta;
        /* applyHouseholderOnTheLeft on the trailing columns */
        if (tau[k] != 0.f) {
            for (int j = k + 1; j < 3; ++j) {
                float tmp = 0.f;
                for (int i = k + 1; i < 5; ++i) tmp = fadd(tmp, fmul(a[i][k], a[i][j]));
                tmp = fadd(tmp, a[k][j]);
                a[k][j] = fsub(a[k][j], fmul(tau[k], tmp));
                for (int i = k + 1; i < 5; ++i) a[i][j] = fsub(a[i][j], fmul(fmul(tau[k], a[i][k]), tmp));
            }
        }
        /* column-norm downdate */
        for (int j = k + 1; j < 3; ++j) {
            if (nu[j] != 0.f) {
                float t = fdiv(fabsf(a[k][j]), nu[j]);
                t = fmul(fadd(1.0f, t), fsub(1.0f, t));
                t = t < 0.f ? 0.f : t;
                const float r = fdiv(nu[j], nd[j]);
                const float t2 = fmul(t, fmul(r, r));
                if (t2 <= downdate_threshold) {
                    float s = 0.f;
                    for (int i = k + 1; i < 5; ++i) s = fadd(s, fmul(a[i][j], a[i][j]));
                    nd[j] = fsqrt(s);
                    nu[j] = nd[j];
                } else {
                    nu[j] = fmul(nu[j], fsqrt(t));
                }
            }
        }
    }
    /* c = Q^T b with b = -1 */
    float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
    n[0] = n[1] = n[2] = 0.f;
    if (nonzero == 0) return;
    for (int k = 0; k < 3; ++k) {
        if (k >= nonzero || tau[k] == 0.f) continue;
        float tmp = 0.f;
        for (int i = k + 1; i < 5; ++i) tmp = fadd(tmp, fmul(a[i][k], c[i]));
        tmp = fadd(tmp, c[k]);
        c[k] = fsub(c[k], fmul(tau[k], tmp));
        for (int i = k + 1; i < 5; ++i) c[i] = fsub(c[i], fmul(fmul(tau[k], a[i][k]), tmp));
    }
    /* back substitution on the leading nonzero x nonzero triangle */
    for (int i = 2; i >= 0; --i) {
        if (i >= nonzero) continue;
        float s = c[i];
        for (int j = i + 1; j < 3; ++j)
            if (j < nonzero) s = fsub(s, fmul(a[i][j], c[j]));
        c[i] = fdiv(s, a[i][i]);
    }
    for (int i = 0; i < 3; ++i)
        if (i < nonzero) {
            /* n[perm[i]] = c[i] without dynamic register indexing */
            if (perm[i] == 0) n[0] = c[i];
            else if (perm[i] == 1) n[1] = c[i];
            else n[2] = c[i];
        }
}

/* estimate_plane (Utils.cpp:32-57) + is_plane (Utils.cpp:59-66).  abcd = (A,B,C,D). */
LV_HD bool plane_fit(const float (*q)[3], float threshold, float* abcd) {
    float nv[3];
    qr_solve_5x3(q, nv);
    const float n = fsqrt(fadd(fadd(fmul(nv[0], nv[0]), fmul(nv[1], nv[1])), fmul(nv[2], nv[2])));
    abcd[0] = fdiv(nv[0], n);
    abcd[1] = fdiv(nv[1], n);
    abcd[2] = fdiv(nv[2], n);
    abcd[3] = fdiv(1.0f, n);   /* (float)(1.0 / n): double rounding is innocuous for division */
    bool ok = true;
    for (int j = 0; j < 5; ++j) {
        const float res = fadd(fadd(fadd(fmul(abcd[0], q[j][0]), fmul(abcd[1], q[j][1])), fmul(abcd[2], q[j][2])), abcd[3]);
        if (fabsf(res) > threshold) ok = false;
    }
    return ok;
}

/* Plane::dist_to_plane (Plane.cpp:27-29) */
LV_HD float plane_dist(const float* abcd, const float* g) {
    return fadd(fadd(fadd(fmul(abcd[0], g[0]), fmul(abcd[1], g[1])), fmul(abcd[2], g[2])), abcd[3]);
}

/* One row of calculate_H (Localizator.cpp:36-56): row[12] (fp64) and h = -distance. */
LV_HD void jacobian_row(const Frame& f, const float* g, const float* abcd, float dist, bool estimate_extrinsics,
                        double* row, double* hval) {
    float pl[3], pi[3];
    rt_apply(f.world_to_lidar, g[0], g[1], g[2], pl);
    rt_apply(f.lidar_to_imu, pl[0], pl[1], pl[2], pi);
    const double nx = (double)abcd[0], ny = (double)abcd[1], nz = (double)abcd[2];
    const double Cx = dot3d(f.R_inv[0], f.R_inv[1], f.R_inv[2], nx, ny, nz);
    const double Cy = dot3d(f.R_inv[3], f.R_inv[4], f.R_inv[5], nx, ny, nz);
    const double Cz = dot3d(f.R_inv[6], f.R_inv[7], f.R_inv[8], nx, ny, nz);
    row[0] = nx; row[1] = ny; row[2] = nz;
    const double ix = (double)pi[0], iy = (double)pi[1], iz = (double)pi[2];
    row[3] = dsub(dmul(iy, Cz), dmul(iz, Cy));     /* A = p_imu x C */
    row[4] = dsub(dmul(iz, Cx), dmul(ix, Cz));
    row[5] = dsub(dmul(ix, Cy), dmul(iy, Cx));
    if (estimate_extrinsics) {
        const double Dx = dot3d(f.RLI_inv[0], f.RLI_inv[1], f.RLI_inv[2], Cx, Cy, Cz);
        const double Dy = dot3d(f.RLI_inv[3], f.RLI_inv[4], f.RLI_inv[5], Cx, Cy, Cz);
        const double Dz = dot3d(f.RLI_inv[6], f.RLI_inv[7], f.RLI_inv[8], Cx, Cy, Cz);
        const double lx = (double)pl[0], ly = (double)pl[1], lz = (double)pl[2];
        row[6] = dsub(dmul(ly, Dz), dmul(lz, Dy));  /* B = p_lidar x (R_LI^T C) */
        row[7] = dsub(dmul(lz, Dx), dmul(lx, Dz));
        row[8] = dsub(dmul(lx, Dy), dmul(ly, Dx));
        row[9] = Cx; row[10] = Cy; row[11] = Cz;
    } else {
        for (int i = 6; i < 12; ++i) row[i] = 0.0;
    }
    *hval = -(double)dist;
}

}  // namespace lv
#endif
