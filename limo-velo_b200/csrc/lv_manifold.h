/*
 * lv_manifold.h — fp64 manifold algebra of the 23-DoF IKFoM state, host+device.
 *
 * Replaces the MTK pieces the hot path touches (paths under include/IKFoM/IKFoM_toolkit/):
 *   compound [+]/[-]           mtk/build_manifold.hpp:192-200
 *   SO3 [+],[-],exp,log        mtk/types/SOn.hpp:233-239,284-297
 *   S2  [+],[-],Bx,Nx_yy,Mx    mtk/types/S2.hpp:136-167,215-231,259-280  (length 9.809, typ 1:
 *                              include/IKFoM/use-ikfom.hpp:8)
 *   vect [+]/[-]               mtk/types/vect.hpp:117-122
 *   hat, A_matrix, exp, log, cos_sinc_sqrt, tolerance   mtk/src/mtkmath.hpp:119-122,142-183,235-288
 * State layout: see include/limovelo_b200.h.
 */
#ifndef LV_MANIFOLD_H_
#define LV_MANIFOLD_H_

#include "lv_hd.h"

namespace lv {

enum StateOffset { kPos = 0, kRot = 3, kOffR = 7, kOffT = 11, kVel = 14, kBg = 17, kBa = 20, kGrav = 23 };
enum { kStateLen = 26, kDof = 23, kMeas = 12 };

#define LV_TOL 1e-11            /* MTK::tolerance<double>() */
#define LV_S2_LEN 9.809         /* 98090 / 10000            */

struct Vec3d { double x, y, z; };
struct Quatd { double x, y, z, w; };
struct Mat3d { double m[9]; };   /* row-major */

LV_HD Mat3d mat3_identity() {
    Mat3d r;
    for (int i = 0; i < 9; ++i) r.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return r;
}
LV_HD Mat3d mat3_mul(const Mat3d& a, const Mat3d& b) {
    Mat3d r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
    return r;
}
LV_HD Mat3d mat3_transpose(const Mat3d& a) {
    Mat3d r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[j * 3 + i];
    return r;
}
LV_HD Vec3d mat3_apply(const Mat3d& a, const Vec3d& v) {
    Vec3d r;
    r.x = a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z;
    r.y = a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z;
    r.z = a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z;
    return r;
}
LV_HD Mat3d hat(const Vec3d& v) {   /* mtkmath.hpp:176-183 */
    Mat3d r;
    r.m[0] = 0;    r.m[1] = -v.z; r.m[2] = v.y;
    r.m[3] = v.z;  r.m[4] = 0;    r.m[5] = -v.x;
    r.m[6] = -v.y; r.m[7] = v.x;  r.m[8] = 0;
    return r;
}

/* Eigen::QuaternionBase::toRotationMatrix.  Every product and sum is rounded separately (dmul / dadd /
 * dsub), as the reference's x86-64 build does: the rotation matrices feed the fp64 -> fp32 casts of
 * State.cpp:53-61 and the Jacobian rows, so a contracted FMA here would show up as 1-ulp differences. */
LV_HD Mat3d quat_to_rot(const Quatd& q) {
    const double tx = dmul(2.0, q.x), ty = dmul(2.0, q.y), tz = dmul(2.0, q.z);
    const double twx = dmul(tx, q.w), twy = dmul(ty, q.w), twz = dmul(tz, q.w);
    const double txx = dmul(tx, q.x), txy = dmul(ty, q.x), txz = dmul(tz, q.x);
    const double tyy = dmul(ty, q.y), tyz = dmul(tz, q.y), tzz = dmul(tz, q.z);
    Mat3d r;
    r.m[0] = dsub(1.0, dadd(tyy, tzz)); r.m[1] = dsub(txy, twz);            r.m[2] = dadd(txz, twy);
    r.m[3] = dadd(txy, twz);            r.m[4] = dsub(1.0, dadd(txx, tzz)); r.m[5] = dsub(tyz, twx);
    r.m[6] = dsub(txz, twy);            r.m[7] = dadd(tyz, twx);            r.m[8] = dsub(1.0, dadd(txx, tyy));
    return r;
}
LV_HD Quatd quat_mul(const Quatd& a, const Quatd& b) {
    Quatd r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
LV_HD Quatd quat_conj(const Quatd& a) {
    Quatd r;
    r.x = -a.x; r.y = -a.y; r.z = -a.z; r.w = a.w;
    return r;
}
/* Eigen quaternion-from-matrix (Shepperd), used by SO3(Matrix3d) in Localizator.cpp:140 */
LV_HD Quatd rot_to_quat(const Mat3d& R) {
    double q[4];
    double t = R.m[0] + R.m[4] + R.m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R.m[7] - R.m[5]) * t;
        q[1] = (R.m[2] - R.m[6]) * t;
        q[2] = (R.m[3] - R.m[1]) * t;
    } else {
        int i = 0;
        if (R.m[4] > R.m[0]) i = 1;
        if (R.m[8] > R.m[i * 4]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R.m[i * 4] - R.m[j * 4] - R.m[k * 4] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R.m[k * 3 + j] - R.m[j * 3 + k]) * t;
        q[j] = (R.m[j * 3 + i] + R.m[i * 3 + j]) * t;
        q[k] = (R.m[k * 3 + i] + R.m[i * 3 + k]) * t;
    }
    Quatd r;
    r.x = q[0]; r.y = q[1]; r.z = q[2]; r.w = q[3];
    return r;
}
LV_HD Quatd load_quat(const double* p) {
    Quatd q;
    q.x = p[0]; q.y = p[1]; q.z = p[2]; q.w = p[3];
    return q;
}
LV_HD void store_quat(double* p, const Quatd& q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
LV_HD Vec3d load_vec3(const double* p) {
    Vec3d v;
    v.x = p[0]; v.y = p[1]; v.z = p[2];
    return v;
}
LV_HD void store_vec3(double* p, const Vec3d& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

/* mtkmath.hpp:142-174: cos(sqrt(x2)), sinc(sqrt(x2)) with the reference's 3-term series */
LV_HD void cos_sinc_sqrt(double x2, double* c, double* s) {
    const double taylor_n_bound = 1.220703125e-4; /* sqrt(sqrt(DBL_EPSILON)) = 2^-13 */
    if (x2 >= taylor_n_bound) {
        double x = sqrt(x2);
        *c = cos(x);
        *s = sin(x) / x;
        return;
    }
    double cosi = 1., sinc = 1.;
    double term = -1 / 2. * x2;
    cosi += term; term *= 1 / 3.; sinc += term; term *= -(1 / 4.) * x2;
    cosi += term; term *= 1 / 5.; sinc += term; term *= -(1 / 6.) * x2;
    cosi += term; term *= 1 / 7.; sinc += term;
    *c = cosi;
    *s = sinc;
}
/* SO3::exp(vec, scale) (SOn.hpp:284-288 -> mtkmath.hpp:249-256); half = scale / 2 */
LV_HD Quatd so3_exp(const Vec3d& v, double half) {
    double c, s;
    cos_sinc_sqrt(half * half * (v.x * v.x + v.y * v.y + v.z * v.z), &c, &s);
    double mult = s * half;
    Quatd q;
    q.x = mult * v.x; q.y = mult * v.y; q.z = mult * v.z; q.w = c;
    return q;
}
/* SO3::log (SOn.hpp:293-297 -> mtkmath.hpp:268-288, scale 2, plus_minus_periodicity) */
LV_HD Vec3d so3_log(const Quatd& q) {
    double nv = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (nv < LV_TOL) nv = LV_TOL;
    double s = 2.0 / nv * atan(nv / q.w);
    Vec3d r;
    r.x = s * q.x; r.y = s * q.y; r.z = s * q.z;
    return r;
}
/* mtkmath.hpp:235-247 */
LV_HD Mat3d A_matrix(const Vec3d& v) {
    double sq = v.x * v.x + v.y * v.y + v.z * v.z;
    double norm = sqrt(sq);
    Mat3d r = mat3_identity();
    if (norm < LV_TOL) return r;
    Mat3d H = hat(v), HH = mat3_mul(H, H);
    double a = (1 - cos(norm)) / sq, b = (1 - sin(norm) / norm) / sq;
    for (int i = 0; i < 9; ++i) r.m[i] += a * H.m[i] + b * HH.m[i];
    return r;
}

/* S2.hpp:215-231, typ-1 branch; B is 3x2 row-major */
LV_HD void s2_Bx(const Vec3d& v, double* B) {
    const double L = LV_S2_LEN;
    if (v.x + L > LV_TOL) {
        /* one reciprocal instead of the reference's ten divisions (a dependent fp64 division costs ~100 cycles in
         * the single-thread pieces of the step kernel); differs from S2.hpp in the last ulp only */
        const double id = 1.0 / (L + v.x), iL = 1.0 / L;
        const double yz = v.z * v.y * id;
        B[0] = -v.y * iL;                   B[1] = -v.z * iL;
        B[2] = (L - v.y * v.y * id) * iL;   B[3] = -yz * iL;
        B[4] = -yz * iL;                    B[5] = (L - v.z * v.z * id) * iL;
    } else {
        for (int i = 0; i < 6; ++i) B[i] = 0;
        B[3] = -1;
        B[4] = 1;
    }
}
LV_HD Vec3d s2_Bx_apply(const double* B, double d0, double d1) {
    Vec3d r;
    r.x = B[0] * d0 + B[1] * d1;
    r.y = B[2] * d0 + B[3] * d1;
    r.z = B[4] * d0 + B[5] * d1;
    return r;
}
/* S2.hpp:136-142 */
LV_HD Vec3d s2_boxplus(const Vec3d& v, double d0, double d1) {
    double B[6];
    s2_Bx(v, B);
    Vec3d Bu = s2_Bx_apply(B, d0, d1);
    return mat3_apply(quat_to_rot(so3_exp(Bu, 0.5)), v);
}
/* S2.hpp:144-167: res = v [-] o */
LV_HD void s2_boxminus(const Vec3d& v, const Vec3d& o, double* res) {
    Vec3d hv = mat3_apply(hat(v), o);
    double v_sin = sqrt(hv.x * hv.x + hv.y * hv.y + hv.z * hv.z);
    double v_cos = v.x * o.x + v.y * o.y + v.z * o.z;
    double theta = atan2(v_sin, v_cos);
    if (v_sin < LV_TOL) {
        if (fabs(theta) > LV_TOL) { res[0] = 3.1415926; res[1] = 0; }
        else { res[0] = 0; res[1] = 0; }
        return;
    }
    double B[6];
    s2_Bx(o, B);
    Vec3d t = mat3_apply(hat(o), v);
    double f = theta / v_sin;
    res[0] = f * (B[0] * t.x + B[2] * t.y + B[4] * t.z);
    res[1] = f * (B[1] * t.x + B[3] * t.y + B[5] * t.z);
}
/* S2.hpp:259-264: Nx (2x3 row-major) = 1/L^2 Bx^T hat(v) */
LV_HD void s2_Nx_yy(const Vec3d& v, double* N) {
    double B[6];
    s2_Bx(v, B);
    Mat3d H = hat(v);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j)
            N[i * 3 + j] = 1 / LV_S2_LEN / LV_S2_LEN * (B[i] * H.m[j] + B[2 + i] * H.m[3 + j] + B[4 + i] * H.m[6 + j]);
}
/* S2.hpp:266-280: Mx (3x2 row-major).  scalar(1/2) is an integer division in the reference,
 * so exp_delta is the identity rotation.                                */
LV_HD void s2_Mx(const Vec3d& v, double d0, double d1, double* Mx) {
    double B[6];
    s2_Bx(v, B);
    Mat3d H = hat(v);
    Mat3d T = H;
    if (!(sqrt(d0 * d0 + d1 * d1) < LV_TOL)) {
        Vec3d Bu = s2_Bx_apply(B, d0, d1);
        T = mat3_mul(H, mat3_transpose(A_matrix(Bu)));
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 2; ++j)
            Mx[i * 2 + j] = -(T.m[i * 3] * B[j] + T.m[i * 3 + 1] * B[2 + j] + T.m[i * 3 + 2] * B[4 + j]);
}
/* the 2x2 block Nx(x_now.grav) * Mx(x_prop.grav, delta) of esekfom.hpp:1689-1691,1803-1805 */
LV_HD void s2_J(const Vec3d& grav_now, const Vec3d& grav_prop, double d0, double d1, double* J) {
    double Nx[6], Mx[6];
    s2_Nx_yy(grav_now, Nx);
    s2_Mx(grav_prop, d0, d1, Mx);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) J[i * 2 + j] = Nx[i * 3] * Mx[j] + Nx[i * 3 + 1] * Mx[2 + j] + Nx[i * 3 + 2] * Mx[4 + j];
}

/* build_manifold.hpp:192-194 */
LV_HD void state_boxplus(double* x, const double* d) {
    for (int i = 0; i < 3; ++i) x[kPos + i] += d[i];
    store_quat(x + kRot, quat_mul(load_quat(x + kRot), so3_exp(load_vec3(d + 3), 0.5)));
    store_quat(x + kOffR, quat_mul(load_quat(x + kOffR), so3_exp(load_vec3(d + 6), 0.5)));
    for (int i = 0; i < 3; ++i) x[kOffT + i] += d[9 + i];
    for (int i = 0; i < 3; ++i) x[kVel + i] += d[12 + i];
    for (int i = 0; i < 3; ++i) x[kBg + i] += d[15 + i];
    for (int i = 0; i < 3; ++i) x[kBa + i] += d[18 + i];
    store_vec3(x + kGrav, s2_boxplus(load_vec3(x + kGrav), d[21], d[22]));
}
/* build_manifold.hpp:198-200: d = x [-] y */
LV_HD void state_boxminus(const double* x, const double* y, double* d) {
    for (int i = 0; i < 3; ++i) d[i] = x[kPos + i] - y[kPos + i];
    store_vec3(d + 3, so3_log(quat_mul(quat_conj(load_quat(y + kRot)), load_quat(x + kRot))));
    store_vec3(d + 6, so3_log(quat_mul(quat_conj(load_quat(y + kOffR)), load_quat(x + kOffR))));
    for (int i = 0; i < 3; ++i) d[9 + i] = x[kOffT + i] - y[kOffT + i];
    for (int i = 0; i < 3; ++i) d[12 + i] = x[kVel + i] - y[kVel + i];
    for (int i = 0; i < 3; ++i) d[15 + i] = x[kBg + i] - y[kBg + i];
    for (int i = 0; i < 3; ++i) d[18 + i] = x[kBa + i] - y[kBa + i];
    s2_boxminus(load_vec3(x + kGrav), load_vec3(y + kGrav), d + 21);
}

}  // namespace lv
#endif
