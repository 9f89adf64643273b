/*
 * lv_ieskf.h — one evaluation of the iterated error-state Kalman update, host+device.
 *
 * Replaces the body of esekf::update_iterated_dyn_share_modified
 * (include/IKFoM/IKFoM_toolkit/esekfom/esekfom.hpp:1620-1823) for n = 23, measurement block 12,
 * in the only branch LIMO-Velo can define (Nm >= 23, esekfom.hpp:1720-1729).  The 23x23 algebra
 * that the reference runs on the host between two h_dyn_share calls runs here as ONE thread block
 * between two measure-kernel launches, so a whole <= (MAX_NUM_ITERS+1)-evaluation update needs no
 * host round trip.
 *
 * The code is written as "parallel phases": LV_PAR(i, n) distributes n independent items over the
 * block and ex.sync() separates phases.  With ExecSerial (one thread, no-op sync) the very same
 * code runs on the host; tests/cpu_shim uses that to unit-test it without a GPU.
 */
#ifndef LV_IESKF_H_
#define LV_IESKF_H_

#include "lv_point_math.h"

namespace lv {

enum { kN = 23, kAug = 46, kMaxEvals = 8 };

struct IeskfParams {
    double R;                 /* LiDAR_noise            (Localizator.cpp:132)   */
    double D;                 /* degeneracy_threshold   (Localizator.cpp:132)   */
    double limits[kN];        /* LIMITS                 (Localizator.cpp:115)   */
    int32_t max_iter;         /* MAX_NUM_ITERS          (Localizator.cpp:114)   */
    int32_t estimate_extrinsics;
};

/* mirrors lv_iter_log of the C ABI (include/limovelo_b200.h) */
struct IterLog {
    int64_t n_matches;
    int32_t converged;
    int32_t degenerate;
    double HTH[144];
    double HTh[12];
    double dx[kN];
    double x_after[kStateLen];
};

/* device-resident state of one update, shared by the measure and the step kernels */
struct UpdateCtrl {
    int32_t iter;             /* loop variable i of esekfom.hpp:1634, starts at -1       */
    int32_t t;                /* number of converged evaluations (esekfom.hpp:1757)      */
    int32_t done;             /* 1: update finished, later launches return immediately   */
    int32_t status;           /* lv_status                                               */
    int32_t n_evals;
    int32_t pad_[3];
    double x_prop[kStateLen];
    double P_prop[kN * kN];
    double x[kStateLen];      /* current iterate; final state when done                  */
    double P[kN * kN];        /* final covariance when done                              */
    Frame frame;              /* transforms of the current iterate for the next measure  */
    IterLog logs[kMaxEvals];
};

/* shared-memory workspace of the step */
struct IeskfWork {
    double x[kStateLen], xp[kStateLen];
    double P[kN * kN];        /* P_ after the J blocks (esekfom.hpp:1655-1697)           */
    double M1[12 * 25];       /* Gauss-Jordan scratch: [I + Q S11 | Q | HTh]              */
    double Kx[kN * 12];
    double Kh[kN];
    double HTH[144], HTh[12];
    double dx[kN], dx_new[kN], dxs[kN], dnd[kN];   /* dx, J*dx, dx_ (solved), masked */
    double J[3][9];           /* J blocks: SO3@3, SO3@6 (3x3), S2@21 (2x2 in the first 4) */
    double L[kN * kN];
    double pivval;
    int64_t n_matches;
    int32_t piv;
    int32_t finish;
    int32_t abort_;
};

/* executors of the "parallel phases" below: ExecSerial (host, one thread) and ExecBlock (device) */
struct ExecSerial {
    int tid, nthreads;
    LV_HD ExecSerial() : tid(0), nthreads(1) {}
    LV_HD void sync() {}
    /* partial pivoting: row p >= k with the largest |M[p][k]| (first one on ties) */
    LV_HD void pivot(const double* M, int w, int k, int n, int32_t* s_piv, double* s_pivval) {
        int p = k;
        double best = fabs(M[k * w + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(M[i * w + k]);
            if (v > best) { best = v; p = i; }
        }
        *s_piv = p;
        *s_pivval = M[p * w + k];
    }
};
#if defined(__CUDACC__)
/* One thread block runs the step.  (A single-warp variant with __syncwarp instead of block barriers
 * was measured slower, 59 us vs 45 us per evaluation: 15 k dependent fp64 / shared-memory
 * instructions issue at ~8 cycles each when one warp has nothing else to switch to.) */
struct ExecBlock {
    int tid, nthreads;
    __device__ __forceinline__ ExecBlock() : tid(threadIdx.x), nthreads(blockDim.x) {}
    __device__ __forceinline__ void sync() { __syncthreads(); }
    /* partial pivoting by warp 0: lane i holds row k+i (n - k <= 32), shuffle arg-max, lowest row on ties */
    __device__ __forceinline__ void pivot(const double* M, int w, int k, int n, int32_t* s_piv, double* s_pivval) {
        if (tid < 32) {
            const int row = k + tid;
            double v = row < n ? fabs(M[row * w + k]) : -1.0;
            int r = row < n ? row : 0x7fffffff;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, s);
                const int orow = __shfl_xor_sync(0xffffffffu, r, s);
                if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; }
            }
            if (tid == 0) { *s_piv = r; *s_pivval = M[r * w + k]; }
        }
    }
};
#endif

#define LV_PAR(i, n) for (int i = ex.tid; i < (n); i += ex.nthreads)

/* Gauss-Jordan elimination with partial pivoting on M = [A | B] (n rows, w >= n columns, row-major):
 * on return the columns n..w-1 hold A^-1 B (the left block is not cleaned up).                   */
template <class Ex>
LV_HD_NOINLINE void gj_solve(Ex& ex, double* M, int n, int w, int32_t* s_piv, double* s_pivval) {
    for (int k = 0; k < n; ++k) {
        ex.pivot(M, w, k, n, s_piv, s_pivval);
        ex.sync();
        const int p = *s_piv;
        const double ipv = 1.0 / *s_pivval;
        LV_PAR(jj, w - k) {
            const int j = k + jj;
            const double a = M[k * w + j], b = M[p * w + j];
            if (p != k) M[p * w + j] = a;
            M[k * w + j] = b * ipv;
        }
        ex.sync();
        const int ncols = w - k - 1;
        LV_PAR(it, (n - 1) * ncols) {
            int i = it / ncols;
            const int j = k + 1 + (it - i * ncols);
            if (i >= k) ++i;
            M[i * w + j] -= M[i * w + k] * M[k * w + j];
        }
        ex.sync();
    }
}

/* rows idx..idx+d-1 of dst (all `cols` columns) = J * the same rows of src */
template <class Ex>
LV_HD void apply_rows(Ex& ex, double* dst, const double* src, int idx, int d, const double* J, int cols, int stride) {
    LV_PAR(i, cols) {
        double v[3], r[3];
        for (int c = 0; c < d; ++c) v[c] = src[(idx + c) * stride + i];
        for (int a = 0; a < d; ++a) {
            double s = 0;
            for (int c = 0; c < d; ++c) s += J[a * d + c] * v[c];
            r[a] = s;
        }
        for (int a = 0; a < d; ++a) dst[(idx + a) * stride + i] = r[a];
    }
}
/* columns idx..idx+d-1 of M (all `rows` rows) = the same columns * J^T */
template <class Ex>
LV_HD void apply_cols(Ex& ex, double* M, int idx, int d, const double* J, int rows, int stride) {
    LV_PAR(i, rows) {
        double v[3], r[3];
        for (int c = 0; c < d; ++c) v[c] = M[i * stride + idx + c];
        for (int a = 0; a < d; ++a) {
            double s = 0;
            for (int c = 0; c < d; ++c) s += v[c] * J[a * d + c];
            r[a] = s;
        }
        for (int a = 0; a < d; ++a) M[i * stride + idx + a] = r[a];
    }
}

/* ---- serial helpers for the (rare) degenerate branch, esekfom.hpp:1736-1744 ---------------- */
/* cyclic Jacobi, eigenvalues ascending, vectors = columns of V, largest component positive.
 * The reference uses Eigen::EigenSolver whose pair order is unspecified: the degenerate branch
 * is "parity unpinned" (DESIGN.md).                                                           */
LV_HD_NOINLINE void sym_eig6(const double* Ain, double* ev, double* V) {
    const int n = 6;
    double A[36];
    for (int i = 0; i < 36; ++i) { A[i] = Ain[i]; V[i] = (i % 7 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < n; ++i) {
            diag += A[i * n + i] * A[i * n + i];
            for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    /* selection sort of the eigenpairs (ascending) + sign normalisation */
    double Vs[36];
    bool used[6] = {false, false, false, false, false, false};
    for (int j = 0; j < n; ++j) {
        int src = -1;
        for (int c = 0; c < n; ++c)
            if (!used[c] && (src < 0 || A[c * n + c] < A[src * n + src])) src = c;
        used[src] = true;
        ev[j] = A[src * n + src];
        int imax = 0;
        for (int k = 1; k < n; ++k)
            if (fabs(V[k * n + src]) > fabs(V[imax * n + src])) imax = k;
        const double sg = V[imax * n + src] < 0 ? -1.0 : 1.0;
        for (int k = 0; k < n; ++k) Vs[k * n + j] = sg * V[k * n + src];
    }
    for (int i = 0; i < 36; ++i) V[i] = Vs[i];
}

/* dnd[0:6] = V^-1 * sel * dx_[0:6]  (esekfom.hpp:1736-1744), executed by one thread */
LV_HD_NOINLINE void degenerate_mask(const double* HTH, double D, const double* dxs, double* dnd) {
    double A6[36], ev[6], V[36], sel[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) A6[i * 6 + j] = HTH[i * 12 + j];
    sym_eig6(A6, ev, V);
    double prod = 1;
    for (int i = 0; i < 6; ++i) prod *= ev[i];
    if (prod < 1e-20)
        for (int i = 0; i < 36; ++i) V[i] = (i % 7 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 36; ++i) sel[i] = V[i];
    for (int j = 0; j < 6; ++j)
        if (ev[j] < D)
            for (int c = 0; c < 6; ++c) sel[j * 6 + c] = 0;
    double t[6];
    for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int j = 0; j < 6; ++j) s += sel[i * 6 + j] * dxs[j];
        t[i] = s;
    }
    /* solve V y = t by Gauss-Jordan with partial pivoting (== V^-1 t) */
    double M[6 * 7];
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) M[i * 7 + j] = V[i * 6 + j];
        M[i * 7 + 6] = t[i];
    }
    for (int k = 0; k < 6; ++k) {
        int p = k;
        for (int i = k + 1; i < 6; ++i)
            if (fabs(M[i * 7 + k]) > fabs(M[p * 7 + k])) p = i;
        if (p != k)
            for (int j = 0; j < 7; ++j) { const double tmp = M[k * 7 + j]; M[k * 7 + j] = M[p * 7 + j]; M[p * 7 + j] = tmp; }
        const double pv = M[k * 7 + k];
        for (int j = k; j < 7; ++j) M[k * 7 + j] /= pv;
        for (int i = 0; i < 6; ++i) {
            if (i == k) continue;
            const double f = M[i * 7 + k];
            for (int j = k; j < 7; ++j) M[i * 7 + j] -= f * M[k * 7 + j];
        }
    }
    for (int i = 0; i < 6; ++i) dnd[i] = M[i * 7 + 6];
}

/* true iff every eigenvalue of HTH[0:6,0:6] exceeds D: Cholesky of (A - D I) succeeds */
LV_HD bool all_eigs_above(const double* HTH, double D) {
    double Lc[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = HTH[i * 12 + j] - (i == j ? D : 0.0);
            for (int k = 0; k < j; ++k) s -= Lc[i * 6 + k] * Lc[j * 6 + k];
            if (i == j) {
                if (!(s > 0.0)) return false;
                Lc[i * 6 + i] = sqrt(s);
            } else {
                Lc[i * 6 + j] = s / Lc[j * 6 + j];
            }
        }
    return true;
}

/*
 * One evaluation (esekfom.hpp:1647-1822).  Inputs: c->x_prop/P_prop/x, w->HTH/HTh/n_matches
 * (already reduced).  Outputs: c->x (new iterate), c->frame, c->logs[], and on exit c->P, c->done.
 */
template <class Ex>
LV_HD_NOINLINE void ieskf_step(Ex& ex, const IeskfParams& prm, UpdateCtrl* c, IeskfWork* w) {
    const int n = kN;
    /* Nm < n: the reference takes esekfom.hpp:1701-1709 and then reads an uninitialised HTH
     * (SURVEY 8c quirk 4).  Here: stop, report LV_TOO_FEW_MATCHES, keep the current iterate. */
    if (ex.tid == 0) {
        w->abort_ = 0;
        if (w->n_matches < n) {
            IterLog* lg = &c->logs[c->n_evals < kMaxEvals ? c->n_evals : kMaxEvals - 1];
            lg->n_matches = w->n_matches;
            c->status = 2; /* LV_TOO_FEW_MATCHES */
            c->done = 1;
            w->abort_ = 1;
        }
    }
    LV_PAR(i, n * n) w->P[i] = c->P_prop[i];
    LV_PAR(i, kStateLen) { w->x[i] = c->x[i]; w->xp[i] = c->x_prop[i]; }
    ex.sync();
    if (w->abort_) return;

    /* dx = x [-] x_prop ; J blocks (esekfom.hpp:1651-1697); three independent serial pieces */
    if (ex.tid == 0) {
        state_boxminus(w->x, w->xp, w->dx);
        for (int i = 0; i < n; ++i) w->dx_new[i] = w->dx[i];
    }
    ex.sync();
    LV_PAR(b, 3) {
        if (b < 2) {
            const int idx = 3 + 3 * b;
            const Mat3d Jt = mat3_transpose(A_matrix(load_vec3(w->dx + idx)));
            for (int i = 0; i < 9; ++i) w->J[b][i] = Jt.m[i];
            const Vec3d r = mat3_apply(Jt, load_vec3(w->dx + idx));
            store_vec3(w->dx_new + idx, r);
        } else {
            s2_J(load_vec3(w->x + kGrav), load_vec3(w->xp + kGrav), w->dx[21], w->dx[22], w->J[2]);
            const double a0 = w->J[2][0] * w->dx[21] + w->J[2][1] * w->dx[22];
            const double a1 = w->J[2][2] * w->dx[21] + w->J[2][3] * w->dx[22];
            w->dx_new[21] = a0;
            w->dx_new[22] = a1;
        }
    }
    ex.sync();
    for (int b = 0; b < 3; ++b) {
        const int idx = b < 2 ? 3 + 3 * b : 21, d = b < 2 ? 3 : 2;
        apply_rows(ex, w->P, w->P, idx, d, w->J[b], n, n);
        ex.sync();
        apply_cols(ex, w->P, idx, d, w->J[b], n, n);
        ex.sync();
    }

    /* Gain (esekfom.hpp:1722-1729).  The reference forms P_inv = ((P/R)^-1 + E^T Q E)^-1 with two 23x23
     * inverses (Q = HTH, E = [I12 0]) and uses only P_inv[:, :12].  With S = P/R the matrix-inversion
     * lemma gives  P_inv[:, :12] = S[:, :12] (I12 + Q S11)^-1  exactly, also for singular Q
     * (estimate_extrinsics = false).  So one 12x12 system with 13 right-hand sides replaces both
     * inverses:  Y = (I + Q S11)^-1 [Q | HTh],  K_x[:, :12] = S[:, :12] Y[:, :12],  K_h = S[:, :12] Y[:, 12].
     * Against exact arithmetic this is at least as accurate as the reference's formulation
     * ("gain formulation").                                                              */
    const double invR = 1.0 / prm.R;
    LV_PAR(it, 12 * 25) {
        const int i = it / 25, j = it - i * 25;
        double v;
        if (j < 12) {
            double s = 0;
            for (int k = 0; k < 12; ++k) s += w->HTH[i * 12 + k] * w->P[k * n + j];
            v = s * invR + (i == j ? 1.0 : 0.0);
        } else if (j < 24) {
            v = w->HTH[i * 12 + (j - 12)];
        } else {
            v = w->HTh[i];
        }
        w->M1[it] = v;
    }
    ex.sync();
    gj_solve(ex, w->M1, 12, 25, &w->piv, &w->pivval);
    LV_PAR(it, n * 13) {
        const int i = it / 13, j = it - i * 13;
        double s = 0;
        for (int k = 0; k < 12; ++k) s += w->P[i * n + k] * w->M1[k * 25 + 12 + j];
        s *= invR;
        if (j < 12) w->Kx[i * 12 + j] = s; else w->Kh[i] = s;
    }
    ex.sync();
    /* dx_ = K_h + (K_x - I) dx_new   (esekfom.hpp:1733) */
    LV_PAR(i, n) {
        double s = 0;
        for (int j = 0; j < 12; ++j) s += w->Kx[i * 12 + j] * w->dx_new[j];
        w->dxs[i] = w->Kh[i] + s - w->dx_new[i];
        w->dnd[i] = w->dxs[i];
    }
    ex.sync();

    {   /* log the reduced measurement of this evaluation (all threads) */
        IterLog* lgp = &c->logs[c->n_evals < kMaxEvals ? c->n_evals : kMaxEvals - 1];
        LV_PAR(i, 144) lgp->HTH[i] = w->HTH[i];
        LV_PAR(i, 12) lgp->HTh[i] = w->HTh[i];
    }
    if (ex.tid == 0) {
        IterLog* lg = &c->logs[c->n_evals < kMaxEvals ? c->n_evals : kMaxEvals - 1];
        /* degeneracy (esekfom.hpp:1736-1744): identity unless an eigenvalue of HTH[0:6,0:6] < D */
        lg->degenerate = 0;
        if (!all_eigs_above(w->HTH, prm.D)) {
            degenerate_mask(w->HTH, prm.D, w->dxs, w->dnd);
            lg->degenerate = 1;
        }
        state_boxplus(w->x, w->dnd);                               /* :1747 */
        int conv = 1;                                              /* :1748-1756 uses dx_ (pre-mask) */
        for (int i = 0; i < n; ++i)
            if (fabs(w->dxs[i]) > prm.limits[i]) { conv = 0; break; }
        if (conv) c->t += 1;                                       /* :1757 */
        if (!c->t && c->iter == prm.max_iter - 2) conv = 1;        /* :1759-1762 */
        lg->n_matches = w->n_matches;
        lg->converged = conv;
        for (int i = 0; i < n; ++i) lg->dx[i] = w->dxs[i];
        for (int i = 0; i < kStateLen; ++i) { lg->x_after[i] = w->x[i]; c->x[i] = w->x[i]; }
        c->n_evals += 1;
        w->finish = (c->t > 1 || c->iter == prm.max_iter - 1) ? 1 : 0;   /* :1764 */
        c->iter += 1;
        if (!w->finish) make_frame(w->x, &c->frame);
    }
    ex.sync();
    if (!w->finish) return;

    /* exit block, esekfom.hpp:1766-1817 */
    LV_PAR(i, n * n) w->L[i] = w->P[i];
    if (ex.tid == 0) {
        for (int b = 0; b < 2; ++b) {
            const Mat3d Jt = mat3_transpose(A_matrix(load_vec3(w->dxs + 3 + 3 * b)));
            for (int i = 0; i < 9; ++i) w->J[b][i] = Jt.m[i];
        }
        s2_J(load_vec3(w->x + kGrav), load_vec3(w->xp + kGrav), w->dxs[21], w->dxs[22], w->J[2]);
    }
    ex.sync();
    for (int b = 0; b < 3; ++b) {
        const int idx = b < 2 ? 3 + 3 * b : 21, d = b < 2 ? 3 : 2;
        apply_rows(ex, w->L, w->P, idx, d, w->J[b], n, n);
        apply_rows(ex, w->Kx, w->Kx, idx, d, w->J[b], 12, 12);
        ex.sync();
        apply_cols(ex, w->L, idx, d, w->J[b], n, n);
        apply_cols(ex, w->P, idx, d, w->J[b], n, n);
        ex.sync();
    }
    LV_PAR(it, n * n) {                                            /* :1817 */
        const int i = it / n, j = it - i * n;
        double s = 0;
        for (int k = 0; k < 12; ++k) s += w->Kx[i * 12 + k] * w->P[k * n + j];
        c->P[it] = w->L[it] - s;
    }
    if (ex.tid == 0) c->done = 1;
    ex.sync();
}

/* start of an update: x_prop = x, P_prop = P, loop counters (esekfom.hpp:1622-1634) */
template <class Ex>
LV_HD void ieskf_begin(Ex& ex, UpdateCtrl* c) {
    LV_PAR(i, kStateLen) c->x_prop[i] = c->x[i];
    LV_PAR(i, kN * kN) c->P_prop[i] = c->P[i];
    if (ex.tid == 0) {
        c->iter = -1;
        c->t = 0;
        c->done = 0;
        c->status = 0;
        c->n_evals = 0;
        make_frame(c->x, &c->frame);
    }
    ex.sync();
}

}  // namespace lv
#endif
