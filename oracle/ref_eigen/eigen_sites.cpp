/*
 * eigen_sites.cpp — TEST INFRASTRUCTURE, built only where Eigen3 is installed (oracle/build_with_eigen.sh).
 *
 * The reference's arithmetic on the hot path that lives in Eigen, at the reference's own call sites, behind a C ABI, so that
 * a reviewer with Eigen can pin what the oracle (and through it the CUDA path) can only restate in this image:
 *   src/Utils/Utils.cpp:34-54      A (5x3 fp32), b = -1:  A.colPivHouseholderQr().solve(b); normalise
 *   esekfom.hpp:1722,1726          (P_ / R).inverse(), then (.. + HTH block).inverse()        (23x23 fp64)
 *   esekfom.hpp:1736               Eigen::EigenSolver<Matrix<double,6,6>> of HTH[0:6,0:6]
 * The statements below are the reference's statements with its variable types (MatrixXf / Matrix<double,23,23>); nothing
 * else of the reference is needed for these three sites.  tests/test_eigen_sites.py uses the library when it exists.
 */
#include <Eigen/Dense>
#include <Eigen/Eigenvalues>

extern "C" {

/* R3Math::estimate_plane, Utils.cpp:32-57: pts = 5 x 3 row-major fp32 -> abcd (A, B, C, D) */
void eig_estimate_plane(const float* pts, int n, float* abcd) {
    Eigen::Matrix<float, Eigen::Dynamic, 3> A(n, 3);
    Eigen::Matrix<float, Eigen::Dynamic, 1> b(n, 1);
    A.setZero();
    b.setOnes();
    b *= -1.0f;
    for (int j = 0; j < n; ++j) { A(j, 0) = pts[3 * j]; A(j, 1) = pts[3 * j + 1]; A(j, 2) = pts[3 * j + 2]; }
    Eigen::Matrix<float, 3, 1> normvec = A.colPivHouseholderQr().solve(b);
    const float norm = normvec.norm();
    abcd[0] = normvec(0) / norm;
    abcd[1] = normvec(1) / norm;
    abcd[2] = normvec(2) / norm;
    abcd[3] = 1.0f / norm;
}

/* esekfom.hpp:1722-1729: P_temp = (P_/R).inverse(); P_temp.block<12,12>(0,0) += HTH; P_inv = P_temp.inverse();
 * K_h = P_inv.block<23,12>(0,0) * HTh; K_x.block<23,12>(0,0) = P_inv.block<23,12>(0,0) * HTH   (row-major in / out) */
void eig_gain(const double* P, double R, const double* HTH, const double* HTh, double* K_h, double* K_x12) {
    typedef Eigen::Matrix<double, 23, 23> Cov;
    Cov P_ = Eigen::Map<const Eigen::Matrix<double, 23, 23, Eigen::RowMajor>>(P);
    Eigen::Matrix<double, 12, 12> Q = Eigen::Map<const Eigen::Matrix<double, 12, 12, Eigen::RowMajor>>(HTH);
    Eigen::Matrix<double, 12, 1> h = Eigen::Map<const Eigen::Matrix<double, 12, 1>>(HTh);
    Cov P_temp = (P_ / R).inverse();
    P_temp.block<12, 12>(0, 0) += Q;
    Cov P_inv = P_temp.inverse();
    Eigen::Matrix<double, 23, 1> Kh = P_inv.block<23, 12>(0, 0) * h;
    Eigen::Matrix<double, 23, 12> Kx = P_inv.block<23, 12>(0, 0) * Q;
    for (int i = 0; i < 23; ++i) {
        K_h[i] = Kh(i);
        for (int j = 0; j < 12; ++j) K_x12[i * 12 + j] = Kx(i, j);
    }
}

/* esekfom.hpp:1736: real parts of the eigenpairs of HTH[0:6,0:6], in the order EigenSolver returns them */
void eig_eigensolver6(const double* A6, double* values, double* vectors) {
    Eigen::Matrix<double, 6, 6> A = Eigen::Map<const Eigen::Matrix<double, 6, 6, Eigen::RowMajor>>(A6);
    Eigen::EigenSolver<Eigen::Matrix<double, 6, 6>> es(A);
    for (int i = 0; i < 6; ++i) {
        values[i] = es.eigenvalues()(i).real();
        for (int j = 0; j < 6; ++j) vectors[i * 6 + j] = es.eigenvectors()(i, j).real();
    }
}

}  // extern "C"
