// oracle/shim/eigen_algorithms.h — the third-party ALGORITHMS the reference's sources call through Eigen
// (general eigenvalues of the 3x3 companion matrix of cvo.cpp:76-92 and of the 6x6 Hessian of cvo.cpp:726-758,
// the matrix logarithm of cvo.cpp:101, a general inverse), implemented in closed form with the methods the
// oracle documents for the same calls (oracle/cvo_oracle.cpp: cubic_real_roots, jacobi, dist_se3).  They stand
// in for Eigen's iterative solvers; agreement with Eigen itself is NOT claimed here.
#pragma once
#include <cmath>
#include <complex>

namespace Eigen {
namespace shim_detail {
inline int cubic_roots(double A, double B, double C, double re[3], double &pair_re, double &pair_im) {
    // monic t^3 + A t^2 + B t + C: one real root in closed form, polished; deflation; the remaining pair
    auto polish = [&](double t) {
        for (int it = 0; it < 4; it++) {
            double f = ((t + A) * t + B) * t + C;
            double fp = (3.0 * t + 2.0 * A) * t + B;
            if (fp == 0.0 || !std::isfinite(fp)) break;
            double tn = t - f / fp;
            if (!std::isfinite(tn)) break;
            t = tn;
        }
        return t;
    };
    double sq = A * A, p = (3.0 * B - sq) / 3.0, q = (2.0 * A * sq - 9.0 * A * B + 27.0 * C) / 27.0;
    double disc = q * q / 4.0 + p * p * p / 27.0, r;
    if (disc > 0) {
        double sd = std::sqrt(disc);
        r = std::cbrt(-q / 2.0 + sd) + std::cbrt(-q / 2.0 - sd) - A / 3.0;
    } else if (p == 0.0) {
        r = -A / 3.0;
    } else {
        double m = 2.0 * std::sqrt(-p / 3.0);
        double arg = std::max(-1.0, std::min(1.0, 3.0 * q / (p * m)));
        double th = std::acos(arg) / 3.0;
        r = 0;
        for (int k = 0; k < 3; k++) {
            double cand = m * std::cos(th - 2.0943951023931954923 * k) - A / 3.0;
            if (std::fabs(cand) >= std::fabs(r)) r = cand;
        }
    }
    r = polish(r);
    double b1, b0;
    if (r != 0.0 && std::fabs(r * r * r) >= std::fabs(C)) { b0 = -C / r; b1 = (b0 - B) / r; }
    else { b1 = A + r; b0 = B + r * b1; }
    int n = 0;
    re[n++] = r;
    double d2 = b1 * b1 - 4.0 * b0;
    if (d2 >= 0) {
        double qq = -0.5 * (b1 + (b1 >= 0 ? 1.0 : -1.0) * std::sqrt(d2));
        re[n++] = polish(qq);
        re[n++] = polish((qq != 0.0) ? b0 / qq : 0.0);
    } else {
        pair_re = -0.5 * b1;
        pair_im = 0.5 * std::sqrt(-d2);
    }
    return n;
}
inline void jacobi_sym(int n, double *a /* n x n, row major, symmetrised by the caller */, double *ev) {
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                if (a[p * n + q] == 0.0) continue;
                double th = (a[q * n + q] - a[p * n + p]) / (2.0 * a[p * n + q]);
                double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) { double akp = a[k * n + p], akq = a[k * n + q]; a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq; }
                for (int k = 0; k < n; k++) { double apk = a[p * n + k], aqk = a[q * n + k]; a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk; }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = a[i * n + i];
}
}  // namespace shim_detail

template <class T>
Matrix<std::complex<T>, Dynamic, 1> shim_eigenvalues(const Dyn<T> &m) {
    const int n = m.rows();
    Matrix<std::complex<T>, Dynamic, 1> out(n);
    bool companion = n == 3;
    if (companion)
        for (int i = 1; i < 3; i++) for (int j = 0; j < 3; j++) companion = companion && m(i, j) == (j == i - 1 ? T(1) : T(0));
    if (companion) {
        // the companion matrix of cvo::poly_solver (cvo.cpp:76-92): characteristic polynomial
        // t^3 - m00 t^2 - m01 t - m02, i.e. the monic cubic the float quotients -(coef / coef(0)) define
        double re[3], pr = 0, pi = 0;
        const double A = -(double)m(0, 0), B = -(double)m(0, 1), C = -(double)m(0, 2);
        int nr = 0;
        if (std::isfinite(A) && std::isfinite(B) && std::isfinite(C)) nr = shim_detail::cubic_roots(A, B, C, re, pr, pi);
        if (nr == 0) { for (int i = 0; i < 3; i++) out(i) = std::complex<T>(std::nan(""), std::nan("")); }   // inf / NaN matrix: no admissible root
        else if (nr == 3) { for (int i = 0; i < 3; i++) out(i) = std::complex<T>((T)re[i], T(0)); }
        else { out(0) = std::complex<T>((T)re[0], T(0)); out(1) = std::complex<T>((T)pr, (T)pi); out(2) = std::complex<T>((T)pr, (T)-pi); }
        return out;
    }
    // otherwise: the (numerically symmetric) 6x6 Hessian of se3_Hessian — cyclic Jacobi on the symmetrised matrix
    std::vector<double> a((size_t)n * n), ev(n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i * n + j] = 0.5 * ((double)m(i, j) + (double)m(j, i));
    shim_detail::jacobi_sym(n, a.data(), ev.data());
    for (int i = 0; i < n; i++) out(i) = std::complex<T>((T)ev[i], T(0));
    return out;
}

// log of a 4x4 rigid transform [R t; 0 1] (cvo::dist_se3, cvo.cpp:94-104): closed-form SE(3) logarithm in double
template <class T>
Dyn<T> shim_matrix_log(const Dyn<T> &m) {
    assert(m.rows() == 4 && m.cols() == 4);
    double r[3][3], t[3];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r[i][j] = m(i, j); t[i] = m(i, 3); }
    double ax = 0.5 * (r[2][1] - r[1][2]), ay = 0.5 * (r[0][2] - r[2][0]), az = 0.5 * (r[1][0] - r[0][1]);
    double s = std::sqrt(ax * ax + ay * ay + az * az), c = 0.5 * (r[0][0] + r[1][1] + r[2][2] - 1.0);
    double theta = std::atan2(s, c), wx, wy, wz;
    if (s < 1e-12) { wx = ax; wy = ay; wz = az; } else { double k = theta / s; wx = ax * k; wy = ay * k; wz = az * k; }
    double coef = theta < 1e-4 ? 1.0 / 12.0 : (1.0 - theta * std::sin(theta) / (2.0 * (1.0 - std::cos(theta)))) / (theta * theta);
    double wt[3] = {wy * t[2] - wz * t[1], wz * t[0] - wx * t[2], wx * t[1] - wy * t[0]};
    double wwt[3] = {wy * wt[2] - wz * wt[1], wz * wt[0] - wx * wt[2], wx * wt[1] - wy * wt[0]};
    Dyn<T> L = Dyn<T>::Zero(4, 4);
    L(0, 1) = (T)-wz; L(0, 2) = (T)wy; L(1, 0) = (T)wz; L(1, 2) = (T)-wx; L(2, 0) = (T)-wy; L(2, 1) = (T)wx;
    for (int i = 0; i < 3; i++) L(i, 3) = (T)(t[i] - 0.5 * wt[i] + coef * wwt[i]);
    return L;
}

// general inverse (Gauss-Jordan with partial pivoting, in the matrix's own scalar type)
template <class T>
Dyn<T> shim_inverse(const Dyn<T> &m) {
    const int n = m.rows();
    Dyn<T> a = m, inv = Dyn<T>::Identity(n, n);
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++) if (std::abs(a(r, c)) > std::abs(a(p, c))) p = r;
        for (int k = 0; k < n; k++) { std::swap(a(c, k), a(p, k)); std::swap(inv(c, k), inv(p, k)); }
        const T d = a(c, c);
        for (int k = 0; k < n; k++) { a(c, k) = a(c, k) / d; inv(c, k) = inv(c, k) / d; }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const T f = a(r, c);
            for (int k = 0; k < n; k++) { a(r, k) = a(r, k) - f * a(c, k); inv(r, k) = inv(r, k) - f * inv(c, k); }
        }
    }
    return inv;
}
}  // namespace Eigen
