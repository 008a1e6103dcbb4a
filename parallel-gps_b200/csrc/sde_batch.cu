// C ABI: pssgp_sde_dim, pssgp_sde_batch, pssgp_grid_loglik (see include/pssgp_b200.h).
//
// Native batched construction of the LTI SDE of a covariance function for B hyper-parameter settings at once — what
// the reference does one setting at a time in Python/TF with a numba + host round trip: get_sde of the Matern family
// (pssgp/kernels/matern/common.py:26-52, matern12.py:18-23, matern32.py:20-28, matern52.py:21-25), RBF
// (rbf.py:14-61, 78-101), Periodic (periodic.py:18-81), sums (kernels/base.py:151-183) and products (:199-244),
// balance_ss (math_utils.py:10-81) and solve_lyap_vec (math_utils.py:84-120).  Host code (a d x d problem per setting,
// microseconds each), spread over the host threads; the time-parallel work of every setting then runs on the GPU
// through pssgp_grid_loglik: discretise + filter + log-likelihood per setting, enqueued back to back from ONE call.
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <array>
#include <complex>
#include <mutex>
#include <thread>
#include <vector>

#include "scan_run.cuh"
#include "workspace.h"

namespace pssgp {
namespace sde {

// ---- scalar types: double, or a forward-mode dual number carrying the derivatives w.r.t. NP hyper-parameters -----
template <int NP> struct Dual {
    double v;
    double g[NP];
    Dual() : v(0.0) { for (int i = 0; i < NP; ++i) g[i] = 0.0; }
    Dual(double x) : v(x) { for (int i = 0; i < NP; ++i) g[i] = 0.0; }
};
inline double val(double x) { return x; }
template <int NP> inline double val(const Dual<NP>& x) { return x.v; }
inline bool is_zero(double x) { return x == 0.0; }
template <int NP> inline bool is_zero(const Dual<NP>& x) {
    if (x.v != 0.0) return false;
    for (int i = 0; i < NP; ++i)
        if (x.g[i] != 0.0) return false;
    return true;
}
#define SDE_DUAL_BIN(OP, VEXPR, GEXPR)                                                            \
    template <int NP> inline Dual<NP> operator OP(const Dual<NP>& a, const Dual<NP>& b) {        \
        Dual<NP> r;                                                                               \
        r.v = VEXPR;                                                                              \
        for (int i = 0; i < NP; ++i) r.g[i] = GEXPR;                                              \
        return r;                                                                                 \
    }                                                                                             \
    template <int NP> inline Dual<NP> operator OP(const Dual<NP>& a, double b) { return a OP Dual<NP>(b); } \
    template <int NP> inline Dual<NP> operator OP(double a, const Dual<NP>& b) { return Dual<NP>(a) OP b; }
SDE_DUAL_BIN(+, a.v + b.v, a.g[i] + b.g[i])
SDE_DUAL_BIN(-, a.v - b.v, a.g[i] - b.g[i])
SDE_DUAL_BIN(*, a.v * b.v, a.g[i] * b.v + a.v * b.g[i])
SDE_DUAL_BIN(/, a.v / b.v, (a.g[i] - (a.v / b.v) * b.g[i]) / b.v)
#undef SDE_DUAL_BIN
template <int NP> inline Dual<NP> operator-(const Dual<NP>& a) { return Dual<NP>(0.0) - a; }
template <int NP, class B> inline Dual<NP>& operator+=(Dual<NP>& a, const B& b) { a = a + b; return a; }
template <int NP, class B> inline Dual<NP>& operator-=(Dual<NP>& a, const B& b) { a = a - b; return a; }
template <int NP, class B> inline Dual<NP>& operator*=(Dual<NP>& a, const B& b) { a = a * b; return a; }
template <int NP, class B> inline Dual<NP>& operator/=(Dual<NP>& a, const B& b) { a = a / b; return a; }
inline double s_sqrt(double x) { return sqrt(x); }
inline double s_exp(double x) { return exp(x); }
inline double s_abs(double x) { return fabs(x); }
inline double s_powi(double x, int k) { return pow(x, k); }
template <int NP> inline Dual<NP> s_sqrt(const Dual<NP>& x) {
    Dual<NP> r(sqrt(x.v));
    for (int i = 0; i < NP; ++i) r.g[i] = 0.5 * x.g[i] / r.v;
    return r;
}
template <int NP> inline Dual<NP> s_exp(const Dual<NP>& x) {
    Dual<NP> r(exp(x.v));
    for (int i = 0; i < NP; ++i) r.g[i] = r.v * x.g[i];
    return r;
}
template <int NP> inline Dual<NP> s_abs(const Dual<NP>& x) { return x.v < 0.0 ? -x : x; }
template <int NP> inline Dual<NP> s_powi(const Dual<NP>& x, int k) {
    Dual<NP> r(pow(x.v, k));
    const double dv = k == 0 ? 0.0 : k * pow(x.v, k - 1);
    for (int i = 0; i < NP; ++i) r.g[i] = dv * x.g[i];
    return r;
}

template <class S> using MatT = std::vector<S>;  // dense row-major
typedef MatT<double> Mat;

template <class S> struct SdeT {
    int d = 0, r = 0;           // state dimension, noise dimension
    MatT<S> P0, F, L, H, Q;     // P0 [d,d], F [d,d], L [d,r], H [d], Q [r,r]
};
typedef SdeT<double> Sde;

template <class S> static MatT<S> eye(int n) {
    MatT<S> I((size_t)n * n, S(0.0));
    for (int i = 0; i < n; ++i) I[(size_t)i * n + i] = S(1.0);
    return I;
}
// C[m,n] = A[m,k] B[k,n]
template <class S> static MatT<S> matmul(const MatT<S>& A, const MatT<S>& B, int m, int k, int n) {
    MatT<S> C((size_t)m * n, S(0.0));
    for (int i = 0; i < m; ++i)
        for (int p = 0; p < k; ++p) {
            const S a = A[(size_t)i * k + p];
            if (is_zero(a)) continue;
            for (int j = 0; j < n; ++j) C[(size_t)i * n + j] += a * B[(size_t)p * n + j];
        }
    return C;
}
template <class S> static MatT<S> transpose(const MatT<S>& A, int m, int n) {
    MatT<S> T((size_t)m * n);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) T[(size_t)j * m + i] = A[(size_t)i * n + j];
    return T;
}
// kron(A[ma,na], B[mb,nb])
template <class S> static MatT<S> kron(const MatT<S>& A, int ma, int na, const MatT<S>& B, int mb, int nb) {
    MatT<S> K((size_t)ma * mb * na * nb, S(0.0));
    const size_t ld = (size_t)na * nb;
    for (int i = 0; i < ma; ++i)
        for (int j = 0; j < na; ++j) {
            const S a = A[(size_t)i * na + j];
            if (is_zero(a)) continue;
            for (int p = 0; p < mb; ++p)
                for (int q = 0; q < nb; ++q) K[((size_t)i * mb + p) * ld + (size_t)j * nb + q] = a * B[(size_t)p * nb + q];
        }
    return K;
}
// L Q L^T  ([d,d])
template <class S> static MatT<S> lqlt(const SdeT<S>& s) {
    MatT<S> LQ = matmul<S>(s.L, s.Q, s.d, s.r, s.r);
    return matmul<S>(LQ, transpose<S>(s.L, s.d, s.r), s.d, s.r, s.d);
}

// The row update is the hot loop of the elimination (n^3 / 3 flops): compiled for AVX2 + FMA where the host has it.
__attribute__((target_clones("avx2,fma", "default"))) static void row_axpy(double* __restrict__ rr, const double* __restrict__ rc,
                                                                            double f, int lo, int n) {
    for (int j = lo; j < n; ++j) rr[j] -= f * rc[j];
}
template <int NP> static void row_axpy(Dual<NP>* rr, const Dual<NP>* rc, const Dual<NP>& f, int lo, int n) {
    for (int j = lo; j < n; ++j) rr[j] -= f * rc[j];
}
// Solve A x = b in place (A [n,n] destroyed, b -> x), Gaussian elimination with partial pivoting (on the values); zero
// multipliers are skipped (the Lyapunov systems below are mostly zeros).  Returns false for a singular matrix.
template <class S> static bool solve_inplace(MatT<S>& A, MatT<S>& b, int n) {
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = fabs(val(A[(size_t)c * n + c]));
        for (int r = c + 1; r < n; ++r) {
            const double v = fabs(val(A[(size_t)r * n + c]));
            if (v > best) best = v, piv = r;
        }
        if (!(best > 0.0)) return false;
        if (piv != c) {
            for (int j = c; j < n; ++j) std::swap(A[(size_t)c * n + j], A[(size_t)piv * n + j]);
            std::swap(b[c], b[piv]);
        }
        const S inv = S(1.0) / A[(size_t)c * n + c];
        for (int r = c + 1; r < n; ++r) {
            const S f = A[(size_t)r * n + c] * inv;
            if (is_zero(f)) continue;
            row_axpy(&A[(size_t)r * n], &A[(size_t)c * n], f, c + 1, n);
            b[r] -= f * b[c];
        }
    }
    for (int c = n - 1; c >= 0; --c) {
        S v = b[c];
        const S* rc = &A[(size_t)c * n];
        for (int j = c + 1; j < n; ++j) v -= rc[j] * b[j];
        b[c] = v / rc[c];
    }
    return true;
}

// X with F X + X F^T = G (math_utils.py:108-118 solves the d^2 x d^2 Kronecker system kron(I,F) + kron(F,I)).
// For a symmetric right-hand side the solution is symmetric, and the system is assembled directly in the
// d(d+1)/2 unknowns of its lower triangle: the same equations, about 8x fewer elimination flops.
template <class S> static bool lyap_solve(const MatT<S>& F, const MatT<S>& G, int d, MatT<S>& X) {
    double gmax = 0.0, asym = 0.0;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            gmax = std::max(gmax, fabs(val(G[(size_t)i * d + j])));
            asym = std::max(asym, fabs(val(G[(size_t)i * d + j]) - val(G[(size_t)j * d + i])));
        }
    X.assign((size_t)d * d, S(0.0));
    if (asym <= 1e-14 * gmax) {
        const int m = d * (d + 1) / 2;
        auto idx = [](int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; };
        MatT<S> op((size_t)m * m, S(0.0)), rhs((size_t)m);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j <= i; ++j) {
                S* row = &op[(size_t)idx(i, j) * m];
                for (int k = 0; k < d; ++k) {
                    row[idx(k, j)] += F[(size_t)i * d + k];  // (F X)_ij
                    row[idx(i, k)] += F[(size_t)j * d + k];  // (X F^T)_ij
                }
                rhs[idx(i, j)] = 0.5 * (G[(size_t)i * d + j] + G[(size_t)j * d + i]);
            }
        if (!solve_inplace<S>(op, rhs, m)) return false;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) X[(size_t)i * d + j] = rhs[idx(i, j)];
        return true;
    }
    MatT<S> I = eye<S>(d);
    MatT<S> op = kron<S>(I, d, d, F, d, d);
    MatT<S> op2 = kron<S>(F, d, d, I, d, d);
    for (size_t i = 0; i < op.size(); ++i) op[i] += op2[i];
    MatT<S> rhs = G;
    if (!solve_inplace<S>(op, rhs, d * d)) return false;
    X = rhs;
    return true;
}

// math_utils.py:84-120: F P + P F^T + L Q L^T = 0, P = -sym(X)
template <class S> static bool solve_lyap_vec(const SdeT<S>& s, MatT<S>& P) {
    const int d = s.d;
    MatT<S> X;
    if (!lyap_solve<S>(s.F, lqlt<S>(s), d, X)) return false;
    P.assign((size_t)d * d, S(0.0));
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) P[(size_t)i * d + j] = -0.5 * (X[(size_t)i * d + j] + X[(size_t)j * d + i]);
    return true;
}

// math_utils.py:32-81 (balance_ss) with the scaling of math_utils.py:10-29 (pssgp_balance_ss, host_utils.cu).  The
// scaling vector is computed from the VALUES of F and is a constant w.r.t. differentiation, as in the reference
// (it crosses tf.numpy_function, math_utils.py:68); the two max-abs normalisations differentiate through their argmax.
template <class S> static void balance_ss(SdeT<S>& s, int n_iter) {
    const int d = s.d;
    std::vector<double> dv(d), Fv((size_t)d * d);
    for (size_t i = 0; i < Fv.size(); ++i) Fv[i] = val(s.F[i]);
    pssgp_balance_ss(Fv.data(), d, n_iter, dv.data());
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) s.F[(size_t)i * d + j] = s.F[(size_t)i * d + j] * dv[j] / dv[i];
    size_t a3 = 0, a4 = 0;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < s.r; ++j) {
            s.L[(size_t)i * s.r + j] /= dv[i];
            if (fabs(val(s.L[(size_t)i * s.r + j])) > fabs(val(s.L[a3]))) a3 = (size_t)i * s.r + j;
        }
    for (int i = 0; i < d; ++i) {
        s.H[i] *= dv[i];
        if (fabs(val(s.H[i])) > fabs(val(s.H[a4]))) a4 = i;
    }
    const S t3 = s_abs(s.L[a3]), t4 = s_abs(s.H[a4]);
    for (auto& v : s.L) v /= t3;
    for (auto& v : s.H) v /= t4;
    for (auto& v : s.Q) v *= (t3 * t3) * (t4 * t4);
}

static double factorial(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= i;
    return f;
}
static double binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double b = 1.0;
    for (int i = 1; i <= k; ++i) b = b * (n - k + i) / i;
    return floor(b + 0.5);
}

// matern/common.py:26-52
template <class S> static SdeT<S> matern_companion(int d, const S& variance, const S& ell) {
    SdeT<S> s;
    s.d = d, s.r = 1;
    const S lam = sqrt(2.0 * d - 1.0) / ell;
    s.F.assign((size_t)d * d, S(0.0));
    for (int i = 0; i + 1 < d; ++i) s.F[(size_t)i * d + i + 1] = S(1.0);
    for (int k = 0; k < d; ++k) s.F[(size_t)(d - 1) * d + k] = -binom(d, k) * s_powi(lam, d - k);
    s.L.assign(d, S(0.0));
    s.L[d - 1] = S(1.0);
    s.H.assign(d, S(0.0));
    s.H[0] = S(1.0);
    s.Q.assign(1, s_powi(2.0 * lam, 2 * d - 1) * variance * factorial(d - 1) * factorial(d - 1) / factorial(2 * d - 2));
    return s;
}

// Roots of g(z) = sum_k a[k] z^k (degree m, real coefficients) by the Aberth-Ehrlich iteration.
static std::vector<std::complex<double>> poly_roots(const std::vector<double>& a) {
    typedef std::complex<double> C;
    const int m = (int)a.size() - 1;
    std::vector<C> z(m);
    const double radius = pow(fabs(a[0] / a[m]), 1.0 / m);
    for (int j = 0; j < m; ++j) z[j] = std::polar(radius, 2.0 * M_PI * j / m + 0.4);
    for (int it = 0; it < 2000; ++it) {
        double worst = 0.0;
        for (int j = 0; j < m; ++j) {
            C p = a[m], dp = 0.0;
            for (int k = m - 1; k >= 0; --k) {
                dp = dp * z[j] + p;
                p = p * z[j] + a[k];
            }
            const C newton = p / dp;
            C rep = 0.0;
            for (int l = 0; l < m; ++l)
                if (l != j) rep += 1.0 / (z[j] - z[l]);
            const C step = newton / (1.0 - newton * rep);
            z[j] -= step;
            worst = std::max(worst, std::abs(step) / std::max(std::abs(z[j]), 1e-300));
        }
        if (worst < 1e-16) break;
    }
    return z;
}

// rbf.py:14-61.  The denominator polynomial is the order-`order` Taylor polynomial of exp(-s^2/2): even in s, so
// its roots are the square roots of the roots of g(z) = sum_k (-z/2)^k / k!; the stable half gives the drift.
// Independent of the hyper-parameters (plain doubles).
static void rbf_unscaled(int order, Mat& F, double& gain, double& q) {
    typedef std::complex<double> C;
    std::vector<double> a(order + 1);
    for (int k = 0; k <= order; ++k) a[k] = pow(-0.5, k) / factorial(k);
    std::vector<C> zr = poly_roots(a);
    std::vector<C> stable;
    for (const C& z : zr) {
        C s = std::sqrt(z);
        if (s.real() > 0.0) s = -s;
        stable.push_back(s);
    }
    // conjugate pairs next to each other, then the monic polynomial prod (s - r_j), highest power first
    std::sort(stable.begin(), stable.end(), [](const C& x, const C& y) {
        if (x.real() != y.real()) return x.real() < y.real();
        return x.imag() < y.imag();
    });
    std::vector<C> poly(1, C(1.0));
    for (const C& r : stable) {
        std::vector<C> np(poly.size() + 1, C(0.0));
        for (size_t i = 0; i < poly.size(); ++i) {
            np[i] += poly[i];
            np[i + 1] -= poly[i] * r;
        }
        poly.swap(np);
    }
    const int n = order;
    std::vector<double> denom(n + 1);
    for (int i = 0; i <= n; ++i) denom[i] = poly[i].real();
    // rbf.py:40-43: normalise the constant term to one, the gain is the inverse leading coefficient, monic again
    const double c0 = denom[n];
    for (auto& v : denom) v /= c0;
    gain = 1.0 / denom[0];
    const double lead = denom[0];
    for (auto& v : denom) v /= lead;
    F.assign((size_t)n * n, 0.0);
    for (int i = 0; i + 1 < n; ++i) F[(size_t)i * n + i + 1] = 1.0;
    for (int j = 0; j < n; ++j) F[(size_t)(n - 1) * n + j] = -denom[n - j];
    q = sqrt(2.0 * M_PI);  // rbf.py:24: sqrt(2 pi) / coeffs[-1], the constant Taylor coefficient being 1
}

enum { K_MATERN12 = 0, K_MATERN32 = 1, K_MATERN52 = 2, K_RBF = 3, K_PERIODIC = 4 };

static int base_dim(int type, int order) {
    switch (type) {
        case K_MATERN12: return 1;
        case K_MATERN32: return 2;
        case K_MATERN52: return 3;
        case K_RBF: return order;
        case K_PERIODIC: return 2 * (order + 1);
    }
    return -1;
}
static int base_nparams(int type) { return type == K_PERIODIC ? 3 : 2; }

struct RbfCache {  // the unscaled RBF SDE depends on the order only: computed once per call, shared by the settings
    int order = -1;
    Mat F;
    double gain = 0.0, q = 0.0;
};

// one base kernel; p = (variance, lengthscale[, period])
template <class S>
static bool base_sde(int type, int order, int bal_iter, const S* p, const std::vector<RbfCache>& rbf, SdeT<S>& s) {
    const S variance = p[0], ell = p[1];
    switch (type) {
        case K_MATERN12: {  // matern12.py:18-23
            s = matern_companion<S>(1, variance, ell);
            s.P0.assign(1, variance);
            return true;
        }
        case K_MATERN32: {  // matern32.py:20-28
            s = matern_companion<S>(2, variance, ell);
            const S lam = sqrt(3.0) / ell;
            s.P0 = {variance, S(0.0), S(0.0), lam * lam * variance};
            return true;
        }
        case K_MATERN52: {  // matern52.py:21-25
            s = matern_companion<S>(3, variance, ell);
            balance_ss<S>(s, bal_iter);
            return solve_lyap_vec<S>(s, s.P0);
        }
        case K_RBF: {  // rbf.py:78-101
            const RbfCache* c = nullptr;
            for (const auto& e : rbf)
                if (e.order == order) c = &e;
            if (!c) return false;
            const int n = order;
            s.d = n, s.r = 1;
            s.F.assign((size_t)n * n, S(0.0));
            for (size_t i = 0; i < s.F.size(); ++i) s.F[i] = S(c->F[i]);
            for (int j = 0; j < n; ++j) s.F[(size_t)(n - 1) * n + j] /= s_powi(ell, n - j);
            s.L.assign(n, S(0.0));
            s.L[n - 1] = S(1.0);
            s.H.assign(n, S(0.0));
            s.H[0] = c->gain / s_powi(ell, n);
            s.Q.assign(1, variance * ell * c->q);
            balance_ss<S>(s, bal_iter);
            return solve_lyap_vec<S>(s, s.P0);
        }
        case K_PERIODIC: {  // periodic.py:18-81
            const int N = order, dim = 2 * (N + 1);
            const S w0 = 2.0 * M_PI / p[2], l2 = 2.0 * ell;  // periodic.py:57: lengthscale x 2
            s.d = dim, s.r = dim;
            s.F.assign((size_t)dim * dim, S(0.0));
            for (int j = 0; j <= N; ++j) {
                s.F[(size_t)(2 * j) * dim + 2 * j + 1] = -w0 * (double)j;
                s.F[(size_t)(2 * j + 1) * dim + 2 * j] = w0 * (double)j;
            }
            s.L = eye<S>(dim);
            s.Q.assign((size_t)dim * dim, S(0.0));
            s.P0.assign((size_t)dim * dim, S(0.0));
            const S il2 = 1.0 / (l2 * l2);
            for (int J = 0; J <= N; ++J) {
                S q2(0.0);  // sum over K of b(K,J) l^-2K / K! exp(-l^-2) 2^-K variance
                for (int K = J; K <= N; K += 2) {
                    const double b = 2.0 * binom(K, (K - J) / 2) / (J == 0 ? 2.0 : 1.0);
                    q2 += b * s_powi(il2, K) / factorial(K) * s_exp(-il2) * pow(2.0, -K) * variance;
                }
                s.P0[(size_t)(2 * J) * dim + 2 * J] = q2;
                s.P0[(size_t)(2 * J + 1) * dim + 2 * J + 1] = q2;
            }
            s.H.assign(dim, S(0.0));
            for (int j = 0; j <= N; ++j) s.H[2 * j] = S(1.0);
            return true;
        }
    }
    return false;
}

// kernels/base.py:199-220 (Kronecker-sum drift, product diffusion and stationary covariance of two factors)
template <class S> static SdeT<S> product2(const SdeT<S>& a, const SdeT<S>& b) {
    SdeT<S> s;
    s.d = a.d * b.d, s.r = s.d;
    MatT<S> Ia = eye<S>(a.d), Ib = eye<S>(b.d);
    s.F = kron<S>(a.F, a.d, a.d, Ib, b.d, b.d);
    MatT<S> t = kron<S>(Ia, a.d, a.d, b.F, b.d, b.d);
    for (size_t i = 0; i < t.size(); ++i) s.F[i] += t[i];
    MatT<S> g1 = lqlt<S>(a), g2 = lqlt<S>(b);
    s.Q = kron<S>(g1, a.d, a.d, b.P0, b.d, b.d);
    t = kron<S>(a.P0, a.d, a.d, g2, b.d, b.d);
    for (size_t i = 0; i < t.size(); ++i) s.Q[i] += t[i];
    s.H = kron<S>(a.H, 1, a.d, b.H, 1, b.d);
    s.P0 = kron<S>(a.P0, a.d, a.d, b.P0, b.d, b.d);
    s.L = eye<S>(s.d);
    return s;
}

// spec: [combine_balancing_iter, n_terms, {n_factors, {type, order, balancing_iter} x n_factors} x n_terms]
struct Spec {
    int comb_iter = 0;
    std::vector<std::vector<std::array<int, 3>>> terms;
    int d = 0, nparams = 0;
};

static bool parse_spec(const int32_t* spec, int len, Spec& out) {
    if (!spec || len < 2) return false;
    int pos = 0;
    out.comb_iter = spec[pos++];
    const int nt = spec[pos++];
    if (nt < 1 || out.comb_iter < 0) return false;
    out.d = 0, out.nparams = 0;
    for (int t = 0; t < nt; ++t) {
        if (pos >= len) return false;
        const int nf = spec[pos++];
        if (nf < 1 || pos + 3 * nf > len) return false;
        std::vector<std::array<int, 3>> fs;
        int td = 1;
        for (int f = 0; f < nf; ++f) {
            std::array<int, 3> e = {spec[pos], spec[pos + 1], spec[pos + 2]};
            pos += 3;
            const int bd = base_dim(e[0], e[1]);
            if (bd < 1 || e[2] < 0) return false;
            td *= bd;
            out.nparams += base_nparams(e[0]);
            fs.push_back(e);
        }
        out.d += td;
        out.terms.push_back(fs);
    }
    return pos == len;
}

template <class S>
static bool build_one(const Spec& sp, const std::vector<RbfCache>& rbf, const S* params, SdeT<S>& out) {
    std::vector<SdeT<S>> terms;
    const S* p = params;
    for (const auto& fs : sp.terms) {
        SdeT<S> term;
        for (size_t f = 0; f < fs.size(); ++f) {
            SdeT<S> b;
            if (!base_sde<S>(fs[f][0], fs[f][1], fs[f][2], p, rbf, b)) return false;
            p += base_nparams(fs[f][0]);
            term = f == 0 ? b : product2<S>(term, b);
        }
        if (fs.size() > 1) {  // kernels/base.py:236-244: the product is balanced and its Pinf solved again
            balance_ss<S>(term, sp.comb_iter);
            if (!solve_lyap_vec<S>(term, term.P0)) return false;
        }
        terms.push_back(std::move(term));
    }
    if (terms.size() == 1) {
        out = std::move(terms[0]);
        return true;
    }
    // kernels/base.py:151-183: block-diagonal sum, balanced, Pinf solved for the whole system
    SdeT<S> s;
    for (const auto& t : terms) s.d += t.d, s.r += t.r;
    s.F.assign((size_t)s.d * s.d, S(0.0));
    s.L.assign((size_t)s.d * s.r, S(0.0));
    s.Q.assign((size_t)s.r * s.r, S(0.0));
    s.H.assign(s.d, S(0.0));
    int od = 0, orr = 0;
    for (const auto& t : terms) {
        for (int i = 0; i < t.d; ++i) {
            for (int j = 0; j < t.d; ++j) s.F[(size_t)(od + i) * s.d + od + j] = t.F[(size_t)i * t.d + j];
            for (int j = 0; j < t.r; ++j) s.L[(size_t)(od + i) * s.r + orr + j] = t.L[(size_t)i * t.r + j];
            s.H[od + i] = t.H[i];
        }
        for (int i = 0; i < t.r; ++i)
            for (int j = 0; j < t.r; ++j) s.Q[(size_t)(orr + i) * s.r + orr + j] = t.Q[(size_t)i * t.r + j];
        od += t.d, orr += t.r;
    }
    balance_ss<S>(s, sp.comb_iter);
    if (!solve_lyap_vec<S>(s, s.P0)) return false;
    out = std::move(s);
    return true;
}

// the tables are functions of the order only: computed once per process (the root iteration costs ~2 ms)
static std::vector<RbfCache> rbf_tables(const Spec& sp) {
    static std::mutex mu;
    static std::vector<RbfCache> known;
    std::lock_guard<std::mutex> lock(mu);
    std::vector<RbfCache> rbf;
    for (const auto& fs : sp.terms)
        for (const auto& f : fs)
            if (f[0] == K_RBF && std::none_of(rbf.begin(), rbf.end(), [&](const RbfCache& c) { return c.order == f[1]; })) {
                auto it = std::find_if(known.begin(), known.end(), [&](const RbfCache& c) { return c.order == f[1]; });
                if (it == known.end()) {
                    RbfCache c;
                    c.order = f[1];
                    rbf_unscaled(c.order, c.F, c.gain, c.q);
                    known.push_back(c);
                    it = known.end() - 1;
                }
                rbf.push_back(*it);
            }
    return rbf;
}

// values and Jacobians w.r.t. the NP >= nparams hyper-parameters of one setting (forward mode: one dual build)
template <int NP>
static bool build_jac(const Spec& sp, const std::vector<RbfCache>& rbf, const double* params, double* F, double* Pinf,
                      double* H, double* dF, double* dPinf, double* dH) {
    typedef Dual<NP> S;
    std::vector<S> p(sp.nparams);
    for (int i = 0; i < sp.nparams; ++i) {
        p[i] = S(params[i]);
        p[i].g[i] = 1.0;
    }
    SdeT<S> s;
    if (!build_one<S>(sp, rbf, p.data(), s) || s.d != sp.d) return false;
    const int d = sp.d, dd = d * d;
    for (int e = 0; e < dd; ++e) {
        F[e] = s.F[e].v;
        Pinf[e] = s.P0[e].v;
        for (int q = 0; q < sp.nparams; ++q) {
            dF[(size_t)q * dd + e] = s.F[e].g[q];
            dPinf[(size_t)q * dd + e] = s.P0[e].g[q];
        }
    }
    for (int e = 0; e < d; ++e) {
        H[e] = s.H[e].v;
        for (int q = 0; q < sp.nparams; ++q) dH[(size_t)q * d + e] = s.H[e].g[q];
    }
    return true;
}

}  // namespace sde
}  // namespace pssgp

using namespace pssgp;

extern "C" {

int pssgp_sde_dim(const int32_t* spec, int spec_len, int* d_out, int* nparams_out) {
    sde::Spec sp;
    if (!sde::parse_spec(spec, spec_len, sp)) return set_err(PSSGP_ERR_INVALID, "sde: malformed kernel spec");
    if (d_out) *d_out = sp.d;
    if (nparams_out) *nparams_out = sp.nparams;
    return PSSGP_OK;
}

int pssgp_lyap_solve(const void* F, const void* G, int d, void* X) {
    if (!F || !G || !X || d < 1) return set_err(PSSGP_ERR_INVALID, "lyap_solve: bad argument");
    const double *Fp = (const double*)F, *Gp = (const double*)G;
    sde::Mat Fm(Fp, Fp + (size_t)d * d), Gm(Gp, Gp + (size_t)d * d), Xm;
    if (!sde::lyap_solve<double>(Fm, Gm, d, Xm))
        return set_err(PSSGP_ERR_INVALID, "lyap_solve: singular system (F and -F share an eigenvalue)");
    std::copy(Xm.begin(), Xm.end(), (double*)X);
    return PSSGP_OK;
}

int pssgp_sde_batch(const int32_t* spec, int spec_len, int64_t batch, const double* params, int64_t params_stride,
                    double* F, double* Pinf, double* H, int nthreads) {
    sde::Spec sp;
    if (!sde::parse_spec(spec, spec_len, sp)) return set_err(PSSGP_ERR_INVALID, "sde: malformed kernel spec");
    if (batch < 1 || !params || !F || !Pinf || !H || params_stride < sp.nparams)
        return set_err(PSSGP_ERR_INVALID, "sde_batch: bad argument");
    const std::vector<sde::RbfCache> rbf = sde::rbf_tables(sp);
    const int d = sp.d;
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 64), batch));
    std::vector<int> failed(nt, 0);
    auto work = [&](int tid) {
        for (int64_t b = tid; b < batch; b += nt) {
            sde::Sde s;
            if (!sde::build_one<double>(sp, rbf, params + b * params_stride, s) || s.d != d) {
                failed[tid] = 1;
                continue;
            }
            std::copy(s.F.begin(), s.F.end(), F + b * d * d);
            std::copy(s.P0.begin(), s.P0.end(), Pinf + b * d * d);
            std::copy(s.H.begin(), s.H.end(), H + b * d);
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    for (int f : failed)
        if (f) return set_err(PSSGP_ERR_INVALID, "sde_batch: singular Lyapunov system or unsupported kernel in a setting");
    return PSSGP_OK;
}

int pssgp_sde_batch_jac(const int32_t* spec, int spec_len, int64_t batch, const double* params, int64_t params_stride,
                        double* F, double* Pinf, double* H, double* dF, double* dPinf, double* dH, int nthreads) {
    sde::Spec sp;
    if (!sde::parse_spec(spec, spec_len, sp)) return set_err(PSSGP_ERR_INVALID, "sde: malformed kernel spec");
    if (batch < 1 || !params || !F || !Pinf || !H || !dF || !dPinf || !dH || params_stride < sp.nparams)
        return set_err(PSSGP_ERR_INVALID, "sde_batch_jac: bad argument");
    if (sp.nparams > 16) return set_err(PSSGP_ERR_UNSUPPORTED, "sde_batch_jac: at most 16 hyper-parameters (got %d)", sp.nparams);
    const std::vector<sde::RbfCache> rbf = sde::rbf_tables(sp);
    const int d = sp.d, np = sp.nparams;
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 64), batch));
    std::vector<int> failed(nt, 0);
    auto work = [&](int tid) {
        for (int64_t b = tid; b < batch; b += nt) {
            const double* pb = params + b * params_stride;
            double *Fb = F + b * d * d, *Pb = Pinf + b * d * d, *Hb = H + b * d;
            double *dFb = dF + b * np * d * d, *dPb = dPinf + b * np * d * d, *dHb = dH + b * np * d;
            bool ok;
            if (np <= 4) ok = sde::build_jac<4>(sp, rbf, pb, Fb, Pb, Hb, dFb, dPb, dHb);
            else if (np <= 8) ok = sde::build_jac<8>(sp, rbf, pb, Fb, Pb, Hb, dFb, dPb, dHb);
            else ok = sde::build_jac<16>(sp, rbf, pb, Fb, Pb, Hb, dFb, dPb, dHb);
            if (!ok) failed[tid] = 1;
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    for (int f : failed)
        if (f) return set_err(PSSGP_ERR_INVALID, "sde_batch_jac: singular Lyapunov system or unsupported kernel in a setting");
    return PSSGP_OK;
}

// Settings are independent: they are dealt out to kGridLanes concurrent lanes (child handle = own workspace, own
// stream), so that the latency-bound aggregate hierarchy of one setting overlaps the SM-filling kernels of the others.
static int grid_lanes_init(pssgp_handle* h, int lanes) {
    cudaError_t e = cudaSuccess;
    if (!h->fork_event) e = cudaEventCreateWithFlags((cudaEvent_t*)&h->fork_event, cudaEventDisableTiming);
    for (int i = 0; i < lanes && e == cudaSuccess; ++i) {
        if (!h->lane[i]) {
            int rc = pssgp_create(&h->lane[i], h->device);
            if (rc) return rc;
        }
        if (!h->lane_stream[i]) e = cudaStreamCreateWithFlags((cudaStream_t*)&h->lane_stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess && !h->lane_event[i])
            e = cudaEventCreateWithFlags((cudaEvent_t*)&h->lane_event[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "grid_loglik: %s", cudaGetErrorString(e));
    return PSSGP_OK;
}

static int grid_impl(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, const void* F, const void* Pinf,
                     const void* H, const void* R, const void* dts, const void* y, void* ll, void* dF, void* dPinf,
                     void* dP0, void* dH, void* dR, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    const bool grad = dF != nullptr;
    if (batch < 1 || !F || !Pinf || !H || !R || !dts || !y || !ll || (grad && (!dPinf || !dP0 || !dH || !dR)))
        return set_err(PSSGP_ERR_INVALID, "grid_loglik: bad argument");
    const size_t es = dtype == PSSGP_F64 ? 8 : 4;
    const size_t nm = (size_t)n * d * d * es, nv = (size_t)n * d * es;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t need = (grad ? 5 : 3) * up(nm) + up(nv) + 256;
    // per-kernel timing (option "timing") brackets launches with events on ONE stream: a single lane then
    const int lanes = h->timing ? 1 : (int)std::min<int64_t>(h->grid_lanes > 0 ? h->grid_lanes : 4, std::min<int64_t>(batch, 8));
    if (lanes > 1 && (rc = grid_lanes_init(h, lanes))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (lanes > 1) {
        cudaEventRecord((cudaEvent_t)h->fork_event, st);
        for (int i = 0; i < lanes; ++i) cudaStreamWaitEvent((cudaStream_t)h->lane_stream[i], (cudaEvent_t)h->fork_event, 0);
    }
    static const double one64 = 1.0;   // static: the asynchronous copies below may read them after this call returns
    static const float one32 = 1.0f;
    for (int64_t b = 0; b < batch; ++b) {
        const int l = (int)(b % lanes);
        pssgp_handle* hl = lanes > 1 ? h->lane[l] : h;
        void* sl = lanes > 1 ? h->lane_stream[l] : stream;
        if (lanes > 1) {
            hl->chunk_opt = h->chunk_opt, hl->mid_warps = h->mid_warps, hl->mid_smem = h->mid_smem;
            hl->force_generic = h->force_generic, hl->pdl = h->pdl;
        }
        if ((rc = ws_reserve(hl, WS_GRID, need))) return rc;
        char* base = (char*)hl->buf[WS_GRID];
        void* g_ll = base;   // device scalar 1 (upstream gradient of the log-likelihood), then the per-setting arrays
        char* arr = base + 256;
        void *Fs = arr, *Qs = arr + up(nm), *fPs = arr + 2 * up(nm), *fms = arr + 3 * up(nm);
        void *dFs = arr + 3 * up(nm) + up(nv), *dQs = arr + 4 * up(nm) + up(nv);
        if (grad && b < lanes) {   // first setting of this lane in this call
            cudaError_t e = cudaMemcpyAsync(g_ll, dtype == PSSGP_F64 ? (const void*)&one64 : (const void*)&one32, es,
                                            cudaMemcpyHostToDevice, (cudaStream_t)sl);
            if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "grid_loglik: %s", cudaGetErrorString(e));
        }
        const char* Fb = (const char*)F + (size_t)b * d * d * es;
        const char* Pb = (const char*)Pinf + (size_t)b * d * d * es;
        const char* Hb = (const char*)H + (size_t)b * d * es;
        const char* Rb = (const char*)R + (size_t)b * es;
        const int64_t l0 = hl->launches;
        if ((rc = pssgp_discretise(hl, dtype, n, d, Fb, Pb, dts, Fs, Qs, sl))) return rc;
        if (!grad) {
            rc = pssgp_pkf(hl, dtype, n, d, Pb, Fs, Qs, Hb, Rb, y, nullptr, 1, fms, fPs, (char*)ll + (size_t)b * es, nullptr, sl);
        } else {
            rc = pssgp_pkfs_grad(hl, dtype, n, d, Pb, Fs, Qs, Hb, Rb, y, g_ll, fms, fPs, (char*)ll + (size_t)b * es, nullptr,
                                 nullptr, (char*)dP0 + (size_t)b * d * d * es, dFs, dQs, (char*)dH + (size_t)b * d * es,
                                 (char*)dR + (size_t)b * es, sl);
            if (!rc)
                rc = pssgp_discretise_backward(hl, dtype, n, d, Fb, Pb, dts, Fs, dFs, dQs, (char*)dF + (size_t)b * d * d * es,
                                               (char*)dPinf + (size_t)b * d * d * es, sl);
        }
        if (rc) return rc;
        if (lanes > 1) h->launches += hl->launches - l0;
    }
    if (lanes > 1)
        for (int i = 0; i < lanes; ++i) {
            cudaEventRecord((cudaEvent_t)h->lane_event[i], (cudaStream_t)h->lane_stream[i]);
            cudaStreamWaitEvent(st, (cudaEvent_t)h->lane_event[i], 0);
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "grid_loglik: %s", cudaGetErrorString(e));
    return PSSGP_OK;
}

int pssgp_grid_loglik(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, const void* F, const void* Pinf,
                      const void* H, const void* R, const void* dts, const void* y, void* ll, void* stream) {
    return grid_impl(h, dtype, batch, n, d, F, Pinf, H, R, dts, y, ll, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

int pssgp_grid_loglik_grad(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, const void* F, const void* Pinf,
                           const void* H, const void* R, const void* dts, const void* y, void* ll, void* dF, void* dPinf,
                           void* dP0, void* dH, void* dR, void* stream) {
    if (!dF) return set_err(PSSGP_ERR_INVALID, "grid_loglik_grad: null gradient output");
    return grid_impl(h, dtype, batch, n, d, F, Pinf, H, R, dts, y, ll, dF, dPinf, dP0, dH, dR, stream);
}

}  // extern "C"
