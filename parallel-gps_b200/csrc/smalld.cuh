// Register-resident linear algebra for tiny state dimensions (D <= 4): every loop is fully
// unrolled so matrices live in registers.  Symmetric matrices are stored packed (lower triangle,
// row-major: (i,j), i>=j  ->  i*(i+1)/2 + j).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pssgp {

#define PSSGP_DEV __device__ __forceinline__

__host__ __device__ constexpr int nsym(int D) { return D * (D + 1) / 2; }
__host__ __device__ constexpr int sidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

template <typename T> PSSGP_DEV T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> PSSGP_DEV bool t_isnan(T x) { return x != x; }
PSSGP_DEV double t_log(double x) { return log(x); }
PSSGP_DEV float t_log(float x) { return logf(x); }
PSSGP_DEV double t_sqrt(double x) { return sqrt(x); }
PSSGP_DEV float t_sqrt(float x) { return sqrtf(x); }

// Reciprocal of a normal, finite number: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps, 5
// instructions instead of the ~20 of an IEEE division with its slow path.  Accurate to about 1 ulp, which
// is all the callers (pivots, innovation variances) need.
PSSGP_DEV double t_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
PSSGP_DEV float t_rcp(float x) { return __frcp_rn(x); }

// Running sum of log(x_i) kept as (mantissa product, exponent sum): one logarithm per chunk instead of
// one per time step (an FP64 log is ~80 instructions).  x must be positive and finite; anything else is
// routed through log() so that NaN / -inf still propagate.
template <typename T> struct LogSum;
template <> struct LogSum<double> {
    double m, extra;
    int e, cnt;
    PSSGP_DEV void init() { m = 1.0; extra = 0.0; e = 0; cnt = 0; }
    PSSGP_DEV void renorm() {
        const long long b = __double_as_longlong(m);
        e += (int)((b >> 52) & 0x7ff) - 1022;
        m = __longlong_as_double((b & 0x800fffffffffffffLL) | 0x3fe0000000000000LL);
    }
    PSSGP_DEV void add(double x) {
        const long long b = __double_as_longlong(x);
        const int be = (int)((b >> 52) & 0x7ff);
        if (b > 0 && be != 0 && be != 0x7ff) {
            e += be - 1022;
            m *= __longlong_as_double((b & 0x000fffffffffffffLL) | 0x3fe0000000000000LL);  // in [0.5, 1)
            if ((++cnt & 511) == 0) renorm();
        } else {
            extra += log(x);
        }
    }
    PSSGP_DEV double value() const { return log(m) + (double)e * 0.69314718055994530942 + extra; }
};
template <> struct LogSum<float> {
    float m, extra;
    int e, cnt;
    PSSGP_DEV void init() { m = 1.0f; extra = 0.0f; e = 0; cnt = 0; }
    PSSGP_DEV void renorm() {
        const int b = __float_as_int(m);
        e += ((b >> 23) & 0xff) - 126;
        m = __int_as_float((b & 0x807fffff) | 0x3f000000);
    }
    PSSGP_DEV void add(float x) {
        const int b = __float_as_int(x);
        const int be = (b >> 23) & 0xff;
        if (b > 0 && be != 0 && be != 0xff) {
            e += be - 126;
            m *= __int_as_float((b & 0x007fffff) | 0x3f000000);
            if ((++cnt & 63) == 0) renorm();
        } else {
            extra += logf(x);
        }
    }
    PSSGP_DEV float value() const { return logf(m) + (float)e * 0.69314718f + extra; }
};

// C(full) = A(full) * B(full)
template <typename T, int D> PSSGP_DEV void mm_ff(const T* A, const T* B, T* C) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fma(A[i * D + k], B[k * D + j], acc);
            C[i * D + j] = acc;
        }
}
// C(full) = A(full) * S(sym packed)
template <typename T, int D> PSSGP_DEV void mm_fs(const T* A, const T* S, T* C) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fma(A[i * D + k], S[sidx(k, j)], acc);
            C[i * D + j] = acc;
        }
}
// C(full) = S(sym) * A(full)
template <typename T, int D> PSSGP_DEV void mm_sf(const T* S, const T* A, T* C) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fma(S[sidx(i, k)], A[k * D + j], acc);
            C[i * D + j] = acc;
        }
}
// C(full) = S(sym) * A(full)^T
template <typename T, int D> PSSGP_DEV void mm_sft(const T* S, const T* A, T* C) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fma(S[sidx(i, k)], A[j * D + k], acc);
            C[i * D + j] = acc;
        }
}
// C(full) = A^T * B
template <typename T, int D> PSSGP_DEV void mm_tf(const T* A, const T* B, T* C) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fma(A[k * D + i], B[k * D + j], acc);
            C[i * D + j] = acc;
        }
}
// S(sym) = lower triangle of X(full) * A(full)^T  + S0(sym)   (X A^T assumed symmetric in exact arithmetic)
template <typename T, int D> PSSGP_DEV void sym_xat_plus(const T* X, const T* A, const T* S0, T* S) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T acc = S0[sidx(i, j)];
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fma(X[i * D + k], A[j * D + k], acc);
            S[sidx(i, j)] = acc;
        }
}
// S(sym) = 0.5*(X + X^T) + S0 where X = A(full) * B(full)
template <typename T, int D> PSSGP_DEV void sym_half_ab_plus(const T* A, const T* B, const T* S0, T* S) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T a1 = T(0), a2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                a1 = fma(A[i * D + k], B[k * D + j], a1);
                a2 = fma(A[j * D + k], B[k * D + i], a2);
            }
            S[sidx(i, j)] = T(0.5) * (a1 + a2) + S0[sidx(i, j)];
        }
}
// S(sym) = 0.5*(X + X^T) + S0 where X = A^T * B
template <typename T, int D> PSSGP_DEV void sym_half_atb_plus(const T* A, const T* B, const T* S0, T* S) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T a1 = T(0), a2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                a1 = fma(A[k * D + i], B[k * D + j], a1);
                a2 = fma(A[k * D + j], B[k * D + i], a2);
            }
            S[sidx(i, j)] = T(0.5) * (a1 + a2) + S0[sidx(i, j)];
        }
}
// y = A x
template <typename T, int D> PSSGP_DEV void mv_f(const T* A, const T* x, T* y) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fma(A[i * D + k], x[k], acc);
        y[i] = acc;
    }
}
// y = A^T x
template <typename T, int D> PSSGP_DEV void mv_t(const T* A, const T* x, T* y) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fma(A[k * D + i], x[k], acc);
        y[i] = acc;
    }
}
// y = S x (S sym packed)
template <typename T, int D> PSSGP_DEV void mv_s(const T* S, const T* x, T* y) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fma(S[sidx(i, k)], x[k], acc);
        y[i] = acc;
    }
}
template <typename T, int D> PSSGP_DEV T dot(const T* a, const T* b) {
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) acc = fma(a[k], b[k], acc);
    return acc;
}

// Solve M Z = B in place (B is D x NR, row-major), Gaussian elimination with partial pivoting,
// all row exchanges done with selects so nothing is dynamically indexed.
template <typename T, int D, int NR> PSSGP_DEV void lu_solve(T* M, T* B) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
#pragma unroll
        for (int r = c + 1; r < D; ++r) {
            bool sw = t_abs(M[r * D + c]) > t_abs(M[c * D + c]);
#pragma unroll
            for (int j = c; j < D; ++j) {
                T a = M[c * D + j], b = M[r * D + j];
                M[c * D + j] = sw ? b : a;
                M[r * D + j] = sw ? a : b;
            }
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                T a = B[c * NR + j], b = B[r * NR + j];
                B[c * NR + j] = sw ? b : a;
                B[r * NR + j] = sw ? a : b;
            }
        }
        T inv = t_rcp(M[c * D + c]);
#pragma unroll
        for (int j = c + 1; j < D; ++j) M[c * D + j] *= inv;
#pragma unroll
        for (int j = 0; j < NR; ++j) B[c * NR + j] *= inv;
#pragma unroll
        for (int r = c + 1; r < D; ++r) {
            T f = M[r * D + c];
#pragma unroll
            for (int j = c + 1; j < D; ++j) M[r * D + j] = fma(-f, M[c * D + j], M[r * D + j]);
#pragma unroll
            for (int j = 0; j < NR; ++j) B[r * NR + j] = fma(-f, B[c * NR + j], B[r * NR + j]);
        }
    }
    // back substitution (unit upper triangular after the row scaling above)
#pragma unroll
    for (int c = D - 1; c > 0; --c)
#pragma unroll
        for (int r = 0; r < c; ++r) {
            T f = M[r * D + c];
#pragma unroll
            for (int j = 0; j < NR; ++j) B[r * NR + j] = fma(-f, B[c * NR + j], B[r * NR + j]);
        }
}

// LDL^T of a packed symmetric positive definite matrix in place: unit lower factor in the strict lower
// triangle, the RECIPROCALS of the pivots d_j on the diagonal (no square roots; three reciprocals serve
// the factorisation and both triangular solves).
template <typename T, int D> PSSGP_DEV void ldl_packed(T* S) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        T d = S[sidx(j, j)];
        T w[D > 1 ? D - 1 : 1];  // w[k] = L[j][k] d_k
#pragma unroll
        for (int k = 0; k < j; ++k) {
            w[k] = S[sidx(j, k)];  // still holds L[j][k] d_k at this point (scaled below)
            d = fma(-w[k], w[k] * S[sidx(k, k)], d);
        }
#pragma unroll
        for (int k = 0; k < j; ++k) S[sidx(j, k)] = w[k] * S[sidx(k, k)];  // L[j][k]
        const T inv = t_rcp(d);
        S[sidx(j, j)] = inv;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            T v = S[sidx(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-S[sidx(i, k)], S[sidx(j, k)], v);  // S[i][k] still holds L[i][k] d_k
            S[sidx(i, j)] = v;  // L[i][j] d_j (scaled when row i is factored)
        }
    }
}
// Solve (L D L^T) X = B in place, B is D x NR row-major, factor from ldl_packed.
template <typename T, int D, int NR> PSSGP_DEV void ldl_solve(const T* L, T* B) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            T v = B[i * NR + j];
#pragma unroll
            for (int k = 0; k < i; ++k) v = fma(-L[sidx(i, k)], B[k * NR + j], v);
            B[i * NR + j] = v;
        }
#pragma unroll
    for (int i = D - 1; i >= 0; --i)
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            T v = B[i * NR + j] * L[sidx(i, i)];
#pragma unroll
            for (int k = i + 1; k < D; ++k) v = fma(-L[sidx(k, i)], B[k * NR + j], v);
            B[i * NR + j] = v;
        }
}

template <typename T> PSSGP_DEV T shfl_up_t(T v, int delta) { return __shfl_up_sync(0xffffffffu, v, delta); }
template <typename T> PSSGP_DEV T shfl_idx_t(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <typename T> PSSGP_DEV T shfl_down_t(T v, int delta) { return __shfl_down_sync(0xffffffffu, v, delta); }

}  // namespace pssgp
