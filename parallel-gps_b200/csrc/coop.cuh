// CTA-cooperative dense linear algebra on small (d <= 64) matrices that live in shared memory.
// Used by the generic state-dimension path: one CTA works on one chunk / one aggregate at a time,
// threads stride over output entries; every routine leaves synchronisation to the caller unless
// stated otherwise ("syncs internally").
#pragma once
#include "smalld.cuh"

namespace pssgp {

struct Coop {
    int tid, nt;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};

// C(i,j) = alpha * sum_k A(i,k) B(k,j) + (C0 ? C0(i,j) : 0);  A(i,k) = A[i*ar + k*ac], B(k,j) = B[k*br + j*bc];
// C, C0 row-major with leading dimension ldc; rows x cols output, inner dimension kk.
template <typename T>
__device__ __forceinline__ void co_mm(const Coop& c, int rows, int cols, int kk, const T* __restrict__ A, int ar, int ac,
                                      const T* __restrict__ B, int br, int bc, T* __restrict__ C, int ldc,
                                      const T* __restrict__ C0, T alpha) {
    for (int idx = c.tid; idx < rows * cols; idx += c.nt) {
        const int i = idx / cols, j = idx - i * cols;
        const T* a = A + i * ar;
        const T* b = B + j * bc;
        T acc0 = T(0), acc1 = T(0);
        int k = 0;
        for (; k + 1 < kk; k += 2) {
            acc0 = fma(a[k * ac], b[k * br], acc0);
            acc1 = fma(a[(k + 1) * ac], b[(k + 1) * br], acc1);
        }
        if (k < kk) acc0 = fma(a[k * ac], b[k * br], acc0);
        T v = alpha * (acc0 + acc1);
        if (C0) v += C0[i * ldc + j];
        C[i * ldc + j] = v;
    }
}

// y(i) = alpha * sum_k A(i,k) x(k) + (y0 ? y0(i) : 0)
template <typename T>
__device__ __forceinline__ void co_mv(const Coop& c, int rows, int kk, const T* __restrict__ A, int ar, int ac,
                                      const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ y0, T alpha) {
    for (int i = c.tid; i < rows; i += c.nt) {
        T acc = T(0);
        for (int k = 0; k < kk; ++k) acc = fma(A[i * ar + k * ac], x[k], acc);
        T v = alpha * acc;
        if (y0) v += y0[i];
        y[i] = v;
    }
}

template <typename T>
__device__ __forceinline__ void co_copy(const Coop& c, int n, const T* __restrict__ src, T* __restrict__ dst) {
    for (int i = c.tid; i < n; i += c.nt) dst[i] = src[i];
}

template <typename T> __device__ __forceinline__ void co_fill(const Coop& c, int n, T* dst, T v) {
    for (int i = c.tid; i < n; i += c.nt) dst[i] = v;
}

template <typename T> __device__ __forceinline__ void co_eye(const Coop& c, int d, T* dst) {
    for (int i = c.tid; i < d * d; i += c.nt) dst[i] = ((i / d) == (i % d)) ? T(1) : T(0);
}

// S <- 0.5 (S + S^T) + (S0 ? S0 : 0), in place, d x d.  Each unordered pair is handled by one thread.
template <typename T>
__device__ __forceinline__ void co_symmetrise(const Coop& c, int d, T* S, const T* S0) {
    for (int idx = c.tid; idx < d * d; idx += c.nt) {
        const int i = idx / d, j = idx - i * d;
        if (j > i) continue;
        T v = T(0.5) * (S[i * d + j] + S[j * d + i]);
        T vij = v, vji = v;
        if (S0) {
            vij += S0[i * d + j];
            vji += S0[j * d + i];
        }
        S[i * d + j] = vij;
        S[j * d + i] = vji;
    }
}

// Gauss-Jordan with partial pivoting: solves M X = B in place (X overwrites B).  M: d x d (ld d), destroyed;
// B: d x nr (ld nr).  Syncs internally (entry state must already be synchronised; exit state is synchronised).
template <typename T>
__device__ __forceinline__ void co_solve(const Coop& c, int d, int nr, T* M, T* B, int* piv) {
    for (int col = 0; col < d; ++col) {
        if (c.tid == 0) {
            int p = col;
            T best = t_abs(M[col * d + col]);
            for (int r = col + 1; r < d; ++r) {
                const T v = t_abs(M[r * d + col]);
                if (v > best) {
                    best = v;
                    p = r;
                }
            }
            *piv = p;
        }
        c.sync();
        const int p = *piv;
        if (p != col) {
            for (int j = c.tid; j < d + nr; j += c.nt) {
                T* base = (j < d) ? (M + j) : (B + (j - d));
                const int ld = (j < d) ? d : nr;
                const T t = base[col * ld];
                base[col * ld] = base[p * ld];
                base[p * ld] = t;
            }
            c.sync();
        }
        const T inv = T(1) / M[col * d + col];
        c.sync();  // everyone has read the pivot before the row is scaled
        for (int j = c.tid; j < d + nr; j += c.nt) {
            if (j < d) {
                if (j >= col) M[col * d + j] *= inv;  // includes the pivot (-> 1); it is not read again
            } else {
                B[col * nr + (j - d)] *= inv;
            }
        }
        c.sync();
        // eliminate column `col` from every other row; columns <= col of M are not needed any more
        const int wcols = (d - col - 1) + nr;
        for (int idx = c.tid; idx < d * wcols; idx += c.nt) {
            const int r = idx / wcols, jj = idx - r * wcols;
            if (r == col) continue;
            const T f = M[r * d + col];
            if (jj < d - col - 1) {
                const int j = col + 1 + jj;
                M[r * d + j] = fma(-f, M[col * d + j], M[r * d + j]);
            } else {
                const int j = jj - (d - col - 1);
                B[r * nr + j] = fma(-f, B[col * nr + j], B[r * nr + j]);
            }
        }
        c.sync();
    }
}

// Deterministic CTA reduction of one value per thread; result valid on thread 0.  Syncs internally.
template <typename T> __device__ __forceinline__ T co_reduce_sum(const Coop& c, T v, T* scratch /* >= 32 */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += shfl_down_t(v, off);
    if ((c.tid & 31) == 0) scratch[c.tid >> 5] = v;
    c.sync();
    T tot = T(0);
    if (c.tid == 0)
        for (int w = 0; w < (c.nt + 31) / 32; ++w) tot += scratch[w];
    c.sync();
    return tot;
}

}  // namespace pssgp
