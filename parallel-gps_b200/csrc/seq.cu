// C ABI: pssgp_kf, pssgp_ks (see include/pssgp_b200.h) — the sequential Kalman filter and RTS smoother of the
// reference (pssgp/kalman/sequential.py:11-47 kf, :50-68 ks), run for `batch` independent series at once.
//
// The recursion over time is serial by definition, so the parallel axis is the batch of series:
//   d <= 4      one THREAD per series, state in registers (packed symmetric covariances, smalld.cuh);
//   5 <= d <= 32 one WARP per series, state in shared memory, the d*d matrix elements dealt out to the lanes,
//               the next step's inputs staged with cp.async while the current step computes.
// For a single long series this is the comparator of the reference (StateSpaceGP(parallel=False)), bound by the
// dependent latency of one step; for many short series it is the throughput path (no scan overhead).
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "scan_run.cuh"
#include "scan_stream.cuh"
#include "smalld.cuh"
#include "workspace.h"

namespace pssgp {

template <typename T>
struct SeqKfArgs {
    const T *P0, *Fs, *Qs, *H, *R, *y;
    T *fms, *fPs, *mps, *Pps, *ll;
    long batch, n;
    int d, lgssm_batched;
};
template <typename T>
struct SeqKsArgs {
    const T *Fs, *fms, *fPs, *mps, *Pps;
    T *sms, *sPs;
    long batch, n;
    int d, lgssm_batched;
};

// Running log-likelihood  sum_k log N(y_k; yp_k, S_k)  (sequential.py:24-28, 37): the quadratic terms are summed as
// they come, the logarithms of the innovation variances as ONE logarithm of their running product (smalld.cuh LogSum).
template <typename T> struct LogLik {
    LogSum<T> ls;
    T quad;
    long cnt;
    PSSGP_DEV void init() { ls.init(); quad = T(0); cnt = 0; }
    PSSGP_DEV void add(T res, T S, T invS) { ls.add(S); quad = fma(res * res, invS, quad); ++cnt; }
    PSSGP_DEV T value() const { return T(-0.5) * (ls.value() + quad + T(cnt) * T(1.8378770664093454835606594728112)); }
};

// ------------------------------------------------------------------------------------------------------------------
// d <= 4: one thread per series
// ------------------------------------------------------------------------------------------------------------------
template <typename T, int D> PSSGP_DEV void ld_sym_full(const T* A, T* S) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) S[sidx(i, j)] = T(0.5) * (A[i * D + j] + A[j * D + i]);
}
template <typename T, int D> PSSGP_DEV void st_sym_full(const T* S, T* A) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) A[i * D + j] = S[sidx(i, j)];
}

// S(sym) = 0.5*(X A^T + (X A^T)^T) + S0  (sequential.py:20-21)
template <typename T, int D> PSSGP_DEV void sym_half_xat_plus(const T* X, const T* A, const T* S0, T* S) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T a1 = T(0), a2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                a1 = fma(X[i * D + k], A[j * D + k], a1);
                a2 = fma(X[j * D + k], A[i * D + k], a2);
            }
            S[sidx(i, j)] = T(0.5) * (a1 + a2) + S0[sidx(i, j)];
        }
}

template <typename T, int D>
__global__ void __launch_bounds__(64) seq_kf_thread(SeqKfArgs<T> a) {
    constexpr int NS = nsym(D);
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    const long lb = a.lgssm_batched ? b : 0;
    const T* Fs = a.Fs + lb * a.n * D * D;
    const T* Qs = a.Qs + lb * a.n * D * D;
    const T* y = a.y + b * a.n;
    T h[D], m[D], P[NS];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        h[i] = a.H[lb * D + i];
        m[i] = T(0);
    }
    ld_sym_full<T, D>(a.P0 + lb * D * D, P);
    const T R = a.R[lb];
    LogLik<T> ell;
    ell.init();
    T F[D * D], Q[NS], yk;
    // software pipeline: the loads of step k + 1 are issued before the arithmetic of step k
    T Fn[D * D], Qn[NS], yn;
#pragma unroll
    for (int e = 0; e < D * D; ++e) Fn[e] = Fs[e];
    ld_sym_full<T, D>(Qs, Qn);
    yn = y[0];
    for (long k = 0; k < a.n; ++k) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) F[e] = Fn[e];
#pragma unroll
        for (int e = 0; e < NS; ++e) Q[e] = Qn[e];
        yk = yn;
        if (k + 1 < a.n) {
#pragma unroll
            for (int e = 0; e < D * D; ++e) Fn[e] = Fs[(k + 1) * D * D + e];
            ld_sym_full<T, D>(Qs + (k + 1) * D * D, Qn);
            yn = y[k + 1];
        }
        T mp[D], X[D * D], Pp[NS];
        mv_f<T, D>(F, m, mp);
        mm_fs<T, D>(F, P, X);
        sym_half_xat_plus<T, D>(X, F, Q, Pp);
        if (a.mps) {
            T* o = a.mps + (b * a.n + k) * D;
#pragma unroll
            for (int i = 0; i < D; ++i) o[i] = mp[i];
            st_sym_full<T, D>(Pp, a.Pps + (b * a.n + k) * D * D);
        }
        if (!t_isnan(yk)) {
            T PH[D];
            mv_s<T, D>(Pp, h, PH);
            const T S = dot<T, D>(h, PH) + R;
            const T res = yk - dot<T, D>(h, mp);
            const T invS = t_rcp(S);
            ell.add(res, S, invS);
            const T g = res * invS;
            T Kg[D];  // gain K = Pp H^T / S
#pragma unroll
            for (int i = 0; i < D; ++i) {
                Kg[i] = PH[i] * invS;
                m[i] = fma(PH[i], g, mp[i]);
            }
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) P[sidx(i, j)] = fma(-Kg[i], PH[j], Pp[sidx(i, j)]);
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) m[i] = mp[i];
#pragma unroll
            for (int e = 0; e < NS; ++e) P[e] = Pp[e];
        }
        T* o = a.fms + (b * a.n + k) * D;
#pragma unroll
        for (int i = 0; i < D; ++i) o[i] = m[i];
        st_sym_full<T, D>(P, a.fPs + (b * a.n + k) * D * D);
    }
    if (a.ll) a.ll[b] = ell.value();
}

template <typename T, int D>
__global__ void __launch_bounds__(64) seq_ks_thread(SeqKsArgs<T> a) {
    constexpr int NS = nsym(D);
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    const long lb = a.lgssm_batched ? b : 0;
    const T* Fs = a.Fs + lb * a.n * D * D;
    const long o1 = b * a.n * D, o2 = b * a.n * D * D;
    T sm[D], sP[NS];
#pragma unroll
    for (int i = 0; i < D; ++i) sm[i] = a.fms[o1 + (a.n - 1) * D + i];
    ld_sym_full<T, D>(a.fPs + o2 + (a.n - 1) * D * D, sP);
#pragma unroll
    for (int i = 0; i < D; ++i) a.sms[o1 + (a.n - 1) * D + i] = sm[i];
    st_sym_full<T, D>(sP, a.sPs + o2 + (a.n - 1) * D * D);
    for (long k = a.n - 2; k >= 0; --k) {
        T F[D * D], m[D], mp[D], P[NS], Pp[NS];
#pragma unroll
        for (int e = 0; e < D * D; ++e) F[e] = Fs[(k + 1) * D * D + e];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            m[i] = a.fms[o1 + k * D + i];
            mp[i] = a.mps[o1 + (k + 1) * D + i];
        }
        ld_sym_full<T, D>(a.fPs + o2 + k * D * D, P);
        ld_sym_full<T, D>(a.Pps + o2 + (k + 1) * D * D, Pp);
        // Ct = Pp^-1 F P  (sequential.py:56-58), dm = sm - mp, Dm = sP - Pp
        T Ct[D * D], dm[D], Dm[NS], X[D * D];
        mm_fs<T, D>(F, P, Ct);
#pragma unroll
        for (int i = 0; i < D; ++i) dm[i] = sm[i] - mp[i];
#pragma unroll
        for (int e = 0; e < NS; ++e) Dm[e] = sP[e] - Pp[e];
        ldl_packed<T, D>(Pp);
        ldl_solve<T, D, D>(Pp, Ct);
        T t[D];
        mv_t<T, D>(Ct, dm, t);
#pragma unroll
        for (int i = 0; i < D; ++i) sm[i] = m[i] + t[i];
        mm_sf<T, D>(Dm, Ct, X);
        sym_half_atb_plus<T, D>(Ct, X, P, sP);
#pragma unroll
        for (int i = 0; i < D; ++i) a.sms[o1 + k * D + i] = sm[i];
        st_sym_full<T, D>(sP, a.sPs + o2 + k * D * D);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 5 <= d <= 32: one warp per series
// ------------------------------------------------------------------------------------------------------------------
template <typename T> PSSGP_DEV T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> PSSGP_DEV void cp_async_elem(T* dst, const T* src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
    if constexpr (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(src) : "memory");
}

// The d*d elements of a matrix are dealt out to the lanes in row-major order: lane l owns elements l, l + 32, ...;
// (i, j) are advanced incrementally so that no integer division runs inside the time loop.
#define SEQ_FOR_ELEMS                                                                                  \
    for (int e = lane, i = i0, j = j0; e < dd; e += 32, i += istep, j += jstep, (j >= d ? (j -= d, ++i) : 0))

constexpr int kSeqWarpThreads = 128;

__host__ __device__ inline size_t seq_kf_warp_elems(int d) {
    const int LD = d | 1;
    size_t v = (size_t)4 * d * d + (size_t)3 * d * LD + (size_t)4 * d;
    return (v + 1) & ~(size_t)1;
}
__host__ __device__ inline size_t seq_ks_stage_elems(int d) { return ((size_t)3 * d * d + 2 * d + 1) & ~(size_t)1; }
__host__ __device__ inline size_t seq_ks_warp_elems(int d) {
    const int LD = d | 1;
    size_t v = 2 * seq_ks_stage_elems(d) + (size_t)5 * d * LD + (size_t)2 * d;
    return (v + 1) & ~(size_t)1;
}

template <typename T>
__global__ void __launch_bounds__(kSeqWarpThreads) seq_kf_warp(SeqKfArgs<T> a) {
    extern __shared__ __align__(16) unsigned char seq_smem[];
    const int d = a.d, dd = d * d, LD = d | 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long b = (long)blockIdx.x * (blockDim.x >> 5) + w;
    if (b >= a.batch) return;  // warp-uniform; no block-level barrier below
    const int i0 = lane / d, j0 = lane % d, istep = 32 / d, jstep = 32 % d;
    T* stage = (T*)seq_smem + (size_t)w * seq_kf_warp_elems(d);  // [2][F | Q] raw row-major
    T* P = stage + 4 * dd;
    T* X = P + d * LD;
    T* Y = X + d * LD;
    T* m = Y + d * LD;
    T* mp = m + d;
    T* H = mp + d;
    T* PH = H + d;
    const long lb = a.lgssm_batched ? b : 0;
    const T* Fs = a.Fs + lb * a.n * dd;
    const T* Qs = a.Qs + lb * a.n * dd;
    const T* y = a.y + b * a.n;
    const T R = a.R[lb];
    SEQ_FOR_ELEMS P[i * LD + j] = a.P0[lb * dd + e];
    if (lane < d) {
        m[lane] = T(0);
        H[lane] = a.H[lb * d + lane];
    }
    auto issue = [&](long k) {
        T* s = stage + (k & 1) * 2 * dd;
        for (int e = lane; e < dd; e += 32) {
            cp_async_elem(s + e, Fs + k * dd + e);
            cp_async_elem(s + dd + e, Qs + k * dd + e);
        }
        cp_async_commit();
    };
    issue(0);
    LogLik<T> ell;
    ell.init();
    for (long k = 0; k < a.n; ++k) {
        if (k + 1 < a.n) {
            issue(k + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const T* F = stage + (k & 1) * 2 * dd;
        const T* Q = F + dd;
        const T yk = y[k];
        // mp = F m;  X = F P
        if (lane < d) {
            T acc = T(0);
            for (int c = 0; c < d; ++c) acc = fma(F[lane * d + c], m[c], acc);
            mp[lane] = acc;
        }
        SEQ_FOR_ELEMS {
            T acc = T(0);
            for (int c = 0; c < d; ++c) acc = fma(F[i * d + c], P[c * LD + j], acc);
            X[i * LD + j] = acc;
        }
        __syncwarp();
        // Y = (X F^T)^T + Q;  Pp = (Y + Y^T) / 2  (sequential.py:20-21)
        SEQ_FOR_ELEMS {
            T acc = Q[e];
            for (int c = 0; c < d; ++c) acc = fma(F[i * d + c], X[j * LD + c], acc);
            Y[i * LD + j] = acc;
        }
        __syncwarp();
        const long ok = b * a.n + k;
        SEQ_FOR_ELEMS {
            const T v = T(0.5) * (Y[i * LD + j] + Y[j * LD + i]);
            P[i * LD + j] = v;
            if (a.Pps) a.Pps[ok * dd + e] = v;
        }
        if (a.mps && lane < d) a.mps[ok * d + lane] = mp[lane];
        __syncwarp();
        if (!t_isnan(yk)) {  // warp-uniform
            T ph = T(0), hl = T(0), ml = T(0);
            if (lane < d) {
                hl = H[lane];
                ml = mp[lane];
                for (int c = 0; c < d; ++c) ph = fma(P[lane * LD + c], H[c], ph);
            }
            const T S = warp_sum(hl * ph) + R;
            const T res = yk - warp_sum(hl * ml);
            const T invS = t_rcp(S);
            ell.add(res, S, invS);
            if (lane < d) {
                PH[lane] = ph;
                m[lane] = fma(ph, res * invS, ml);
            }
            __syncwarp();
            // symmetric to the last bit: the product PH[i] PH[j] is formed first
            SEQ_FOR_ELEMS P[i * LD + j] = fma(-(PH[i] * PH[j]), invS, P[i * LD + j]);
        } else if (lane < d) {
            m[lane] = mp[lane];
        }
        __syncwarp();
        if (lane < d) a.fms[ok * d + lane] = m[lane];
        SEQ_FOR_ELEMS a.fPs[ok * dd + e] = P[i * LD + j];
    }
    if (a.ll && lane == 0) a.ll[b] = ell.value();
}

template <typename T>
__global__ void __launch_bounds__(kSeqWarpThreads) seq_ks_warp(SeqKsArgs<T> a) {
    extern __shared__ __align__(16) unsigned char seq_smem[];
    const int d = a.d, dd = d * d, LD = d | 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long b = (long)blockIdx.x * (blockDim.x >> 5) + w;
    if (b >= a.batch) return;
    const int i0 = lane / d, j0 = lane % d, istep = 32 / d, jstep = 32 % d;
    const size_t SE = seq_ks_stage_elems(d);
    T* stage = (T*)seq_smem + (size_t)w * seq_ks_warp_elems(d);  // [2][F | P | Pp | m | mp] raw
    T* Lc = stage + 2 * SE;
    T* G = Lc + d * LD;
    T* Dm = G + d * LD;
    T* X = Dm + d * LD;
    T* sP = X + d * LD;
    T* sm = sP + d * LD;
    T* dv = sm + d;
    const long lb = a.lgssm_batched ? b : 0;
    const T* Fs = a.Fs + lb * a.n * dd;
    const long o1 = b * a.n * d, o2 = b * a.n * dd;
    auto issue = [&](long k) {  // inputs of the step that produces the smoothed state at k
        T* s = stage + (k & 1) * SE;
        for (int e = lane; e < dd; e += 32) {
            cp_async_elem(s + e, Fs + (k + 1) * dd + e);
            cp_async_elem(s + dd + e, a.fPs + o2 + k * dd + e);
            cp_async_elem(s + 2 * dd + e, a.Pps + o2 + (k + 1) * dd + e);
        }
        if (lane < d) {
            cp_async_elem(s + 3 * dd + lane, a.fms + o1 + k * d + lane);
            cp_async_elem(s + 3 * dd + d + lane, a.mps + o1 + (k + 1) * d + lane);
        }
        cp_async_commit();
    };
    if (a.n >= 2) issue(a.n - 2);
    SEQ_FOR_ELEMS {
        const T v = a.fPs[o2 + (a.n - 1) * dd + e];
        sP[i * LD + j] = v;
        a.sPs[o2 + (a.n - 1) * dd + e] = v;
    }
    if (lane < d) {
        const T v = a.fms[o1 + (a.n - 1) * d + lane];
        sm[lane] = v;
        a.sms[o1 + (a.n - 1) * d + lane] = v;
    }
    for (long k = a.n - 2; k >= 0; --k) {
        if (k >= 1) {
            issue(k - 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const T* F = stage + (k & 1) * SE;
        const T* P = F + dd;
        const T* Pp = P + dd;
        const T* m = Pp + dd;
        const T* mp = m + d;
        // Lc = Pp, Dm = sP - Pp, G = F P, dv = sm - mp
        SEQ_FOR_ELEMS {
            const T pp = Pp[e];
            Lc[i * LD + j] = pp;
            Dm[i * LD + j] = sP[i * LD + j] - pp;
            T acc = T(0);
            for (int c = 0; c < d; ++c) acc = fma(F[i * d + c], P[c * d + j], acc);
            G[i * LD + j] = acc;
        }
        if (lane < d) dv[lane] = sm[lane] - mp[lane];
        __syncwarp();
        // Cholesky of Pp, lower factor in place, one column per round (lane = row)
        for (int c = 0; c < d; ++c) {
            T s = T(0);
            if (lane >= c && lane < d) {
                s = Lc[lane * LD + c];
                for (int q = 0; q < c; ++q) s = fma(-Lc[lane * LD + q], Lc[c * LD + q], s);
            }
            const T dg = t_sqrt(__shfl_sync(0xffffffffu, s, c));
            if (lane == c)
                Lc[c * LD + c] = dg;
            else if (lane > c && lane < d)
                Lc[lane * LD + c] = s / dg;
            __syncwarp();
        }
        // Ct = Pp^-1 G in place: lane j solves column j (forward, then backward substitution)
        if (lane < d) {
            for (int r = 0; r < d; ++r) {
                T v = G[r * LD + lane];
                for (int q = 0; q < r; ++q) v = fma(-Lc[r * LD + q], G[q * LD + lane], v);
                G[r * LD + lane] = v / Lc[r * LD + r];
            }
            for (int r = d - 1; r >= 0; --r) {
                T v = G[r * LD + lane];
                for (int q = r + 1; q < d; ++q) v = fma(-Lc[q * LD + r], G[q * LD + lane], v);
                G[r * LD + lane] = v / Lc[r * LD + r];
            }
        }
        __syncwarp();
        // sm = m + Ct^T dv;  X = Dm Ct
        if (lane < d) {
            T acc = m[lane];
            for (int c = 0; c < d; ++c) acc = fma(G[c * LD + lane], dv[c], acc);
            sm[lane] = acc;
            a.sms[o1 + k * d + lane] = acc;
        }
        SEQ_FOR_ELEMS {
            T acc = T(0);
            for (int c = 0; c < d; ++c) acc = fma(Dm[i * LD + c], G[c * LD + j], acc);
            X[i * LD + j] = acc;
        }
        __syncwarp();
        // Dm <- P + Ct^T X;  sP = (Dm + Dm^T) / 2  (sequential.py:59-61)
        SEQ_FOR_ELEMS {
            T acc = P[e];
            for (int c = 0; c < d; ++c) acc = fma(G[c * LD + i], X[c * LD + j], acc);
            Dm[i * LD + j] = acc;
        }
        __syncwarp();
        SEQ_FOR_ELEMS {
            const T v = T(0.5) * (Dm[i * LD + j] + Dm[j * LD + i]);
            sP[i * LD + j] = v;
            a.sPs[o2 + k * dd + e] = v;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------
template <typename T, int D>
int kf_thread_impl(pssgp_handle* h, const SeqKfArgs<T>& a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.batch + 63) / 64);
    PSSGP_LAUNCH(h, "seq_kf", st, (seq_kf_thread<T, D><<<grid, 64, 0, st>>>(a)));
    return check_launch(h, "seq_kf", 1);
}
template <typename T, int D>
int ks_thread_impl(pssgp_handle* h, const SeqKsArgs<T>& a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.batch + 63) / 64);
    PSSGP_LAUNCH(h, "seq_ks", st, (seq_ks_thread<T, D><<<grid, 64, 0, st>>>(a)));
    return check_launch(h, "seq_ks", 1);
}

// warps per CTA such that the per-warp shared-memory slices fit (and small batches still spread over the SMs)
inline int seq_warps(size_t bytes_per_warp, long batch, int num_sms) {
    int wpb = kSeqWarpThreads / 32;
    while (wpb > 1 && (bytes_per_warp * wpb > (size_t)200 * 1024 || (batch + wpb - 1) / wpb < num_sms)) --wpb;
    return wpb;
}

template <typename T>
int kf_warp_impl(pssgp_handle* h, const SeqKfArgs<T>& a, cudaStream_t st) {
    const size_t per_warp = seq_kf_warp_elems(a.d) * sizeof(T);
    const int wpb = seq_warps(per_warp, a.batch, h->num_sms);
    const size_t smem = per_warp * wpb;
    cudaError_t e = cudaFuncSetAttribute(seq_kf_warp<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "seq_kf: %s", cudaGetErrorString(e));
    const unsigned grid = (unsigned)((a.batch + wpb - 1) / wpb);
    PSSGP_LAUNCH(h, "seq_kf", st, (seq_kf_warp<T><<<grid, wpb * 32, smem, st>>>(a)));
    return check_launch(h, "seq_kf", 1);
}
template <typename T>
int ks_warp_impl(pssgp_handle* h, const SeqKsArgs<T>& a, cudaStream_t st) {
    const size_t per_warp = seq_ks_warp_elems(a.d) * sizeof(T);
    const int wpb = seq_warps(per_warp, a.batch, h->num_sms);
    const size_t smem = per_warp * wpb;
    cudaError_t e = cudaFuncSetAttribute(seq_ks_warp<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "seq_ks: %s", cudaGetErrorString(e));
    const unsigned grid = (unsigned)((a.batch + wpb - 1) / wpb);
    PSSGP_LAUNCH(h, "seq_ks", st, (seq_ks_warp<T><<<grid, wpb * 32, smem, st>>>(a)));
    return check_launch(h, "seq_ks", 1);
}

template <typename T>
int kf_typed(pssgp_handle* h, int64_t batch, int64_t n, int d, int lgssm_batched, const void* P0, const void* Fs,
             const void* Qs, const void* H, const void* R, const void* y, void* fms, void* fPs, void* mps, void* Pps,
             void* ll, cudaStream_t st) {
    SeqKfArgs<T> a{(const T*)P0, (const T*)Fs, (const T*)Qs, (const T*)H, (const T*)R, (const T*)y,
                   (T*)fms,      (T*)fPs,      (T*)mps,      (T*)Pps,     (T*)ll,      (long)batch,
                   (long)n,      d,            lgssm_batched};
    switch (d) {
        case 1: return kf_thread_impl<T, 1>(h, a, st);
        case 2: return kf_thread_impl<T, 2>(h, a, st);
        case 3: return kf_thread_impl<T, 3>(h, a, st);
        case 4: return kf_thread_impl<T, 4>(h, a, st);
    }
    return kf_warp_impl<T>(h, a, st);
}
template <typename T>
int ks_typed(pssgp_handle* h, int64_t batch, int64_t n, int d, int lgssm_batched, const void* Fs, const void* fms,
             const void* fPs, const void* mps, const void* Pps, void* sms, void* sPs, cudaStream_t st) {
    SeqKsArgs<T> a{(const T*)Fs, (const T*)fms, (const T*)fPs, (const T*)mps, (const T*)Pps, (T*)sms,
                   (T*)sPs,      (long)batch,   (long)n,       d,             lgssm_batched};
    switch (d) {
        case 1: return ks_thread_impl<T, 1>(h, a, st);
        case 2: return ks_thread_impl<T, 2>(h, a, st);
        case 3: return ks_thread_impl<T, 3>(h, a, st);
        case 4: return ks_thread_impl<T, 4>(h, a, st);
    }
    return ks_warp_impl<T>(h, a, st);
}

}  // namespace pssgp

using namespace pssgp;

extern "C" {

int pssgp_kf(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, int lgssm_batched, const void* P0,
             const void* Fs, const void* Qs, const void* H, const void* R, const void* y, void* fms, void* fPs, void* mps,
             void* Pps, void* ll, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (batch < 1) return set_err(PSSGP_ERR_INVALID, "batch must be >= 1 (got %lld)", (long long)batch);
    if (d > 32) return set_err(PSSGP_ERR_UNSUPPORTED, "kf: d <= 32 (got %d)", d);
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if ((mps == nullptr) != (Pps == nullptr)) return set_err(PSSGP_ERR_INVALID, "kf: mps and Pps come together");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PSSGP_F64)
        return kf_typed<double>(h, batch, n, d, lgssm_batched != 0, P0, Fs, Qs, H, R, y, fms, fPs, mps, Pps, ll, st);
    return kf_typed<float>(h, batch, n, d, lgssm_batched != 0, P0, Fs, Qs, H, R, y, fms, fPs, mps, Pps, ll, st);
}

int pssgp_ks(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, int lgssm_batched, const void* Fs,
             const void* fms, const void* fPs, const void* mps, const void* Pps, void* sms, void* sPs, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (batch < 1) return set_err(PSSGP_ERR_INVALID, "batch must be >= 1 (got %lld)", (long long)batch);
    if (d > 32) return set_err(PSSGP_ERR_UNSUPPORTED, "ks: d <= 32 (got %d)", d);
    if (!Fs || !fms || !fPs || !mps || !Pps || !sms || !sPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PSSGP_F64) return ks_typed<double>(h, batch, n, d, lgssm_batched != 0, Fs, fms, fPs, mps, Pps, sms, sPs, st);
    return ks_typed<float>(h, batch, n, d, lgssm_batched != 0, Fs, fms, fPs, mps, Pps, sms, sPs, st);
}

}  // extern "C"
