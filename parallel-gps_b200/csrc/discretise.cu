// SDE discretisation and its adjoint.
//
// Replaces the reference's pssgp/kernels/base.py:29-47 (_get_ssm): Fs = expm(dt F) and Qs (the
// reference uses a matrix-fraction decomposition through a second, 2d x 2d expm; here the
// stationary identity Q = Pinf - A Pinf A^T is used).
//
// Because F is the same for every time step, expm(F dt_k) is evaluated as a scaled Taylor polynomial
// whose matrix coefficients C_j = (F/||F||_1)^j / j! are computed ONCE (setup kernel):
//     x = ||F||_1 |dt| / 2^s <= theta,  A_h = sum_j C_j x^j (Horner, d^2 FMAs per term, no matmul),
//     A = A_h^(2^s) by s squarings.
// The adjoint needs no per-step matmul when s = 0 either: dF = 1/||F|| sum_p 1/p! sum_i (G^T)^i W_p (G^T)^(p-1-i)
// with the moment matrices W_p = sum_k x_k^p dA_h,k, which are plain reductions over time.
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <stdlib.h>

#include "coop.cuh"
#include "workspace.h"

namespace pssgp {


template <typename T> struct Taylor;
template <> struct Taylor<double> { static constexpr int DEG = 18; };
template <> struct Taylor<float> { static constexpr int DEG = 10; };
constexpr int kMaxSquarings = 24;
constexpr int kDiscThreads = 128;

// coef layout: [0]: ||F||_1 ; then C_j (j = 0..DEG), each d*d, starting at offset 8 (keeps 16B alignment)
template <typename T>
__global__ void taylor_setup_kernel(const T* __restrict__ F, int d, T* __restrict__ coef, int transpose) {
    extern __shared__ unsigned char smem_raw[];
    T* G = (T*)smem_raw;          // d*d
    T* cur = G + d * d;           // d*d
    T* nxt = cur + d * d;         // d*d
    __shared__ T normF;
    const int dd = d * d;
    if (threadIdx.x == 0) {
        T best = T(0);
        for (int j = 0; j < d; ++j) {
            T cs = T(0);
            for (int i = 0; i < d; ++i) cs += t_abs(F[i * d + j]);
            best = cs > best ? cs : best;
        }
        normF = best;
        coef[0] = best;
    }
    __syncthreads();
    const T inv = normF > T(0) ? T(1) / normF : T(0);
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {
        const int i = e / d, j = e % d;
        G[e] = (transpose ? F[j * d + i] : F[e]) * inv;
        cur[e] = (i == j) ? T(1) : T(0);
        coef[8 + e] = cur[e];
    }
    __syncthreads();
    const int DEG = Taylor<T>::DEG;
    for (int p = 1; p <= DEG; ++p) {
        for (int e = threadIdx.x; e < dd; e += blockDim.x) {
            const int i = e / d, j = e % d;
            T acc = T(0);
            for (int k = 0; k < d; ++k) acc = fma(G[i * d + k], cur[k * d + j], acc);
            nxt[e] = acc / T(p);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < dd; e += blockDim.x) {
            cur[e] = nxt[e];
            coef[8 + (size_t)p * dd + e] = nxt[e];
        }
        __syncthreads();
    }
}

template <typename T> PSSGP_DEV int pick_squarings(T nrm, T& x) {
    // x = nrm / 2^s <= 1
    int s = 0;
    x = nrm;
    while (x > T(1) && s < kMaxSquarings) {
        x *= T(0.5);
        ++s;
    }
    return s;
}

template <typename T, int D>
__global__ void __launch_bounds__(kDiscThreads)
discretise_small_kernel(const T* __restrict__ coef, const T* __restrict__ Pinf, const T* __restrict__ dts, long n,
                        T* __restrict__ Fs, T* __restrict__ Qs) {
    constexpr int DEG = Taylor<T>::DEG;
    constexpr int DD = D * D;
    __shared__ T sC[(DEG + 1) * DD];
    __shared__ T sP[DD];
    __shared__ T sOut[2 * kDiscThreads * DD];
    for (int e = threadIdx.x; e < (DEG + 1) * DD; e += blockDim.x) sC[e] = coef[8 + e];
    for (int e = threadIdx.x; e < DD; e += blockDim.x) sP[e] = Pinf[e];
    const T normF = coef[0];
    __syncthreads();
    const long base = (long)blockIdx.x * kDiscThreads;
    const long k = base + threadIdx.x;
    if (k < n) {
        const T dt = dts[k];
        T x;
        const int s = pick_squarings<T>(normF * t_abs(dt), x);
        if (dt < T(0)) x = -x;
        T A[DD];
#pragma unroll
        for (int e = 0; e < DD; ++e) A[e] = sC[DEG * DD + e];
#pragma unroll 1
        for (int p = DEG - 1; p >= 0; --p) {
#pragma unroll
            for (int e = 0; e < DD; ++e) A[e] = fma(A[e], x, sC[p * DD + e]);
        }
        for (int i = 0; i < s; ++i) {
            T B[DD];
            mm_ff<T, D>(A, A, B);
#pragma unroll
            for (int e = 0; e < DD; ++e) A[e] = B[e];
        }
        // Q = Pinf - A Pinf A^T, symmetrised
        T P[nsym(D)];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) P[sidx(i, j)] = T(0.5) * (sP[i * D + j] + sP[j * D + i]);
        T AP[DD];
        mm_fs<T, D>(A, P, AP);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                T a1 = T(0), a2 = T(0);
#pragma unroll
                for (int kk = 0; kk < D; ++kk) {
                    a1 = fma(AP[i * D + kk], A[j * D + kk], a1);
                    a2 = fma(AP[j * D + kk], A[i * D + kk], a2);
                }
                sOut[kDiscThreads * DD + threadIdx.x * DD + i * D + j] = P[sidx(i, j)] - T(0.5) * (a1 + a2);
            }
#pragma unroll
        for (int e = 0; e < DD; ++e) sOut[threadIdx.x * DD + e] = A[e];
    }
    __syncthreads();
    long cnt = n - base;
    if (cnt > kDiscThreads) cnt = kDiscThreads;
    const long tot = cnt * DD;
    T* gF = Fs + base * DD;
    T* gQ = Qs + base * DD;
    for (long e = threadIdx.x; e < tot; e += blockDim.x) {
        gF[e] = sOut[e];
        gQ[e] = sOut[kDiscThreads * DD + e];
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Per-block partial sums: part[b][ (DEG)*DD (W_1..W_DEG) + DD (dPinf) ].
template <typename T, int D> __host__ __device__ constexpr int bwd_tile() { return sizeof(T) * D * D > 100 ? 64 : 128; }

template <typename T, int D>
__global__ void __launch_bounds__((bwd_tile<T, D>()))
discretise_bwd_small_kernel(const T* __restrict__ coef, const T* __restrict__ Pinf, const T* __restrict__ dts,
                            long n, const T* __restrict__ Fs, const T* __restrict__ dFs,
                            const T* __restrict__ dQs, T* __restrict__ part) {
    constexpr int DEG = Taylor<T>::DEG;
    constexpr int TB = bwd_tile<T, D>();
    constexpr int DD = D * D;
    constexpr int NOUT = (DEG + 1) * DD;
    constexpr int PER = (NOUT + TB - 1) / TB;
    __shared__ T sC[(DEG + 1) * DD];
    __shared__ T sP[DD];
    __shared__ T sA[TB * DD];        // dA_h per step
    __shared__ T sQ[TB * DD];        // dPinf contribution per step
    __shared__ T sXp[DEG * (TB + 1)]; // x^p per step, p-major
    __shared__ T sMax[TB / 32];
    for (int e = threadIdx.x; e < (DEG + 1) * DD; e += blockDim.x) sC[e] = coef[8 + e];
    for (int e = threadIdx.x; e < DD; e += blockDim.x) sP[e] = Pinf[e];
    const T normF = coef[0];
    __syncthreads();
    T P[nsym(D)];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) P[sidx(i, j)] = T(0.5) * (sP[i * D + j] + sP[j * D + i]);
    T accv[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) accv[q] = T(0);
    const long ntiles = (n + TB - 1) / TB;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long k = tile * TB + threadIdx.x;
        T dA[DD], dPi[DD];
        T x = T(0);
        if (k < n) {
            const T dt = dts[k];
            const int s = pick_squarings<T>(normF * t_abs(dt), x);
            if (dt < T(0)) x = -x;
            T A[DD], dQ[nsym(D)];
#pragma unroll
            for (int e = 0; e < DD; ++e) A[e] = Fs[k * DD + e];
#pragma unroll
            for (int e = 0; e < DD; ++e) dA[e] = dFs[k * DD + e];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j)
                    dQ[sidx(i, j)] = T(0.5) * (dQs[k * DD + i * D + j] + dQs[k * DD + j * D + i]);
            // dA_tot = dA - 2 dQ A Pinf ; dPinf += dQ - A^T dQ A
            T QA[DD], QAP[DD];
            mm_sf<T, D>(dQ, A, QA);
            mm_fs<T, D>(QA, P, QAP);
#pragma unroll
            for (int e = 0; e < DD; ++e) dA[e] -= T(2) * QAP[e];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    T a1 = dQ[sidx(i, j)];
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) a1 = fma(-A[kk * D + i], QA[kk * D + j], a1);
                    dPi[i * D + j] = a1;
                }
            if (s > 0) {
                // recompute the squaring chain A_0 = A_h, A_{i+1} = A_i^2 and back-propagate through it
                T chain[kMaxSquarings][DD];
                T B[DD];
#pragma unroll
                for (int e = 0; e < DD; ++e) B[e] = sC[DEG * DD + e];
#pragma unroll 1
                for (int p = DEG - 1; p >= 0; --p) {
#pragma unroll
                    for (int e = 0; e < DD; ++e) B[e] = fma(B[e], x, sC[p * DD + e]);
                }
                for (int i = 0; i < s; ++i) {
#pragma unroll
                    for (int e = 0; e < DD; ++e) chain[i][e] = B[e];
                    T B2[DD];
                    mm_ff<T, D>(B, B, B2);
#pragma unroll
                    for (int e = 0; e < DD; ++e) B[e] = B2[e];
                }
                for (int i = s - 1; i >= 0; --i) {
                    // dA_i = A_i^T dA_{i+1} + dA_{i+1} A_i^T
                    T t1[DD];
#pragma unroll
                    for (int e = 0; e < DD; ++e) B[e] = chain[i][e];
                    mm_tf<T, D>(B, dA, t1);
#pragma unroll
                    for (int r = 0; r < D; ++r)
#pragma unroll
                        for (int c = 0; c < D; ++c) {
                            T a1 = t1[r * D + c];
#pragma unroll
                            for (int kk = 0; kk < D; ++kk) a1 = fma(dA[r * D + kk], B[c * D + kk], a1);
                            t1[r * D + c] = a1;
                        }
#pragma unroll
                    for (int e = 0; e < DD; ++e) dA[e] = t1[e];
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < DD; ++e) {
                dA[e] = T(0);
                dPi[e] = T(0);
            }
        }
        __syncthreads();  // previous tile's readers are done
#pragma unroll
        for (int e = 0; e < DD; ++e) {
            sA[threadIdx.x * DD + e] = dA[e];
            sQ[threadIdx.x * DD + e] = dPi[e];
        }
        {
            // power table x^1..x^DEG (p-major, padded against bank conflicts) and the tile's max |x|
            T xp = T(1);
#pragma unroll 1
            for (int p = 1; p <= DEG; ++p) {
                xp *= x;
                sXp[(p - 1) * (TB + 1) + threadIdx.x] = xp;
            }
            T ax = t_abs(x);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                T o = shfl_down_t(ax, off);
                ax = o > ax ? o : ax;
            }
            if ((threadIdx.x & 31) == 0) sMax[threadIdx.x >> 5] = ax;
        }
        __syncthreads();
        // degrees whose weight x^(p-1)/p! is below 1e-19 (f64) / 1e-10 (f32) for every step of the tile are skipped
        int pmax = DEG;
        {
            T ax = sMax[0];
            for (int w = 1; w < TB / 32; ++w) ax = sMax[w] > ax ? sMax[w] : ax;
            const T tol = sizeof(T) == 8 ? T(1e-19) : T(1e-10);
            T wgt = T(1);
            for (int p = 1; p <= DEG; ++p) {
                if (p > 1) wgt *= ax / T(p);
                if (wgt < tol) {
                    pmax = p - 1;
                    break;
                }
            }
            if (pmax < 1) pmax = 1;
        }
        // re-partition: thread owns outputs o = threadIdx.x + q*blockDim.x ; o = p*DD + e (p = 0 -> dPinf, p >= 1 -> W_p)
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int o = threadIdx.x + q * TB;
            if (o < NOUT) {
                const int p = o / DD, e = o % DD;
                T acc = accv[q];
                if (p == 0) {
                    for (int t = 0; t < TB; ++t) acc += sQ[t * DD + e];
                } else if (p <= pmax) {
                    const T* xp = sXp + (p - 1) * (TB + 1);
#pragma unroll 4
                    for (int t = 0; t < TB; ++t) acc = fma(xp[t], sA[t * DD + e], acc);
                }
                accv[q] = acc;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int o = threadIdx.x + q * TB;
        if (o < NOUT) part[(long)blockIdx.x * NOUT + o] = accv[q];
    }
}

// Sums the per-block partials and assembles dF and dPinf.  One block; generic d.
// coefT holds C_j of G^T (setup kernel called with transpose = 1).  W: NOUT scalars, V: DEG * d * d scalars of
// global scratch.  Every phase is element-parallel with one barrier between phases (the first version walked the
// (i, l) terms one after the other with two barriers each: 169 us at d = 3, all of it latency).
template <typename T>
__global__ void discretise_bwd_final_kernel(const T* __restrict__ coefT, const T* __restrict__ part, int nparts, int d,
                                            T* __restrict__ W, T* __restrict__ V, T* __restrict__ dF,
                                            T* __restrict__ dPinf) {
    const int DEG = Taylor<T>::DEG;
    const int dd = d * d;
    const int NOUT = (DEG + 1) * dd;
    extern __shared__ unsigned char smem_raw[];
    T* red = (T*)smem_raw;  // blockDim.x scalars
    // phase 0: W[o] = sum over the partials, G thread groups share the partials of an output (fixed order)
    const int G = NOUT < (int)blockDim.x ? (int)blockDim.x / NOUT : 1;
    if (G > 1) {
        const int g = threadIdx.x / NOUT, o = threadIdx.x - g * NOUT;
        T sacc = T(0);
        if (g < G)
            for (int b = g; b < nparts; b += G) sacc += part[(long)b * NOUT + o];
        red[threadIdx.x] = sacc;
        __syncthreads();
        if ((int)threadIdx.x < NOUT) {
            T tot = T(0);
            for (int gg = 0; gg < G; ++gg) tot += red[gg * NOUT + threadIdx.x];
            W[threadIdx.x] = tot;
        }
    } else {
        for (int o = threadIdx.x; o < NOUT; o += blockDim.x) {
            T sacc = T(0);
            for (int b = 0; b < nparts; ++b) sacc += part[(long)b * NOUT + o];
            W[o] = sacc;
        }
    }
    __syncthreads();
    const T normF = coefT[0];
    const T inv = normF > T(0) ? T(1) / normF : T(0);
    // dF = inv * sum_{p} (1/p!) sum_{i=0}^{p-1} B^i W_p B^(p-1-i) = inv * sum_i C_i V_i   with C_j = B^j / j! and
    // V_i = sum_l [ i! l! / (i+l+1)! ] W_{i+l+1} C_l
    // phase 1: one thread per element of every V_i
    for (int idx = threadIdx.x; idx < DEG * dd; idx += blockDim.x) {
        const int i = idx / dd, e = idx - i * dd;
        const int r = e / d, c = e - r * d;
        T v = T(0);
        T wgt = T(1) / T(i + 1);  // i! 0! / (i+1)!
        for (int l = 0; i + l + 1 <= DEG; ++l) {
            const T* Wp = W + (size_t)(i + l + 1) * dd;
            const T* Cl = coefT + 8 + (size_t)l * dd;
            T a1 = T(0);
            for (int k = 0; k < d; ++k) a1 = fma(Wp[r * d + k], Cl[k * d + c], a1);
            v = fma(wgt, a1, v);
            wgt = wgt * T(l + 1) / T(i + l + 2);  // i!(l+1)!/(i+l+2)!
        }
        V[idx] = v;
    }
    __syncthreads();
    // phase 2: one thread per element of dF (and dPinf = W_0)
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {
        const int r = e / d, c = e - r * d;
        T acc = T(0);
        for (int i = 0; i < DEG; ++i) {
            const T* Ci = coefT + 8 + (size_t)i * dd;
            const T* Vi = V + (size_t)i * dd;
            T a1 = T(0);
            for (int k = 0; k < d; ++k) a1 = fma(Ci[r * d + k], Vi[k * d + c], a1);
            acc += a1;
        }
        dF[e] = acc * inv;
        dPinf[e] = W[e];
    }
}

int discretise_frag_f64(pssgp_handle* h, int64_t n, int d, const double* coef, const double* Pinf, const double* dts,
                        double* Fs, double* Qs, cudaStream_t st);
int discretise_bwd_frag_f64(pssgp_handle* h, int64_t n, int d, const double* coef, const double* Pinf, const double* dts,
                            const double* Fs, const double* dFs, const double* dQs, double* part, long* grid_out,
                            int* per_cta, cudaStream_t st);

constexpr size_t coef_count(int deg, int d) { return 8 + (size_t)(deg + 1) * d * d; }

template <typename T>
int setup_coef(pssgp_handle* h, const void* F, int d, int transpose, T** coef_out, size_t slot_offset, cudaStream_t st) {
    const size_t cnt = coef_count(Taylor<T>::DEG, d);
    *coef_out = (T*)h->buf[WS_MISC] + slot_offset;
    PSSGP_LAUNCH(h, "taylor_setup", st,
                 (taylor_setup_kernel<T><<<1, 256, 3 * d * d * sizeof(T), st>>>((const T*)F, d, *coef_out, transpose)));
    (void)cnt;
    return check_launch(h, "taylor_setup", 0);
}

template <typename T, int D>
int discretise_impl(pssgp_handle* h, int64_t n, const void* F, const void* Pinf, const void* dts, void* Fs, void* Qs,
                    cudaStream_t st) {
    int rc;
    const size_t cnt = coef_count(Taylor<T>::DEG, D);
    if ((rc = ws_reserve(h, WS_MISC, sizeof(T) * cnt * 2))) return rc;
    T* coef;
    if ((rc = setup_coef<T>(h, F, D, 0, &coef, 0, st))) return rc;
    const unsigned grid = (unsigned)((n + kDiscThreads - 1) / kDiscThreads);
    PSSGP_LAUNCH(h, "discretise", st,
                 (discretise_small_kernel<T, D><<<grid, kDiscThreads, 0, st>>>(coef, (const T*)Pinf, (const T*)dts, n,
                                                                              (T*)Fs, (T*)Qs)));
    return check_launch(h, "discretise", 2);
}

template <typename T, int D>
int discretise_bwd_impl(pssgp_handle* h, int64_t n, const void* F, const void* Pinf, const void* dts, const void* Fs,
                        const void* dFs, const void* dQs, void* dF, void* dPinf, cudaStream_t st) {
    int rc;
    constexpr int DEG = Taylor<T>::DEG;
    const size_t cnt = coef_count(DEG, D);
    const int NOUT = (DEG + 1) * D * D;
    constexpr int TB = bwd_tile<T, D>();
    const long ntiles = (n + TB - 1) / TB;
    int grid = h->num_sms * 4;
    if (grid > ntiles) grid = (int)ntiles;
    if (grid < 1) grid = 1;
    if ((rc = ws_reserve(h, WS_MISC, sizeof(T) * (cnt * 2 + 2 * NOUT)))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * (size_t)NOUT * grid))) return rc;
    T *coef, *coefT;
    if ((rc = setup_coef<T>(h, F, D, 0, &coef, 0, st))) return rc;
    if ((rc = setup_coef<T>(h, F, D, 1, &coefT, cnt, st))) return rc;
    T* W = (T*)h->buf[WS_MISC] + 2 * cnt;
    T* V = W + NOUT;
    T* part = (T*)h->buf[WS_PART];
    PSSGP_LAUNCH(h, "discretise_bwd", st,
                 (discretise_bwd_small_kernel<T, D><<<grid, TB, 0, st>>>(coef, (const T*)Pinf, (const T*)dts, n,
                                                                        (const T*)Fs, (const T*)dFs, (const T*)dQs, part)));
    PSSGP_LAUNCH(h, "discretise_bwd_final", st,
                 (discretise_bwd_final_kernel<T><<<1, 1024, 1024 * sizeof(T), st>>>(coefT, part, grid, D, W, V, (T*)dF,
                                                                                       (T*)dPinf)));
    return check_launch(h, "discretise_backward", 4);
}

// ---------------------------------------------------------------------------------------------
// generic state dimension: one CTA per time step (grid-stride), matrices in shared memory
// ---------------------------------------------------------------------------------------------
// One thread per matrix element (i, j); matrices in shared memory with an odd leading dimension, every product reads one
// operand as a row broadcast and the other along consecutive columns (no bank conflicts, no division in the loop).
// FP64 with d <= 32 runs discretise_frag.cu instead; this kernel serves FP32 and the option "force_generic".
template <typename T>
__global__ void __launch_bounds__(1024) discretise_generic_kernel(const T* __restrict__ coef, const T* __restrict__ Pinf, int d,
                                          const T* __restrict__ dts, long n, T* __restrict__ Fs, T* __restrict__ Qs) {
    constexpr int DEG = Taylor<T>::DEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dd = d * d, LD = d | 1, MS = d * LD;
    T* A = (T*)smem_raw;
    T* B = A + MS;
    T* P = B + MS;
    const int e = threadIdx.x;
    const bool on = e < dd;
    const int i = on ? e / d : 0, j = on ? e - i * d : 0;
    const int ij = i * LD + j;
    T pij = T(0);
    if (on) {
        pij = T(0.5) * (Pinf[e] + Pinf[j * d + i]);
        P[ij] = pij;
    }
    const T normF = coef[0];
    const T* C = coef + 8;
    __syncthreads();
    for (long k = blockIdx.x; k < n; k += gridDim.x) {
        const T dt = dts[k];
        T x;
        const int s = pick_squarings<T>(normF * t_abs(dt), x);
        if (dt < T(0)) x = -x;
        T a = T(0);
        if (on) {
            a = C[(size_t)DEG * dd + e];
            for (int p = DEG - 1; p >= 0; --p) a = fma(a, x, C[(size_t)p * dd + e]);
            A[ij] = a;
        }
        __syncthreads();
        T* cur = A;
        T* nxt = B;
        for (int q = 0; q < s; ++q) {
            if (on) {
                a = T(0);
                for (int kk = 0; kk < d; ++kk) a = fma(cur[i * LD + kk], cur[kk * LD + j], a);
                nxt[ij] = a;
            }
            __syncthreads();
            T* t = cur;
            cur = nxt;
            nxt = t;
        }
        if (on) {
            Fs[k * dd + e] = a;
            T ap = T(0);   // (A P)_ij
            for (int kk = 0; kk < d; ++kk) ap = fma(cur[i * LD + kk], P[kk * LD + j], ap);
            nxt[ij] = ap;
        }
        __syncthreads();
        if (on) {
            // Q_ij = P_ij - ((A P A^T)_ij + (A P A^T)_ji) / 2
            T a1 = T(0), a2 = T(0);
            for (int kk = 0; kk < d; ++kk) {
                a1 = fma(nxt[i * LD + kk], cur[j * LD + kk], a1);
                a2 = fma(nxt[j * LD + kk], cur[i * LD + kk], a2);
            }
            Qs[k * dd + e] = pij - T(0.5) * (a1 + a2);
        }
        __syncthreads();
    }
}

// part[cta][p*dd + e]: p = 0 -> dPinf, p >= 1 -> W_p.  ONE THREAD PER MATRIX ELEMENT (i, j): its DEG + 1 moment
// accumulators live in registers for the whole grid-stride loop and are written once at the end; matrices sit in shared
// memory with an odd leading dimension, and every product reads one operand as a row broadcast and the other along
// consecutive columns (no bank conflicts, no integer division inside the loop).
template <typename T, int MAXNT>
__global__ void __launch_bounds__(MAXNT)
discretise_bwd_generic_kernel(const T* __restrict__ coef, const T* __restrict__ Pinf, int d,
                                              const T* __restrict__ dts, long n, const T* __restrict__ Fs,
                                              const T* __restrict__ dFs, const T* __restrict__ dQs,
                                              T* __restrict__ part) {
    constexpr int DEG = Taylor<T>::DEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dd = d * d, LD = d | 1, MS = d * LD;
    T* A = (T*)smem_raw;   // A_k
    T* dA = A + MS;
    T* dQ = dA + MS;
    T* T1 = dQ + MS;
    T* T2 = T1 + MS;
    T* P = T2 + MS;
    T* Ah = P + MS;
    const int e = threadIdx.x;
    const bool on = e < dd;
    const int i = on ? e / d : 0, j = on ? e - i * d : 0;
    const int ij = i * LD + j, eT = j * d + i;
    T acc[DEG + 1];
#pragma unroll
    for (int p = 0; p <= DEG; ++p) acc[p] = T(0);
    if (on) P[ij] = T(0.5) * (Pinf[e] + Pinf[eT]);
    const T normF = coef[0];
    const T* C = coef + 8;
    __syncthreads();
    for (long k = blockIdx.x; k < n; k += gridDim.x) {
        const T dt = dts[k];
        T x;
        const int s = pick_squarings<T>(normF * t_abs(dt), x);
        if (dt < T(0)) x = -x;
        T dq = T(0), da = T(0);
        if (on) {
            A[ij] = Fs[k * dd + e];
            da = dFs[k * dd + e];
            dq = T(0.5) * (dQs[k * dd + e] + dQs[k * dd + eT]);
            dQ[ij] = dq;
        }
        __syncthreads();
        if (on) {  // T1 = dQ A
            T a1 = T(0);
            for (int kk = 0; kk < d; ++kk) a1 = fma(dQ[i * LD + kk], A[kk * LD + j], a1);
            T1[ij] = a1;
        }
        __syncthreads();
        if (on) {
            T a1 = T(0), a2 = T(0);
            for (int kk = 0; kk < d; ++kk) {
                a1 = fma(T1[i * LD + kk], P[kk * LD + j], a1);   // (dQ A P)_ij
                a2 = fma(A[kk * LD + i], T1[kk * LD + j], a2);   // (A^T dQ A)_ij
            }
            da = fma(T(-2), a1, da);                             // dA -= 2 dQ A P
            acc[0] += dq - a2;                                   // dPinf += dQ - A^T dQ A
        }
        if (s > 0) {  // block-uniform: back-propagate through the squarings A = Ah^(2^s)
            if (on) {
                dA[ij] = da;
                T a = C[(size_t)DEG * dd + e];
                for (int p = DEG - 1; p >= 0; --p) a = fma(a, x, C[(size_t)p * dd + e]);
                Ah[ij] = a;
            }
            __syncthreads();
            for (int lvl = s - 1; lvl >= 0; --lvl) {
                // A_lvl = Ah^(2^lvl) by lvl squarings (recomputed: no per-level storage)
                const T* cur = Ah;
                T* nxt = T1;
                for (int q = 0; q < lvl; ++q) {
                    if (on) {
                        T a1 = T(0);
                        for (int kk = 0; kk < d; ++kk) a1 = fma(cur[i * LD + kk], cur[kk * LD + j], a1);
                        nxt[ij] = a1;
                    }
                    __syncthreads();
                    cur = nxt;
                    nxt = (nxt == T1) ? T2 : T1;
                }
                // dA <- A_lvl^T dA + dA A_lvl^T
                if (on) {
                    T a1 = T(0);
                    for (int kk = 0; kk < d; ++kk) {
                        a1 = fma(cur[kk * LD + i], dA[kk * LD + j], a1);
                        a1 = fma(dA[i * LD + kk], cur[j * LD + kk], a1);
                    }
                    da = a1;
                }
                __syncthreads();
                if (on) dA[ij] = da;
                __syncthreads();
            }
        }
        // moments W_p += x^p dA_h, skipping degrees whose weight x^(p-1)/p! is negligible
        {
            const T ax = t_abs(x);
            const T tol = sizeof(T) == 8 ? T(1e-19) : T(1e-10);
            int pmax = DEG;
            T wgt = T(1);
            for (int p = 1; p <= DEG; ++p) {
                if (p > 1) wgt *= ax / T(p);
                if (wgt < tol) {
                    pmax = p - 1;
                    break;
                }
            }
            if (pmax < 1) pmax = 1;
            T xp = T(1);
#pragma unroll
            for (int p = 1; p <= DEG; ++p) {
                xp *= x;
                if (p <= pmax) acc[p] = fma(xp, da, acc[p]);
            }
        }
        __syncthreads();
    }
    if (on) {
        T* mine = part + (size_t)blockIdx.x * (DEG + 1) * dd;
#pragma unroll
        for (int p = 0; p <= DEG; ++p) mine[(size_t)p * dd + e] = acc[p];
    }
}

// out[g][o] = sum of part[b][o] over the partials b = g, g + G, ... (fixed order: deterministic), G = gridDim.y
template <typename T>
__global__ void partials_reduce_kernel(const T* __restrict__ part, int nparts, int nout, T* __restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nout) return;
    T s = T(0);
    for (int b = blockIdx.y; b < nparts; b += gridDim.y) s += part[(size_t)b * nout + o];
    out[(size_t)blockIdx.y * nout + o] = s;
}

template <typename T>
int discretise_generic_impl(pssgp_handle* h, int64_t n, int d, const void* F, const void* Pinf, const void* dts,
                            void* Fs, void* Qs, cudaStream_t st) {
    int rc;
    const size_t cnt = coef_count(Taylor<T>::DEG, d);
    if ((rc = ws_reserve(h, WS_MISC, sizeof(T) * cnt * 2))) return rc;
    T* coef;
    if ((rc = setup_coef<T>(h, F, d, 0, &coef, 0, st))) return rc;
    if constexpr (sizeof(T) == 8) {
        // FP64, d <= 32: warp-per-step tensor-core kernel (discretise_frag.cu)
        if (d <= 32 && !h->force_generic)
            return discretise_frag_f64(h, n, d, (const double*)coef, (const double*)Pinf, (const double*)dts, (double*)Fs,
                                       (double*)Qs, st);
    }
    const int nt = (d * d + 31) / 32 * 32;   // one thread per matrix element
    if (nt > 1024) return set_err(PSSGP_ERR_UNSUPPORTED, "discretise: d <= 32 (got %d)", d);
    const size_t sm = sizeof(T) * 3 * d * (d | 1);
    int per_sm = 2048 / nt;
    if (per_sm > 16) per_sm = 16;
    if (const char* e = getenv("PSSGP_DISC_CTAS")) per_sm = atoi(e);   // tuning aid
    long grid = (long)h->num_sms * per_sm;
    if (grid > n) grid = n;
    if (sm > 48 * 1024) cudaFuncSetAttribute(discretise_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    PSSGP_LAUNCH(h, "discretise", st,
                 (discretise_generic_kernel<T><<<(unsigned)grid, nt, sm, st>>>(coef, (const T*)Pinf, d, (const T*)dts, n,
                                                                              (T*)Fs, (T*)Qs)));
    return check_launch(h, "discretise", 2);
}

template <typename T>
int discretise_bwd_generic_impl(pssgp_handle* h, int64_t n, int d, const void* F, const void* Pinf, const void* dts,
                                const void* Fs, const void* dFs, const void* dQs, void* dF, void* dPinf,
                                cudaStream_t st) {
    int rc;
    constexpr int DEG = Taylor<T>::DEG;
    constexpr int kGroups = 16;
    const size_t cnt = coef_count(DEG, d);
    const int NOUT = (DEG + 1) * d * d;
    // FP64, d <= 24: warp-per-step tensor-core products + thread-per-element moments (discretise_frag.cu)
    bool frag = false;
    long fgrid = 0;
    int per_cta = 1;
    if constexpr (sizeof(T) == 8) {
        if (d <= 24 && !h->force_generic) {
            frag = true;
            if ((rc = discretise_bwd_frag_f64(h, n, d, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &fgrid,
                                              &per_cta, st)))
                return rc;
        }
    }
    const int nt = (d * d + 31) / 32 * 32;   // one thread per matrix element
    if (nt > 1024) return set_err(PSSGP_ERR_UNSUPPORTED, "discretise_backward: d <= 32 (got %d)", d);
    const size_t sm = sizeof(T) * 7 * d * (d | 1);
    int per_sm = 2048 / nt;
    if (per_sm > 16) per_sm = 16;
    if ((size_t)per_sm * (sm + 1024) > (size_t)200 * 1024) per_sm = (int)((size_t)200 * 1024 / (sm + 1024));
    if (per_sm < 1) per_sm = 1;
    if (const char* e = getenv("PSSGP_DISCB_CTAS")) per_sm = atoi(e);  // tuning aid
    long grid = (long)h->num_sms * per_sm;
    if (grid > n) grid = n;
    if (frag) grid = fgrid * per_cta;   // number of partials
    if ((rc = ws_reserve(h, WS_MISC, sizeof(T) * (cnt * 2 + 2 * NOUT)))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * (size_t)NOUT * (grid + kGroups)))) return rc;
    T *coef, *coefT;
    if ((rc = setup_coef<T>(h, F, d, 0, &coef, 0, st))) return rc;
    if ((rc = setup_coef<T>(h, F, d, 1, &coefT, cnt, st))) return rc;
    T* W = (T*)h->buf[WS_MISC] + 2 * cnt;
    T* V = W + NOUT;
    T* part = (T*)h->buf[WS_PART];
    T* grouped = part + (size_t)NOUT * grid;
    if (frag) {
        if constexpr (sizeof(T) == 8) {
            if ((rc = discretise_bwd_frag_f64(h, n, d, (const double*)coef, (const double*)Pinf, (const double*)dts,
                                              (const double*)Fs, (const double*)dFs, (const double*)dQs, (double*)part, &fgrid,
                                              &per_cta, st)))
                return rc;
        }
    } else {
    // register budget follows the CTA size: 256 threads keep the 19 moment accumulators in registers, the larger
    // CTAs of d > 16 are compiled for 128 / 64 registers per thread (the accumulators spill to local memory there)
    auto launch = [&](auto kern) {
        if (sm > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        PSSGP_LAUNCH(h, "discretise_bwd", st,
                     (kern<<<(unsigned)grid, nt, sm, st>>>(coef, (const T*)Pinf, d, (const T*)dts, n, (const T*)Fs,
                                                           (const T*)dFs, (const T*)dQs, part)));
    };
    if (nt <= 256)
        launch(discretise_bwd_generic_kernel<T, 256>);
    else if (nt <= 512)
        launch(discretise_bwd_generic_kernel<T, 512>);
    else
        launch(discretise_bwd_generic_kernel<T, 1024>);
    }
    const int groups = grid < kGroups ? (int)grid : kGroups;
    PSSGP_LAUNCH(h, "discretise_bwd_reduce", st,
                 (partials_reduce_kernel<T><<<dim3((NOUT + 127) / 128, groups), 128, 0, st>>>(part, (int)grid, NOUT, grouped)));
    PSSGP_LAUNCH(h, "discretise_bwd_final", st,
                 (discretise_bwd_final_kernel<T><<<1, 1024, 1024 * sizeof(T), st>>>(coefT, grouped, groups, d, W, V,
                                                                                   (T*)dF, (T*)dPinf)));
    return check_launch(h, "discretise_backward", 5);
}

}  // namespace pssgp

using namespace pssgp;

#define DISPATCH_SMALL(FN, ...)                                                                   \
    do {                                                                                          \
        if (dtype == PSSGP_F64) {                                                                 \
            switch (d) {                                                                          \
                case 1: return FN<double, 1>(__VA_ARGS__);                                        \
                case 2: return FN<double, 2>(__VA_ARGS__);                                        \
                case 3: return FN<double, 3>(__VA_ARGS__);                                        \
                case 4: return FN<double, 4>(__VA_ARGS__);                                        \
            }                                                                                     \
        } else if (dtype == PSSGP_F32) {                                                          \
            switch (d) {                                                                          \
                case 1: return FN<float, 1>(__VA_ARGS__);                                         \
                case 2: return FN<float, 2>(__VA_ARGS__);                                         \
                case 3: return FN<float, 3>(__VA_ARGS__);                                         \
                case 4: return FN<float, 4>(__VA_ARGS__);                                         \
            }                                                                                     \
        }                                                                                         \
    } while (0)

extern "C" {

int pssgp_discretise(pssgp_handle* h, int dtype, int64_t n, int d, const void* F, const void* Pinf, const void* dts,
                     void* Fs, void* Qs, void* stream) {
    if (!h) return set_err(PSSGP_ERR_INVALID, "null handle");
    if (n < 1 || d < 1) return set_err(PSSGP_ERR_INVALID, "discretise: n and d must be >= 1");
    if (!F || !Pinf || !dts || !Fs || !Qs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != PSSGP_F64 && dtype != PSSGP_F32) return set_err(PSSGP_ERR_INVALID, "bad dtype %d", dtype);
    // Fs / Qs are about to be overwritten: chunk aggregates built from them (pending *_summary calls) are stale
    for (int kind = 0; kind < 3; ++kind) pending_clear(h, kind);
    DISPATCH_SMALL(discretise_impl, h, n, F, Pinf, dts, Fs, Qs, st);
    if (d > 64) return set_err(PSSGP_ERR_UNSUPPORTED, "discretise: state dimension %d > 64", d);
    if (dtype == PSSGP_F64) return discretise_generic_impl<double>(h, n, d, F, Pinf, dts, Fs, Qs, st);
    return discretise_generic_impl<float>(h, n, d, F, Pinf, dts, Fs, Qs, st);
}

int pssgp_discretise_backward(pssgp_handle* h, int dtype, int64_t n, int d, const void* F, const void* Pinf,
                              const void* dts, const void* Fs, const void* dFs, const void* dQs, void* dF, void* dPinf,
                              void* stream) {
    if (!h) return set_err(PSSGP_ERR_INVALID, "null handle");
    if (n < 1 || d < 1) return set_err(PSSGP_ERR_INVALID, "discretise_backward: n and d must be >= 1");
    if (!F || !Pinf || !dts || !Fs || !dFs || !dQs || !dF || !dPinf)
        return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != PSSGP_F64 && dtype != PSSGP_F32) return set_err(PSSGP_ERR_INVALID, "bad dtype %d", dtype);
    DISPATCH_SMALL(discretise_bwd_impl, h, n, F, Pinf, dts, Fs, dFs, dQs, dF, dPinf, st);
    if (d > 64) return set_err(PSSGP_ERR_UNSUPPORTED, "discretise_backward: state dimension %d > 64", d);
    if (dtype == PSSGP_F64) return discretise_bwd_generic_impl<double>(h, n, d, F, Pinf, dts, Fs, dFs, dQs, dF, dPinf, st);
    return discretise_bwd_generic_impl<float>(h, n, d, F, Pinf, dts, Fs, dFs, dQs, dF, dPinf, st);
}

}  // extern "C"
