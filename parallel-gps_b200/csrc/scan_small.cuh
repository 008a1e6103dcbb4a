// Three-kernel chunked scan for register-resident state dimensions (D <= 4).
//
//   K1 reduce : one thread per chunk of L consecutive time steps builds the chunk aggregate by
//               sequential "append" (cheap conditional recursion, no generic operator), then a
//               warp-level Kogge-Stone scan with the generic associative operator. It stores the
//               lane-exclusive aggregate of every chunk and the total of every warp.
//   K2 mid    : one CTA scans the warp totals and turns them into the *state* entering each warp
//               (prefixes that start at the sequence origin collapse to a state: (m,P) for the
//               filter, (ms,Ps) for the smoother, (dm,dP) for the adjoint).
//   K3 apply  : every thread applies state o lane-exclusive-aggregate to get the state entering
//               its chunk and re-runs the cheap seeded recursion over its L steps, writing outputs.
//
// An "Algebra" supplies: NAGG, NSTATE, NACC, Params, identity, append, combine, apply, step,
// load_init, finish.  Direction (forward/reverse in time) is the Algebra's business: the framework
// only sees logical indices 0..n-1.
#pragma once
#include "smalld.cuh"

namespace pssgp {

constexpr int kReduceThreads = 128;
constexpr int kMidThreads = 256;

template <typename Alg>
__global__ void __launch_bounds__(kReduceThreads)
scan_reduce_kernel(typename Alg::Params p, long n, int L, long nChunksPad,
                   typename Alg::scalar* __restrict__ lane_excl,
                   typename Alg::scalar* __restrict__ wagg, long nW) {
    using T = typename Alg::scalar;
    const long chunk = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    long k0 = chunk * (long)L;
    long k1 = k0 + L;
    if (k1 > n) k1 = n;
    T a[Alg::NAGG];
    Alg::identity(a);
    for (long k = k0; k < k1; ++k) Alg::append(a, k, p);
    // warp inclusive scan (earlier lane is the left operand)
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        T o[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(a[e], off);
        if (lane >= off) {
            T r[Alg::NAGG];
            Alg::combine(o, a, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    // block level: every warp folds the totals of the warps before it (<= 3 combines)
    __shared__ T shw[(kReduceThreads / 32) * Alg::NAGG];
    const int wid = threadIdx.x >> 5;
    if (lane == 31) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) shw[wid * Alg::NAGG + e] = a[e];
    }
    T ex[Alg::NAGG];
#pragma unroll
    for (int e = 0; e < Alg::NAGG; ++e) ex[e] = shfl_up_t(a[e], 1);
    if (lane == 0) Alg::identity(ex);
    __syncthreads();
    if (wid > 0) {
        T wp[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) wp[e] = shw[e];
#pragma unroll 1
        for (int w = 1; w < wid; ++w) {
            T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) b[e] = shw[w * Alg::NAGG + e];
            Alg::combine(wp, b, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) wp[e] = r[e];
        }
        T r[Alg::NAGG];
        Alg::combine(wp, ex, r);
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) ex[e] = r[e];
    }
    if (threadIdx.x != 0 && k0 < n) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) lane_excl[(long)e * nChunksPad + chunk] = ex[e];
    }
    if (threadIdx.x == kReduceThreads - 1) {
        // block total = exclusive prefix of the last thread o its own chunk aggregate... the last
        // lane's inclusive warp aggregate `a` already covers its warp, so fold the warp prefix in.
        T tot[Alg::NAGG];
        if (wid > 0) {
            T wp[Alg::NAGG];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) wp[e] = shw[e];
#pragma unroll 1
            for (int w = 1; w < wid; ++w) {
                T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) b[e] = shw[w * Alg::NAGG + e];
                Alg::combine(wp, b, r);
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) wp[e] = r[e];
            }
            Alg::combine(wp, a, tot);
        } else {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) tot[e] = a[e];
        }
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) wagg[(long)e * nW + blockIdx.x] = tot[e];
    }
}

// Single CTA.  wstate[s*nW + w] = state entering warp w.  final_state = state after everything.
template <typename Alg>
__global__ void __launch_bounds__(kMidThreads)
scan_mid_kernel(typename Alg::Params p, const typename Alg::scalar* __restrict__ wagg, long nW,
                typename Alg::scalar* __restrict__ wstate,
                typename Alg::scalar* __restrict__ final_state) {
    using T = typename Alg::scalar;
    __shared__ T sh[32 * Alg::NAGG];
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int nthreads = blockDim.x;
    const long per = (nW + nthreads - 1) / nthreads;
    long i0 = (long)tid * per, i1 = i0 + per;
    if (i0 > nW) i0 = nW;
    if (i1 > nW) i1 = nW;
    T a[Alg::NAGG];
    Alg::identity(a);
    for (long i = i0; i < i1; ++i) {
        T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) b[e] = wagg[(long)e * nW + i];
        if (i == i0) {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = b[e];
        } else {
            Alg::combine(a, b, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    // block-level exclusive scan of the per-thread aggregates
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        T o[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(a[e], off);
        if (lane >= off) {
            T r[Alg::NAGG];
            Alg::combine(o, a, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) sh[e * 32 + wid] = a[e];
    }
    T ex[Alg::NAGG];  // lane-exclusive within the warp
#pragma unroll
    for (int e = 0; e < Alg::NAGG; ++e) ex[e] = shfl_up_t(a[e], 1);
    if (lane == 0) Alg::identity(ex);
    __syncthreads();
    const int nwarps = nthreads >> 5;
    if (wid == 0) {
        T w[Alg::NAGG];
        if (lane < nwarps) {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) w[e] = sh[e * 32 + lane];
        } else {
            Alg::identity(w);
        }
#pragma unroll 1
        for (int off = 1; off < 32; off <<= 1) {
            T o[Alg::NAGG];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(w[e], off);
            if (lane >= off) {
                T r[Alg::NAGG];
                Alg::combine(o, w, r);
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) w[e] = r[e];
            }
        }
        // exclusive over warps
        T wx[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) wx[e] = shfl_up_t(w[e], 1);
        if (lane == 0) Alg::identity(wx);
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) sh[e * 32 + lane] = wx[e];
    }
    __syncthreads();
    T s[Alg::NSTATE];
    Alg::load_init(p, s);
    if (wid > 0) {
        T w[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) w[e] = sh[e * 32 + wid];
        T s2[Alg::NSTATE];
        Alg::apply(s, w, s2);
#pragma unroll
        for (int e = 0; e < Alg::NSTATE; ++e) s[e] = s2[e];
    }
    if (lane > 0) {
        T s2[Alg::NSTATE];
        Alg::apply(s, ex, s2);
#pragma unroll
        for (int e = 0; e < Alg::NSTATE; ++e) s[e] = s2[e];
    }
    for (long i = i0; i < i1; ++i) {
#pragma unroll
        for (int e = 0; e < Alg::NSTATE; ++e) wstate[(long)e * nW + i] = s[e];
        T b[Alg::NAGG], s2[Alg::NSTATE];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) b[e] = wagg[(long)e * nW + i];
        Alg::apply(s, b, s2);
#pragma unroll
        for (int e = 0; e < Alg::NSTATE; ++e) s[e] = s2[e];
    }
    // the thread that owns the last warp total holds the final state
    if (final_state != nullptr && i1 == nW && i0 < nW) Alg::expand_state(s, final_state);
    if (final_state != nullptr && nW == 0 && tid == 0) Alg::expand_state(s, final_state);
}

template <typename Alg>
__global__ void __launch_bounds__(kReduceThreads)
scan_apply_kernel(typename Alg::Params p, long n, int L, long nChunksPad,
                  const typename Alg::scalar* __restrict__ lane_excl,
                  const typename Alg::scalar* __restrict__ wstate, long nW,
                  typename Alg::scalar* __restrict__ acc_part, unsigned int* __restrict__ ticket,
                  typename Alg::scalar* __restrict__ acc_out) {
    using T = typename Alg::scalar;
    const long chunk = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    long k0 = chunk * (long)L;
    long k1 = k0 + L;
    if (k1 > n) k1 = n;
    T acc[Alg::NACC > 0 ? Alg::NACC : 1];
#pragma unroll
    for (int e = 0; e < (Alg::NACC > 0 ? Alg::NACC : 1); ++e) acc[e] = T(0);
    if (k0 < n) {
        T s[Alg::NSTATE];
#pragma unroll
        for (int e = 0; e < Alg::NSTATE; ++e) s[e] = wstate[(long)e * nW + blockIdx.x];
        if (threadIdx.x != 0) {
            T ex[Alg::NAGG], s2[Alg::NSTATE];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) ex[e] = lane_excl[(long)e * nChunksPad + chunk];
            Alg::apply(s, ex, s2);
#pragma unroll
            for (int e = 0; e < Alg::NSTATE; ++e) s[e] = s2[e];
        }
        for (long k = k0; k < k1; ++k) Alg::step(s, k, p, acc);
    }
    if (Alg::NACC > 0) {
        // deterministic grid reduction: warp shuffle -> smem -> per-block partial -> last block sums in order
        __shared__ T red[(kReduceThreads / 32) * (Alg::NACC > 0 ? Alg::NACC : 1)];
        __shared__ bool is_last;
#pragma unroll
        for (int e = 0; e < Alg::NACC; ++e) {
            T v = acc[e];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += shfl_down_t(v, off);
            if (lane == 0) red[(threadIdx.x >> 5) * Alg::NACC + e] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int e = 0; e < Alg::NACC; ++e) {
                T v = T(0);
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w * Alg::NACC + e];
                acc_part[(long)blockIdx.x * Alg::NACC + e] = v;
            }
            __threadfence();
            unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            const int nb = gridDim.x;
            for (int e = 0; e < Alg::NACC; ++e) {
                // fixed-shape tree: thread t sums blocks t, t+128, ... then ordered smem tree
                T v = T(0);
                for (int b = threadIdx.x; b < nb; b += blockDim.x)
                    v += ((volatile T*)acc_part)[(long)b * Alg::NACC + e];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += shfl_down_t(v, off);
                __syncthreads();
                if (lane == 0) red[threadIdx.x >> 5] = v;
                __syncthreads();
                if (threadIdx.x == 0) {
                    T tot = T(0);
                    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
                    Alg::finish(p, e, tot, acc_out);
                }
            }
            if (threadIdx.x == 0) *ticket = 0u;
        }
    }
}

// Single CTA: out[NAGG] = wagg[0] o wagg[1] o ... o wagg[nW-1]  (shard summary for time sharding).
template <typename Alg>
__global__ void __launch_bounds__(kMidThreads)
scan_total_kernel(const typename Alg::scalar* __restrict__ wagg, long nW, typename Alg::scalar* __restrict__ out) {
    using T = typename Alg::scalar;
    __shared__ T sh[(kMidThreads / 32) * Alg::NAGG];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nthreads = blockDim.x;
    const long per = (nW + nthreads - 1) / nthreads;
    long i0 = (long)tid * per, i1 = i0 + per;
    if (i0 > nW) i0 = nW;
    if (i1 > nW) i1 = nW;
    T a[Alg::NAGG];
    Alg::identity(a);
    for (long i = i0; i < i1; ++i) {
        T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) b[e] = wagg[(long)e * nW + i];
        Alg::combine(a, b, r);
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
    }
    // ordered tree reduction inside the warp: lane l <- a[l] o a[l+off]
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        T o[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_down_t(a[e], off);
        if ((lane & (2 * off - 1)) == 0) {
            T r[Alg::NAGG];
            Alg::combine(a, o, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) sh[wid * Alg::NAGG + e] = a[e];
    }
    __syncthreads();
    if (tid == 0) {
        const int nwarps = nthreads >> 5;
        for (int w = 1; w < nwarps; ++w) {
            T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) b[e] = sh[w * Alg::NAGG + e];
            Alg::combine(a, b, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) out[e] = a[e];
    }
}

// One thread: s = init; for i in 0..count-1: s = s o summaries[i*stride ...]; out = Alg::expand(s).
template <typename Alg>
__global__ void scan_fold_kernel(typename Alg::Params p, const typename Alg::scalar* __restrict__ summaries,
                                 int count, long stride, typename Alg::scalar* __restrict__ out) {
    using T = typename Alg::scalar;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T s[Alg::NSTATE];
    Alg::load_init(p, s);
    for (int i = 0; i < count; ++i) {
        T b[Alg::NAGG], s2[Alg::NSTATE];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) b[e] = summaries[(long)i * stride + e];
        Alg::apply(s, b, s2);
#pragma unroll
        for (int e = 0; e < Alg::NSTATE; ++e) s[e] = s2[e];
    }
    Alg::expand_state(s, out);
}

}  // namespace pssgp
