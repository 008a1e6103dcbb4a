// Single-CTA kernels of the chunked scan for register-resident state dimensions (D <= 4): the scan
// over CTA totals (K2 of scan_stream.cuh), the shard summary and the fold of gathered summaries (time
// sharding).  The streaming kernels K1 / K3 live in scan_stream.cuh.
#pragma once
#include "smalld.cuh"

namespace pssgp {

constexpr int kMidThreads = 256;

#ifdef PSSGP_PHASES  // tuning aid: %globaltimer stamps of the phases of K1 and of the scan over CTA totals (thread 0 of every CTA)
static __device__ unsigned long long g_phase[1024 * 8];
PSSGP_DEV void phase_stamp(int slot) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_phase[blockIdx.x * 8 + slot] = t;
    }
}
#define PSSGP_PHASE(slot) phase_stamp(slot)
#else
#define PSSGP_PHASE(slot)
#endif


// One CTA (all of its threads).  wstate[s*nW + w] = state entering CTA w of K1/K3.  final_state = state after
// everything.  sh: 32 * NAGG scalars of shared memory.  Runs either as its own kernel (scan_mid_kernel) or at
// the end of K1 in the CTA that finishes last (scan_stream.cuh).
// Barrier over a group of `nthreads` threads (multiple of 32) of the CTA: id 0 with the whole CTA is
// __syncthreads(); other ids let two halves of a CTA run independent scans side by side.
__device__ __forceinline__ void group_sync(int bar_id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
}

// tid / nthreads: this thread's index in, and the size of, the thread group that runs the scan (whole warps).
template <typename Alg>
__device__ __forceinline__ void scan_mid_body(const typename Alg::Params& p, const typename Alg::scalar* wagg, long nW,
                                              typename Alg::scalar* wstate, typename Alg::scalar* final_state,
                                              typename Alg::scalar* sh, int tid, int nthreads, int bar_id) {
    using T = typename Alg::scalar;
    const int lane = tid & 31, wid = tid >> 5;
    const long per = (nW + nthreads - 1) / nthreads;
    long i0 = (long)tid * per, i1 = i0 + per;
    if (i0 > nW) i0 = nW;
    if (i1 > nW) i1 = nW;
    T a[Alg::NAGG];
    Alg::identity(a);
    for (long i = i0; i < i1; ++i) {
        T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) b[e] = __ldcg(wagg + (long)e * nW + i);
        if (i == i0) {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = b[e];
        } else {
            Alg::combine(a, b, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    PSSGP_PHASE(4);
    // Block-level exclusive scan of the per-thread aggregates.  The warp level (5 shuffle steps) and the level over
    // warp totals (done by warp 0) run through ONE loop body, and all the applies below through another one: every
    // inlined copy of combine / apply is several KB of straight-line code that would be fetched cold from the
    // instruction cache for a handful of executions (measured ~3.5 us per cold copy of FilterAlg::combine, ~0.5 us warm).
    const int nwarps = nthreads >> 5;
    int logw = 0;
    while ((1 << logw) < nwarps) ++logw;
    T ex[Alg::NAGG];  // lane-exclusive within the warp
#pragma unroll 1
    for (int lvl = 0;; ++lvl) {
        if (lvl == 5) {
            PSSGP_PHASE(5);
            if (lane == 31) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) sh[e * 32 + wid] = a[e];
            }
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) ex[e] = shfl_up_t(a[e], 1);
            if (lane == 0) Alg::identity(ex);
            group_sync(bar_id, nthreads);
            if (wid != 0) break;
            if (lane < nwarps) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) a[e] = sh[e * 32 + lane];
            } else {
                Alg::identity(a);
            }
        }
        if (lvl == 5 + logw) break;
        const int off = 1 << (lvl < 5 ? lvl : lvl - 5);
        T o[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(a[e], off);
        if (lane >= off && (lvl < 5 || lane < nwarps)) {
            T r[Alg::NAGG];
            Alg::combine(o, a, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    if (wid == 0) {
        // exclusive over warps
        T wx[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) wx[e] = shfl_up_t(a[e], 1);
        if (lane == 0) Alg::identity(wx);
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) sh[e * 32 + lane] = wx[e];
    }
    group_sync(bar_id, nthreads);
    PSSGP_PHASE(6);
    T s[Alg::NSTATE];
    Alg::load_init(p, s);
    // s = init o (prefix of the earlier warps) o (prefix of the earlier lanes), then item by item
    const bool want_final = final_state != nullptr && i1 == nW;
    const long nsteps = 2 + (i1 - i0);
#pragma unroll 1
    for (long j = 0; j < nsteps; ++j) {
        T b[Alg::NAGG];
        bool act;
        if (j == 0) {
            act = wid > 0;
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) b[e] = sh[e * 32 + wid];
        } else if (j == 1) {
            act = lane > 0;
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) b[e] = ex[e];
        } else {
            const long i = i0 + (j - 2);
#pragma unroll
            for (int e = 0; e < Alg::NSTATE; ++e) wstate[(long)e * nW + i] = s[e];
            act = (i + 1 < i1) || want_final;  // the state after the last item is only needed as the final state
            if (act) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) b[e] = __ldcg(wagg + (long)e * nW + i);
            }
        }
        if (act) {
            T s2[Alg::NSTATE];
            Alg::apply(s, b, s2);
#pragma unroll
            for (int e = 0; e < Alg::NSTATE; ++e) s[e] = s2[e];
        }
    }
    // the thread that owns the last warp total holds the final state
    if (final_state != nullptr && i1 == nW && i0 < nW) Alg::expand_state(s, final_state);
    if (final_state != nullptr && nW == 0 && tid == 0) Alg::expand_state(s, final_state);
}

template <typename Alg>
__global__ void __launch_bounds__(kMidThreads)
scan_mid_kernel(typename Alg::Params p, const typename Alg::scalar* __restrict__ wagg, long nW,
                typename Alg::scalar* __restrict__ wstate, typename Alg::scalar* __restrict__ final_state) {
    __shared__ typename Alg::scalar sh[32 * Alg::NAGG];
    scan_mid_body<Alg>(p, wagg, nW, wstate, final_state, sh, (int)threadIdx.x, (int)blockDim.x, 0);
}

// Time sharding: like scan_mid_body, but the state entering the shard is not known yet (it depends on the other
// shards' summaries), so what is produced is the exclusive prefix AGGREGATE of every CTA, wprefix[e*nW + w] =
// wagg[0] o ... o wagg[w-1] (identity for w = 0), plus the shard summary total_out = wagg[0] o ... o wagg[nW-1].
// The apply kernel then starts from  state_in o wprefix[w]  (stream_apply_kernel, wprefix != nullptr).
template <typename Alg>
__device__ __forceinline__ void scan_prefix_body(const typename Alg::scalar* wagg, long nW, typename Alg::scalar* wprefix,
                                                 typename Alg::scalar* total_out, typename Alg::scalar* sh, int tid,
                                                 int nthreads, int bar_id) {
    using T = typename Alg::scalar;
    const int lane = tid & 31, wid = tid >> 5;
    const long per = (nW + nthreads - 1) / nthreads;
    long i0 = (long)tid * per, i1 = i0 + per;
    if (i0 > nW) i0 = nW;
    if (i1 > nW) i1 = nW;
    T a[Alg::NAGG];
    Alg::identity(a);
    for (long i = i0; i < i1; ++i) {
        T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) b[e] = __ldcg(wagg + (long)e * nW + i);
        if (i == i0) {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = b[e];
        } else {
            Alg::combine(a, b, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    const int nwarps = nthreads >> 5;
    int logw = 0;
    while ((1 << logw) < nwarps) ++logw;
    T ex[Alg::NAGG];  // lane-exclusive within the warp
#pragma unroll 1
    for (int lvl = 0;; ++lvl) {
        if (lvl == 5) {
            if (lane == 31) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) sh[e * 32 + wid] = a[e];
            }
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) ex[e] = shfl_up_t(a[e], 1);
            group_sync(bar_id, nthreads);
            if (wid != 0) break;
            if (lane < nwarps) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) a[e] = sh[e * 32 + lane];
            } else {
                Alg::identity(a);
            }
        }
        if (lvl == 5 + logw) break;
        const int off = 1 << (lvl < 5 ? lvl : lvl - 5);
        T o[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(a[e], off);
        if (lane >= off && (lvl < 5 || lane < nwarps)) {
            T r[Alg::NAGG];
            Alg::combine(o, a, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    if (wid == 0) {
        // exclusive over warps (the value of lane 0 is never used)
        T wx[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) wx[e] = shfl_up_t(a[e], 1);
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) sh[e * 32 + lane] = wx[e];
    }
    group_sync(bar_id, nthreads);
    // g = (prefix of the earlier warps) o (prefix of the earlier lanes), then item by item
    T g[Alg::NAGG];
    Alg::identity(g);
    bool have = false;
    const long nsteps = 2 + (i1 - i0);
#pragma unroll 1
    for (long j = 0; j < nsteps; ++j) {
        T b[Alg::NAGG];
        bool act;
        if (j == 0) {
            act = wid > 0;
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) b[e] = sh[e * 32 + wid];
        } else if (j == 1) {
            act = lane > 0;
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) b[e] = ex[e];
        } else {
            const long i = i0 + (j - 2);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) wprefix[(long)e * nW + i] = g[e];
            act = (i + 1 < i1) || (i1 == nW);  // the aggregate after the last item of all is the shard summary
            if (act) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) b[e] = __ldcg(wagg + (long)e * nW + i);
            }
        }
        if (act) {
            if (!have) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) g[e] = b[e];
                have = true;
            } else {
                T r[Alg::NAGG];
                Alg::combine(g, b, r);
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) g[e] = r[e];
            }
        }
    }
    if (total_out != nullptr && i1 == nW && i0 < nW) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) total_out[e] = g[e];
    }
}

// out[NAGG] = wagg[0] o wagg[1] o ... o wagg[nW-1]  (shard summary for time sharding), by a group of `nthreads`
// threads of one CTA (see scan_mid_body for tid / nthreads / bar_id).  sh: (nthreads / 32) * NAGG scalars.
template <typename Alg>
__device__ __forceinline__ void scan_total_body(const typename Alg::scalar* wagg, long nW, typename Alg::scalar* out,
                                                typename Alg::scalar* sh, int tid, int nthreads, int bar_id) {
    using T = typename Alg::scalar;
    const int lane = tid & 31, wid = tid >> 5;
    const long per = (nW + nthreads - 1) / nthreads;
    long i0 = (long)tid * per, i1 = i0 + per;
    if (i0 > nW) i0 = nW;
    if (i1 > nW) i1 = nW;
    T a[Alg::NAGG];
    Alg::identity(a);
    for (long i = i0; i < i1; ++i) {
        T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) b[e] = __ldcg(wagg + (long)e * nW + i);
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
    group_sync(bar_id, nthreads);
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

template <typename Alg>
__global__ void __launch_bounds__(kMidThreads)
scan_total_kernel(const typename Alg::scalar* __restrict__ wagg, long nW, typename Alg::scalar* __restrict__ out) {
    __shared__ typename Alg::scalar sh[(kMidThreads / 32) * Alg::NAGG];
    scan_total_body<Alg>(wagg, nW, out, sh, (int)threadIdx.x, (int)blockDim.x, 0);
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
