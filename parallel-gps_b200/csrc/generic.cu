// Generic state dimension (d > 4): CTA-cooperative chunked scan with a recursive fan-out hierarchy.
//
//   reduce : one CTA per chunk of L time steps appends the steps to the chunk aggregate (level 0)
//   up     : one CTA per group of Gf aggregates combines them sequentially (level l -> l+1), repeated
//            until few aggregates are left
//   top    : one CTA walks the remaining aggregates from the initial state (or reduces them to the
//            shard summary for time sharding)
//   down   : one CTA per group turns the group-entry state into the entry state of each member
//   apply  : one CTA per chunk re-runs the seeded recursion over its steps and writes the outputs
//
// The generic associative operator (with its d x d solve) is evaluated ~N/L * (1 + 1/Gf + ...) times only.
#include "generic_algebras.cuh"
#include "scan_run.cuh"
#include "mid_host.h"

namespace pssgp {

constexpr int kTopMax = 16;
constexpr int kMaxLevels = 12;

template <typename T> __device__ __forceinline__ GScratch<T> make_scratch(T* tail) {
    GScratch<T> g;
    g.red = tail;
    g.piv = (int*)(tail + 32);
    return g;
}
constexpr int kScratch = 34;  // 32 reduction slots + pivot index (rounded)

template <class G>
__global__ void g_reduce_kernel(typename G::Params p, long n, int L, typename G::scalar* __restrict__ aggs) {
    using T = typename G::scalar;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* agg = (T*)smem_raw;
    const int d = p.d, NA = G::nagg(d);
    T* work = agg + NA;
    GScratch<T> sc = make_scratch<T>(work + G::nwork(d));
    Coop c{(int)threadIdx.x, (int)blockDim.x};
    const long chunk = blockIdx.x;
    long k0 = chunk * (long)L, k1 = k0 + L;
    if (k1 > n) k1 = n;
    G::identity(c, d, agg);
    for (long k = k0; k < k1; ++k) G::append(c, p, agg, k, work, sc);
    co_copy(c, NA, agg, aggs + chunk * NA);
}

template <class G>
__global__ void g_up_kernel(int d, const typename G::scalar* __restrict__ in, long nin, int Gf,
                            typename G::scalar* __restrict__ out) {
    using T = typename G::scalar;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NA = G::nagg(d);
    T* a = (T*)smem_raw;
    T* b = a + NA;
    T* o = b + NA;
    T* work = o + NA;
    GScratch<T> sc = make_scratch<T>(work + G::nwork(d));
    Coop c{(int)threadIdx.x, (int)blockDim.x};
    const long g = blockIdx.x;
    long i0 = g * (long)Gf, i1 = i0 + Gf;
    if (i1 > nin) i1 = nin;
    co_copy(c, NA, in + i0 * NA, a);
    c.sync();
    for (long i = i0 + 1; i < i1; ++i) {
        co_copy(c, NA, in + i * NA, b);
        c.sync();
        G::combine(c, d, a, b, o, work, sc);
        T* t = a;
        a = o;
        o = t;
    }
    co_copy(c, NA, a, out + g * NA);
}

// Single CTA.  summary != nullptr: reduce the n aggregates to one (shard summary).  Otherwise walk them
// from the initial state: states[i] = state entering aggregate i; final_state = state after all.
template <class G>
__global__ void g_top_kernel(typename G::Params p, const typename G::scalar* __restrict__ aggs, long n,
                             typename G::scalar* __restrict__ states, typename G::scalar* __restrict__ final_state,
                             typename G::scalar* __restrict__ summary) {
    using T = typename G::scalar;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = p.d, NA = G::nagg(d), NS = G::nstate(d);
    Coop c{(int)threadIdx.x, (int)blockDim.x};
    if (summary != nullptr) {
        T* a = (T*)smem_raw;
        T* b = a + NA;
        T* o = b + NA;
        T* work = o + NA;
        GScratch<T> sc = make_scratch<T>(work + G::nwork(d));
        co_copy(c, NA, aggs, a);
        c.sync();
        for (long i = 1; i < n; ++i) {
            co_copy(c, NA, aggs + i * NA, b);
            c.sync();
            G::combine(c, d, a, b, o, work, sc);
            T* t = a;
            a = o;
            o = t;
        }
        co_copy(c, NA, a, summary);
        return;
    }
    T* a = (T*)smem_raw;
    T* s = a + NA;
    T* s2 = s + NS;
    T* work = s2 + NS;
    GScratch<T> sc = make_scratch<T>(work + G::nwork(d));
    G::load_init(c, p, s);
    for (long i = 0; i < n; ++i) {
        co_copy(c, NS, s, states + i * NS);
        co_copy(c, NA, aggs + i * NA, a);
        c.sync();
        G::apply(c, d, s, a, s2, work, sc);
        T* t = s;
        s = s2;
        s2 = t;
    }
    if (final_state != nullptr) G::expand_state(c, d, s, final_state);
}

template <class G>
__global__ void g_down_kernel(int d, const typename G::scalar* __restrict__ aggs, long n, int Gf,
                              const typename G::scalar* __restrict__ gstates, typename G::scalar* __restrict__ states) {
    using T = typename G::scalar;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NA = G::nagg(d), NS = G::nstate(d);
    T* a = (T*)smem_raw;
    T* s = a + NA;
    T* s2 = s + NS;
    T* work = s2 + NS;
    GScratch<T> sc = make_scratch<T>(work + G::nwork(d));
    Coop c{(int)threadIdx.x, (int)blockDim.x};
    const long g = blockIdx.x;
    long i0 = g * (long)Gf, i1 = i0 + Gf;
    if (i1 > n) i1 = n;
    co_copy(c, NS, gstates + g * NS, s);
    c.sync();
    for (long i = i0; i < i1; ++i) {
        co_copy(c, NS, s, states + i * NS);
        if (i + 1 < i1) {
            co_copy(c, NA, aggs + i * NA, a);
            c.sync();
            G::apply(c, d, s, a, s2, work, sc);
            T* t = s;
            s = s2;
            s2 = t;
        }
    }
}

template <class G>
__global__ void g_apply_kernel(typename G::Params p, long n, int L, const typename G::scalar* __restrict__ states,
                               int nacc, typename G::scalar* __restrict__ part) {
    using T = typename G::scalar;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = p.d, NS = G::nstate(d);
    T* s = (T*)smem_raw;
    T* work = s + NS;
    GScratch<T> sc = make_scratch<T>(work + G::nwork(d));
    Coop c{(int)threadIdx.x, (int)blockDim.x};
    const long chunk = blockIdx.x;
    long k0 = chunk * (long)L, k1 = k0 + L;
    if (k1 > n) k1 = n;
    co_copy(c, NS, states + chunk * NS, s);
    c.sync();
    T acc[66];
    for (int e = 0; e < nacc; ++e) acc[e] = T(0);
    for (long k = k0; k < k1; ++k) G::step(c, p, s, k, work, sc, acc);
    if (nacc > 0 && c.tid == 0)
        for (int e = 0; e < nacc; ++e) part[chunk * nacc + e] = acc[e];
}

template <class G>
__global__ void g_finish_kernel(typename G::Params p, const typename G::scalar* __restrict__ part, long nparts,
                                int nacc, typename G::scalar* __restrict__ acc_out) {
    using T = typename G::scalar;
    __shared__ T red[32];
    Coop c{(int)threadIdx.x, (int)blockDim.x};
    for (int e = 0; e < nacc; ++e) {
        T v = T(0);
        for (long b = c.tid; b < nparts; b += c.nt) v += part[b * nacc + e];
        const T tot = co_reduce_sum(c, v, red);
        if (c.tid == 0) G::finish(p, e, tot, acc_out);
    }
}

template <class K> int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return PSSGP_OK;
}

int pick_chunk_generic(const pssgp_handle* h, int64_t n, int d) {
    if (h->chunk_opt > 0) return (int)h->chunk_opt;
    // aim at >= 8 chunks per SM; longer chunks amortise the aggregate traffic and the operator hierarchy
    int64_t target = (int64_t)h->num_sms * 8;
    int L = 16;
    while (L < 256 && n / (2 * L) >= target) L *= 2;
    (void)d;
    return L;
}

// Level sizes of the fan-out hierarchy over cnt0 level-0 aggregates.
struct HierLevels {
    int64_t cnt[kMaxLevels];
    size_t off[kMaxLevels];
    int nl;
    size_t tot;
};
static HierLevels hier_levels(int64_t cnt0, int Gf) {
    HierLevels hl;
    hl.nl = 0;
    hl.cnt[hl.nl++] = cnt0;
    while (hl.cnt[hl.nl - 1] > kTopMax && hl.nl < kMaxLevels) {
        hl.cnt[hl.nl] = (hl.cnt[hl.nl - 1] + Gf - 1) / Gf;
        ++hl.nl;
    }
    hl.tot = 0;
    for (int l = 0; l < hl.nl; ++l) {
        hl.off[l] = hl.tot;
        hl.tot += (size_t)hl.cnt[l];
    }
    return hl;
}
constexpr int kGf = 16;
size_t hier_total(int64_t cnt0) { return hier_levels(cnt0, kGf).tot; }

template <class G> int hier_configure(int d) {
    using T = typename G::scalar;
    const int NA = G::nagg(d), NS = G::nstate(d), NWK = G::nwork(d) + kScratch + 8;
    const size_t sm_up = sizeof(T) * (size_t)(3 * NA + NWK);
    const size_t sm_top = sm_up > sizeof(T) * (size_t)(NA + 2 * NS + NWK) ? sm_up : sizeof(T) * (size_t)(NA + 2 * NS + NWK);
    const size_t sm_down = sizeof(T) * (size_t)(NA + 2 * NS + NWK);
    if (sm_top > 227 * 1024)
        return set_err(PSSGP_ERR_UNSUPPORTED, "state dimension %d needs %zu B of shared memory per CTA (max 232448)", d, sm_top);
    int rc;
    if ((rc = set_smem(g_up_kernel<G>, sm_up))) return rc;
    if ((rc = set_smem(g_top_kernel<G>, sm_top))) return rc;
    if ((rc = set_smem(g_down_kernel<G>, sm_down))) return rc;
    return PSSGP_OK;
}

// Hierarchy over the level-0 aggregates aggs[0 .. cnt0) (the buffers have room for hier_total(cnt0) entries):
// up-sweep (unless `have_up`: the upper levels are still valid), then either the shard summary (summary != nullptr)
// or the top walk + down-sweep that leaves the state entering every level-0 aggregate in states[0 .. cnt0).
template <class G>
int run_hierarchy(pssgp_handle* h, const typename G::Params& p, int d, int64_t cnt0, typename G::scalar* aggs,
                  typename G::scalar* states, typename G::scalar* final_state, typename G::scalar* summary, bool have_up,
                  cudaStream_t st, int* launches) {
    using T = typename G::scalar;
    const int NA = G::nagg(d), NS = G::nstate(d), NWK = G::nwork(d) + kScratch + 8;
    const int nt = d <= 8 ? 64 : (d <= 16 ? 128 : 256);
    const int Gf = kGf;
    const size_t sm_up = sizeof(T) * (size_t)(3 * NA + NWK);
    const size_t sm_top = sm_up > sizeof(T) * (size_t)(NA + 2 * NS + NWK) ? sm_up : sizeof(T) * (size_t)(NA + 2 * NS + NWK);
    const size_t sm_down = sizeof(T) * (size_t)(NA + 2 * NS + NWK);
    int rc;
    if ((rc = hier_configure<G>(d))) return rc;
    const HierLevels hl = hier_levels(cnt0, Gf);
    const int nl = hl.nl;
    if (!have_up) {
        for (int l = 0; l + 1 < nl; ++l) {
            PSSGP_LAUNCH(h, G::name(1), st,
                         (g_up_kernel<G><<<(unsigned)hl.cnt[l + 1], nt, sm_up, st>>>(d, aggs + hl.off[l] * NA, hl.cnt[l], Gf,
                                                                                      aggs + hl.off[l + 1] * NA)));
            ++*launches;
        }
    }
    if (summary != nullptr) {
        PSSGP_LAUNCH(h, G::name(2), st,
                     (g_top_kernel<G><<<1, nt, sm_top, st>>>(p, aggs + hl.off[nl - 1] * NA, hl.cnt[nl - 1], nullptr, nullptr,
                                                            summary)));
        ++*launches;
        return PSSGP_OK;
    }
    PSSGP_LAUNCH(h, G::name(2), st,
                 (g_top_kernel<G><<<1, nt, sm_top, st>>>(p, aggs + hl.off[nl - 1] * NA, hl.cnt[nl - 1],
                                                        states + hl.off[nl - 1] * NS, final_state, nullptr)));
    ++*launches;
    for (int l = nl - 2; l >= 0; --l) {
        PSSGP_LAUNCH(h, G::name(3), st,
                     (g_down_kernel<G><<<(unsigned)hl.cnt[l + 1], nt, sm_down, st>>>(d, aggs + hl.off[l] * NA, hl.cnt[l], Gf,
                                                                                     states + hl.off[l + 1] * NS,
                                                                                     states + hl.off[l] * NS)));
        ++*launches;
    }
    return PSSGP_OK;
}

template <class G>
int run_generic(pssgp_handle* h, typename G::Params p, int64_t n, int d, int nacc, typename G::scalar* acc_out,
                typename G::scalar* final_state, typename G::scalar* summary, uint64_t key, cudaStream_t st) {
    using T = typename G::scalar;
    const int NA = G::nagg(d), NS = G::nstate(d), NWK = G::nwork(d) + kScratch + 8;
    const int nt = d <= 8 ? 64 : (d <= 16 ? 128 : 256);
    const int L = pick_chunk_generic(h, n, d);
    const size_t sm_reduce = sizeof(T) * (size_t)(NA + NWK);
    const size_t sm_apply = sizeof(T) * (size_t)(NS + NWK);
    int rc;
    if ((rc = hier_configure<G>(d))) return rc;
    if ((rc = set_smem(g_reduce_kernel<G>, sm_reduce))) return rc;
    if ((rc = set_smem(g_apply_kernel<G>, sm_apply))) return rc;
    const int64_t cnt0 = (n + L - 1) / L;
    const size_t tot = hier_total(cnt0);
    constexpr int kind = G::KIND;
    const bool reuse = (summary == nullptr && key != 0 && h->pending_key[kind] == key && h->pending_n[kind] == n &&
                        h->pending_L[kind] == L);
    pending_clear(h, kind);
    if (!reuse) {
        if ((rc = ws_reserve(h, WS_LANE + kind, sizeof(T) * tot * NA))) return rc;
    }
    if ((rc = ws_reserve(h, WS_WAGG + kind, sizeof(T) * tot * NS))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * (size_t)cnt0 * (nacc > 0 ? nacc : 1)))) return rc;
    T* aggs = (T*)h->buf[WS_LANE + kind];
    T* states = (T*)h->buf[WS_WAGG + kind];
    T* part = (T*)h->buf[WS_PART];
    int launches = 0;
    if (!reuse) {
        PSSGP_LAUNCH(h, G::name(0), st, (g_reduce_kernel<G><<<(unsigned)cnt0, nt, sm_reduce, st>>>(p, n, L, aggs)));
        ++launches;
    }
    if ((rc = run_hierarchy<G>(h, p, d, cnt0, aggs, states, final_state, summary, reuse, st, &launches))) return rc;
    if (summary != nullptr) {
        h->pending_key[kind] = key;
        h->pending_n[kind] = n;
        h->pending_L[kind] = L;
        return check_launch(h, "generic summary", launches);
    }
    PSSGP_LAUNCH(h, G::name(4), st,
                 (g_apply_kernel<G><<<(unsigned)cnt0, nt, sm_apply, st>>>(p, n, L, states, nacc, part)));
    ++launches;
    if (nacc > 0) {
        PSSGP_LAUNCH(h, "g_finish", st, (g_finish_kernel<G><<<1, 256, 0, st>>>(p, part, cnt0, nacc, acc_out)));
        ++launches;
    }
    return check_launch(h, "generic scan", launches);
}

// One-CTA fold of shard summaries (time sharding): s = init; for i: s = s o summaries[i*stride]; out = expand(s)
template <class G>
__global__ void g_fold_kernel(typename G::Params p, const typename G::scalar* __restrict__ summaries, int count,
                              long stride, typename G::scalar* __restrict__ out) {
    using T = typename G::scalar;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = p.d, NA = G::nagg(d), NS = G::nstate(d);
    T* a = (T*)smem_raw;
    T* s = a + NA;
    T* s2 = s + NS;
    T* work = s2 + NS;
    GScratch<T> sc = make_scratch<T>(work + G::nwork(d));
    Coop c{(int)threadIdx.x, (int)blockDim.x};
    G::load_init(c, p, s);
    for (int i = 0; i < count; ++i) {
        co_copy(c, NA, summaries + (long)i * stride, a);
        c.sync();
        G::apply(c, d, s, a, s2, work, sc);
        T* t = s;
        s = s2;
        s2 = t;
    }
    G::expand_state(c, d, s, out);
}

template <class G>
int run_fold_generic(pssgp_handle* h, typename G::Params p, int d, const typename G::scalar* summaries, int count,
                     long stride, typename G::scalar* out, cudaStream_t st) {
    using T = typename G::scalar;
    const size_t sm = sizeof(T) * (size_t)(G::nagg(d) + 2 * G::nstate(d) + G::nwork(d) + kScratch + 8);
    int rc;
    if ((rc = set_smem(g_fold_kernel<G>, sm))) return rc;
    const int nt = d <= 8 ? 64 : (d <= 16 ? 128 : 256);
    PSSGP_LAUNCH(h, "g_fold", st, (g_fold_kernel<G><<<1, nt, sm, st>>>(p, summaries, count, stride, out)));
    return check_launch(h, "generic fold", 1);
}

#define GEN_DISPATCH(EXPR_F64, EXPR_F32)          \
    do {                                          \
        if (dtype == PSSGP_F64) {                 \
            using T = double;                     \
            EXPR_F64;                             \
        } else {                                  \
            using T = float;                      \
            EXPR_F32;                             \
        }                                         \
    } while (0)

template <typename T>
int pkf_generic_t(pssgp_handle* h, int64_t n, int d, const void* P0, const void* Fs, const void* Qs, const void* H,
                  const void* R, const void* y, const void* m0, int first_special, void* fms, void* fPs, void* ll,
                  void* final_state, void* summary, cudaStream_t st) {
    typename GFilter<T>::Params p;
    p.Fs = (const T*)Fs; p.Qs = (const T*)Qs; p.y = (const T*)y; p.H = (const T*)H; p.R = (const T*)R;
    p.P0 = (const T*)P0; p.m0 = (const T*)m0; p.fms = (T*)fms; p.fPs = (T*)fPs;
    p.n = n; p.d = d; p.first_special = first_special;
    return run_generic<GFilter<T>>(h, p, n, d, summary ? 0 : 1, (T*)ll, (T*)final_state, (T*)summary,
                                   filter_sig(sizeof(T), d, n, Fs, Qs, y, H, R, first_special), st);
}

int pkf_generic(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                const void* H, const void* R, const void* y, const void* m0, int first_special, void* fms, void* fPs,
                void* ll, void* final_state, void* summary, cudaStream_t st) {
    if (dtype == PSSGP_F64 && mid::supported(d) && !h->force_generic)
        return mid::pkf_dispatch(d, h, n, (const double*)P0, (const double*)Fs, (const double*)Qs, (const double*)H,
                                 (const double*)R, (const double*)y, (const double*)m0, first_special, (double*)fms,
                                 (double*)fPs, (double*)ll, (double*)final_state, (double*)summary, st);
    if (dtype == PSSGP_F32 && mid::supported(d) && !h->force_generic && m0 == nullptr && first_special &&
        final_state == nullptr && summary == nullptr)
        return mid::f32_pkfs_grad(h, n, d, (const float*)P0, (const float*)Fs, (const float*)Qs, (const float*)H,
                                  (const float*)R, (const float*)y, nullptr, (float*)fms, (float*)fPs, (float*)ll, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, true, st);
    if (dtype == PSSGP_F64)
        return pkf_generic_t<double>(h, n, d, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, final_state, summary, st);
    return pkf_generic_t<float>(h, n, d, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, final_state, summary, st);
}

template <typename T>
int filter_fold_t(pssgp_handle* h, int d, int count, const void* P0, const void* m0, const void* summaries,
                  void* state_out, cudaStream_t st) {
    typename GFilter<T>::Params p = {};
    p.P0 = (const T*)P0; p.m0 = (const T*)m0; p.d = d;
    return run_fold_generic<GFilter<T>>(h, p, d, (const T*)summaries, count, GFilter<T>::nagg(d), (T*)state_out, st);
}

int filter_fold_generic(pssgp_handle* h, int dtype, int d, int count, const void* P0, const void* m0,
                        const void* summaries, void* state_out, cudaStream_t st) {
    if (dtype == PSSGP_F64) return filter_fold_t<double>(h, d, count, P0, m0, summaries, state_out, st);
    return filter_fold_t<float>(h, d, count, P0, m0, summaries, state_out, st);
}

template <typename T>
int pks_generic_t(pssgp_handle* h, int64_t n, int d, const void* Fs, const void* Qs, const void* fms, const void* fPs,
                  int last_special, const void* Fnext, const void* Qnext, const void* init, void* sms, void* sPs,
                  void* first_state, void* summary, cudaStream_t st) {
    typename GSmoother<T>::Params p;
    p.Fs = (const T*)Fs; p.Qs = (const T*)Qs; p.fms = (const T*)fms; p.fPs = (const T*)fPs;
    p.sms = (T*)sms; p.sPs = (T*)sPs; p.n = n; p.d = d; p.last_special = last_special;
    p.Fnext = (const T*)Fnext; p.Qnext = (const T*)Qnext; p.init = (const T*)init;
    return run_generic<GSmoother<T>>(h, p, n, d, 0, nullptr, (T*)first_state, (T*)summary,
                                     smoother_sig(sizeof(T), d, n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext), st);
}

int pks_generic(pssgp_handle* h, int dtype, int64_t n, int d, const void* Fs, const void* Qs, const void* fms,
                const void* fPs, int last_special, const void* Fnext, const void* Qnext, const void* init, void* sms,
                void* sPs, void* first_state, void* summary, cudaStream_t st) {
    if (dtype == PSSGP_F64)
        return pks_generic_t<double>(h, n, d, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, init, sms, sPs, first_state, summary, st);
    return pks_generic_t<float>(h, n, d, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, init, sms, sPs, first_state, summary, st);
}

template <typename T>
int smoother_fold_t(pssgp_handle* h, int d, int count, const void* summaries, void* state_out, cudaStream_t st) {
    typename GSmoother<T>::Params p = {};
    p.d = d;
    const long NA = GSmoother<T>::nagg(d);
    return run_fold_generic<GSmoother<T>>(h, p, d, (const T*)summaries + (long)(count - 1) * NA, count, -NA,
                                          (T*)state_out, st);
}

int smoother_fold_generic(pssgp_handle* h, int dtype, int d, int count, const void* summaries, void* state_out,
                          cudaStream_t st) {
    if (dtype == PSSGP_F64) return smoother_fold_t<double>(h, d, count, summaries, state_out, st);
    return smoother_fold_t<float>(h, d, count, summaries, state_out, st);
}

template <typename T>
int pkf_bwd_generic_t(pssgp_handle* h, int64_t n, int d, const void* P0, const void* m0, const void* Fs, const void* Qs,
                      const void* H, const void* R, const void* y, const void* fms, const void* fPs, const void* g_ll,
                      int first_special, const void* adj_init, void* dP0, void* dFs, void* dQs, void* dH, void* dR,
                      void* adj_first, void* summary, cudaStream_t st) {
    typename GAdjoint<T>::Params p;
    p.Fs = (const T*)Fs; p.Qs = (const T*)Qs; p.y = (const T*)y; p.H = (const T*)H; p.R = (const T*)R;
    p.P0 = (const T*)P0; p.m0 = (const T*)m0; p.fms = (const T*)fms; p.fPs = (const T*)fPs; p.g = (const T*)g_ll;
    p.init = (const T*)adj_init; p.dFs = (T*)dFs; p.dQs = (T*)dQs; p.dP0 = (T*)dP0; p.dH = (T*)dH; p.dR = (T*)dR;
    p.n = n; p.d = d; p.first_special = first_special;
    return run_generic<GAdjoint<T>>(h, p, n, d, summary ? 0 : 1 + d, (T*)dR, (T*)adj_first, (T*)summary,
                                    adjoint_sig(sizeof(T), d, n, Fs, Qs, y, H, R, fms, fPs, first_special), st);
}

int pkf_bwd_generic(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* m0, const void* Fs,
                    const void* Qs, const void* H, const void* R, const void* y, const void* fms, const void* fPs,
                    const void* g_ll, int first_special, const void* adj_init, void* dP0, void* dFs, void* dQs,
                    void* dH, void* dR, void* adj_first, void* summary, cudaStream_t st) {
    if (d > 64) return set_err(PSSGP_ERR_UNSUPPORTED, "pkf_backward: state dimension %d > 64", d);
    if (dtype == PSSGP_F64 && mid::supported(d) && !h->force_generic && summary == nullptr && adj_init == nullptr &&
        adj_first == nullptr)
        return mid::pkf_backward_dispatch(d, h, n, (const double*)P0, (const double*)m0, (const double*)Fs,
                                          (const double*)Qs, (const double*)H, (const double*)R, (const double*)y,
                                          (const double*)fms, (const double*)fPs, (const double*)g_ll, first_special,
                                          (double*)dP0, (double*)dFs, (double*)dQs, (double*)dH, (double*)dR, st);
    if (dtype == PSSGP_F64)
        return pkf_bwd_generic_t<double>(h, n, d, P0, m0, Fs, Qs, H, R, y, fms, fPs, g_ll, first_special, adj_init, dP0,
                                         dFs, dQs, dH, dR, adj_first, summary, st);
    return pkf_bwd_generic_t<float>(h, n, d, P0, m0, Fs, Qs, H, R, y, fms, fPs, g_ll, first_special, adj_init, dP0, dFs,
                                    dQs, dH, dR, adj_first, summary, st);
}

template <typename T>
int adjoint_fold_t(pssgp_handle* h, int d, int count, const void* summaries, void* state_out, cudaStream_t st) {
    typename GAdjoint<T>::Params p = {};
    p.d = d;
    const long NA = GAdjoint<T>::nagg(d);
    return run_fold_generic<GAdjoint<T>>(h, p, d, (const T*)summaries + (long)(count - 1) * NA, count, -NA,
                                         (T*)state_out, st);
}

int adjoint_fold_generic(pssgp_handle* h, int dtype, int d, int count, const void* summaries, void* state_out,
                         cudaStream_t st) {
    if (dtype == PSSGP_F64) return adjoint_fold_t<double>(h, d, count, summaries, state_out, st);
    return adjoint_fold_t<float>(h, d, count, summaries, state_out, st);
}

// Hierarchy / finish entry points used by the warp-level d > 4 kernels (mid_inst.cu, one translation unit per d).
int hier_filter_f64(pssgp_handle* h, const GFilter<double>::Params& p, int d, int64_t cnt0, double* aggs, double* states,
                    double* final_state, double* summary, bool have_up, cudaStream_t st, int* launches) {
    return run_hierarchy<GFilter<double>>(h, p, d, cnt0, aggs, states, final_state, summary, have_up, st, launches);
}
int hier_rev_f64(pssgp_handle* h, const GRev<double>::Params& p, int d, int64_t cnt0, double* aggs, double* states,
                 double* final_state, double* summary, bool have_up, cudaStream_t st, int* launches) {
    return run_hierarchy<GRev<double>>(h, p, d, cnt0, aggs, states, final_state, summary, have_up, st, launches);
}
int finish_filter_f64(pssgp_handle* h, const GFilter<double>::Params& p, const double* part, int64_t nparts, double* ll,
                      cudaStream_t st) {
    PSSGP_LAUNCH(h, "g_finish", st, (g_finish_kernel<GFilter<double>><<<1, 256, 0, st>>>(p, part, nparts, 1, ll)));
    return PSSGP_OK;
}
int finish_rev_f64(pssgp_handle* h, const GRev<double>::Params& p, const double* part, int64_t nparts, cudaStream_t st) {
    PSSGP_LAUNCH(h, "g_finish", st, (g_finish_kernel<GRev<double>><<<1, 256, 0, st>>>(p, part, nparts, 1 + p.d, p.dR)));
    return PSSGP_OK;
}

}  // namespace pssgp
