// Explicit instantiation of the warp-level d > 4 path for ONE state dimension (compile with -DMID_D=<D>).
#include "mid.cuh"
#include "mid_frag.cuh"
#include "mid_hier.cuh"
#include "mid_host.h"

#ifndef MID_D
#error "compile with -DMID_D=<state dimension>"
#endif

namespace pssgp {
namespace mid {

template <class Kern> static int set_smem_attr(Kern kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
    return PSSGP_OK;
}


// Kernel family per state dimension: fragment-resident (registers, one warp per chunk) up to D = 16, shared-memory
// tiles with groups of warps above.  Option "mid_smem" forces the shared-memory family (tests / tuning).
template <int D> constexpr bool has_frag() { return D <= 24; }

static thread_local int g_mid_warps = 0;  // option "mid_warps" of the handle in use (0 = compile-time default)
// measured (scripts/sweep_mid_warps.py, N = 1e6): d = 6 gains up to the 24 warps that fit; d = 9 / d = 16 are fastest with 8
static int cap_warps(int wpc) {
    const int lim = g_mid_warps > 0 ? g_mid_warps : (MID_D > 9 ? 8 : 1000);
    return lim < wpc ? lim : wpc;
}

template <int D> static int k1_groups(bool fr) {
    if constexpr (has_frag<D>()) if (fr) return cap_warps(frag::FK1<D>::WPC);
    return K1<D, default_wg<D>()>::GPC;
}
template <int D, bool REV, bool STORED> static int k2_groups(bool fr) {
    if constexpr (has_frag<D>()) if (fr) return cap_warps(frag::FK2<D, REV, STORED>::WPC);
    return K2<D, default_wg<D>(), REV, STORED>::GPC;
}
template <int D, bool SMOOTH, bool ADJ> static int k3_groups(bool fr) {
    if constexpr (has_frag<D>()) if (fr) return cap_warps(frag::FK3<D, SMOOTH, ADJ>::WPC);
    return K3<D, default_wg<D>(), SMOOTH, ADJ>::GPC;
}

// gpc: groups (chunks) per CTA shared by the kernels of one call — the chunk count is sized for ONE resident wave of
// num_sms * gpc chunks, so every kernel launches num_sms CTAs (a kernel that could pack more groups per CTA would
// otherwise leave SMs idle)
template <int D> static int launch_k1(pssgp_handle* h, bool fr, const Params& p, int L, int64_t nchunks, double* aggs, cudaStream_t st, int gpc) {
    int rc;
    if constexpr (has_frag<D>()) {
        if (fr) {
            using KA = frag::FK1<D>;
            if ((rc = set_smem_attr(frag::fk1_filter_reduce<D>, KA::SMEM))) return rc;
            const int wpc = gpc < cap_warps(KA::WPC) ? gpc : cap_warps(KA::WPC);
            const unsigned grid = (unsigned)((nchunks + wpc - 1) / wpc);
            PSSGP_LAUNCH(h, "mid_filter_reduce", st, (frag::fk1_filter_reduce<D><<<grid, wpc * 32, wpc * KA::WARP_SMEM, st>>>(p, L, nchunks, aggs)));
            return PSSGP_OK;
        }
    }
    constexpr int WG = default_wg<D>();
    using KA = K1<D, WG>;
    if ((rc = set_smem_attr(k1_filter_reduce<D, WG>, (size_t)KA::GPC * KA::GROUP_DOUBLES * 8))) return rc;
    const int gp = gpc < KA::GPC ? gpc : KA::GPC;
    const unsigned grid = (unsigned)((nchunks + gp - 1) / gp);
    PSSGP_LAUNCH(h, "mid_filter_reduce", st,
                 (k1_filter_reduce<D, WG><<<grid, gp * WG * 32, (size_t)gp * KA::GROUP_DOUBLES * 8, st>>>(p, L, nchunks, aggs)));
    return PSSGP_OK;
}

template <int D, bool REV, bool STORED>
static int launch_k2(pssgp_handle* h, bool fr, const char* name, const Params& p, int L, int64_t nchunks, const double* fstates,
                     double* part, double* raggs, cudaStream_t st, int gpc) {
    int rc;
    if constexpr (has_frag<D>()) {
        if (fr) {
            using KB = frag::FK2<D, REV, STORED>;
            if ((rc = set_smem_attr(frag::fk2_forward<D, REV, STORED>, KB::SMEM))) return rc;
            const int wpc = gpc < cap_warps(KB::WPC) ? gpc : cap_warps(KB::WPC);
            const unsigned grid = (unsigned)((nchunks + wpc - 1) / wpc);
            PSSGP_LAUNCH(h, name, st,
                         (frag::fk2_forward<D, REV, STORED><<<grid, wpc * 32, wpc * KB::WARP_SMEM, st>>>(p, L, nchunks, fstates, part, raggs)));
            return PSSGP_OK;
        }
    }
    constexpr int WG = default_wg<D>();
    using KB = K2<D, WG, REV, STORED>;
    if ((rc = set_smem_attr(k2_forward<D, WG, REV, STORED>, (size_t)KB::GPC * KB::GROUP_DOUBLES * 8))) return rc;
    const int gp = gpc < KB::GPC ? gpc : KB::GPC;
    const unsigned grid = (unsigned)((nchunks + gp - 1) / gp);
    PSSGP_LAUNCH(h, name, st,
                 (k2_forward<D, WG, REV, STORED><<<grid, gp * WG * 32, (size_t)gp * KB::GROUP_DOUBLES * 8, st>>>(
                     p, L, nchunks, fstates, part, raggs)));
    return PSSGP_OK;
}

template <int D, bool SMOOTH, bool ADJ>
static int launch_k3(pssgp_handle* h, bool fr, const Params& p, int L, int64_t nchunks, const double* rstates, double* part,
                     cudaStream_t st, int gpc) {
    int rc;
    if constexpr (has_frag<D>()) {
        if (fr) {
            using KC = frag::FK3<D, SMOOTH, ADJ>;
            if ((rc = set_smem_attr(frag::fk3_reverse<D, SMOOTH, ADJ>, KC::SMEM))) return rc;
            const int wpc = gpc < cap_warps(KC::WPC) ? gpc : cap_warps(KC::WPC);
            const unsigned grid = (unsigned)((nchunks + wpc - 1) / wpc);
            PSSGP_LAUNCH(h, "mid_reverse", st,
                         (frag::fk3_reverse<D, SMOOTH, ADJ><<<grid, wpc * 32, wpc * KC::WARP_SMEM, st>>>(p, L, nchunks, rstates, part)));
            return PSSGP_OK;
        }
    }
    constexpr int WG = default_wg<D>();
    using KC = K3<D, WG, SMOOTH, ADJ>;
    if ((rc = set_smem_attr(k3_reverse<D, WG, SMOOTH, ADJ>, (size_t)KC::GPC * KC::GROUP_DOUBLES * 8))) return rc;
    const int gp = gpc < KC::GPC ? gpc : KC::GPC;
    const unsigned grid = (unsigned)((nchunks + gp - 1) / gp);
    PSSGP_LAUNCH(h, "mid_reverse", st,
                 (k3_reverse<D, WG, SMOOTH, ADJ><<<grid, gp * WG * 32, (size_t)gp * KC::GROUP_DOUBLES * 8, st>>>(
                     p, L, nchunks, rstates, part)));
    return PSSGP_OK;
}

template <int D>
static int hier_filter(pssgp_handle* h, const Params& p, int64_t nchunks, double* aggs, double* states, double* final_state,
                       double* summary, bool have_up, cudaStream_t st, int* launches) {
    static const char* const names[3] = {"mhier_filter_up", "mhier_filter_top", "mhier_filter_down"};
    const typename hier::FilterH<D>::Init in = {p.P0, p.m0};
    return hier::run<hier::FilterH<D>>(h, in, nchunks, aggs, states, final_state, summary, have_up, names, st, launches);
}
template <int D>
static int hier_rev(pssgp_handle* h, int64_t nchunks, double* raggs, double* rstates, const double* init, double* summary,
                    bool have_up, cudaStream_t st, int* launches) {
    static const char* const names[3] = {"mhier_rev_up", "mhier_rev_top", "mhier_rev_down"};
    const typename hier::RevH<D>::Init in = {init};
    return hier::run<hier::RevH<D>>(h, in, nchunks, raggs, rstates, nullptr, summary, have_up, names, st, launches);
}

// chunk length: one resident wave of groups of the most shared-memory-hungry kernel of the call
static int pick_len(const pssgp_handle* h, int64_t n, int gpc) {
    if (h->chunk_opt > 0) return (int)h->chunk_opt;
    // whole waves of `slots` resident groups, chunks of at most kMaxLen steps: every wave is full (a last partial wave
    // would cost as much as a full one)
    constexpr int64_t kMaxLen = 4096;
    const int64_t slots = (int64_t)(h->num_sms > 0 ? h->num_sms : 1) * gpc;
    const int64_t waves = (n + slots * kMaxLen - 1) / (slots * kMaxLen);
    int64_t L = (n + slots * waves - 1) / (slots * waves);
    if (L < 16) L = 16;
    return (int)L;
}

static Params base_params(int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H, const double* R,
                          const double* y, const double* m0, int first_special) {
    Params p = {};
    p.Fs = Fs; p.Qs = Qs; p.y = y; p.H = H; p.R = R; p.P0 = P0; p.m0 = m0;
    p.n = n; p.first_special = first_special;
    return p;
}

static GFilter<double>::Params gfilter_params(const Params& p, int d) {
    GFilter<double>::Params q = {};
    q.Fs = p.Fs; q.Qs = p.Qs; q.y = p.y; q.H = p.H; q.R = p.R; q.P0 = p.P0; q.m0 = p.m0; q.fms = p.fms; q.fPs = p.fPs;
    q.n = p.n; q.d = d; q.first_special = p.first_special;
    return q;
}

static GRev<double>::Params grev_params(const Params& p, int d, double* dH, double* dR) {
    GRev<double>::Params q = {};
    q.Fs = p.Fs; q.Qs = p.Qs; q.y = p.y; q.H = p.H; q.R = p.R; q.P0 = p.P0; q.m0 = p.m0; q.fms = p.fms_in; q.fPs = p.fPs_in;
    q.g = p.g; q.init = nullptr; q.sms = p.sms; q.sPs = p.sPs; q.dFs = p.dFs; q.dQs = p.dQs; q.dP0 = p.dP0; q.dH = dH; q.dR = dR;
    q.n = p.n; q.d = d; q.first_special = p.first_special;
    return q;
}

template <int D>
int pkf(pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H, const double* R,
        const double* y, const double* m0, int first_special, double* fms, double* fPs, double* ll, double* final_state,
        double* summary, cudaStream_t st) {
    g_mid_warps = h->mid_warps;
    const bool fr = has_frag<D>() && !h->mid_smem;
    int rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, m0, first_special);
    p.fms = fms; p.fPs = fPs;
    const int g1 = k1_groups<D>(fr), g2 = k2_groups<D, false, false>(fr);
    const int gpc = g2 < g1 ? g2 : g1;
    const int L = pick_len(h, n, gpc);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier::total(nchunks);
    const int NA = 3 * D * D + 2 * D, NS = D + D * D;
    const uint64_t key = filter_sig(8, D, n, Fs, Qs, y, H, R, first_special);
    const bool reuse = (summary == nullptr && h->pending_key[KIND_FILTER] == key && h->pending_n[KIND_FILTER] == n &&
                        h->pending_L[KIND_FILTER] == L);
    pending_clear(h, KIND_FILTER);
    if (!reuse)
        if ((rc = ws_reserve(h, WS_LANE + KIND_FILTER, sizeof(double) * tot * NA))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_FILTER, sizeof(double) * tot * NS))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(double) * (size_t)nchunks))) return rc;
    double* aggs = (double*)h->buf[WS_LANE + KIND_FILTER];
    double* states = (double*)h->buf[WS_WAGG + KIND_FILTER];
    double* part = (double*)h->buf[WS_PART];
    int launches = 0;
    if (!reuse) {
        if ((rc = launch_k1<D>(h, fr, p, L, nchunks, aggs, st, gpc))) return rc;
        ++launches;
    }
    const GFilter<double>::Params gp = gfilter_params(p, D);
    if ((rc = hier_filter<D>(h, p, nchunks, aggs, states, final_state, summary, reuse, st, &launches))) return rc;
    if (summary != nullptr) {
        h->pending_key[KIND_FILTER] = key;
        h->pending_n[KIND_FILTER] = n;
        h->pending_L[KIND_FILTER] = L;
        return check_launch(h, "mid pkf summary", launches);
    }
    if ((rc = launch_k2<D, false, false>(h, fr, "mid_forward", p, L, nchunks, states, part, nullptr, st, gpc))) return rc;
    ++launches;
    if (ll != nullptr) {
        finish_filter_f64(h, gp, part, nchunks, ll, st);
        ++launches;
    }
    return check_launch(h, "mid pkf", launches);
}

// reverse part shared by pkfs_grad and pkf_backward: hierarchy over the reverse aggregates + K3 + finish
template <int D, bool SMOOTH, bool ADJ>
static int run_reverse(pssgp_handle* h, bool fr, const Params& p, int L, int64_t nchunks, double* raggs, double* rstates,
                       double* part, double* dH, double* dR, cudaStream_t st, int* launches, int gpc,
                       const double* rev_init = nullptr, bool have_up = false) {
    int rc;
    const GRev<double>::Params rp = grev_params(p, D, dH, dR);
    if ((rc = hier_rev<D>(h, nchunks, raggs, rstates, rev_init, nullptr, have_up, st, launches))) return rc;
    if ((rc = launch_k3<D, SMOOTH, ADJ>(h, fr, p, L, nchunks, rstates, part, st, gpc))) return rc;
    ++*launches;
    if (ADJ) {
        finish_rev_f64(h, rp, part, nchunks, st);
        ++*launches;
    }
    return PSSGP_OK;
}

template <int D>
int pkfs_grad(pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H,
              const double* R, const double* y, const double* g_ll, double* fms, double* fPs, double* ll, double* sms,
              double* sPs, double* dP0, double* dFs, double* dQs, double* dH, double* dR, cudaStream_t st) {
    g_mid_warps = h->mid_warps;
    const bool fr = has_frag<D>() && !h->mid_smem;
    double* proj = (double*)h->mid_proj;   // set by pssgp_pkfs: (H sm, H sP H^T) per step instead of sms / sPs
    h->mid_proj = nullptr;
    if (proj != nullptr && !fr)
        return set_err(PSSGP_ERR_UNSUPPORTED, "pkfs: projected output needs the fragment-resident kernels (d <= 24)");
    const bool smooth = sms != nullptr || proj != nullptr, adj = dFs != nullptr;
    int rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, nullptr, 1);
    p.fms = fms; p.fPs = fPs; p.fms_in = fms; p.fPs_in = fPs; p.g = g_ll;
    p.sms = sms; p.sPs = sPs; p.dFs = dFs; p.dQs = dQs; p.dP0 = dP0; p.proj = proj;
    int gpc = k3_groups<D, true, true>(fr);
    if (k2_groups<D, true, false>(fr) < gpc) gpc = k2_groups<D, true, false>(fr);
    if (k1_groups<D>(fr) < gpc) gpc = k1_groups<D>(fr);
    const int L = pick_len(h, n, gpc);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier::total(nchunks);
    const int NAF = 3 * D * D + 2 * D, NSF = D + D * D, NAR = 3 * D * D + D, NSR = 2 * D * D + 2 * D;
    for (int kind = 0; kind < 3; ++kind) pending_clear(h, kind);
    if ((rc = ws_reserve(h, WS_LANE + KIND_FILTER, sizeof(double) * tot * NAF))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_FILTER, sizeof(double) * tot * NSF))) return rc;
    if ((rc = ws_reserve(h, WS_LANE + KIND_ADJOINT, sizeof(double) * tot * NAR))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_ADJOINT, sizeof(double) * tot * NSR))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(double) * (size_t)nchunks * (1 + D)))) return rc;
    double* aggs = (double*)h->buf[WS_LANE + KIND_FILTER];
    double* states = (double*)h->buf[WS_WAGG + KIND_FILTER];
    double* raggs = (double*)h->buf[WS_LANE + KIND_ADJOINT];
    double* rstates = (double*)h->buf[WS_WAGG + KIND_ADJOINT];
    double* part = (double*)h->buf[WS_PART];
    int launches = 0;
    if ((rc = launch_k1<D>(h, fr, p, L, nchunks, aggs, st, gpc))) return rc;
    ++launches;
    const GFilter<double>::Params gp = gfilter_params(p, D);
    if ((rc = hier_filter<D>(h, p, nchunks, aggs, states, nullptr, nullptr, false, st, &launches))) return rc;
    if ((rc = launch_k2<D, true, false>(h, fr, "mid_forward_rev", p, L, nchunks, states, part, raggs, st, gpc))) return rc;
    ++launches;
    if (ll != nullptr) {
        finish_filter_f64(h, gp, part, nchunks, ll, st);
        ++launches;
    }
    if (smooth && adj) rc = run_reverse<D, true, true>(h, fr, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches, gpc);
    else if (smooth) rc = run_reverse<D, true, false>(h, fr, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches, gpc);
    else rc = run_reverse<D, false, true>(h, fr, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches, gpc);
    if (rc) return rc;
    return check_launch(h, "mid pkfs_grad", launches);
}

template <int D>
int pkf_backward(pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs, const double* Qs,
                 const double* H, const double* R, const double* y, const double* fms, const double* fPs,
                 const double* g_ll, int first_special, double* dP0, double* dFs, double* dQs, double* dH, double* dR,
                 cudaStream_t st) {
    g_mid_warps = h->mid_warps;
    const bool fr = has_frag<D>() && !h->mid_smem;
    int rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, m0, first_special);
    p.fms_in = fms; p.fPs_in = fPs; p.g = g_ll; p.dFs = dFs; p.dQs = dQs; p.dP0 = dP0;
    const int g2 = k2_groups<D, true, true>(fr), g3 = k3_groups<D, false, true>(fr);
    const int gpc = g2 < g3 ? g2 : g3;
    const int L = pick_len(h, n, gpc);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier::total(nchunks);
    const int NAR = 3 * D * D + D, NSR = 2 * D * D + 2 * D;
    pending_clear(h, KIND_ADJOINT);
    if ((rc = ws_reserve(h, WS_LANE + KIND_ADJOINT, sizeof(double) * tot * NAR))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_ADJOINT, sizeof(double) * tot * NSR))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(double) * (size_t)nchunks * (1 + D)))) return rc;
    double* raggs = (double*)h->buf[WS_LANE + KIND_ADJOINT];
    double* rstates = (double*)h->buf[WS_WAGG + KIND_ADJOINT];
    double* part = (double*)h->buf[WS_PART];
    int launches = 0;
    if ((rc = launch_k2<D, true, true>(h, fr, "mid_forward_stored", p, L, nchunks, nullptr, nullptr, raggs, st, gpc))) return rc;
    ++launches;
    if ((rc = run_reverse<D, false, true>(h, fr, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches, gpc))) return rc;
    return check_launch(h, "mid pkf_backward", launches);
}


// ---- time sharding: one contiguous shard of the series per GPU ------------------------------------------------
// forward phase of one shard: filter seeded by (m0, P0) = filtered state entering the shard (rank 0: the prior,
// first_special = 1), + the shard summary of the combined reverse scan.  If pssgp_pkf_summary ran on the same arrays
// just before, its chunk aggregates are reused (no second reduce pass).
template <int D>
int shard_forward(pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H,
                  const double* R, const double* y, const double* m0, int first_special, double* fms, double* fPs,
                  double* ll, double* rev_summary, cudaStream_t st) {
    g_mid_warps = h->mid_warps;
    const bool fr = has_frag<D>() && !h->mid_smem;
    int rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, m0, first_special);
    p.fms = fms; p.fPs = fPs; p.fms_in = fms; p.fPs_in = fPs;
    int gpc = k3_groups<D, true, true>(fr);
    if (k2_groups<D, true, false>(fr) < gpc) gpc = k2_groups<D, true, false>(fr);
    if (k1_groups<D>(fr) < gpc) gpc = k1_groups<D>(fr);
    const int L = pick_len(h, n, gpc);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier::total(nchunks);
    const int NAF = 3 * D * D + 2 * D, NSF = D + D * D, NAR = 3 * D * D + D, NSR = 2 * D * D + 2 * D;
    const uint64_t key = filter_sig(8, D, n, Fs, Qs, y, H, R, first_special);
    const bool reuse = (h->pending_key[KIND_FILTER] == key && h->pending_n[KIND_FILTER] == n && h->pending_L[KIND_FILTER] == L);
    for (int kind = 0; kind < 3; ++kind) pending_clear(h, kind);
    if (!reuse)
        if ((rc = ws_reserve(h, WS_LANE + KIND_FILTER, sizeof(double) * tot * NAF))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_FILTER, sizeof(double) * tot * NSF))) return rc;
    if ((rc = ws_reserve(h, WS_LANE + KIND_ADJOINT, sizeof(double) * tot * NAR))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_ADJOINT, sizeof(double) * tot * NSR))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(double) * (size_t)nchunks * (1 + D)))) return rc;
    double* aggs = (double*)h->buf[WS_LANE + KIND_FILTER];
    double* states = (double*)h->buf[WS_WAGG + KIND_FILTER];
    double* raggs = (double*)h->buf[WS_LANE + KIND_ADJOINT];
    double* rstates = (double*)h->buf[WS_WAGG + KIND_ADJOINT];
    double* part = (double*)h->buf[WS_PART];
    int launches = 0;
    if (!reuse) {
        if ((rc = launch_k1<D>(h, fr, p, L, nchunks, aggs, st, gpc))) return rc;
        ++launches;
    }
    const GFilter<double>::Params gp = gfilter_params(p, D);
    if ((rc = hier_filter<D>(h, p, nchunks, aggs, states, nullptr, nullptr, reuse, st, &launches))) return rc;
    if ((rc = launch_k2<D, true, false>(h, fr, "mid_forward_rev", p, L, nchunks, states, part, raggs, st, gpc))) return rc;
    ++launches;
    if (ll != nullptr) {
        finish_filter_f64(h, gp, part, nchunks, ll, st);
        ++launches;
    }
    if ((rc = hier_rev<D>(h, nchunks, raggs, rstates, nullptr, rev_summary, false, st, &launches))) return rc;
    h->pending_key[KIND_ADJOINT] = adjoint_sig(8, D, n, Fs, Qs, y, H, R, fms, fPs, first_special);
    h->pending_n[KIND_ADJOINT] = n;
    h->pending_L[KIND_ADJOINT] = L;
    return check_launch(h, "mid shard_forward", launches);
}

// reverse phase of one shard: smoother (sms != nullptr) and gradient (dFs != nullptr) seeded by rev_init = state of
// the combined reverse scan entering the shard from the following shards (nullptr on the last rank).
template <int D>
int shard_reverse(pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs, const double* Qs,
                  const double* H, const double* R, const double* y, const double* fms, const double* fPs,
                  const double* g_ll, int first_special, const double* rev_init, double* sms, double* sPs, double* dP0,
                  double* dFs, double* dQs, double* dH, double* dR, cudaStream_t st) {
    g_mid_warps = h->mid_warps;
    const bool fr = has_frag<D>() && !h->mid_smem;
    const bool smooth = sms != nullptr, adj = dFs != nullptr;
    int rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, m0, first_special);
    p.fms_in = fms; p.fPs_in = fPs; p.g = g_ll; p.sms = sms; p.sPs = sPs; p.dFs = dFs; p.dQs = dQs; p.dP0 = dP0;
    int gpc = k3_groups<D, true, true>(fr);
    if (k2_groups<D, true, false>(fr) < gpc) gpc = k2_groups<D, true, false>(fr);
    if (k1_groups<D>(fr) < gpc) gpc = k1_groups<D>(fr);
    const int L = pick_len(h, n, gpc);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier::total(nchunks);
    const int NAR = 3 * D * D + D, NSR = 2 * D * D + 2 * D;
    const uint64_t key = adjoint_sig(8, D, n, Fs, Qs, y, H, R, fms, fPs, first_special);
    const bool reuse = (h->pending_key[KIND_ADJOINT] == key && h->pending_n[KIND_ADJOINT] == n && h->pending_L[KIND_ADJOINT] == L);
    pending_clear(h, KIND_ADJOINT);
    if (!reuse)
        if ((rc = ws_reserve(h, WS_LANE + KIND_ADJOINT, sizeof(double) * tot * NAR))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_ADJOINT, sizeof(double) * tot * NSR))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(double) * (size_t)nchunks * (1 + D)))) return rc;
    double* raggs = (double*)h->buf[WS_LANE + KIND_ADJOINT];
    double* rstates = (double*)h->buf[WS_WAGG + KIND_ADJOINT];
    double* part = (double*)h->buf[WS_PART];
    int launches = 0;
    if (!reuse) {
        // no aggregates left by shard_forward for these arrays: rebuild them from the stored moments
        if ((rc = launch_k2<D, true, true>(h, fr, "mid_forward_stored", p, L, nchunks, nullptr, nullptr, raggs, st, gpc))) return rc;
        ++launches;
    }
    if (smooth && adj) rc = run_reverse<D, true, true>(h, fr, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches, gpc, rev_init, reuse);
    else if (smooth) rc = run_reverse<D, true, false>(h, fr, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches, gpc, rev_init, reuse);
    else rc = run_reverse<D, false, true>(h, fr, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches, gpc, rev_init, reuse);
    if (rc) return rc;
    return check_launch(h, "mid shard_reverse", launches);
}

// folds the reverse-scan summaries of the `count` FOLLOWING shards (rank order, `stride` doubles apart) into the state
// entering this shard: state = 0 o summary[count-1] o ... o summary[0]
template <int D>
int rev_fold(pssgp_handle* h, const double* summaries, int count, int64_t stride, double* state_out, cudaStream_t st) {
    const typename hier::RevH<D>::Init in = {nullptr};
    int rc = hier::fold<hier::RevH<D>>(h, in, summaries + (int64_t)(count - 1) * stride, count, -(long)stride, state_out,
                                       "mhier_rev_fold", st);
    if (rc) return rc;
    return check_launch(h, "mid rev_fold", 1);
}

template int pkf<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*, const double*,
                        const double*, const double*, int, double*, double*, double*, double*, double*, cudaStream_t);
template int pkfs_grad<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*,
                              const double*, const double*, const double*, double*, double*, double*, double*, double*,
                              double*, double*, double*, double*, double*, cudaStream_t);
template int pkf_backward<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*,
                                 const double*, const double*, const double*, const double*, const double*, const double*,
                                 int, double*, double*, double*, double*, double*, cudaStream_t);

template int shard_forward<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*,
                                  const double*, const double*, const double*, int, double*, double*, double*, double*,
                                  cudaStream_t);
template int shard_reverse<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*,
                                  const double*, const double*, const double*, const double*, const double*, const double*,
                                  int, const double*, double*, double*, double*, double*, double*, double*, double*,
                                  cudaStream_t);
template int rev_fold<MID_D>(pssgp_handle*, const double*, int, int64_t, double*, cudaStream_t);

}  // namespace mid
}  // namespace pssgp
