// Explicit instantiation of the warp-level d > 4 path for ONE state dimension (compile with -DMID_D=<D>).
#include "mid.cuh"
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

// chunk length: one resident wave of groups of the most shared-memory-hungry kernel of the call
static int pick_len(const pssgp_handle* h, int64_t n, int gpc) {
    if (h->chunk_opt > 0) return (int)h->chunk_opt;
    const int64_t slots = (int64_t)(h->num_sms > 0 ? h->num_sms : 1) * gpc;
    int64_t L = (n + slots - 1) / slots;
    if (L < 16) L = 16;
    if (L > 4096) L = 4096;
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
    constexpr int WG = default_wg<D>();
    using KA = K1<D, WG>;
    using KB = K2<D, WG, false, false>;
    int rc;
    if ((rc = set_smem_attr(k1_filter_reduce<D, WG>, (size_t)KA::GPC * KA::GROUP_DOUBLES * 8))) return rc;
    if ((rc = set_smem_attr(k2_forward<D, WG, false, false>, (size_t)KB::GPC * KB::GROUP_DOUBLES * 8))) return rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, m0, first_special);
    p.fms = fms; p.fPs = fPs;
    const int L = pick_len(h, n, KB::GPC < KA::GPC ? KB::GPC : KA::GPC);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier_total(nchunks);
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
        const unsigned grid = (unsigned)((nchunks + KA::GPC - 1) / KA::GPC);
        PSSGP_LAUNCH(h, "mid_filter_reduce", st,
                     (k1_filter_reduce<D, WG><<<grid, KA::GPC * WG * 32, (size_t)KA::GPC * KA::GROUP_DOUBLES * 8, st>>>(
                         p, L, nchunks, aggs)));
        ++launches;
    }
    const GFilter<double>::Params gp = gfilter_params(p, D);
    if ((rc = hier_filter_f64(h, gp, D, nchunks, aggs, states, final_state, summary, reuse, st, &launches))) return rc;
    if (summary != nullptr) {
        h->pending_key[KIND_FILTER] = key;
        h->pending_n[KIND_FILTER] = n;
        h->pending_L[KIND_FILTER] = L;
        return check_launch(h, "mid pkf summary", launches);
    }
    {
        const unsigned grid = (unsigned)((nchunks + KB::GPC - 1) / KB::GPC);
        PSSGP_LAUNCH(h, "mid_forward", st,
                     (k2_forward<D, WG, false, false><<<grid, KB::GPC * WG * 32, (size_t)KB::GPC * KB::GROUP_DOUBLES * 8, st>>>(
                         p, L, nchunks, states, part, nullptr)));
        ++launches;
    }
    if (ll != nullptr) {
        finish_filter_f64(h, gp, part, nchunks, ll, st);
        ++launches;
    }
    return check_launch(h, "mid pkf", launches);
}

// reverse part shared by pkfs_grad and pkf_backward: hierarchy over the reverse aggregates + K3 + finish
template <int D, bool SMOOTH, bool ADJ>
static int run_reverse(pssgp_handle* h, const Params& p, int L, int64_t nchunks, double* raggs, double* rstates,
                       double* part, double* dH, double* dR, cudaStream_t st, int* launches) {
    constexpr int WG = default_wg<D>();
    using KC = K3<D, WG, SMOOTH, ADJ>;
    int rc;
    if ((rc = set_smem_attr(k3_reverse<D, WG, SMOOTH, ADJ>, (size_t)KC::GPC * KC::GROUP_DOUBLES * 8))) return rc;
    const GRev<double>::Params rp = grev_params(p, D, dH, dR);
    if ((rc = hier_rev_f64(h, rp, D, nchunks, raggs, rstates, nullptr, nullptr, false, st, launches))) return rc;
    const unsigned grid = (unsigned)((nchunks + KC::GPC - 1) / KC::GPC);
    PSSGP_LAUNCH(h, "mid_reverse", st,
                 (k3_reverse<D, WG, SMOOTH, ADJ><<<grid, KC::GPC * WG * 32, (size_t)KC::GPC * KC::GROUP_DOUBLES * 8, st>>>(
                     p, L, nchunks, rstates, part)));
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
    constexpr int WG = default_wg<D>();
    using KA = K1<D, WG>;
    using KB = K2<D, WG, true, false>;
    using KC = K3<D, WG, true, true>;
    const bool smooth = sms != nullptr, adj = dFs != nullptr;
    int rc;
    if ((rc = set_smem_attr(k1_filter_reduce<D, WG>, (size_t)KA::GPC * KA::GROUP_DOUBLES * 8))) return rc;
    if ((rc = set_smem_attr(k2_forward<D, WG, true, false>, (size_t)KB::GPC * KB::GROUP_DOUBLES * 8))) return rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, nullptr, 1);
    p.fms = fms; p.fPs = fPs; p.fms_in = fms; p.fPs_in = fPs; p.g = g_ll;
    p.sms = sms; p.sPs = sPs; p.dFs = dFs; p.dQs = dQs; p.dP0 = dP0;
    int gpc = KC::GPC;
    if (KB::GPC < gpc) gpc = KB::GPC;
    if (KA::GPC < gpc) gpc = KA::GPC;
    const int L = pick_len(h, n, gpc);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier_total(nchunks);
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
    {
        const unsigned grid = (unsigned)((nchunks + KA::GPC - 1) / KA::GPC);
        PSSGP_LAUNCH(h, "mid_filter_reduce", st,
                     (k1_filter_reduce<D, WG><<<grid, KA::GPC * WG * 32, (size_t)KA::GPC * KA::GROUP_DOUBLES * 8, st>>>(
                         p, L, nchunks, aggs)));
        ++launches;
    }
    const GFilter<double>::Params gp = gfilter_params(p, D);
    if ((rc = hier_filter_f64(h, gp, D, nchunks, aggs, states, nullptr, nullptr, false, st, &launches))) return rc;
    {
        const unsigned grid = (unsigned)((nchunks + KB::GPC - 1) / KB::GPC);
        PSSGP_LAUNCH(h, "mid_forward_rev", st,
                     (k2_forward<D, WG, true, false><<<grid, KB::GPC * WG * 32, (size_t)KB::GPC * KB::GROUP_DOUBLES * 8, st>>>(
                         p, L, nchunks, states, part, raggs)));
        ++launches;
    }
    if (ll != nullptr) {
        finish_filter_f64(h, gp, part, nchunks, ll, st);
        ++launches;
    }
    if (smooth && adj) rc = run_reverse<D, true, true>(h, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches);
    else if (smooth) rc = run_reverse<D, true, false>(h, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches);
    else rc = run_reverse<D, false, true>(h, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches);
    if (rc) return rc;
    return check_launch(h, "mid pkfs_grad", launches);
}

template <int D>
int pkf_backward(pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs, const double* Qs,
                 const double* H, const double* R, const double* y, const double* fms, const double* fPs,
                 const double* g_ll, int first_special, double* dP0, double* dFs, double* dQs, double* dH, double* dR,
                 cudaStream_t st) {
    constexpr int WG = default_wg<D>();
    using KB = K2<D, WG, true, true>;
    using KC = K3<D, WG, false, true>;
    int rc;
    if ((rc = set_smem_attr(k2_forward<D, WG, true, true>, (size_t)KB::GPC * KB::GROUP_DOUBLES * 8))) return rc;
    Params p = base_params(n, P0, Fs, Qs, H, R, y, m0, first_special);
    p.fms_in = fms; p.fPs_in = fPs; p.g = g_ll; p.dFs = dFs; p.dQs = dQs; p.dP0 = dP0;
    const int L = pick_len(h, n, KB::GPC < KC::GPC ? KB::GPC : KC::GPC);
    const int64_t nchunks = (n + L - 1) / L;
    const size_t tot = hier_total(nchunks);
    const int NAR = 3 * D * D + D, NSR = 2 * D * D + 2 * D;
    pending_clear(h, KIND_ADJOINT);
    if ((rc = ws_reserve(h, WS_LANE + KIND_ADJOINT, sizeof(double) * tot * NAR))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG + KIND_ADJOINT, sizeof(double) * tot * NSR))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(double) * (size_t)nchunks * (1 + D)))) return rc;
    double* raggs = (double*)h->buf[WS_LANE + KIND_ADJOINT];
    double* rstates = (double*)h->buf[WS_WAGG + KIND_ADJOINT];
    double* part = (double*)h->buf[WS_PART];
    int launches = 0;
    {
        const unsigned grid = (unsigned)((nchunks + KB::GPC - 1) / KB::GPC);
        PSSGP_LAUNCH(h, "mid_forward_stored", st,
                     (k2_forward<D, WG, true, true><<<grid, KB::GPC * WG * 32, (size_t)KB::GPC * KB::GROUP_DOUBLES * 8, st>>>(
                         p, L, nchunks, nullptr, nullptr, raggs)));
        ++launches;
    }
    if ((rc = run_reverse<D, false, true>(h, p, L, nchunks, raggs, rstates, part, dH, dR, st, &launches))) return rc;
    return check_launch(h, "mid pkf_backward", launches);
}

template int pkf<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*, const double*,
                        const double*, const double*, int, double*, double*, double*, double*, double*, cudaStream_t);
template int pkfs_grad<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*,
                              const double*, const double*, const double*, double*, double*, double*, double*, double*,
                              double*, double*, double*, double*, double*, cudaStream_t);
template int pkf_backward<MID_D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*,
                                 const double*, const double*, const double*, const double*, const double*, const double*,
                                 int, double*, double*, double*, double*, double*, cudaStream_t);

}  // namespace mid
}  // namespace pssgp
