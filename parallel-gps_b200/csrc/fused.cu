// C ABI: pssgp_pkfs_grad — filter + log-likelihood + smoother + gradient in one call (see include/pssgp_b200.h).
#include "fused_small.cuh"
#include "scan_run.cuh"
#include "generic_algebras.cuh"
#include "mid_host.h"

namespace pssgp {

constexpr int kNotFused = -12345;  // this (dtype, d) has no common partition: run the three scans one after the other

// pdl: launch with programmatic stream serialization (the kernel waits for its predecessor with
// griddepcontrol.wait before it reads anything, see scan_stream.cuh) so that its CTAs can become resident while the
// predecessor's last CTA is still scanning the CTA totals.
template <typename Alg>
int launch_apply(pssgp_handle* h, const typename Alg::Params& p, const StreamPart& sp, const typename Alg::scalar* lane,
                 const typename Alg::scalar* wexcl, const typename Alg::scalar* wstate, typename Alg::scalar* part,
                 typename Alg::scalar* acc_out, cudaStream_t st, bool pdl = false,
                 const typename Alg::scalar* wprefix = nullptr) {
    using Lay = StreamLayout<Alg>;
    constexpr int NW = Lay::NW;
    const long nChunksPad = (long)sp.nCta * NW * 32;
    if (pdl && !h->timing) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)sp.nCta);
        cfg.blockDim = dim3(NW * 32);
        cfg.dynamicSmemBytes = NW * Lay::WARP_BYTES_APPLY;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, stream_apply_kernel<Alg>, p, sp, nChunksPad, lane, wexcl, wstate, part, h->ticket,
                           acc_out, wprefix);
        return PSSGP_OK;
    }
    PSSGP_LAUNCH(h, Alg::name_apply(), st,
                 (stream_apply_kernel<Alg><<<(unsigned)sp.nCta, NW * 32, NW * Lay::WARP_BYTES_APPLY, st>>>(
                     p, sp, nChunksPad, lane, wexcl, wstate, part, h->ticket, acc_out, wprefix)));
    return PSSGP_OK;
}

// SMOOTH = false: filter + log-likelihood + gradient only (sms / sPs not produced, the smoother's chunk aggregates
// are not built): the training step of StateSpaceGP.maximum_log_likelihood_objective.
template <typename T, int D, bool SMOOTH>
int pkfs_grad_impl_t(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H,
                     const void* R, const void* y, const void* g_ll, void* fms, void* fPs, void* ll, void* sms, void* sPs,
                     void* dP0, void* dFs, void* dQs, void* dH, void* dR, cudaStream_t st) {
    using FA = FilterAlg<T, D>;
    using FF = FusedFwdAlg<T, D, SMOOTH, true>;
    using SA = SmootherAlg<T, D>;
    using AA = AdjointAlg<T, D>;
    using FR = FusedRevAlg<T, D>;
    constexpr int NW = StreamLayout<FA>::NW;
    constexpr int LS = StreamLayout<FA>::LS;
    // the fused step needs one partition of the time axis for all of its kernels
    if constexpr (StreamLayout<FF>::NW != NW || StreamLayout<SA>::NW != NW || StreamLayout<AA>::NW != NW) {
        return kNotFused;
    } else {
    int rc;
    if ((rc = stream_configure<FA>(h->device))) return rc;
    if ((rc = stream_configure_apply<FF>(h->device))) return rc;
    if ((rc = stream_configure<SA>(h->device))) return rc;
    if ((rc = stream_configure<AA>(h->device))) return rc;
    const void* arrs[] = {Fs, Qs, y, fms, fPs, sms, sPs, dFs, dQs};
    for (const void* a : arrs)
        if (a != nullptr && !aligned16(a)) return set_err(PSSGP_ERR_INVALID, "pkfs_grad: arrays must be 16-byte aligned");
    const StreamPart sp = make_partition<NW, LS>(h, n);
    const int64_t nCta = sp.nCta;
    const int64_t nChunksPad = nCta * NW * 32;
    for (int kind = 0; kind < 3; ++kind) pending_clear(h, kind);
    // the two reverse scans share one workspace (smoother rows first) so that K3' sees one aggregate / state
    const int NAGG[3] = {FA::NAGG, SA::NAGG + AA::NAGG, AA::NAGG};
    for (int kind = 0; kind < 3; ++kind) {
        if ((rc = ws_reserve(h, WS_LANE + kind, sizeof(T) * NAGG[kind] * (size_t)nChunksPad))) return rc;
        if ((rc = ws_reserve(h, WS_WAGG + kind, sizeof(T) * NAGG[kind] * (size_t)nCta))) return rc;
        if ((rc = ws_reserve(h, WS_WEXCL + kind, sizeof(T) * NAGG[kind] * (size_t)nCta * NW))) return rc;
    }
    if ((rc = ws_reserve(h, WS_WSTATE, sizeof(T) * FA::NSTATE * (size_t)nCta))) return rc;
    if ((rc = ws_reserve(h, WS_WSTATE_S, sizeof(T) * (SA::NSTATE + AA::NSTATE) * (size_t)nCta))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * (AA::NACC + 1) * (size_t)nCta))) return rc;
    T* part = (T*)h->buf[WS_PART];

    typename FF::Params fp;
    fp.Fs = (const T*)Fs;
    fp.Qs = (const T*)Qs;
    fp.y = (const T*)y;
    fp.H = (const T*)H;
    fp.R = (const T*)R;
    fp.P0 = (const T*)P0;
    fp.m0 = nullptr;
    fp.fms = (T*)fms;
    fp.fPs = (T*)fPs;
    fp.first_special = 1;
    fp.n = n;
    fp.sm = {(T*)h->buf[WS_LANE + KIND_SMOOTHER], (T*)h->buf[WS_WEXCL + KIND_SMOOTHER],
             (T*)h->buf[WS_WAGG + KIND_SMOOTHER], (T*)h->buf[WS_WSTATE_S]};
    fp.ad = {fp.sm.lane_excl + (size_t)SA::NAGG * nChunksPad, fp.sm.warp_excl + (size_t)SA::NAGG * nCta * NW,
             (T*)h->buf[WS_WAGG + KIND_ADJOINT], fp.sm.wstate + (size_t)SA::NSTATE * nCta};
    fp.side_ticket = h->ticket + 1;
    fp.last_special = 1;
    fp.Fnext = fp.Qnext = nullptr;
    fp.sm_summary = fp.ad_summary = nullptr;

    // K1: chunk aggregates of the filter + scan over the CTA totals
    {
        const typename FA::Params& bp = fp;
        using Lay = StreamLayout<FA>;
        PSSGP_LAUNCH(h, FA::name_reduce(), st,
                     (stream_reduce_kernel<FA><<<(unsigned)nCta, NW * 32, NW * Lay::WARP_BYTES_REDUCE, st>>>(
                         bp, sp, nChunksPad, (T*)h->buf[WS_LANE + KIND_FILTER], (T*)h->buf[WS_WEXCL + KIND_FILTER],
                         (T*)h->buf[WS_WAGG + KIND_FILTER], (T*)h->buf[WS_WSTATE], (T*)nullptr, h->ticket + 1,
                         (T*)nullptr, (T*)nullptr)));
    }
    // K2': seeded filter recursion + chunk aggregates and CTA-level scans of both reverse scans
    launch_apply<FF>(h, fp, sp, (const T*)h->buf[WS_LANE + KIND_FILTER], (const T*)h->buf[WS_WEXCL + KIND_FILTER],
                     (const T*)h->buf[WS_WSTATE], part, (T*)ll, st, h->pdl != 0);
    // K3: smoother and adjoint recursions seeded by the states K2' produced
    typename SA::Params sp_;
    sp_.Fs = (const T*)Fs;
    sp_.Qs = (const T*)Qs;
    sp_.fms = (const T*)fms;
    sp_.fPs = (const T*)fPs;
    sp_.sms = (T*)sms;
    sp_.sPs = (T*)sPs;
    sp_.n = n;
    sp_.last_special = 1;
    sp_.Fnext = sp_.Qnext = sp_.init = nullptr;
    typename AA::Params ap;
    ap.Fs = (const T*)Fs;
    ap.Qs = (const T*)Qs;
    ap.y = (const T*)y;
    ap.H = (const T*)H;
    ap.R = (const T*)R;
    ap.P0 = (const T*)P0;
    ap.m0 = nullptr;
    ap.fms = (const T*)fms;
    ap.fPs = (const T*)fPs;
    ap.g = (const T*)g_ll;
    ap.init = nullptr;
    ap.dFs = (T*)dFs;
    ap.dQs = (T*)dQs;
    ap.dP0 = (T*)dP0;
    ap.dH = (T*)dH;
    ap.dR = (T*)dR;
    ap.first_state = nullptr;
    ap.n = n;
    ap.first_special = 1;
    if constexpr (SMOOTH && StreamLayout<FR>::NW == NW) {
        // option "fused_reverse": one kernel for both reverse recursions (one read of F, Q, y, fms, fPs instead of
        // two).  Off by default: at d = 3 / FP64 the fused kernel is issue-bound at 254 registers per thread and
        // takes longer (198 us) than the two separate ones (80 + 98 us); see DESIGN.md.
        if (h->fused_reverse) {
            if ((rc = stream_configure_apply<FR>(h->device))) return rc;
            typename FR::Params rp;
            rp.s = sp_;
            rp.a = ap;
            launch_apply<FR>(h, rp, sp, fp.sm.lane_excl, fp.sm.warp_excl, fp.sm.wstate, part, (T*)dR, st);
            return check_launch(h, "pkfs_grad", 3);
        }
    }
    if constexpr (SMOOTH)
        launch_apply<SA>(h, sp_, sp, fp.sm.lane_excl, fp.sm.warp_excl, fp.sm.wstate, part, (T*)nullptr, st, h->pdl != 0);
    launch_apply<AA>(h, ap, sp, fp.ad.lane_excl, fp.ad.warp_excl, fp.ad.wstate, part, (T*)dR, st, h->pdl != 0);
    return check_launch(h, "pkfs_grad", SMOOTH ? 4 : 3);
    }
}

template <typename T, int D>
int pkfs_grad_impl(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H,
                   const void* R, const void* y, const void* g_ll, void* fms, void* fPs, void* ll, void* sms, void* sPs,
                   void* dP0, void* dFs, void* dQs, void* dH, void* dR, cudaStream_t st) {
    if (sms != nullptr)
        return pkfs_grad_impl_t<T, D, true>(h, n, P0, Fs, Qs, H, R, y, g_ll, fms, fPs, ll, sms, sPs, dP0, dFs, dQs, dH, dR, st);
    return pkfs_grad_impl_t<T, D, false>(h, n, P0, Fs, Qs, H, R, y, g_ll, fms, fPs, ll, sms, sPs, dP0, dFs, dQs, dH, dR, st);
}

// One shard of a time-sharded series: seeded filter recursion (pssgp_pkf) + chunk aggregates and shard summaries of
// both reverse scans, left in the workspace for the pssgp_pks / pssgp_pkf_backward calls that follow.
template <typename T, int D>
int pkf_with_summaries_impl(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H,
                            const void* R, const void* y, const void* m0, int first_special, int last_special,
                            const void* Fnext, const void* Qnext, void* fms, void* fPs, void* ll, void* sm_summary,
                            void* ad_summary, cudaStream_t st) {
    using FA = FilterAlg<T, D>;
    using FF = FusedFwdAlg<T, D>;
    using SA = SmootherAlg<T, D>;
    using AA = AdjointAlg<T, D>;
    constexpr int NW = StreamLayout<FA>::NW;
    constexpr int LS = StreamLayout<FA>::LS;
    if constexpr (StreamLayout<FF>::NW != NW || StreamLayout<SA>::NW != NW || StreamLayout<AA>::NW != NW) {
        return kNotFused;
    } else {
    int rc;
    if ((rc = stream_configure<FA>(h->device))) return rc;
    if ((rc = stream_configure_apply<FF>(h->device))) return rc;
    const void* arrs[] = {Fs, Qs, y, fms, fPs};
    for (const void* a : arrs)
        if (!aligned16(a)) return set_err(PSSGP_ERR_INVALID, "pkf_with_summaries: arrays must be 16-byte aligned");
    const StreamPart sp = make_partition<NW, LS>(h, n);
    const int64_t nCta = sp.nCta;
    const int64_t nChunksPad = nCta * NW * 32;
    const bool reuse = h->pending_key[KIND_FILTER] == filter_sig(sizeof(T), D, n, Fs, Qs, y, H, R, first_special) &&
                       h->pending_n[KIND_FILTER] == n && h->pending_L[KIND_FILTER] == sp.L;
    const bool have_prefix = reuse && h->pending_prefix[KIND_FILTER];
    for (int kind = 0; kind < 3; ++kind) pending_clear(h, kind);
    const int NAGG[3] = {FA::NAGG, SA::NAGG, AA::NAGG};
    for (int kind = reuse ? 1 : 0; kind < 3; ++kind) {
        if ((rc = ws_reserve(h, WS_LANE + kind, sizeof(T) * NAGG[kind] * (size_t)nChunksPad))) return rc;
        if ((rc = ws_reserve(h, WS_WAGG + kind, sizeof(T) * NAGG[kind] * (size_t)nCta))) return rc;
        if ((rc = ws_reserve(h, WS_WEXCL + kind, sizeof(T) * NAGG[kind] * (size_t)nCta * NW))) return rc;
    }
    for (int kind = 1; kind < 3; ++kind)
        if ((rc = ws_reserve(h, WS_WPREFIX + kind, sizeof(T) * NAGG[kind] * (size_t)nCta))) return rc;
    if ((rc = ws_reserve(h, WS_WSTATE, sizeof(T) * FA::NSTATE * (size_t)nCta))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * (AA::NACC + 1) * (size_t)nCta))) return rc;
    typename FF::Params fp;
    fp.Fs = (const T*)Fs;
    fp.Qs = (const T*)Qs;
    fp.y = (const T*)y;
    fp.H = (const T*)H;
    fp.R = (const T*)R;
    fp.P0 = (const T*)P0;
    fp.m0 = (const T*)m0;
    fp.fms = (T*)fms;
    fp.fPs = (T*)fPs;
    fp.first_special = first_special;
    fp.n = n;
    fp.sm = {(T*)h->buf[WS_LANE + KIND_SMOOTHER], (T*)h->buf[WS_WEXCL + KIND_SMOOTHER],
             (T*)h->buf[WS_WAGG + KIND_SMOOTHER], (T*)h->buf[WS_WPREFIX + KIND_SMOOTHER]};
    fp.ad = {(T*)h->buf[WS_LANE + KIND_ADJOINT], (T*)h->buf[WS_WEXCL + KIND_ADJOINT],
             (T*)h->buf[WS_WAGG + KIND_ADJOINT], (T*)h->buf[WS_WPREFIX + KIND_ADJOINT]};
    fp.side_ticket = h->ticket + 1;
    fp.last_special = last_special;
    fp.Fnext = (const T*)Fnext;
    fp.Qnext = (const T*)Qnext;
    fp.sm_summary = (T*)sm_summary;
    fp.ad_summary = (T*)ad_summary;
    // filter summaries of the previous shards registered with pssgp_set_fold: folded onto (m0, P0) in K2' itself
    if (h->fold_count[KIND_FILTER] > 0) {
        if (!have_prefix) {
            fold_clear(h, KIND_FILTER);
            return set_err(PSSGP_ERR_INVALID, "pkf_with_summaries: pssgp_set_fold(kind 0) needs the aggregates of a "
                                              "pssgp_pkf_summary call on the same arrays");
        }
        fp.fold = (const T*)h->fold_ptr[KIND_FILTER];
        fp.fold_count = h->fold_count[KIND_FILTER];
        fp.fold_stride = (long)h->fold_stride[KIND_FILTER];
        fp.state_in_out = (T*)h->fold_state_out;
        h->fold_count[KIND_FILTER] = 0;
        h->fold_ptr[KIND_FILTER] = nullptr;
    }
    const typename FA::Params& bp = fp;
    int nl = 1;
    if (!reuse) {
        using Lay = StreamLayout<FA>;
        PSSGP_LAUNCH(h, FA::name_reduce(), st,
                     (stream_reduce_kernel<FA><<<(unsigned)nCta, NW * 32, NW * Lay::WARP_BYTES_REDUCE, st>>>(
                         bp, sp, nChunksPad, (T*)h->buf[WS_LANE + KIND_FILTER], (T*)h->buf[WS_WEXCL + KIND_FILTER],
                         (T*)h->buf[WS_WAGG + KIND_FILTER], (T*)h->buf[WS_WSTATE], (T*)nullptr, h->ticket + 1,
                         (T*)nullptr, (T*)nullptr)));
        ++nl;
    } else if (!have_prefix) {
        // the chunk aggregates are those pssgp_pkf_summary left behind: only the scan over the CTA totals is missing
        int midThreads = kMidThreads;
        if (nCta < kMidThreads) midThreads = (int)(((nCta + 31) / 32) * 32);
        PSSGP_LAUNCH(h, FA::name_mid(), st,
                     (scan_mid_kernel<FA><<<1, midThreads, 0, st>>>(bp, (const T*)h->buf[WS_WAGG + KIND_FILTER], nCta,
                                                                   (T*)h->buf[WS_WSTATE], (T*)nullptr)));
        ++nl;
    }
    // with prefix aggregates pending, K2' starts from (m0, P0) o prefix[CTA]: no scan over the CTA totals at all
    launch_apply<FF>(h, fp, sp, (const T*)h->buf[WS_LANE + KIND_FILTER], (const T*)h->buf[WS_WEXCL + KIND_FILTER],
                     (const T*)h->buf[WS_WSTATE], (T*)h->buf[WS_PART], (T*)ll, st, false,
                     have_prefix ? (const T*)h->buf[WS_WPREFIX + KIND_FILTER] : (const T*)nullptr);
    // pssgp_pks (key: fPs) and pssgp_pkf_backward (key: fms) on the same arrays skip their reduce kernels
    h->pending_key[KIND_SMOOTHER] = smoother_sig(sizeof(T), D, n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext);
    h->pending_key[KIND_ADJOINT] = adjoint_sig(sizeof(T), D, n, Fs, Qs, y, H, R, fms, fPs, first_special);
    h->pending_n[KIND_SMOOTHER] = h->pending_n[KIND_ADJOINT] = n;
    h->pending_L[KIND_SMOOTHER] = h->pending_L[KIND_ADJOINT] = sp.L;
    h->pending_prefix[KIND_SMOOTHER] = h->pending_prefix[KIND_ADJOINT] = 1;
    return check_launch(h, "pkf_with_summaries", nl);
    }
}

// Filter (+ log-likelihood) + RTS smoother of one whole series, fused: K1, K2' (filter apply + smoother aggregates
// and scans), then the seeded smoother recursion writing (sms, sPs) or, when proj != nullptr, only the projection
// (H m, H P H^T) of every smoothed state.
template <typename T, int D>
int pkfs_impl(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H, const void* R,
              const void* y, void* fms, void* fPs, void* ll, void* sms, void* sPs, void* proj, cudaStream_t st) {
    using FA = FilterAlg<T, D>;
    using FF = FusedFwdAlg<T, D, true, false>;
    using SA = SmootherAlg<T, D>;
    using SP = SmootherProjAlg<T, D>;
    constexpr int NW = StreamLayout<FA>::NW;
    constexpr int LS = StreamLayout<FA>::LS;
    if constexpr (StreamLayout<FF>::NW != NW || StreamLayout<SA>::NW != NW || StreamLayout<SP>::NW != NW) {
        return kNotFused;
    } else {
    int rc;
    if ((rc = stream_configure<FA>(h->device))) return rc;
    if ((rc = stream_configure_apply<FF>(h->device))) return rc;
    if ((rc = stream_configure<SA>(h->device))) return rc;
    if ((rc = stream_configure_apply<SP>(h->device))) return rc;
    const void* arrs[] = {Fs, Qs, y, fms, fPs, sms, sPs, proj};
    for (const void* a : arrs)
        if (!aligned16(a)) return set_err(PSSGP_ERR_INVALID, "pkfs: arrays must be 16-byte aligned");
    const StreamPart sp = make_partition<NW, LS>(h, n);
    const int64_t nCta = sp.nCta;
    const int64_t nChunksPad = nCta * NW * 32;
    for (int kind = 0; kind < 3; ++kind) pending_clear(h, kind);
    const int NAGG[2] = {FA::NAGG, SA::NAGG};
    for (int kind = 0; kind < 2; ++kind) {
        if ((rc = ws_reserve(h, WS_LANE + kind, sizeof(T) * NAGG[kind] * (size_t)nChunksPad))) return rc;
        if ((rc = ws_reserve(h, WS_WAGG + kind, sizeof(T) * NAGG[kind] * (size_t)nCta))) return rc;
        if ((rc = ws_reserve(h, WS_WEXCL + kind, sizeof(T) * NAGG[kind] * (size_t)nCta * NW))) return rc;
    }
    if ((rc = ws_reserve(h, WS_WSTATE, sizeof(T) * FA::NSTATE * (size_t)nCta))) return rc;
    if ((rc = ws_reserve(h, WS_WSTATE_S, sizeof(T) * SA::NSTATE * (size_t)nCta))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * 2 * (size_t)nCta))) return rc;
    T* part = (T*)h->buf[WS_PART];
    typename FF::Params fp;
    fp.Fs = (const T*)Fs;
    fp.Qs = (const T*)Qs;
    fp.y = (const T*)y;
    fp.H = (const T*)H;
    fp.R = (const T*)R;
    fp.P0 = (const T*)P0;
    fp.m0 = nullptr;
    fp.fms = (T*)fms;
    fp.fPs = (T*)fPs;
    fp.first_special = 1;
    fp.n = n;
    fp.sm = {(T*)h->buf[WS_LANE + KIND_SMOOTHER], (T*)h->buf[WS_WEXCL + KIND_SMOOTHER],
             (T*)h->buf[WS_WAGG + KIND_SMOOTHER], (T*)h->buf[WS_WSTATE_S]};
    fp.ad = {nullptr, nullptr, nullptr, nullptr};
    fp.side_ticket = h->ticket + 1;
    fp.last_special = 1;
    fp.Fnext = fp.Qnext = nullptr;
    fp.sm_summary = fp.ad_summary = nullptr;
    {
        const typename FA::Params& bp = fp;
        using Lay = StreamLayout<FA>;
        PSSGP_LAUNCH(h, FA::name_reduce(), st,
                     (stream_reduce_kernel<FA><<<(unsigned)nCta, NW * 32, NW * Lay::WARP_BYTES_REDUCE, st>>>(
                         bp, sp, nChunksPad, (T*)h->buf[WS_LANE + KIND_FILTER], (T*)h->buf[WS_WEXCL + KIND_FILTER],
                         (T*)h->buf[WS_WAGG + KIND_FILTER], (T*)h->buf[WS_WSTATE], (T*)nullptr, h->ticket + 1,
                         (T*)nullptr, (T*)nullptr)));
    }
    launch_apply<FF>(h, fp, sp, (const T*)h->buf[WS_LANE + KIND_FILTER], (const T*)h->buf[WS_WEXCL + KIND_FILTER],
                     (const T*)h->buf[WS_WSTATE], part, (T*)ll, st, h->pdl != 0);
    typename SP::Params sp_;
    sp_.Fs = (const T*)Fs;
    sp_.Qs = (const T*)Qs;
    sp_.fms = (const T*)fms;
    sp_.fPs = (const T*)fPs;
    sp_.sms = (T*)sms;
    sp_.sPs = (T*)sPs;
    sp_.n = n;
    sp_.last_special = 1;
    sp_.Fnext = sp_.Qnext = sp_.init = nullptr;
    sp_.H = (const T*)H;
    sp_.proj = (T*)proj;
    if (proj != nullptr) {
        launch_apply<SP>(h, sp_, sp, fp.sm.lane_excl, fp.sm.warp_excl, fp.sm.wstate, part, (T*)nullptr, st, h->pdl != 0);
    } else {
        const typename SA::Params& bsp = sp_;
        launch_apply<SA>(h, bsp, sp, fp.sm.lane_excl, fp.sm.warp_excl, fp.sm.wstate, part, (T*)nullptr, st, h->pdl != 0);
    }
    return check_launch(h, "pkfs", 3);
    }
}

static int fused_dispatch(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                          const void* H, const void* R, const void* y, const void* g_ll, void* fms, void* fPs, void* ll,
                          void* sms, void* sPs, void* dP0, void* dFs, void* dQs, void* dH, void* dR, cudaStream_t st) {
    DISPATCH_SMALL(pkfs_grad_impl, h, n, P0, Fs, Qs, H, R, y, g_ll, fms, fPs, ll, sms, sPs, dP0, dFs, dQs, dH, dR, st);
    return kNotFused;
}

static int pkfs_dispatch(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                         const void* H, const void* R, const void* y, void* fms, void* fPs, void* ll, void* sms, void* sPs,
                         void* proj, cudaStream_t st) {
    DISPATCH_SMALL(pkfs_impl, h, n, P0, Fs, Qs, H, R, y, fms, fPs, ll, sms, sPs, proj, st);
    return kNotFused;
}

// d <= 4 without a common partition (pkf_with_summaries falls back to separate scans): the registered filter fold is
// done by the one-thread fold kernel, reading the summaries at their stride.
template <typename T, int D>
int filter_fold_strided_impl(pssgp_handle* h, int count, const void* P0, const void* m0, const void* summaries,
                             long stride, void* state_out, cudaStream_t st) {
    typename FilterAlg<T, D>::Params p;
    p.Fs = p.Qs = p.y = p.H = p.R = nullptr;
    p.P0 = (const T*)P0;
    p.m0 = (const T*)m0;
    p.fms = p.fPs = nullptr;
    p.first_special = 0;
    return run_fold<FilterAlg<T, D>>(h, p, (const T*)summaries, count, stride, (T*)state_out, st);
}

static int filter_fold_strided_dispatch(pssgp_handle* h, int dtype, int d, int count, const void* P0, const void* m0,
                                        const void* summaries, long stride, void* state_out, cudaStream_t st) {
    DISPATCH_SMALL(filter_fold_strided_impl, h, count, P0, m0, summaries, stride, state_out, st);
    return set_err(PSSGP_ERR_UNSUPPORTED, "pssgp_set_fold is implemented for d <= 4 (d = %d)", d);
}

static int with_summaries_dispatch(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs,
                                   const void* Qs, const void* H, const void* R, const void* y, const void* m0,
                                   int first_special, int last_special, const void* Fnext, const void* Qnext, void* fms,
                                   void* fPs, void* ll, void* sm_summary, void* ad_summary, cudaStream_t st) {
    DISPATCH_SMALL(pkf_with_summaries_impl, h, n, P0, Fs, Qs, H, R, y, m0, first_special, last_special, Fnext, Qnext, fms,
                   fPs, ll, sm_summary, ad_summary, st);
    return kNotFused;
}

}  // namespace pssgp

using namespace pssgp;

extern "C" {

int pssgp_pkfs_grad(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                    const void* H, const void* R, const void* y, const void* g_ll, void* fms, void* fPs, void* ll,
                    void* sms, void* sPs, void* dP0, void* dFs, void* dQs, void* dH, void* dR, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !g_ll || !fms || !fPs || !dP0 || !dFs || !dQs || !dH || !dR)
        return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if ((sms == nullptr) != (sPs == nullptr)) return set_err(PSSGP_ERR_INVALID, "pkfs_grad: sms and sPs go together");
    cudaStream_t st = (cudaStream_t)stream;
    rc = fused_dispatch(h, dtype, n, d, P0, Fs, Qs, H, R, y, g_ll, fms, fPs, ll, sms, sPs, dP0, dFs, dQs, dH, dR, st);
    if (rc != kNotFused) return rc;
    if (dtype == PSSGP_F64 && mid::supported(d) && !h->force_generic)
        return mid::pkfs_grad_dispatch(d, h, n, (const double*)P0, (const double*)Fs, (const double*)Qs, (const double*)H,
                                       (const double*)R, (const double*)y, (const double*)g_ll, (double*)fms, (double*)fPs,
                                       (double*)ll, (double*)sms, (double*)sPs, (double*)dP0, (double*)dFs, (double*)dQs,
                                       (double*)dH, (double*)dR, st);
    if (dtype == PSSGP_F32 && mid::supported(d) && !h->force_generic)
        return mid::f32_pkfs_grad(h, n, d, (const float*)P0, (const float*)Fs, (const float*)Qs, (const float*)H,
                                  (const float*)R, (const float*)y, (const float*)g_ll, (float*)fms, (float*)fPs, (float*)ll,
                                  (float*)sms, (float*)sPs, (float*)dP0, (float*)dFs, (float*)dQs, (float*)dH, (float*)dR,
                                  false, st);
    // generic state dimension (or no common partition): the three scans one after the other
    if ((rc = pssgp_pkf(h, dtype, n, d, P0, Fs, Qs, H, R, y, nullptr, 1, fms, fPs, ll, nullptr, stream))) return rc;
    if (sms != nullptr)
        if ((rc = pssgp_pks(h, dtype, n, d, Fs, Qs, fms, fPs, 1, nullptr, nullptr, nullptr, sms, sPs, nullptr, stream)))
            return rc;
    return pssgp_pkf_backward(h, dtype, n, d, P0, nullptr, Fs, Qs, H, R, y, fms, fPs, g_ll, 1, nullptr, dP0, dFs, dQs,
                              dH, dR, nullptr, stream);
}

int pssgp_pkf_with_summaries(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs,
                             const void* Qs, const void* H, const void* R, const void* y, const void* m0,
                             int first_special, int last_special, const void* Fnext, const void* Qnext, void* fms,
                             void* fPs, void* ll, void* sm_summary, void* ad_summary, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs || !sm_summary || !ad_summary)
        return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if (!last_special && (!Fnext || !Qnext))
        return set_err(PSSGP_ERR_INVALID, "pkf_with_summaries: Fnext/Qnext required when last_special == 0");
    cudaStream_t st = (cudaStream_t)stream;
    rc = with_summaries_dispatch(h, dtype, n, d, P0, Fs, Qs, H, R, y, m0, first_special, last_special, Fnext, Qnext, fms,
                                 fPs, ll, sm_summary, ad_summary, st);
    if (rc != kNotFused) return rc;
    // generic state dimension (or no common partition): filter, then the two summaries from their own reduce passes
    if (h->fold_count[KIND_FILTER] > 0) {
        // a registered filter fold: done here explicitly; the folded state replaces (m0, P0)
        const int cnt = h->fold_count[KIND_FILTER];
        const void* sums = h->fold_ptr[KIND_FILTER];
        const long stride = (long)h->fold_stride[KIND_FILTER];
        void* out = h->fold_state_out;
        h->fold_count[KIND_FILTER] = 0;
        h->fold_ptr[KIND_FILTER] = nullptr;
        const size_t esz = dtype == PSSGP_F64 ? 8 : 4;
        if (!out) {
            if ((rc = ws_reserve(h, WS_GEN3, esz * (size_t)(d + d * d)))) return rc;
            out = h->buf[WS_GEN3];
        }
        if ((rc = filter_fold_strided_dispatch(h, dtype, d, cnt, P0, m0, sums, stride, out, st))) return rc;
        m0 = out;
        P0 = (const char*)out + esz * (size_t)d;
    }
    if ((rc = pssgp_pkf(h, dtype, n, d, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, nullptr, stream))) return rc;
    if ((rc = pssgp_pks_summary(h, dtype, n, d, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, sm_summary, stream)))
        return rc;
    return pssgp_pkf_backward_summary(h, dtype, n, d, P0, m0, Fs, Qs, H, R, y, fms, fPs, first_special, ad_summary, stream);
}

int pssgp_shard_forward(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                        const void* H, const void* R, const void* y, const void* m0, int first_special, void* fms,
                        void* fPs, void* ll, void* rev_summary, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs || !rev_summary)
        return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if (dtype != PSSGP_F64 || !mid::supported(d))
        return set_err(PSSGP_ERR_UNSUPPORTED, "shard_forward: FP64 and 5 <= d <= 32 (d = %d); d <= 4 uses "
                                              "pssgp_pkf_with_summaries", d);
    return mid::shard_forward_dispatch(d, h, n, (const double*)P0, (const double*)Fs, (const double*)Qs, (const double*)H,
                                       (const double*)R, (const double*)y, (const double*)m0, first_special, (double*)fms,
                                       (double*)fPs, (double*)ll, (double*)rev_summary, (cudaStream_t)stream);
}

int pssgp_rev_fold(pssgp_handle* h, int dtype, int d, int nshards_after, const void* summaries, int64_t stride,
                   void* state_out, void* stream) {
    int rc = check_common(h, dtype, 1, d);
    if (rc) return rc;
    if (!state_out || nshards_after < 1 || !summaries) return set_err(PSSGP_ERR_INVALID, "rev_fold: bad argument");
    if (dtype != PSSGP_F64 || !mid::supported(d))
        return set_err(PSSGP_ERR_UNSUPPORTED, "rev_fold: FP64 and 5 <= d <= 32 (d = %d)", d);
    return mid::rev_fold_dispatch(d, h, (const double*)summaries, nshards_after, stride, (double*)state_out,
                                  (cudaStream_t)stream);
}

int pssgp_shard_reverse(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* m0, const void* Fs,
                        const void* Qs, const void* H, const void* R, const void* y, const void* fms, const void* fPs,
                        const void* g_ll, int first_special, const void* rev_init, void* sms, void* sPs, void* dP0,
                        void* dFs, void* dQs, void* dH, void* dR, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if ((!sms || !sPs) && !dFs) return set_err(PSSGP_ERR_INVALID, "shard_reverse: nothing to compute");
    if (dFs && (!g_ll || !dQs || !dH || !dR)) return set_err(PSSGP_ERR_INVALID, "shard_reverse: null gradient argument");
    if (dtype != PSSGP_F64 || !mid::supported(d))
        return set_err(PSSGP_ERR_UNSUPPORTED, "shard_reverse: FP64 and 5 <= d <= 32 (d = %d)", d);
    return mid::shard_reverse_dispatch(d, h, n, (const double*)P0, (const double*)m0, (const double*)Fs, (const double*)Qs,
                                       (const double*)H, (const double*)R, (const double*)y, (const double*)fms,
                                       (const double*)fPs, (const double*)g_ll, first_special, (const double*)rev_init,
                                       (double*)sms, (double*)sPs, (double*)dP0, (double*)dFs, (double*)dQs, (double*)dH,
                                       (double*)dR, (cudaStream_t)stream);
}

int pssgp_pkfs(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
               const void* H, const void* R, const void* y, void* fms, void* fPs, void* ll, void* sms, void* sPs,
               void* proj, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if (!proj && (!sms || !sPs)) return set_err(PSSGP_ERR_INVALID, "pkfs: sms and sPs, or proj, must be given");
    cudaStream_t st = (cudaStream_t)stream;
    rc = pkfs_dispatch(h, dtype, n, d, P0, Fs, Qs, H, R, y, fms, fPs, ll, sms, sPs, proj, st);
    if (rc != kNotFused) return rc;
    if (dtype == PSSGP_F64 && mid::supported(d) && !h->force_generic && (!proj || (d <= 24 && !h->mid_smem))) {
        // proj: the fragment-resident reverse kernel emits (H sm, H sP H^T) per step instead of sms / sPs
        h->mid_proj = proj;
        rc = mid::pkfs_grad_dispatch(d, h, n, (const double*)P0, (const double*)Fs, (const double*)Qs, (const double*)H,
                                     (const double*)R, (const double*)y, nullptr, (double*)fms, (double*)fPs, (double*)ll,
                                     proj ? nullptr : (double*)sms, proj ? nullptr : (double*)sPs, nullptr, nullptr, nullptr,
                                     nullptr, nullptr, st);
        h->mid_proj = nullptr;
        return rc;
    }
    if (!proj && dtype == PSSGP_F32 && mid::supported(d) && !h->force_generic)
        return mid::f32_pkfs_grad(h, n, d, (const float*)P0, (const float*)Fs, (const float*)Qs, (const float*)H,
                                  (const float*)R, (const float*)y, nullptr, (float*)fms, (float*)fPs, (float*)ll, (float*)sms,
                                  (float*)sPs, nullptr, nullptr, nullptr, nullptr, nullptr, false, st);
    if (proj) return set_err(PSSGP_ERR_UNSUPPORTED, "pkfs: projected output is implemented for the fused d <= 4 kernels and the "
                                                    "FP64 fragment-resident kernels (5 <= d <= 24); d = %d: pass sms / sPs", d);
    if ((rc = pssgp_pkf(h, dtype, n, d, P0, Fs, Qs, H, R, y, nullptr, 1, fms, fPs, ll, nullptr, stream))) return rc;
    return pssgp_pks(h, dtype, n, d, Fs, Qs, fms, fPs, 1, nullptr, nullptr, nullptr, sms, sPs, nullptr, stream);
}

}  // extern "C"
