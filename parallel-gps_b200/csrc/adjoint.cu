// C ABI: pssgp_pkf_backward, pssgp_pkf_backward_summary, pssgp_adjoint_fold (see include/pssgp_b200.h).
#include "adjoint_small.cuh"
#include "scan_run.cuh"

namespace pssgp {

template <typename T, int D>
typename AdjointAlg<T, D>::Params adjoint_params(int64_t n, const void* P0, const void* m0, const void* Fs,
                                                 const void* Qs, const void* H, const void* R, const void* y,
                                                 const void* fms, const void* fPs, const void* g_ll,
                                                 int first_special, const void* init, void* dP0, void* dFs, void* dQs,
                                                 void* dH, void* dR) {
    typename AdjointAlg<T, D>::Params p;
    p.Fs = (const T*)Fs;
    p.Qs = (const T*)Qs;
    p.y = (const T*)y;
    p.H = (const T*)H;
    p.R = (const T*)R;
    p.P0 = (const T*)P0;
    p.m0 = (const T*)m0;
    p.fms = (const T*)fms;
    p.fPs = (const T*)fPs;
    p.g = (const T*)g_ll;
    p.init = (const T*)init;
    p.dFs = (T*)dFs;
    p.dQs = (T*)dQs;
    p.dP0 = (T*)dP0;
    p.dH = (T*)dH;
    p.dR = (T*)dR;
    p.first_state = nullptr;
    p.n = n;
    p.first_special = first_special;
    return p;
}

template <typename T, int D>
int pkf_bwd_impl(pssgp_handle* h, int64_t n, const void* P0, const void* m0, const void* Fs, const void* Qs,
                 const void* H, const void* R, const void* y, const void* fms, const void* fPs, const void* g_ll,
                 int first_special, const void* init, void* dP0, void* dFs, void* dQs, void* dH, void* dR,
                 void* first_state, cudaStream_t st) {
    auto p = adjoint_params<T, D>(n, P0, m0, Fs, Qs, H, R, y, fms, fPs, g_ll, first_special, init, dP0, dFs, dQs, dH, dR);
    // summaries registered with pssgp_set_fold: folded onto the initial state inside the scan's own kernels
    p.fold = (const T*)h->fold_ptr[KIND_ADJOINT];
    p.fold_count = h->fold_count[KIND_ADJOINT];
    p.fold_stride = (long)h->fold_stride[KIND_ADJOINT];
    fold_clear(h, KIND_ADJOINT);
    return run_scan<AdjointAlg<T, D>>(h, p, n, (T*)dR, (T*)first_state, st, SCAN_FULL, nullptr,
                                      dFs ? adjoint_sig(sizeof(T), D, n, Fs, Qs, y, H, R, fms, fPs, first_special) : 0);
}

template <typename T, int D>
int pkf_bwd_summary_impl(pssgp_handle* h, int64_t n, const void* P0, const void* m0, const void* Fs, const void* Qs,
                         const void* H, const void* R, const void* y, const void* fms, const void* fPs,
                         int first_special, void* summary, cudaStream_t st) {
    auto p = adjoint_params<T, D>(n, P0, m0, Fs, Qs, H, R, y, fms, fPs, nullptr, first_special, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, nullptr);
    return run_scan<AdjointAlg<T, D>>(h, p, n, nullptr, nullptr, st, SCAN_SUMMARY, (T*)summary,
                                      adjoint_sig(sizeof(T), D, n, Fs, Qs, y, H, R, fms, fPs, first_special));
}

template <typename T, int D>
int adjoint_fold_impl(pssgp_handle* h, int count, const void* summaries, void* state_out, cudaStream_t st) {
    auto p = adjoint_params<T, D>(0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    const long NA = AdjointAlg<T, D>::NAGG;
    return run_fold<AdjointAlg<T, D>>(h, p, (const T*)summaries + (long)(count - 1) * NA, count, -NA, (T*)state_out, st);
}

}  // namespace pssgp

using namespace pssgp;

extern "C" {

int pssgp_pkf_backward(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* m0, const void* Fs,
                       const void* Qs, const void* H, const void* R, const void* y, const void* fms, const void* fPs,
                       const void* g_ll, int first_special, const void* adj_init, void* dP0, void* dFs, void* dQs,
                       void* dH, void* dR, void* adj_first, void* stream) {
    int rc = check_common(h, dtype, n, d);
    const bool have_fold = h && h->fold_count[KIND_ADJOINT] > 0;
    const bool null_arg = !P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs || !g_ll || !dFs || !dQs || !dH || !dR;
    if (rc || null_arg || (have_fold && d > 4)) {
        // a registered fold is consumed by this call whether it succeeds or not (never left armed for a later scan)
        if (h) fold_clear(h, KIND_ADJOINT);
        if (rc) return rc;
        if (null_arg) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
        return set_err(PSSGP_ERR_UNSUPPORTED, "pssgp_set_fold is implemented for d <= 4: use pssgp_adjoint_fold");
    }
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pkf_bwd_impl, h, n, P0, m0, Fs, Qs, H, R, y, fms, fPs, g_ll, first_special, adj_init, dP0, dFs, dQs,
                   dH, dR, adj_first, st);
    return pkf_bwd_generic(h, dtype, n, d, P0, m0, Fs, Qs, H, R, y, fms, fPs, g_ll, first_special, adj_init, dP0, dFs,
                           dQs, dH, dR, adj_first, nullptr, st);
}

int pssgp_pkf_backward_summary(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* m0,
                               const void* Fs, const void* Qs, const void* H, const void* R, const void* y,
                               const void* fms, const void* fPs, int first_special, void* summary, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs || !summary)
        return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pkf_bwd_summary_impl, h, n, P0, m0, Fs, Qs, H, R, y, fms, fPs, first_special, summary, st);
    return pkf_bwd_generic(h, dtype, n, d, P0, m0, Fs, Qs, H, R, y, fms, fPs, nullptr, first_special, nullptr, nullptr,
                           nullptr, nullptr, nullptr, nullptr, nullptr, summary, st);
}

int pssgp_adjoint_fold(pssgp_handle* h, int dtype, int d, int nshards_after, const void* summaries, void* state_out,
                       void* stream) {
    int rc = check_common(h, dtype, 1, d);
    if (rc) return rc;
    if (!state_out || nshards_after < 1 || !summaries) return set_err(PSSGP_ERR_INVALID, "adjoint_fold: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(adjoint_fold_impl, h, nshards_after, summaries, state_out, st);
    return adjoint_fold_generic(h, dtype, d, nshards_after, summaries, state_out, st);
}

}  // extern "C"
