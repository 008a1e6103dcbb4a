// C ABI: pssgp_pkf, pssgp_pkf_summary, pssgp_filter_fold (see include/pssgp_b200.h).
#include "filter_small.cuh"
#include "scan_run.cuh"

namespace pssgp {

template <typename T, int D>
typename FilterAlg<T, D>::Params filter_params(const void* P0, const void* Fs, const void* Qs, const void* H,
                                               const void* R, const void* y, const void* m0, int first_special,
                                               void* fms, void* fPs) {
    typename FilterAlg<T, D>::Params p;
    p.Fs = (const T*)Fs;
    p.Qs = (const T*)Qs;
    p.y = (const T*)y;
    p.H = (const T*)H;
    p.R = (const T*)R;
    p.P0 = (const T*)P0;
    p.m0 = (const T*)m0;
    p.fms = (T*)fms;
    p.fPs = (T*)fPs;
    p.first_special = first_special;
    return p;
}

template <typename T, int D>
int pkf_impl(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H,
             const void* R, const void* y, const void* m0, int first_special, void* fms, void* fPs, void* ll,
             void* final_state, cudaStream_t st) {
    auto p = filter_params<T, D>(P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs);
    return run_scan<FilterAlg<T, D>>(h, p, n, (T*)ll, (T*)final_state, st, SCAN_FULL, nullptr,
                                     filter_sig(sizeof(T), D, n, Fs, Qs, y, H, R, first_special));
}

template <typename T, int D>
int pkf_summary_impl(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H,
                     const void* R, const void* y, int first_special, void* summary, cudaStream_t st) {
    auto p = filter_params<T, D>(P0, Fs, Qs, H, R, y, nullptr, first_special, nullptr, nullptr);
    return run_scan<FilterAlg<T, D>>(h, p, n, nullptr, nullptr, st, SCAN_SUMMARY, (T*)summary,
                                     filter_sig(sizeof(T), D, n, Fs, Qs, y, H, R, first_special));
}

template <typename T, int D>
int filter_fold_impl(pssgp_handle* h, int count, const void* P0, const void* m0, const void* summaries,
                     void* state_out, cudaStream_t st) {
    auto p = filter_params<T, D>(P0, nullptr, nullptr, nullptr, nullptr, nullptr, m0, 0, nullptr, nullptr);
    return run_fold<FilterAlg<T, D>>(h, p, (const T*)summaries, count, FilterAlg<T, D>::NAGG, (T*)state_out, st);
}

}  // namespace pssgp

using namespace pssgp;

extern "C" {

int pssgp_pkf(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
              const void* H, const void* R, const void* y, const void* m0, int first_special, void* fms, void* fPs,
              void* ll, void* final_state, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    // fms / fPs are about to be overwritten: reverse-scan aggregates built from them are stale
    pending_clear(h, KIND_SMOOTHER);
    pending_clear(h, KIND_ADJOINT);
    DISPATCH_SMALL(pkf_impl, h, n, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, final_state, st);
    return pkf_generic(h, dtype, n, d, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, final_state, nullptr, st);
}

int pssgp_pkf_summary(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                      const void* H, const void* R, const void* y, int first_special, void* summary, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!Fs || !Qs || !H || !R || !y || !summary) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pkf_summary_impl, h, n, P0, Fs, Qs, H, R, y, first_special, summary, st);
    return pkf_generic(h, dtype, n, d, P0, Fs, Qs, H, R, y, nullptr, first_special, nullptr, nullptr, nullptr, nullptr,
                       summary, st);
}

int pssgp_filter_fold(pssgp_handle* h, int dtype, int d, int nshards_before, const void* P0, const void* m0,
                      const void* summaries, void* state_out, void* stream) {
    int rc = check_common(h, dtype, 1, d);
    if (rc) return rc;
    if (!P0 || !state_out || nshards_before < 0 || (nshards_before > 0 && !summaries))
        return set_err(PSSGP_ERR_INVALID, "filter_fold: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(filter_fold_impl, h, nshards_before, P0, m0, summaries, state_out, st);
    return filter_fold_generic(h, dtype, d, nshards_before, P0, m0, summaries, state_out, st);
}

}  // extern "C"

#ifdef PSSGP_PHASES
extern "C" int pssgp_debug_phases(unsigned long long* out, int count) {
    return (int)cudaMemcpyFromSymbol(out, pssgp::g_phase, sizeof(unsigned long long) * count);
}
#endif
