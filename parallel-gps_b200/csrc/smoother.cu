// C ABI: pssgp_pks, pssgp_pks_summary, pssgp_smoother_fold (see include/pssgp_b200.h).
#include "smoother_small.cuh"
#include "scan_run.cuh"

namespace pssgp {

template <typename T, int D>
typename SmootherAlg<T, D>::Params smoother_params(int64_t n, const void* Fs, const void* Qs, const void* fms,
                                                   const void* fPs, int last_special, const void* Fnext,
                                                   const void* Qnext, const void* init, void* sms, void* sPs) {
    typename SmootherAlg<T, D>::Params p;
    p.Fs = (const T*)Fs;
    p.Qs = (const T*)Qs;
    p.fms = (const T*)fms;
    p.fPs = (const T*)fPs;
    p.sms = (T*)sms;
    p.sPs = (T*)sPs;
    p.n = n;
    p.last_special = last_special;
    p.Fnext = (const T*)Fnext;
    p.Qnext = (const T*)Qnext;
    p.init = (const T*)init;
    return p;
}

template <typename T, int D>
int pks_impl(pssgp_handle* h, int64_t n, const void* Fs, const void* Qs, const void* fms, const void* fPs,
             int last_special, const void* Fnext, const void* Qnext, const void* init, void* sms, void* sPs,
             void* first_state, cudaStream_t st) {
    auto p = smoother_params<T, D>(n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, init, sms, sPs);
    // summaries registered with pssgp_set_fold: folded onto the initial state inside the scan's own kernels
    p.fold = (const T*)h->fold_ptr[KIND_SMOOTHER];
    p.fold_count = h->fold_count[KIND_SMOOTHER];
    p.fold_stride = (long)h->fold_stride[KIND_SMOOTHER];
    fold_clear(h, KIND_SMOOTHER);
    return run_scan<SmootherAlg<T, D>>(h, p, n, nullptr, (T*)first_state, st, SCAN_FULL, nullptr,
                                       smoother_sig(sizeof(T), D, n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext));
}

template <typename T, int D>
int pks_summary_impl(pssgp_handle* h, int64_t n, const void* Fs, const void* Qs, const void* fms, const void* fPs,
                     int last_special, const void* Fnext, const void* Qnext, void* summary, cudaStream_t st) {
    auto p = smoother_params<T, D>(n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, nullptr, nullptr, nullptr);
    return run_scan<SmootherAlg<T, D>>(h, p, n, nullptr, nullptr, st, SCAN_SUMMARY, (T*)summary,
                                       smoother_sig(sizeof(T), D, n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext));
}

template <typename T, int D>
int smoother_fold_impl(pssgp_handle* h, int count, const void* summaries, void* state_out, cudaStream_t st) {
    // summaries points at the LAST shard's summary; walk towards this shard with a negative stride
    auto p = smoother_params<T, D>(0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    const long NA = SmootherAlg<T, D>::NAGG;
    return run_fold<SmootherAlg<T, D>>(h, p, (const T*)summaries + (long)(count - 1) * NA, count, -NA, (T*)state_out, st);
}

}  // namespace pssgp

using namespace pssgp;

extern "C" {

int pssgp_pks(pssgp_handle* h, int dtype, int64_t n, int d, const void* Fs, const void* Qs, const void* fms,
              const void* fPs, int last_special, const void* Fnext, const void* Qnext, const void* init, void* sms,
              void* sPs, void* first_state, void* stream) {
    int rc = check_common(h, dtype, n, d);
    const bool have_fold = h && h->fold_count[KIND_SMOOTHER] > 0;
    if (rc || !Fs || !Qs || !fms || !fPs || !sms || !sPs || (!last_special && (!Fnext || !Qnext || (!init && !have_fold))) ||
        (have_fold && d > 4)) {
        // a registered fold is consumed by this call whether it succeeds or not (never left armed for a later scan)
        if (h) fold_clear(h, KIND_SMOOTHER);
        if (rc) return rc;
        if (!Fs || !Qs || !fms || !fPs || !sms || !sPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
        if (have_fold && d > 4)
            return set_err(PSSGP_ERR_UNSUPPORTED, "pssgp_set_fold is implemented for d <= 4: use pssgp_smoother_fold");
        return set_err(PSSGP_ERR_INVALID, "pks: Fnext/Qnext and init (or pssgp_set_fold) required when last_special == 0");
    }
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pks_impl, h, n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, init, sms, sPs, first_state, st);
    return pks_generic(h, dtype, n, d, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, init, sms, sPs, first_state,
                       nullptr, st);
}

int pssgp_pks_summary(pssgp_handle* h, int dtype, int64_t n, int d, const void* Fs, const void* Qs, const void* fms,
                      const void* fPs, int last_special, const void* Fnext, const void* Qnext, void* summary,
                      void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!Fs || !Qs || !fms || !fPs || !summary) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if (!last_special && (!Fnext || !Qnext))
        return set_err(PSSGP_ERR_INVALID, "pks_summary: Fnext/Qnext required when last_special == 0");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pks_summary_impl, h, n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, summary, st);
    return pks_generic(h, dtype, n, d, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, nullptr, nullptr, nullptr, nullptr,
                       summary, st);
}

int pssgp_smoother_fold(pssgp_handle* h, int dtype, int d, int nshards_after, const void* summaries, void* state_out,
                        void* stream) {
    int rc = check_common(h, dtype, 1, d);
    if (rc) return rc;
    if (!state_out || nshards_after < 1 || !summaries) return set_err(PSSGP_ERR_INVALID, "smoother_fold: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(smoother_fold_impl, h, nshards_after, summaries, state_out, st);
    return smoother_fold_generic(h, dtype, d, nshards_after, summaries, state_out, st);
}

}  // extern "C"
