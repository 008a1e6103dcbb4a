// C ABI of pssgp_b200 (see include/pssgp_b200.h).  Dispatches on (dtype, d) to the templated
// kernels; owns the reusable device workspace.
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "adjoint_small.cuh"
#include "filter_small.cuh"
#include "scan_small.cuh"
#include "smoother_small.cuh"
#include "workspace.h"

namespace pssgp {

thread_local char g_err[512] = "";

int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int ws_reserve(pssgp_handle* h, int slot, size_t bytes) {
    if (bytes <= h->cap[slot]) return PSSGP_OK;
    if (h->buf[slot]) {
        cudaError_t e = cudaFree(h->buf[slot]);
        if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaFree: %s", cudaGetErrorString(e));
        h->buf[slot] = nullptr;
        h->cap[slot] = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&h->buf[slot], want);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    h->cap[slot] = want;
    return PSSGP_OK;
}

int check_launch(pssgp_handle* h, const char* what, int nlaunches) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    h->launches += nlaunches;
    return PSSGP_OK;
}

int pick_chunk(const pssgp_handle* h, int64_t n) {
    if (h->chunk_opt > 0) return (int)h->chunk_opt;
    // enough chunks to give every SM ~16 warps, but never shorter than 8 steps (amortises the
    // generic-operator warp scan) nor longer than 64.
    int64_t target_chunks = (int64_t)h->num_sms * 16 * 32;
    int64_t L = n / (target_chunks > 0 ? target_chunks : 1);
    int c = 8;
    while (c < 64 && c * 2 <= L) c *= 2;
    return c;
}

// Runs K1/K2/K3 for an algebra.  summary != nullptr: only K1 + total (for time sharding).
template <typename Alg>
int run_scan(pssgp_handle* h, typename Alg::Params p, int64_t n, typename Alg::scalar* acc_out,
             typename Alg::scalar* final_state, cudaStream_t st) {
    using T = typename Alg::scalar;
    const int L = pick_chunk(h, n);
    const int64_t nChunks = (n + L - 1) / L;
    const int64_t nBlocks = (nChunks + kReduceThreads - 1) / kReduceThreads;
    const int64_t nW = nBlocks;  // one aggregate per CTA of K1
    const int64_t nChunksPad = nBlocks * kReduceThreads;
    int rc;
    if ((rc = ws_reserve(h, WS_LANE, sizeof(T) * Alg::NAGG * (size_t)nChunksPad))) return rc;
    if ((rc = ws_reserve(h, WS_WAGG, sizeof(T) * Alg::NAGG * (size_t)nW))) return rc;
    if ((rc = ws_reserve(h, WS_WSTATE, sizeof(T) * Alg::NSTATE * (size_t)nW))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * (Alg::NACC > 0 ? Alg::NACC : 1) * (size_t)nBlocks))) return rc;
    T* lane = (T*)h->buf[WS_LANE];
    T* wagg = (T*)h->buf[WS_WAGG];
    T* wstate = (T*)h->buf[WS_WSTATE];
    T* part = (T*)h->buf[WS_PART];
    scan_reduce_kernel<Alg><<<(unsigned)nBlocks, kReduceThreads, 0, st>>>(p, n, L, nChunksPad, lane, wagg, nW);
    int midThreads = kMidThreads;
    if (nW < kMidThreads) midThreads = (int)(((nW + 31) / 32) * 32);
    if (midThreads < 32) midThreads = 32;
    scan_mid_kernel<Alg><<<1, midThreads, 0, st>>>(p, wagg, nW, wstate, final_state);
    scan_apply_kernel<Alg><<<(unsigned)nBlocks, kReduceThreads, 0, st>>>(p, n, L, nChunksPad, lane, wstate, nW,
                                                                         part, h->ticket, acc_out);
    return check_launch(h, "scan", 3);
}

template <typename T, int D>
int pkf_impl(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H,
             const void* R, const void* y, const void* m0, int first_special, void* fms, void* fPs, void* ll,
             void* final_state, cudaStream_t st) {
    using Alg = FilterAlg<T, D>;
    typename Alg::Params p;
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
    return run_scan<Alg>(h, p, n, (T*)ll, (T*)final_state, st);
}

template <typename T, int D>
int pks_impl(pssgp_handle* h, int64_t n, const void* Fs, const void* Qs, const void* fms, const void* fPs,
             int last_special, const void* Fnext, const void* Qnext, const void* init, void* sms, void* sPs,
             void* first_state, cudaStream_t st) {
    using Alg = SmootherAlg<T, D>;
    typename Alg::Params p;
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
    return run_scan<Alg>(h, p, n, nullptr, (T*)first_state, st);
}

template <typename T, int D>
int pkf_bwd_impl(pssgp_handle* h, int64_t n, const void* P0, const void* Fs, const void* Qs, const void* H,
                 const void* R, const void* y, const void* fms, const void* fPs, const void* g_ll, void* dP0,
                 void* dFs, void* dQs, void* dH, void* dR, cudaStream_t st) {
    using Alg = AdjointAlg<T, D>;
    typename Alg::Params p;
    p.Fs = (const T*)Fs;
    p.Qs = (const T*)Qs;
    p.y = (const T*)y;
    p.H = (const T*)H;
    p.R = (const T*)R;
    p.P0 = (const T*)P0;
    p.m0 = nullptr;
    p.fms = (const T*)fms;
    p.fPs = (const T*)fPs;
    p.g = (const T*)g_ll;
    p.init = nullptr;
    p.dFs = (T*)dFs;
    p.dQs = (T*)dQs;
    p.dP0 = (T*)dP0;
    p.dH = (T*)dH;
    p.dR = (T*)dR;
    p.first_state = nullptr;
    p.n = n;
    p.first_special = 1;
    return run_scan<Alg>(h, p, n, (T*)dR, nullptr, st);
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

static int check_common(pssgp_handle* h, int dtype, int64_t n, int d) {
    if (!h) return set_err(PSSGP_ERR_INVALID, "null handle");
    if (dtype != PSSGP_F64 && dtype != PSSGP_F32) return set_err(PSSGP_ERR_INVALID, "bad dtype %d", dtype);
    if (n < 1) return set_err(PSSGP_ERR_INVALID, "n must be >= 1 (got %lld)", (long long)n);
    if (d < 1) return set_err(PSSGP_ERR_INVALID, "d must be >= 1 (got %d)", d);
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
    return PSSGP_OK;
}

extern "C" {

int pssgp_version(void) { return 100; }

const char* pssgp_last_error(void) { return g_err; }

int pssgp_create(pssgp_handle** out, int device) {
    if (!out) return set_err(PSSGP_ERR_INVALID, "null out");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    pssgp_handle* h = new (std::nothrow) pssgp_handle();
    if (!h) return set_err(PSSGP_ERR_INVALID, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->device = device;
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        delete h;
        return set_err(PSSGP_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    }
    h->num_sms = sms;
    e = cudaMalloc((void**)&h->ticket, 64);
    if (e == cudaSuccess) e = cudaMemset(h->ticket, 0, 64);
    if (e != cudaSuccess) {
        delete h;
        return set_err(PSSGP_ERR_CUDA, "workspace init: %s", cudaGetErrorString(e));
    }
    *out = h;
    return PSSGP_OK;
}

int pssgp_destroy(pssgp_handle* h) {
    if (!h) return PSSGP_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < WS_COUNT; ++i)
        if (h->buf[i]) cudaFree(h->buf[i]);
    if (h->ticket) cudaFree(h->ticket);
    delete h;
    return PSSGP_OK;
}

int pssgp_set_option(pssgp_handle* h, const char* name, int64_t value) {
    if (!h || !name) return set_err(PSSGP_ERR_INVALID, "null argument");
    if (strcmp(name, "chunk") == 0) {
        if (value < 0 || value > 4096) return set_err(PSSGP_ERR_INVALID, "chunk out of range");
        h->chunk_opt = value;
        return PSSGP_OK;
    }
    return set_err(PSSGP_ERR_INVALID, "unknown option '%s'", name);
}

int64_t pssgp_launch_count(const pssgp_handle* h) { return h ? h->launches : 0; }

int pssgp_pkf(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
              const void* H, const void* R, const void* y, const void* m0, int first_special, void* fms, void* fPs,
              void* ll, void* final_state, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pkf_impl, h, n, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, final_state, st);
    return set_err(PSSGP_ERR_UNSUPPORTED, "pkf: state dimension %d not supported yet", d);
}

int pssgp_pks(pssgp_handle* h, int dtype, int64_t n, int d, const void* Fs, const void* Qs, const void* fms,
              const void* fPs, int last_special, const void* Fnext, const void* Qnext, const void* init, void* sms,
              void* sPs, void* first_state, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!Fs || !Qs || !fms || !fPs || !sms || !sPs) return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    if (!last_special && (!Fnext || !Qnext || !init))
        return set_err(PSSGP_ERR_INVALID, "pks: Fnext/Qnext/init required when last_special == 0");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pks_impl, h, n, Fs, Qs, fms, fPs, last_special, Fnext, Qnext, init, sms, sPs, first_state, st);
    return set_err(PSSGP_ERR_UNSUPPORTED, "pks: state dimension %d not supported yet", d);
}

}  // extern "C"

// ---- entry points that are implemented in other translation units are declared there; the
// ---- ones below are placeholders until their kernels land (they fail loudly, never fall back).
extern "C" {
#define PSSGP_TODO(name) return set_err(PSSGP_ERR_UNSUPPORTED, name ": not implemented yet")
#define PSSGP_HAVE_DISCRETISE
#ifndef PSSGP_HAVE_DISCRETISE
int pssgp_discretise(pssgp_handle*, int, int64_t, int, const void*, const void*, const void*, void*, void*, void*) {
    PSSGP_TODO("pssgp_discretise");
}
#endif
#ifndef PSSGP_HAVE_SHARDED
int pssgp_pkf_summary(pssgp_handle*, int, int64_t, int, const void*, const void*, const void*, const void*,
                      const void*, const void*, int, void*, void*) {
    PSSGP_TODO("pssgp_pkf_summary");
}
int pssgp_filter_fold(pssgp_handle*, int, int, int, const void*, const void*, const void*, void*, void*) {
    PSSGP_TODO("pssgp_filter_fold");
}
int pssgp_pks_summary(pssgp_handle*, int, int64_t, int, const void*, const void*, const void*, const void*, int,
                      const void*, const void*, void*, void*) {
    PSSGP_TODO("pssgp_pks_summary");
}
int pssgp_smoother_fold(pssgp_handle*, int, int, int, const void*, void*, void*) {
    PSSGP_TODO("pssgp_smoother_fold");
}
#endif
#ifndef PSSGP_HAVE_BACKWARD
int pssgp_pkf_backward(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                       const void* H, const void* R, const void* y, const void* fms, const void* fPs, const void* g_ll,
                       void* dP0, void* dFs, void* dQs, void* dH, void* dR, void* stream) {
    int rc = check_common(h, dtype, n, d);
    if (rc) return rc;
    if (!P0 || !Fs || !Qs || !H || !R || !y || !fms || !fPs || !g_ll || !dP0 || !dFs || !dQs || !dH || !dR)
        return set_err(PSSGP_ERR_INVALID, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_SMALL(pkf_bwd_impl, h, n, P0, Fs, Qs, H, R, y, fms, fPs, g_ll, dP0, dFs, dQs, dH, dR, st);
    return set_err(PSSGP_ERR_UNSUPPORTED, "pkf_backward: state dimension %d not supported yet", d);
}
#endif
}
