// Host-side launcher of the three-kernel chunked scan (scan_stream.cuh, scan_small.cuh) + (dtype, d) dispatch helpers.
#pragma once
#include <cuda_runtime.h>

#include "../../include/pssgp_b200.h"
#include <stdint.h>

#include "scan_small.cuh"
#include "scan_stream.cuh"
#include "workspace.h"

namespace pssgp {

// time steps per thread-chunk: one resident wave of CTAs (threads_per_cta each), multiple of `ls`
int pick_chunk(const pssgp_handle* h, int64_t n, int threads_per_cta, int ls);
int check_common(pssgp_handle* h, int dtype, int64_t n, int d);

// generic state dimension (warp-cooperative path, generic.cu); summary != nullptr selects summary mode
int pkf_generic(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* Fs, const void* Qs,
                const void* H, const void* R, const void* y, const void* m0, int first_special, void* fms, void* fPs,
                void* ll, void* final_state, void* summary, cudaStream_t st);
int filter_fold_generic(pssgp_handle* h, int dtype, int d, int count, const void* P0, const void* m0,
                        const void* summaries, void* state_out, cudaStream_t st);
int pks_generic(pssgp_handle* h, int dtype, int64_t n, int d, const void* Fs, const void* Qs, const void* fms,
                const void* fPs, int last_special, const void* Fnext, const void* Qnext, const void* init, void* sms,
                void* sPs, void* first_state, void* summary, cudaStream_t st);
int smoother_fold_generic(pssgp_handle* h, int dtype, int d, int count, const void* summaries, void* state_out,
                          cudaStream_t st);
int pkf_bwd_generic(pssgp_handle* h, int dtype, int64_t n, int d, const void* P0, const void* m0, const void* Fs,
                    const void* Qs, const void* H, const void* R, const void* y, const void* fms, const void* fPs,
                    const void* g_ll, int first_special, const void* adj_init, void* dP0, void* dFs, void* dQs,
                    void* dH, void* dR, void* adj_first, void* summary, cudaStream_t st);
int adjoint_fold_generic(pssgp_handle* h, int dtype, int d, int count, const void* summaries, void* state_out,
                         cudaStream_t st);

enum ScanMode { SCAN_FULL = 0, SCAN_SUMMARY = 1 };

// Opts the streaming kernels of an algebra into their (> 48 KB) dynamic shared memory, once per process.
template <typename Alg>
int stream_configure(int device) {
    using Lay = StreamLayout<Alg>;
    static unsigned long long done = 0ull;  // one bit per device (the attribute is per device)
    const unsigned long long bit = 1ull << (device & 63);
    if (done & bit) return PSSGP_OK;
    cudaError_t e = cudaFuncSetAttribute(stream_reduce_kernel<Alg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Lay::NW * Lay::WARP_BYTES_REDUCE);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(stream_reduce_kernel<Alg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Lay::NW * Lay::WARP_BYTES_REDUCE);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(stream_apply_kernel<Alg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Lay::NW * Lay::WARP_BYTES_APPLY);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
    done |= bit;
    return PSSGP_OK;
}

// Same for an algebra that only ever runs K3 (the fused algebras of fused_small.cuh).
template <typename Alg>
int stream_configure_apply(int device) {
    using Lay = StreamLayout<Alg>;
    static unsigned long long done = 0ull;
    const unsigned long long bit = 1ull << (device & 63);
    if (done & bit) return PSSGP_OK;
    cudaError_t e = cudaFuncSetAttribute(stream_apply_kernel<Alg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Lay::NW * Lay::WARP_BYTES_APPLY);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
    done |= bit;
    return PSSGP_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Partition of the time axis for CTAs of NW warps (see StreamPart): one resident wave, all SMs but one busy
// with complete chunks of two lengths mixed inside every CTA, at most one short tail CTA.
template <int NW, int LS>
StreamPart make_partition(const pssgp_handle* h, int64_t n) {
    StreamPart sp;
    sp.n = n;
    const int64_t cta_chunks = (int64_t)NW * 32;
    const int64_t C = h->num_sms > 1 ? h->num_sms - 1 : 1;  // main CTAs of the single wave (+ 1 tail CTA)
    int64_t tail;
    if (h->chunk_opt <= 0 && n >= C * cta_chunks * 8) {
        const int64_t units = n / (32 * LS);               // warp-rows of LS rows each
        const int64_t base_units = units / (C * NW);       // every warp gets at least this many
        const int64_t n_long = units - base_units * C * NW;  // warps with one more: < C * NW
        sp.L = (int)((base_units + 1) * LS);
        sp.wl = (int)(n_long / C);
        sp.cl = (int)(n_long % C);
        sp.nMain = (int)C;
        tail = n - units * 32 * LS;
    } else {
        sp.L = pick_chunk(h, n, NW * 32, LS);
        sp.wl = NW;  // every warp "long"
        sp.cl = 0;
        sp.nMain = (int)(n / (cta_chunks * sp.L));
        tail = n - (int64_t)sp.nMain * cta_chunks * sp.L;
    }
    sp.Ltail = (int)((((tail + cta_chunks - 1) / cta_chunks) + LS - 1) / LS * LS);
    if (sp.Ltail < LS) sp.Ltail = LS;
    sp.nCta = sp.nMain + (tail > 0 ? 1 : 0);
    return sp;
}

// Runs K1/K2/K3 for an algebra.  SCAN_SUMMARY: K1 + total only (shard summary for time sharding); the
// chunk aggregates stay in the workspace and the next SCAN_FULL call with the same key skips K1.
template <typename Alg>
int run_scan(pssgp_handle* h, typename Alg::Params p, int64_t n, typename Alg::scalar* acc_out,
             typename Alg::scalar* final_state, cudaStream_t st, int mode = SCAN_FULL,
             typename Alg::scalar* summary = nullptr, uint64_t key = 0) {
    using T = typename Alg::scalar;
    using Lay = StreamLayout<Alg>;
    constexpr int NW = Lay::NW;
    int rc;
    if ((rc = stream_configure<Alg>(h->device))) return rc;
    for (int a = 0; a < Alg::NIN; ++a)
        if (!aligned16(Alg::in_ptr(p, a))) return set_err(PSSGP_ERR_INVALID, "input array %d is not 16-byte aligned", a);
    if (mode == SCAN_FULL)
        for (int a = 0; a < Alg::NOUT; ++a)
            if (!aligned16(Alg::out_ptr(p, a)))
                return set_err(PSSGP_ERR_INVALID, "output array %d is not 16-byte aligned", a);
    const StreamPart sp = make_partition<NW, Lay::LS>(h, n);
    const int L = sp.L;
    const int64_t nBlocks = sp.nCta;
    const int64_t nW = nBlocks;  // one aggregate per CTA of K1
    const int64_t nChunksPad = nBlocks * NW * 32;
    constexpr int kind = Alg::KIND;
    const bool reuse = (mode == SCAN_FULL && key != 0 && h->pending_key[kind] == key &&
                        h->pending_n[kind] == n && h->pending_L[kind] == L);
    // the pending aggregates came with per-CTA prefix aggregates: no scan over the CTA totals is needed (unless
    // the caller wants the state after the last step, which only that scan produces)
    const bool have_prefix = reuse && h->pending_prefix[kind] && final_state == nullptr;
    pending_clear(h, kind);
    if (!reuse) {
        if ((rc = ws_reserve(h, WS_LANE + kind, sizeof(T) * Alg::NAGG * (size_t)nChunksPad))) return rc;
        if ((rc = ws_reserve(h, WS_WAGG + kind, sizeof(T) * Alg::NAGG * (size_t)nW))) return rc;
        if ((rc = ws_reserve(h, WS_WEXCL + kind, sizeof(T) * Alg::NAGG * (size_t)nW * NW))) return rc;
        if (mode == SCAN_SUMMARY)
            if ((rc = ws_reserve(h, WS_WPREFIX + kind, sizeof(T) * Alg::NAGG * (size_t)nW))) return rc;
    }
    if ((rc = ws_reserve(h, WS_WSTATE, sizeof(T) * Alg::NSTATE * (size_t)nW))) return rc;
    if ((rc = ws_reserve(h, WS_PART, sizeof(T) * (Alg::NACC > 0 ? Alg::NACC : 1) * (size_t)nBlocks))) return rc;
    T* lane = (T*)h->buf[WS_LANE + kind];
    T* wagg = (T*)h->buf[WS_WAGG + kind];
    T* wexcl = (T*)h->buf[WS_WEXCL + kind];
    T* wstate = (T*)h->buf[WS_WSTATE];
    T* wprefix = (T*)h->buf[WS_WPREFIX + kind];
    T* part = (T*)h->buf[WS_PART];
    int nl = 0;
    bool fused_mid = false;
    if (!reuse) {
        // SCAN_FULL: the last CTA turns the CTA totals into states; SCAN_SUMMARY: into prefix aggregates + summary
        if (mode == SCAN_SUMMARY) {
            PSSGP_LAUNCH(h, Alg::name_reduce(), st,
                         (stream_reduce_kernel<Alg, true><<<(unsigned)nBlocks, NW * 32, NW * Lay::WARP_BYTES_REDUCE, st>>>(
                             p, sp, nChunksPad, lane, wexcl, wagg, (T*)nullptr, (T*)nullptr, h->ticket + 1, wprefix,
                             summary)));
        } else {
            PSSGP_LAUNCH(h, Alg::name_reduce(), st,
                         (stream_reduce_kernel<Alg, false><<<(unsigned)nBlocks, NW * 32, NW * Lay::WARP_BYTES_REDUCE, st>>>(
                             p, sp, nChunksPad, lane, wexcl, wagg, wstate, final_state, h->ticket + 1, (T*)nullptr,
                             (T*)nullptr)));
        }
        ++nl;
        fused_mid = (mode == SCAN_FULL);
    }
    int midThreads = kMidThreads;
    if (nW < kMidThreads) midThreads = (int)(((nW + 31) / 32) * 32);
    if (midThreads < 32) midThreads = 32;
    if (mode == SCAN_SUMMARY) {
        h->pending_key[kind] = key;
        h->pending_n[kind] = n;
        h->pending_L[kind] = L;
        h->pending_prefix[kind] = 1;
        return check_launch(h, "scan summary", nl);
    }
    if (!fused_mid && !have_prefix) {
        PSSGP_LAUNCH(h, Alg::name_mid(), st, (scan_mid_kernel<Alg><<<1, midThreads, 0, st>>>(p, wagg, nW, wstate, final_state)));
        ++nl;
    }
    PSSGP_LAUNCH(h, Alg::name_apply(), st,
                 (stream_apply_kernel<Alg><<<(unsigned)nBlocks, NW * 32, NW * Lay::WARP_BYTES_APPLY, st>>>(
                     p, sp, nChunksPad, lane, wexcl, wstate, part, h->ticket, acc_out,
                     have_prefix ? (const T*)wprefix : (const T*)nullptr)));
    return check_launch(h, "scan", nl + 1);
}

template <typename Alg>
int run_fold(pssgp_handle* h, typename Alg::Params p, const typename Alg::scalar* summaries, int count, long stride,
             typename Alg::scalar* out, cudaStream_t st) {
    PSSGP_LAUNCH(h, "fold", st, (scan_fold_kernel<Alg><<<1, 32, 0, st>>>(p, summaries, count, stride, out)));
    return check_launch(h, "fold", 1);
}

}  // namespace pssgp

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

