// Handle = reusable device workspace + options.  Internal header (not part of the C ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>

// chunk-aggregate slots exist once per scan kind (filter / smoother / adjoint) so that several
// *_summary calls can be pending at the same time.
enum { WS_LANE = 0, WS_WAGG = 3, WS_WEXCL = 6, WS_WSTATE = 9, WS_PART, WS_MISC, WS_GEN0, WS_GEN1, WS_GEN2, WS_GEN3, WS_WSTATE_S, WS_WPREFIX, WS_WPREFIX1, WS_WPREFIX2, WS_F32, WS_GRID, WS_COUNT };
enum { KIND_FILTER = 0, KIND_SMOOTHER = 1, KIND_ADJOINT = 2 };

struct pssgp_handle {
    int device;
    int num_sms;
    void* buf[WS_COUNT];
    size_t cap[WS_COUNT];
    unsigned int* ticket;
    int64_t chunk_opt;
    int pdl;            // option "pdl": programmatic dependent launch between the kernels of pkfs_grad
    int mid_warps;      // option "mid_warps": cap on the warps per CTA of the fragment-resident d > 4 kernels (0 = default)
    int mid_smem;       // option "mid_smem": 5 <= d <= 16 runs the shared-memory tile kernels instead of the fragment-resident ones
    int force_generic;  // option "force_generic": d > 4 runs the CTA-cooperative kernels of generic.cu (tuning / tests)
    int grid_lanes;     // option "grid_lanes": concurrent settings in pssgp_grid_loglik (0 = default 4, at most 8)
    int fused_reverse;  // option "fused_reverse": pkfs_grad runs smoother + adjoint recursions in one kernel
    void* mid_proj;     // internal: projected output [n,2] requested by pssgp_pkfs for its next mid::pkfs_grad call
    int64_t launches;
    // chunk aggregates left in the workspace by a *_summary call (time sharding)
    // pending_key = signature (dtype size, d, n, every input pointer, the first/last flags) of the call that built
    // them, 0 = none; a consumer must present the same signature.  Calls that overwrite arrays the aggregates were
    // built from (fms / fPs, Fs / Qs) clear the keys of the kinds concerned on entry.
    uint64_t pending_key[3];
    int64_t pending_n[3];
    int pending_L[3];
    // the pending aggregates come with the exclusive prefix aggregate of every CTA (WS_WPREFIX + kind): the full
    // call needs no scan over the CTA totals, only state o prefix in the prologue of its apply kernel
    int pending_prefix[3];
    // shard summaries to fold onto the initial state of the next scan of each kind (pssgp_set_fold): the fold then
    // runs inside that scan's kernels instead of in a launch of its own
    const void* fold_ptr[3];
    int fold_count[3];
    int64_t fold_stride[3];
    void* fold_state_out;  // kind 0 only: where the folded state entering the shard is written
    // optional per-kernel CUDA-event timing (option "timing" = 1)
    // pssgp_grid_loglik: child handles (own workspace), streams and events of the concurrent lanes, created lazily
    pssgp_handle* lane[8];
    void* lane_stream[8];
    void* lane_event[8];
    void* fork_event;
    int timing;
    int n_rec, cap_rec;
    struct pssgp_timing_rec* recs;
};

struct pssgp_timing_rec {
    const char* name;
    void* ev0;
    void* ev1;
};

namespace pssgp {
inline uint64_t sig_val(const void* p) { return (uint64_t)(uintptr_t)p; }
inline uint64_t sig_val(long long v) { return (uint64_t)v; }
inline uint64_t sig_val(long v) { return (uint64_t)v; }
inline uint64_t sig_val(int v) { return (uint64_t)(long long)v; }
inline uint64_t sig_val(unsigned long v) { return (uint64_t)v; }
template <typename... A> inline uint64_t make_sig(A... a) {
    uint64_t s = 0x243f6a8885a308d3ull;
    ((s ^= sig_val(a) + 0x9e3779b97f4a7c15ull + (s << 6) + (s >> 2)), ...);
    return s ? s : 1;
}
// signatures of the three kinds of chunk aggregates (same argument lists on the producing and the consuming side)
inline uint64_t filter_sig(size_t esz, int d, long long n, const void* Fs, const void* Qs, const void* y, const void* H,
                           const void* R, int first_special) {
    return make_sig(0, esz, d, n, Fs, Qs, y, H, R, first_special != 0);
}
inline uint64_t smoother_sig(size_t esz, int d, long long n, const void* Fs, const void* Qs, const void* fms,
                             const void* fPs, int last_special, const void* Fnext, const void* Qnext) {
    return make_sig(1, esz, d, n, Fs, Qs, fms, fPs, last_special != 0, last_special ? nullptr : Fnext,
                    last_special ? nullptr : Qnext);
}
inline uint64_t adjoint_sig(size_t esz, int d, long long n, const void* Fs, const void* Qs, const void* y, const void* H,
                            const void* R, const void* fms, const void* fPs, int first_special) {
    return make_sig(2, esz, d, n, Fs, Qs, y, H, R, fms, fPs, first_special != 0);
}
inline void pending_clear(pssgp_handle* h, int kind) {
    h->pending_key[kind] = 0;
    h->pending_prefix[kind] = 0;
}
inline void fold_clear(pssgp_handle* h, int kind) {
    h->fold_ptr[kind] = nullptr;
    h->fold_count[kind] = 0;
}
int set_err(int code, const char* fmt, ...);
int ws_reserve(pssgp_handle* h, int slot, size_t bytes);
int check_launch(pssgp_handle* h, const char* what, int nlaunches);
void timing_begin(pssgp_handle* h, const char* name, void* stream);
void timing_end(pssgp_handle* h, void* stream);
}  // namespace pssgp

// Launch wrapper: PSSGP_LAUNCH(h, "name", stream, kernel<<<grid, block, smem, stream>>>(args...));
#define PSSGP_LAUNCH(h, name, st, ...)                  \
    do {                                                \
        if ((h)->timing) pssgp::timing_begin((h), (name), (void*)(st)); \
        __VA_ARGS__;                                    \
        if ((h)->timing) pssgp::timing_end((h), (void*)(st));           \
    } while (0)
