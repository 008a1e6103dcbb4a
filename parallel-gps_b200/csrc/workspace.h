// Handle = reusable device workspace + options.  Internal header (not part of the C ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>

// chunk-aggregate slots exist once per scan kind (filter / smoother / adjoint) so that several
// *_summary calls can be pending at the same time.
enum { WS_LANE = 0, WS_WAGG = 3, WS_WEXCL = 6, WS_WSTATE = 9, WS_PART, WS_MISC, WS_GEN0, WS_GEN1, WS_GEN2, WS_GEN3, WS_WSTATE_S, WS_WPREFIX, WS_WPREFIX1, WS_WPREFIX2, WS_COUNT };
enum { KIND_FILTER = 0, KIND_SMOOTHER = 1, KIND_ADJOINT = 2 };

struct pssgp_handle {
    int device;
    int num_sms;
    void* buf[WS_COUNT];
    size_t cap[WS_COUNT];
    unsigned int* ticket;
    int64_t chunk_opt;
    int pdl;            // option "pdl": programmatic dependent launch between the kernels of pkfs_grad
    int fused_reverse;  // option "fused_reverse": pkfs_grad runs smoother + adjoint recursions in one kernel
    int64_t launches;
    // chunk aggregates left in the workspace by a *_summary call (time sharding)
    const void* pending_key[3];
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
