// C ABI of pssgp_b200 (see include/pssgp_b200.h).  Dispatches on (dtype, d) to the templated
// kernels; owns the reusable device workspace.
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

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

void timing_begin(pssgp_handle* h, const char* name, void* stream) {
    if (h->n_rec == h->cap_rec) {
        int ncap = h->cap_rec ? h->cap_rec * 2 : 256;
        pssgp_timing_rec* nr = (pssgp_timing_rec*)realloc(h->recs, sizeof(pssgp_timing_rec) * ncap);
        if (!nr) return;
        h->recs = nr;
        h->cap_rec = ncap;
    }
    pssgp_timing_rec& r = h->recs[h->n_rec];
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    r.name = name;
    r.ev0 = e0;
    r.ev1 = e1;
    cudaEventRecord(e0, (cudaStream_t)stream);
}

void timing_end(pssgp_handle* h, void* stream) {
    if (h->n_rec >= h->cap_rec) return;
    cudaEventRecord((cudaEvent_t)h->recs[h->n_rec].ev1, (cudaStream_t)stream);
    h->n_rec++;
}

int pick_chunk(const pssgp_handle* h, int64_t n, int threads_per_cta, int ls) {
    int64_t L;
    if (h->chunk_opt > 0) {
        L = h->chunk_opt;
    } else {
        // one CTA per SM, all resident at once: the grid is a single wave and every thread streams one chunk.
        // Never shorter than 8 steps (amortises the generic-operator warp scan).
        const int64_t threads = (int64_t)(h->num_sms > 0 ? h->num_sms : 1) * threads_per_cta;
        L = (n + threads - 1) / threads;
        if (L < 8) L = 8;
    }
    L = ((L + ls - 1) / ls) * ls;
    if (L > (1 << 22)) L = 1 << 22;  // 32-bit byte strides in the streaming kernels
    return (int)L;
}

}  // namespace pssgp

using namespace pssgp;

namespace pssgp {
int check_common(pssgp_handle* h, int dtype, int64_t n, int d) {
    if (!h) return set_err(PSSGP_ERR_INVALID, "null handle");
    if (dtype != PSSGP_F64 && dtype != PSSGP_F32) return set_err(PSSGP_ERR_INVALID, "bad dtype %d", dtype);
    if (n < 1) return set_err(PSSGP_ERR_INVALID, "n must be >= 1 (got %lld)", (long long)n);
    if (d < 1) return set_err(PSSGP_ERR_INVALID, "d must be >= 1 (got %d)", d);
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
    return PSSGP_OK;
}
}  // namespace pssgp

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
    if (const char* env = getenv("PSSGP_CHUNK")) h->chunk_opt = atoll(env);  // tuning aid: same as option "chunk"
    if (const char* env = getenv("PSSGP_FUSED_REVERSE")) h->fused_reverse = atoi(env) != 0;  // same as the option
    h->pdl = 1;
    if (const char* env = getenv("PSSGP_PDL")) h->pdl = atoi(env) != 0;     // same as option "pdl"
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
    for (int i = 0; i < h->n_rec; ++i) {
        cudaEventDestroy((cudaEvent_t)h->recs[i].ev0);
        cudaEventDestroy((cudaEvent_t)h->recs[i].ev1);
    }
    free(h->recs);
    if (h->ticket) cudaFree(h->ticket);
    for (int i = 0; i < 8; ++i) {
        if (h->lane[i]) pssgp_destroy(h->lane[i]);
        if (h->lane_stream[i]) cudaStreamDestroy((cudaStream_t)h->lane_stream[i]);
        if (h->lane_event[i]) cudaEventDestroy((cudaEvent_t)h->lane_event[i]);
    }
    if (h->fork_event) cudaEventDestroy((cudaEvent_t)h->fork_event);
    delete h;
    return PSSGP_OK;
}

int pssgp_set_option(pssgp_handle* h, const char* name, int64_t value) {
    if (!h || !name) return set_err(PSSGP_ERR_INVALID, "null argument");
    if (strcmp(name, "timing") == 0) {
        h->timing = value != 0;
        return PSSGP_OK;
    }
    if (strcmp(name, "chunk") == 0) {
        if (value < 0 || value > 4096) return set_err(PSSGP_ERR_INVALID, "chunk out of range");
        h->chunk_opt = value;
        return PSSGP_OK;
    }
    if (strcmp(name, "pdl") == 0) {
        h->pdl = value != 0;
        return PSSGP_OK;
    }
    if (strcmp(name, "grid_lanes") == 0) {
        if (value < 0 || value > 8) return set_err(PSSGP_ERR_INVALID, "grid_lanes out of range (0..8)");
        h->grid_lanes = (int)value;
        return PSSGP_OK;
    }
    if (strcmp(name, "mid_warps") == 0) {
        h->mid_warps = (int)value;
        return PSSGP_OK;
    }
    if (strcmp(name, "mid_smem") == 0) {
        h->mid_smem = value != 0;
        return PSSGP_OK;
    }
    if (strcmp(name, "force_generic") == 0) {
        h->force_generic = value != 0;
        return PSSGP_OK;
    }
    if (strcmp(name, "fused_reverse") == 0) {
        h->fused_reverse = value != 0;
        return PSSGP_OK;
    }
    return set_err(PSSGP_ERR_INVALID, "unknown option '%s'", name);
}

int pssgp_set_fold(pssgp_handle* h, int kind, const void* summaries, int count, int64_t stride, void* state_out) {
    if (!h) return set_err(PSSGP_ERR_INVALID, "null handle");
    if (kind < KIND_FILTER || kind > KIND_ADJOINT) return set_err(PSSGP_ERR_INVALID, "set_fold: kind must be 0, 1 or 2");
    if (count < 0 || (count > 0 && !summaries)) return set_err(PSSGP_ERR_INVALID, "set_fold: bad argument");
    h->fold_ptr[kind] = count > 0 ? summaries : nullptr;
    h->fold_count[kind] = count;
    h->fold_stride[kind] = stride;
    if (kind == KIND_FILTER) h->fold_state_out = count > 0 ? state_out : nullptr;
    return PSSGP_OK;
}

int64_t pssgp_launch_count(const pssgp_handle* h) { return h ? h->launches : 0; }

int pssgp_timing_report(pssgp_handle* h, char* buf, int64_t buflen) {
    // "name count total_ms\n" per kernel name, aggregated over all records since the last report; clears them.
    if (!h || !buf || buflen < 2) return set_err(PSSGP_ERR_INVALID, "timing_report: bad argument");
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    const int MAXN = 64;
    const char* names[MAXN];
    double tot[MAXN];
    long cnt[MAXN];
    int nn = 0;
    for (int i = 0; i < h->n_rec; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, (cudaEvent_t)h->recs[i].ev0, (cudaEvent_t)h->recs[i].ev1);
        int j = 0;
        for (; j < nn; ++j)
            if (strcmp(names[j], h->recs[i].name) == 0) break;
        if (j == nn) {
            if (nn == MAXN) continue;
            names[nn] = h->recs[i].name;
            tot[nn] = 0;
            cnt[nn] = 0;
            nn++;
        }
        tot[j] += ms;
        cnt[j] += 1;
        cudaEventDestroy((cudaEvent_t)h->recs[i].ev0);
        cudaEventDestroy((cudaEvent_t)h->recs[i].ev1);
    }
    h->n_rec = 0;
    int64_t off = 0;
    buf[0] = 0;
    for (int j = 0; j < nn; ++j) {
        int w = snprintf(buf + off, (size_t)(buflen - off), "%s %ld %.6f\n", names[j], cnt[j], tot[j]);
        if (w < 0 || off + w >= buflen) break;
        off += w;
    }
    return PSSGP_OK;
}



}  // extern "C"

