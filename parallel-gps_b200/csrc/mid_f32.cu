// FP32 opt-in mode for 5 <= d <= 32: FP32 STORAGE at the C ABI, FP64 ARITHMETIC in the warp-level DMMA kernels
// (mid_frag.cuh / mid.cuh).  The inputs are widened into the handle's workspace, the FP64 path runs, the outputs are
// narrowed into the caller's arrays.  (The FP64 tensor-core path is the only MMA path with enough mantissa for the
// covariance recursions: TF32 / BF16 tensor cores are not an option; a DFMA-free FP32 variant would need its own
// register-tile kernels.)  Results are at least as accurate as an all-FP32 evaluation; the extra conversion traffic is
// 12 bytes per array element against the 8 + 8 the FP64 path moves anyway.
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>

#include "generic_algebras.cuh"
#include "mid_host.h"
#include "workspace.h"

namespace pssgp {
namespace mid {

template <typename TI, typename TO> __global__ void cvt_kernel(const TI* __restrict__ in, TO* __restrict__ out, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = (TO)in[i];
}
template <typename TI, typename TO> static void cvt(pssgp_handle* h, const TI* in, TO* out, long n, cudaStream_t st) {
    if (in == nullptr || out == nullptr || n <= 0) return;
    long blocks = (n + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    PSSGP_LAUNCH(h, "mid_f32_convert", st, (cvt_kernel<TI, TO><<<(unsigned)blocks, 256, 0, st>>>(in, out, n)));
    h->launches += 1;
}

// bump allocator over one workspace slot
struct Bump {
    char* base;
    size_t off;
    double* take(size_t ndoubles) {
        double* p = (double*)(base + off);
        off += ((ndoubles * 8 + 255) / 256) * 256;
        return p;
    }
};
static size_t padded(size_t ndoubles) { return ((ndoubles * 8 + 255) / 256) * 256; }

// filter (+ll) [+ smoother] [+ gradient] in FP32 storage; sms / dFs / g_ll may be null as in the FP64 entry points
int f32_pkfs_grad(pssgp_handle* h, int64_t n, int d, const float* P0, const float* Fs, const float* Qs, const float* H,
                  const float* R, const float* y, const float* g_ll, float* fms, float* fPs, float* ll, float* sms,
                  float* sPs, float* dP0, float* dFs, float* dQs, float* dH, float* dR, bool filter_only,
                  cudaStream_t st) {
    const size_t nd = (size_t)n * d, ndd = nd * d, dd = (size_t)d * d;
    const bool smooth = sms != nullptr, adj = dFs != nullptr;
    size_t need = 2 * padded(ndd) + padded(n) + 2 * padded(dd) + 4 * padded(d) + 8 * 256   // inputs + small
                  + padded(nd) + padded(ndd);                                              // fms, fPs
    if (smooth) need += padded(nd) + padded(ndd);
    if (adj) need += 2 * padded(ndd);
    int rc;
    if ((rc = ws_reserve(h, WS_F32, need))) return rc;
    Bump b{(char*)h->buf[WS_F32], 0};
    double *Fs64 = b.take(ndd), *Qs64 = b.take(ndd), *y64 = b.take(n), *P064 = b.take(dd), *H64 = b.take(d), *R64 = b.take(1);
    double *g64 = b.take(1), *ll64 = b.take(1), *dP064 = b.take(dd), *dH64 = b.take(d), *dR64 = b.take(1);
    double *fms64 = b.take(nd), *fPs64 = b.take(ndd);
    double *sms64 = smooth ? b.take(nd) : nullptr, *sPs64 = smooth ? b.take(ndd) : nullptr;
    double *dFs64 = adj ? b.take(ndd) : nullptr, *dQs64 = adj ? b.take(ndd) : nullptr;
    cvt(h, Fs, Fs64, (long)ndd, st);
    cvt(h, Qs, Qs64, (long)ndd, st);
    cvt(h, y, y64, (long)n, st);
    cvt(h, P0, P064, (long)dd, st);
    cvt(h, H, H64, (long)d, st);
    cvt(h, R, R64, 1, st);
    if (adj) cvt(h, g_ll, g64, 1, st);
    if (filter_only)
        rc = pkf_dispatch(d, h, n, P064, Fs64, Qs64, H64, R64, y64, nullptr, 1, fms64, fPs64, ll ? ll64 : nullptr, nullptr,
                          nullptr, st);
    else
        rc = pkfs_grad_dispatch(d, h, n, P064, Fs64, Qs64, H64, R64, y64, adj ? g64 : nullptr, fms64, fPs64,
                                ll ? ll64 : nullptr, sms64, sPs64, adj ? dP064 : nullptr, dFs64, dQs64, adj ? dH64 : nullptr,
                                adj ? dR64 : nullptr, st);
    if (rc) return rc;
    cvt(h, fms64, fms, (long)nd, st);
    cvt(h, fPs64, fPs, (long)ndd, st);
    if (ll) cvt(h, ll64, ll, 1, st);
    if (smooth) {
        cvt(h, sms64, sms, (long)nd, st);
        cvt(h, sPs64, sPs, (long)ndd, st);
    }
    if (adj) {
        cvt(h, dFs64, dFs, (long)ndd, st);
        cvt(h, dQs64, dQs, (long)ndd, st);
        cvt(h, dP064, dP0, (long)dd, st);
        cvt(h, dH64, dH, (long)d, st);
        cvt(h, dR64, dR, 1, st);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "mid f32: %s", cudaGetErrorString(e));
    return PSSGP_OK;
}

}  // namespace mid
}  // namespace pssgp
