// Host entry points of the warp-level d > 4 kernels (mid.cuh); one explicit instantiation per supported D, each in
// its own translation unit (mid_inst.cu compiled with -DMID_D=<D>).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace pssgp {
namespace mid {

// filter (+ log-likelihood); summary != nullptr: shard summary only (time sharding)
template <int D>
int pkf(pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H, const double* R,
        const double* y, const double* m0, int first_special, double* fms, double* fPs, double* ll, double* final_state,
        double* summary, cudaStream_t st);
// fused filter + smoother (sms != nullptr) + gradient (dFs != nullptr) of one whole series
template <int D>
int pkfs_grad(pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H,
              const double* R, const double* y, const double* g_ll, double* fms, double* fPs, double* ll, double* sms,
              double* sPs, double* dP0, double* dFs, double* dQs, double* dH, double* dR, cudaStream_t st);
// gradient from stored filtered moments (one whole shard, no incoming adjoint)
template <int D>
int pkf_backward(pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs, const double* Qs,
                 const double* H, const double* R, const double* y, const double* fms, const double* fPs,
                 const double* g_ll, int first_special, double* dP0, double* dFs, double* dQs, double* dH, double* dR,
                 cudaStream_t st);

template <int D>
int shard_forward(pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H,
                  const double* R, const double* y, const double* m0, int first_special, double* fms, double* fPs,
                  double* ll, double* rev_summary, cudaStream_t st);
template <int D>
int shard_reverse(pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs, const double* Qs,
                  const double* H, const double* R, const double* y, const double* fms, const double* fPs,
                  const double* g_ll, int first_special, const double* rev_init, double* sms, double* sPs, double* dP0,
                  double* dFs, double* dQs, double* dH, double* dR, cudaStream_t st);
template <int D>
int rev_fold(pssgp_handle* h, const double* summaries, int count, int64_t stride, double* state_out, cudaStream_t st);

bool supported(int d);
// FP32 storage, FP64 arithmetic (mid_f32.cu)
int f32_pkfs_grad(pssgp_handle* h, int64_t n, int d, const float* P0, const float* Fs, const float* Qs, const float* H,
                  const float* R, const float* y, const float* g_ll, float* fms, float* fPs, float* ll, float* sms,
                  float* sPs, float* dP0, float* dFs, float* dQs, float* dH, float* dR, bool filter_only,
                  cudaStream_t st);
int shard_forward_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs,
                           const double* H, const double* R, const double* y, const double* m0, int first_special,
                           double* fms, double* fPs, double* ll, double* rev_summary, cudaStream_t st);
int shard_reverse_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs,
                           const double* Qs, const double* H, const double* R, const double* y, const double* fms,
                           const double* fPs, const double* g_ll, int first_special, const double* rev_init, double* sms,
                           double* sPs, double* dP0, double* dFs, double* dQs, double* dH, double* dR, cudaStream_t st);
int rev_fold_dispatch(int d, pssgp_handle* h, const double* summaries, int count, int64_t stride, double* state_out,
                      cudaStream_t st);
int pkf_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H,
                 const double* R, const double* y, const double* m0, int first_special, double* fms, double* fPs,
                 double* ll, double* final_state, double* summary, cudaStream_t st);
int pkfs_grad_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs,
                       const double* H, const double* R, const double* y, const double* g_ll, double* fms, double* fPs,
                       double* ll, double* sms, double* sPs, double* dP0, double* dFs, double* dQs, double* dH,
                       double* dR, cudaStream_t st);
int pkf_backward_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs,
                          const double* Qs, const double* H, const double* R, const double* y, const double* fms,
                          const double* fPs, const double* g_ll, int first_special, double* dP0, double* dFs,
                          double* dQs, double* dH, double* dR, cudaStream_t st);

}  // namespace mid

// hierarchy over chunk aggregates (generic.cu)
size_t hier_total(int64_t cnt0);
int hier_filter_f64(pssgp_handle* h, const GFilter<double>::Params& p, int d, int64_t cnt0, double* aggs, double* states,
                    double* final_state, double* summary, bool have_up, cudaStream_t st, int* launches);
int hier_rev_f64(pssgp_handle* h, const GRev<double>::Params& p, int d, int64_t cnt0, double* aggs, double* states,
                 double* final_state, double* summary, bool have_up, cudaStream_t st, int* launches);
int finish_filter_f64(pssgp_handle* h, const GFilter<double>::Params& p, const double* part, int64_t nparts, double* ll,
                      cudaStream_t st);
int finish_rev_f64(pssgp_handle* h, const GRev<double>::Params& p, const double* part, int64_t nparts, cudaStream_t st);

}  // namespace pssgp
