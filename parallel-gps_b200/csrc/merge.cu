// C ABI: pssgp_merge_queries (see include/pssgp_b200.h) — the data preparation of predict_f
// (pssgp/model.py:15-55 _merge_sorted, :99 NaN padding, :104 the merged time grid): merge the sorted query times into
// the sorted training times, carry the observations along (NaN at the queries), and difference the merged grid into
// the time steps the discretisation takes (kernels/base.py:31-35).  Two launches instead of a dozen framework
// kernels: every element finds its merged position by one binary search in the OTHER array (a merge path evaluated
// pointwise: position = own index + number of elements of the other array that precede it).
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "scan_run.cuh"
#include "workspace.h"

namespace pssgp {

template <typename T> __device__ __forceinline__ T nan_value();
template <> __device__ __forceinline__ double nan_value<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <> __device__ __forceinline__ float nan_value<float>() { return __int_as_float(0x7fc00000); }

// number of elements of sorted v[0..n) that are < x (STRICT) or <= x (!STRICT)
template <typename T, bool STRICT>
__device__ __forceinline__ long count_before(const T* __restrict__ v, long n, T x) {
    long lo = 0, hi = n;
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        const T e = v[mid];
        const bool before = STRICT ? (e < x) : (e <= x);
        if (before)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// ties: the SHORTER array goes first (the reference scatters the shorter array at searchsorted(longer, shorter) and
// fills the remaining slots with the longer one in order, model.py:27-41)
template <typename T>
__global__ void merge_scatter_kernel(const T* __restrict__ ts, const T* __restrict__ ys, long n, const T* __restrict__ q,
                                     long K, int queries_first, T* __restrict__ t_all, T* __restrict__ y_all,
                                     long long* __restrict__ q_idx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const T t = ts[i];
        // queries_first: a query equal to t precedes it -> count queries <= t
        const long pos = i + (queries_first ? count_before<T, false>(q, K, t) : count_before<T, true>(q, K, t));
        t_all[pos] = t;
        y_all[pos] = ys[i];
    } else if (i < n + K) {
        const long j = i - n;
        const T t = q[j];
        const long pos = j + (queries_first ? count_before<T, true>(ts, n, t) : count_before<T, false>(ts, n, t));
        t_all[pos] = t;
        y_all[pos] = nan_value<T>();
        q_idx[j] = pos;
    }
}

template <typename T>
__global__ void merge_diff_kernel(const T* __restrict__ t_all, long m, T t0, T* __restrict__ dts) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) dts[k] = t_all[k] - (k == 0 ? t0 : t_all[k - 1]);
}

template <typename T>
int merge_impl(pssgp_handle* h, int64_t n, int64_t K, const void* ts, const void* ys, const void* q, double t0, void* t_all,
               void* y_all, void* dts, void* q_idx, cudaStream_t st) {
    const long m = (long)(n + K);
    const unsigned grid = (unsigned)((m + 255) / 256);
    PSSGP_LAUNCH(h, "merge_scatter", st,
                 (merge_scatter_kernel<T><<<grid, 256, 0, st>>>((const T*)ts, (const T*)ys, (long)n, (const T*)q, (long)K,
                                                                K <= n ? 1 : 0, (T*)t_all, (T*)y_all, (long long*)q_idx)));
    int rc = check_launch(h, "merge_scatter", 1);
    if (rc) return rc;
    PSSGP_LAUNCH(h, "merge_diff", st, (merge_diff_kernel<T><<<grid, 256, 0, st>>>((const T*)t_all, m, (T)t0, (T*)dts)));
    return check_launch(h, "merge_diff", 1);
}

}  // namespace pssgp

using namespace pssgp;

extern "C" int pssgp_merge_queries(pssgp_handle* h, int dtype, int64_t n, int64_t K, const void* ts, const void* ys,
                                   const void* q, double t0, void* t_all, void* y_all, void* dts, void* q_idx,
                                   void* stream) {
    int rc = check_common(h, dtype, n, 1);
    if (rc) return rc;
    if (K < 1 || !ts || !ys || !q || !t_all || !y_all || !dts || !q_idx)
        return set_err(PSSGP_ERR_INVALID, "merge_queries: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PSSGP_F64) return merge_impl<double>(h, n, K, ts, ys, q, t0, t_all, y_all, dts, q_idx, st);
    return merge_impl<float>(h, n, K, ts, ys, q, t0, t_all, y_all, dts, q_idx, st);
}
