// Run-time dispatch on the state dimension for the warp-level d > 4 path (mid.cuh); the instantiations live in
// mid_inst.cu, one object per dimension (see the Makefile).
#include "../../include/pssgp_b200.h"
#include "generic_algebras.cuh"
#include "mid_host.h"

namespace pssgp {
namespace mid {

#define MID_FOR_EACH_D(X) \
    X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) \
    X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32)

#define MID_DECL(D)                                                                                                       \
    extern template int pkf<D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*,        \
                               const double*, const double*, const double*, int, double*, double*, double*, double*,      \
                               double*, cudaStream_t);                                                                    \
    extern template int pkfs_grad<D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*, \
                                     const double*, const double*, const double*, double*, double*, double*, double*,     \
                                     double*, double*, double*, double*, double*, double*, cudaStream_t);                 \
    extern template int pkf_backward<D>(pssgp_handle*, int64_t, const double*, const double*, const double*,             \
                                        const double*, const double*, const double*, const double*, const double*,        \
                                        const double*, const double*, int, double*, double*, double*, double*, double*,   \
                                        cudaStream_t);
MID_FOR_EACH_D(MID_DECL)
#define MID_DECL2(D)                                                                                                       \
    extern template int shard_forward<D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*, \
                                         const double*, const double*, const double*, int, double*, double*, double*,      \
                                         double*, cudaStream_t);                                                            \
    extern template int shard_reverse<D>(pssgp_handle*, int64_t, const double*, const double*, const double*, const double*, \
                                         const double*, const double*, const double*, const double*, const double*,         \
                                         const double*, int, const double*, double*, double*, double*, double*, double*,    \
                                         double*, double*, cudaStream_t);                                                   \
    extern template int rev_fold<D>(pssgp_handle*, const double*, int, int64_t, double*, cudaStream_t);
MID_FOR_EACH_D(MID_DECL2)

bool supported(int d) { return d >= 5 && d <= 32; }

int pkf_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs, const double* H,
                 const double* R, const double* y, const double* m0, int first_special, double* fms, double* fPs,
                 double* ll, double* final_state, double* summary, cudaStream_t st) {
    switch (d) {
#define MID_CASE(D) \
    case D: return pkf<D>(h, n, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, final_state, summary, st);
        MID_FOR_EACH_D(MID_CASE)
#undef MID_CASE
    }
    return set_err(PSSGP_ERR_UNSUPPORTED, "mid path: state dimension %d", d);
}

int pkfs_grad_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs,
                       const double* H, const double* R, const double* y, const double* g_ll, double* fms, double* fPs,
                       double* ll, double* sms, double* sPs, double* dP0, double* dFs, double* dQs, double* dH,
                       double* dR, cudaStream_t st) {
    switch (d) {
#define MID_CASE(D) \
    case D: return pkfs_grad<D>(h, n, P0, Fs, Qs, H, R, y, g_ll, fms, fPs, ll, sms, sPs, dP0, dFs, dQs, dH, dR, st);
        MID_FOR_EACH_D(MID_CASE)
#undef MID_CASE
    }
    return set_err(PSSGP_ERR_UNSUPPORTED, "mid path: state dimension %d", d);
}

int pkf_backward_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs,
                          const double* Qs, const double* H, const double* R, const double* y, const double* fms,
                          const double* fPs, const double* g_ll, int first_special, double* dP0, double* dFs,
                          double* dQs, double* dH, double* dR, cudaStream_t st) {
    switch (d) {
#define MID_CASE(D)                                                                                                 \
    case D:                                                                                                         \
        return pkf_backward<D>(h, n, P0, m0, Fs, Qs, H, R, y, fms, fPs, g_ll, first_special, dP0, dFs, dQs, dH, dR, st);
        MID_FOR_EACH_D(MID_CASE)
#undef MID_CASE
    }
    return set_err(PSSGP_ERR_UNSUPPORTED, "mid path: state dimension %d", d);
}

int shard_forward_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* Fs, const double* Qs,
                           const double* H, const double* R, const double* y, const double* m0, int first_special,
                           double* fms, double* fPs, double* ll, double* rev_summary, cudaStream_t st) {
    switch (d) {
#define MID_CASE(D) \
    case D: return shard_forward<D>(h, n, P0, Fs, Qs, H, R, y, m0, first_special, fms, fPs, ll, rev_summary, st);
        MID_FOR_EACH_D(MID_CASE)
#undef MID_CASE
    }
    return set_err(PSSGP_ERR_UNSUPPORTED, "mid path: state dimension %d", d);
}

int shard_reverse_dispatch(int d, pssgp_handle* h, int64_t n, const double* P0, const double* m0, const double* Fs,
                           const double* Qs, const double* H, const double* R, const double* y, const double* fms,
                           const double* fPs, const double* g_ll, int first_special, const double* rev_init, double* sms,
                           double* sPs, double* dP0, double* dFs, double* dQs, double* dH, double* dR, cudaStream_t st) {
    switch (d) {
#define MID_CASE(D)                                                                                                     \
    case D:                                                                                                             \
        return shard_reverse<D>(h, n, P0, m0, Fs, Qs, H, R, y, fms, fPs, g_ll, first_special, rev_init, sms, sPs, dP0, dFs, \
                                dQs, dH, dR, st);
        MID_FOR_EACH_D(MID_CASE)
#undef MID_CASE
    }
    return set_err(PSSGP_ERR_UNSUPPORTED, "mid path: state dimension %d", d);
}

int rev_fold_dispatch(int d, pssgp_handle* h, const double* summaries, int count, int64_t stride, double* state_out,
                      cudaStream_t st) {
    switch (d) {
#define MID_CASE(D) \
    case D: return rev_fold<D>(h, summaries, count, stride, state_out, st);
        MID_FOR_EACH_D(MID_CASE)
#undef MID_CASE
    }
    return set_err(PSSGP_ERR_UNSUPPORTED, "mid path: state dimension %d", d);
}

}  // namespace mid
}  // namespace pssgp
