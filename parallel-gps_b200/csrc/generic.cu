// Generic state dimension (d > 4): warp-cooperative kernels.  Placeholder: fails loudly until implemented.
#include "scan_run.cuh"

namespace pssgp {

int pkf_generic(pssgp_handle*, int, int64_t, int d, const void*, const void*, const void*, const void*, const void*,
                const void*, const void*, int, void*, void*, void*, void*, void*, cudaStream_t) {
    return set_err(PSSGP_ERR_UNSUPPORTED, "pkf: state dimension %d not supported yet", d);
}
int filter_fold_generic(pssgp_handle*, int, int d, int, const void*, const void*, const void*, void*, cudaStream_t) {
    return set_err(PSSGP_ERR_UNSUPPORTED, "filter_fold: state dimension %d not supported yet", d);
}
int pks_generic(pssgp_handle*, int, int64_t, int d, const void*, const void*, const void*, const void*, int,
                const void*, const void*, const void*, void*, void*, void*, void*, cudaStream_t) {
    return set_err(PSSGP_ERR_UNSUPPORTED, "pks: state dimension %d not supported yet", d);
}
int smoother_fold_generic(pssgp_handle*, int, int d, int, const void*, void*, cudaStream_t) {
    return set_err(PSSGP_ERR_UNSUPPORTED, "smoother_fold: state dimension %d not supported yet", d);
}
int pkf_bwd_generic(pssgp_handle*, int, int64_t, int d, const void*, const void*, const void*, const void*,
                    const void*, const void*, const void*, const void*, const void*, const void*, int, const void*,
                    void*, void*, void*, void*, void*, void*, void*, cudaStream_t) {
    return set_err(PSSGP_ERR_UNSUPPORTED, "pkf_backward: state dimension %d not supported yet", d);
}
int adjoint_fold_generic(pssgp_handle*, int, int d, int, const void*, void*, cudaStream_t) {
    return set_err(PSSGP_ERR_UNSUPPORTED, "adjoint_fold: state dimension %d not supported yet", d);
}

}  // namespace pssgp
