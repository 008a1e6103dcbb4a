// Host-side compiled helpers of the C ABI.
#include "../../include/pssgp_b200.h"

#include <math.h>

#include <vector>

#include "workspace.h"

extern "C" int pssgp_balance_ss(const void* F_in, int d, int n_iter, void* d_out) {
    // Iterated row/column-norm balancing; replaces the numba routine of the reference
    // (pssgp/kernels/math_utils.py:10-29).  F_in: host [d,d] float64 row-major; d_out: host [d] float64.
    // A state without off-diagonal coupling (r == 0 or c == 0) is left unscaled; the reference divides 0/0 there.
    if (!F_in || !d_out || d < 1 || n_iter < 0) return pssgp::set_err(PSSGP_ERR_INVALID, "balance_ss: bad argument");
    const double* Fi = (const double*)F_in;
    double* dv = (double*)d_out;
    std::vector<double> F(Fi, Fi + (size_t)d * d);
    for (int i = 0; i < d; ++i) dv[i] = 1.0;
    for (int it = 0; it < n_iter; ++it) {
        for (int i = 0; i < d; ++i) {
            double c = 0.0, r = 0.0;
            for (int k = 0; k < d; ++k) {
                if (k == i) continue;
                c += F[(size_t)k * d + i] * F[(size_t)k * d + i];
                r += F[(size_t)i * d + k] * F[(size_t)i * d + k];
            }
            c = sqrt(c);
            r = sqrt(r);
            if (!(c > 0.0) || !(r > 0.0)) continue;
            const double f = sqrt(r / c);
            dv[i] *= f;
            for (int k = 0; k < d; ++k) F[(size_t)k * d + i] *= f;
            for (int k = 0; k < d; ++k) F[(size_t)i * d + k] /= f;
        }
    }
    return PSSGP_OK;
}
