/*
 * pssgp_b200 — C ABI of the B200-native temporally-parallel state-space GP inference path.
 *
 * The reference (EEA-sensors/parallel-gps) is pure Python on TensorFlow: its "plugin boundary" for
 * this path is the Python API of pssgp/kalman/parallel.py (pkf :121, pks :187, pkfs :199) and
 * pssgp/kernels/base.py (_get_ssm :29-47).  Each entry point below is what a Python/TF binding of
 * that API would bind (ctypes / tf.py_function + DLPack).  All pointers are
 * DEVICE pointers unless stated otherwise; matrices are dense row-major; no torch/TF types cross
 * this boundary.  Every function returns 0 on success, non-zero on error; the message is
 * available from pssgp_last_error().  Inputs are borrowed, outputs are caller-allocated.
 *
 * dtype: PSSGP_F64 (reference default, gpflow default_float) or PSSGP_F32.
 * stream: a cudaStream_t passed as void* (NULL = legacy default stream).
 */
#ifndef PSSGP_B200_H
#define PSSGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSSGP_F64 0
#define PSSGP_F32 1

#define PSSGP_OK 0
#define PSSGP_ERR_INVALID 1
#define PSSGP_ERR_CUDA 2
#define PSSGP_ERR_UNSUPPORTED 3

typedef struct pssgp_handle pssgp_handle;

/* Library version (major*10000 + minor*100 + patch). */
int pssgp_version(void);
/* Message of the last error raised on the calling thread. */
const char* pssgp_last_error(void);

/* A handle owns the reusable device workspace (chunk aggregates, warp totals, partial sums). */
int pssgp_create(pssgp_handle** out, int device);
int pssgp_destroy(pssgp_handle* h);
/* Tuning knobs: "chunk" = time steps per thread-chunk (0 = heuristic). */
int pssgp_set_option(pssgp_handle* h, const char* name, int64_t value);
/* Number of kernels launched through this handle since creation (bench.py's gpu_launches). */
int64_t pssgp_launch_count(const pssgp_handle* h);

/*
 * Discretisation of the LTI SDE.  Replaces pssgp/kernels/base.py:29-47 (_get_ssm):
 *   Fs[k] = expm(F * dts[k]);   Qs[k] = Pinf - Fs[k] Pinf Fs[k]^T   (stationary form, symmetrised)
 * F, Pinf: [d,d]; dts: [n]; Fs, Qs: [n,d,d].
 */
int pssgp_discretise(pssgp_handle* h, int dtype, int64_t n, int d,
                     const void* F, const void* Pinf, const void* dts,
                     void* Fs, void* Qs, void* stream);

/*
 * Parallel Kalman filter + log-likelihood.  Replaces pssgp/kalman/parallel.py:121-152 (pkf) including
 * element construction (:13-97), the associative operator (:100-118) and the log-likelihood block
 * (:135-151).  P0 [d,d]; Fs,Qs [n,d,d]; H [d] (the reference's H is [1,d]); R [1]; y [n] (NaN =
 * missing); m0 [d] or NULL (zeros, as the reference).  first_special != 0: step 0 is the global first
 * step (update on (m0,P0) without prediction, parallel.py:24-30); 0: (m0,P0) is the filtered state
 * just before this shard (time-sharded multi-GPU use).
 * Outputs: fms [n,d], fPs [n,d,d], ll [1] or NULL, final_state [d + d(d+1)/2] or NULL
 * (filtered mean | packed lower-triangular covariance after the last step).
 */
int pssgp_pkf(pssgp_handle* h, int dtype, int64_t n, int d,
              const void* P0, const void* Fs, const void* Qs, const void* H, const void* R,
              const void* y, const void* m0, int first_special,
              void* fms, void* fPs, void* ll, void* final_state, void* stream);

/*
 * Parallel RTS smoother.  Replaces pssgp/kalman/parallel.py:187-196 (pks) incl. :155-184.
 * last_special != 0: time n-1 is the global last step.  Otherwise Fnext/Qnext [d,d] are F, Q of
 * time n (first step of the next shard) and init [d + d(d+1)/2] is the smoothed state at time n.
 * Outputs sms [n,d], sPs [n,d,d], first_state (smoothed state at time 0, packed) or NULL.
 */
int pssgp_pks(pssgp_handle* h, int dtype, int64_t n, int d,
              const void* Fs, const void* Qs, const void* fms, const void* fPs,
              int last_special, const void* Fnext, const void* Qnext, const void* init,
              void* sms, void* sPs, void* first_state, void* stream);

/*
 * Time-sharded scans (one shard per GPU).  *_summary runs the local reduce and writes the shard's
 * aggregate element ((A,b,C,J,eta) packed: d*d + 2d + d(d+1) values; (E,g,L) packed: d*d + d +
 * d(d+1)/2 values) so that ranks can all-gather them; pssgp_filter_fold / pssgp_smoother_fold turn
 * the gathered summaries of the preceding (following) shards into the state entering this shard.
 */
int pssgp_pkf_summary(pssgp_handle* h, int dtype, int64_t n, int d,
                      const void* P0, const void* Fs, const void* Qs, const void* H, const void* R,
                      const void* y, int first_special, void* summary, void* stream);
int pssgp_filter_fold(pssgp_handle* h, int dtype, int d, int nshards_before,
                      const void* P0, const void* m0, const void* summaries, void* state_out,
                      void* stream);
int pssgp_pks_summary(pssgp_handle* h, int dtype, int64_t n, int d,
                      const void* Fs, const void* Qs, const void* fms, const void* fPs,
                      int last_special, const void* Fnext, const void* Qnext,
                      void* summary, void* stream);
int pssgp_smoother_fold(pssgp_handle* h, int dtype, int d, int nshards_after,
                        const void* summaries, void* state_out, void* stream);

/*
 * Gradient of the log-likelihood w.r.t. the LGSSM fields (hand-written adjoint scan; replaces TF
 * autodiff through pkf, cf. tests/test_gp_vs_kfs.py:53-67).  g_ll [1] is the upstream gradient.
 * Outputs: dP0 [d,d], dFs [n,d,d], dQs [n,d,d], dH [d], dR [1].
 */
int pssgp_pkf_backward(pssgp_handle* h, int dtype, int64_t n, int d,
                       const void* P0, const void* Fs, const void* Qs, const void* H, const void* R,
                       const void* y, const void* fms, const void* fPs, const void* g_ll,
                       void* dP0, void* dFs, void* dQs, void* dH, void* dR, void* stream);

/*
 * Adjoint of pssgp_discretise: (dFs, dQs) -> (dF, dPinf).  dF, dPinf: [d,d].
 */
int pssgp_discretise_backward(pssgp_handle* h, int dtype, int64_t n, int d,
                              const void* F, const void* Pinf, const void* dts,
                              const void* Fs, const void* dFs, const void* dQs,
                              void* dF, void* dPinf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSSGP_B200_H */
