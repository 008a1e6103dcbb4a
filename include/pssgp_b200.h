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
/* Options: "chunk" = time steps per thread-chunk (0 = heuristic); "timing" = 1 brackets every
 * kernel launch with CUDA events on its stream (read back with pssgp_timing_report);
 * "fused_reverse" = 1 makes pssgp_pkfs_grad run the smoother and adjoint recursions in one kernel;
 * "pdl" = 0 turns off programmatic dependent launch between the kernels of pssgp_pkfs_grad (default 1);
 * "grid_lanes" = settings in flight in pssgp_grid_loglik (0 = default 4, at most 8: measured +2 % at 8); tuning / test switches:
 * "mid_warps" (warps per CTA of the 5 <= d <= 24 kernels), "mid_smem" = 1 (shared-memory tile kernels also for
 * d <= 24), "force_generic" = 1 (CTA-cooperative kernels for d > 4, element-per-thread discretisation). */
int pssgp_set_option(pssgp_handle* h, const char* name, int64_t value);
/* Number of kernels launched through this handle since creation (bench.py's gpu_launches). */
int64_t pssgp_launch_count(const pssgp_handle* h);
/* Writes "name count total_ms\n" per kernel name for the launches recorded since the last report
 * (synchronises the device) into buf (host, buflen bytes) and clears the records. */
int pssgp_timing_report(pssgp_handle* h, char* buf, int64_t buflen);

/*
 * HOST routine (host pointers, float64): diagonal balancing of the SDE drift matrix.  Replaces the
 * reference's only natively compiled function, _numba_balance_ss (pssgp/kernels/math_utils.py:10-29).
 * F [d,d] row-major; d_out [d] receives the scaling (F_balanced = D^-1 F D).
 */
int pssgp_balance_ss(const void* F, int d, int n_iter, void* d_out);

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
 * Outputs: fms [n,d], fPs [n,d,d], ll [1] or NULL, final_state [d + d*d] or NULL (filtered mean |
 * full covariance after the last step: usable as (m0, P0 = m0 + d) of the next shard).
 */
int pssgp_pkf(pssgp_handle* h, int dtype, int64_t n, int d,
              const void* P0, const void* Fs, const void* Qs, const void* H, const void* R,
              const void* y, const void* m0, int first_special,
              void* fms, void* fPs, void* ll, void* final_state, void* stream);

/*
 * Parallel RTS smoother.  Replaces pssgp/kalman/parallel.py:187-196 (pks) incl. :155-184.
 * last_special != 0: time n-1 is the global last step.  Otherwise Fnext/Qnext [d,d] are F, Q of
 * time n (first step of the next shard) and init [d + d(d+1)/2] is the smoothed state at time n
 * (mean | packed lower-triangular covariance).
 * Outputs sms [n,d], sPs [n,d,d], first_state (smoothed state at time 0, packed) or NULL.
 */
int pssgp_pks(pssgp_handle* h, int dtype, int64_t n, int d,
              const void* Fs, const void* Qs, const void* fms, const void* fPs,
              int last_special, const void* Fnext, const void* Qnext, const void* init,
              void* sms, void* sPs, void* first_state, void* stream);

/*
 * Gradient of the log-likelihood w.r.t. the LGSSM fields (hand-written adjoint scan; replaces TF
 * autodiff through pkf, cf. tests/test_gp_vs_kfs.py:53-67).  g_ll [1] is the upstream gradient;
 * fms/fPs are pkf's outputs.  (P0, m0, first_special) as in pssgp_pkf.  adj_init [d + d(d+1)/2] or
 * NULL: adjoint w.r.t. the filtered moments of this shard's last step coming from the next shard.
 * Outputs: dP0 [d,d] (written only when first_special; may be NULL otherwise), dFs [n,d,d],
 * dQs [n,d,d], dH [d], dR [1] (this shard's contributions), adj_first [d + d(d+1)/2] or NULL
 * (adjoint w.r.t. the filtered moments entering this shard).
 */
int pssgp_pkf_backward(pssgp_handle* h, int dtype, int64_t n, int d,
                       const void* P0, const void* m0, const void* Fs, const void* Qs, const void* H,
                       const void* R, const void* y, const void* fms, const void* fPs, const void* g_ll,
                       int first_special, const void* adj_init,
                       void* dP0, void* dFs, void* dQs, void* dH, void* dR, void* adj_first, void* stream);

/*
 * Fused filter + log-likelihood + smoother + gradient over ONE shard that holds the whole series
 * (first step and last step are the global ones): the outputs of pssgp_pkf (fms, fPs, ll), pssgp_pks
 * (sms, sPs) and pssgp_pkf_backward (dP0, dFs, dQs, dH, dR) for the same LGSSM in one call.  This is the
 * pkfs (pssgp/kalman/parallel.py:199-201) + TF-autodiff training step of the reference with the three
 * scans sharing their passes over (Fs, Qs, y): for d <= 4 the forward filter pass also builds the chunk
 * aggregates of both reverse scans, so the LGSSM is read three times instead of six
 * (csrc/fused_small.cuh); 5 <= d <= 32 in FP64 runs the warp-level DMMA kernels of csrc/mid.cuh / mid_frag.cuh
 * (smoother in modified Bryson-Frazier form sharing one reverse scan with the adjoint).  Other cases run the three
 * scans one after the other.  m0 = 0.  sms = sPs = NULL: no smoother (filter + log-likelihood + gradient: the training
 * step).
 */
int pssgp_pkfs_grad(pssgp_handle* h, int dtype, int64_t n, int d,
                    const void* P0, const void* Fs, const void* Qs, const void* H, const void* R,
                    const void* y, const void* g_ll,
                    void* fms, void* fPs, void* ll, void* sms, void* sPs,
                    void* dP0, void* dFs, void* dQs, void* dH, void* dR, void* stream);

/*
 * Time-sharded scans (one contiguous shard of the time axis per GPU).  *_summary runs the local
 * reduce and writes the shard's aggregate element so that ranks can all-gather them:
 *   filter   (A,b,C,J,eta): d*d + 2d + d(d+1) values      smoother (E,g,L): d*d + d + d(d+1)/2
 *   adjoint  (Abar,a,B):    d*d + d + d(d+1)/2
 * The chunk aggregates stay in the handle's workspace: the matching full call (same arrays) that
 * follows skips its reduce kernel.  *_fold turns gathered summaries into the state entering this
 * shard: pssgp_filter_fold folds the `nshards_before` summaries (rank order) onto (m0, P0) and writes
 * m [d] | P [d,d]; pssgp_smoother_fold / pssgp_adjoint_fold fold the `nshards_after` summaries of the
 * following shards (given in rank order, applied last-to-first) and write the packed state.
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
int pssgp_pkf_backward_summary(pssgp_handle* h, int dtype, int64_t n, int d,
                               const void* P0, const void* m0, const void* Fs, const void* Qs, const void* H,
                               const void* R, const void* y, const void* fms, const void* fPs,
                               int first_special, void* summary, void* stream);
int pssgp_adjoint_fold(pssgp_handle* h, int dtype, int d, int nshards_after,
                       const void* summaries, void* state_out, void* stream);

/*
 * Filter (+ log-likelihood) + RTS smoother of one whole series in one call: pkfs of the reference
 * (pssgp/kalman/parallel.py:199-201).  For d <= 4 the filter pass also builds the smoother's chunk aggregates
 * (no second reduce pass).  Outputs fms, fPs, ll (or NULL) as pssgp_pkf, and either sms [n,d] + sPs [n,d,d], or —
 * proj != NULL, d <= 4 only — proj [n,2] = (H m_k, H P_k H^T) of every smoothed state, which is all that
 * predict_f keeps (pssgp/model.py:107-111); sms / sPs may then be NULL.
 */
int pssgp_pkfs(pssgp_handle* h, int dtype, int64_t n, int d,
               const void* P0, const void* Fs, const void* Qs, const void* H, const void* R, const void* y,
               void* fms, void* fPs, void* ll, void* sms, void* sPs, void* proj, void* stream);

/*
 * Time sharding, fused: pssgp_pkf on one shard (same arguments) that also builds, in the same pass, the chunk
 * aggregates of the smoother and adjoint scans and writes their shard summaries (layouts as pssgp_pks_summary /
 * pssgp_pkf_backward_summary) — i.e. pssgp_pkf + pssgp_pks_summary + pssgp_pkf_backward_summary without the two
 * extra passes over (Fs, Qs, y, fms, fPs).  last_special / Fnext / Qnext as in pssgp_pks.  The pssgp_pks and
 * pssgp_pkf_backward calls that follow on the same arrays skip their reduce kernels.  If pssgp_pkf_summary ran on
 * the same (Fs, n) just before, its chunk aggregates are reused (no second filter reduce pass).
 */
int pssgp_pkf_with_summaries(pssgp_handle* h, int dtype, int64_t n, int d,
                             const void* P0, const void* Fs, const void* Qs, const void* H, const void* R,
                             const void* y, const void* m0, int first_special, int last_special,
                             const void* Fnext, const void* Qnext,
                             void* fms, void* fPs, void* ll, void* sm_summary, void* ad_summary, void* stream);

/*
 * Time sharding without fold launches (d <= 4): registers shard summaries (rank order, `stride` scalars between
 * consecutive summaries, e.g. rows of an all-gather buffer) for the next scan of that kind on this handle, which
 * folds them onto its initial state inside its own kernels:
 *   kind = 0  filter summaries of the shards BEFORE this one, for pssgp_pkf_with_summaries (which must follow
 *             pssgp_pkf_summary on the same arrays; P0 / m0 are then the global prior); state_out (m [d] | P [d,d],
 *             may be NULL) receives the folded state entering the shard, e.g. for pssgp_pkf_backward's P0 / m0;
 *   kind = 1  smoother summaries of the shards AFTER this one, for pssgp_pks (init may be NULL);
 *   kind = 2  adjoint summaries of the shards AFTER this one, for pssgp_pkf_backward (adj_init may be NULL).
 * Equivalent to pssgp_filter_fold / pssgp_smoother_fold / pssgp_adjoint_fold + passing their result.  count = 0 clears.
 */
int pssgp_set_fold(pssgp_handle* h, int kind, const void* summaries, int count, int64_t stride, void* state_out);

/*
 * Time sharding for 5 <= d <= 32 (FP64): the smoother (modified Bryson-Frazier form, equal to
 * pssgp/kalman/parallel.py:155-196 in exact arithmetic) and the log-likelihood adjoint run as ONE combined reverse
 * scan, so a shard exchanges one reverse summary  Abar[d,d] | Ba[d,d] | Bm[d,d] | a[d]  (3 d^2 + d values) and
 * one state  dm[d] | lam[d] | dP[d,d] | Lam[d,d]  (2 d^2 + 2 d values):
 *   pssgp_pkf_summary   -> all-gather -> pssgp_filter_fold                     (filter summaries, as for any d)
 *   pssgp_shard_forward    filter of the shard seeded by the folded state (m0, P0; rank 0: the prior, first_special)
 *                          + rev_summary of the shard; fms, fPs, ll as pssgp_pkf
 *   all-gather rev_summary -> pssgp_rev_fold(nshards_after, summaries of the FOLLOWING shards in rank order,
 *                          stride scalars apart) -> state entering this shard from above (not on the last rank)
 *   pssgp_shard_reverse    sms, sPs (pssgp_pks outputs; may be NULL) and dP0, dFs, dQs, dH, dR (pssgp_pkf_backward
 *                          outputs; dFs NULL: no gradient) of the shard; rev_init = the folded state or NULL.
 * The chunk aggregates pssgp_pkf_summary / pssgp_shard_forward leave in the workspace are reused by the call that
 * follows on the same arrays.
 */
int pssgp_shard_forward(pssgp_handle* h, int dtype, int64_t n, int d,
                        const void* P0, const void* Fs, const void* Qs, const void* H, const void* R, const void* y,
                        const void* m0, int first_special,
                        void* fms, void* fPs, void* ll, void* rev_summary, void* stream);
int pssgp_rev_fold(pssgp_handle* h, int dtype, int d, int nshards_after, const void* summaries, int64_t stride,
                   void* state_out, void* stream);
int pssgp_shard_reverse(pssgp_handle* h, int dtype, int64_t n, int d,
                        const void* P0, const void* m0, const void* Fs, const void* Qs, const void* H, const void* R,
                        const void* y, const void* fms, const void* fPs, const void* g_ll, int first_special,
                        const void* rev_init,
                        void* sms, void* sPs, void* dP0, void* dFs, void* dQs, void* dH, void* dR, void* stream);

/*
 * All-gather of one small message per rank over NVLink peer memory (time sharding; replaces the latency-bound NCCL
 * all-gather of shard summaries the SURVEY.md §8e design calls for).  peer_bufs[j] / peer_flags[j] (HOST arrays of
 * `world` device pointers) are rank j's copy of a symmetric buffer mapped into this process (torch symmetric memory or
 * cudaIpc): doubles and 64-bit flags respectively.  The kernel stores msg [nvals doubles] into row `rank`
 * (row_stride doubles per row) at slot_offset doubles of EVERY rank's buffer, publishes a sequence number in flag
 * flag_offset + rank of every rank, and waits until the flags flag_offset + j of its own buffer show that every rank's
 * row has landed: afterwards rows 0 .. world-1 of the local slot hold all messages.  seq_counter: one 64-bit device
 * counter per slot in local memory, zero-initialised, owned by the caller (incremented by the kernel: no host-side
 * state, the launch is CUDA-graph replayable).  err_counter: 64-bit device counter incremented when a wait gives up after
 * ~10 s (a peer never published its row): the caller checks it instead of the GPU hanging.  Every rank must issue the same sequence of exchanges; a slot may be
 * reused once every rank has issued a later exchange.  FP64 messages (reinterpret other payloads).
 */
int pssgp_peer_exchange(pssgp_handle* h, const void* msg, int64_t nvals, void* const* peer_bufs, void* const* peer_flags,
                        int world, int rank, int64_t slot_offset, int64_t row_stride, int64_t flag_offset,
                        void* seq_counter, void* err_counter, void* stream);

/*
 * HOST routines (host pointers, float64): native batched construction of the LTI SDE of a covariance function for
 * `batch` hyper-parameter settings.  Replaces, per setting, get_sde of the Matern family
 * (pssgp/kernels/matern/common.py:26-52, matern12.py:18-23, matern32.py:20-28, matern52.py:21-25), RBF
 * (rbf.py:14-101), Periodic (periodic.py:18-81), SDESum (kernels/base.py:151-183), SDEProduct (:199-244), balance_ss
 * (math_utils.py:10-81) and solve_lyap_vec (math_utils.py:84-120).
 * spec (int32): [balancing iterations of Sum/Product, n_terms, then per term: n_factors, then per factor:
 *   type (0 Matern12, 1 Matern32, 2 Matern52, 3 RBF, 4 Periodic over a squared-exponential), order (RBF / Periodic),
 *   balancing iterations of the factor].  The kernel is the SUM of the terms, each term the PRODUCT of its factors.
 * params [batch, params_stride]: per factor in spec order (variance, lengthscale[, period for Periodic]).
 * Outputs: F [batch,d,d] (balanced drift), Pinf [batch,d,d] (stationary covariance = P0), H [batch,d].
 * nthreads: host threads over the settings (0 = all).
 */
/*
 * Data preparation of predict_f.  Replaces pssgp/model.py:15-55 (_merge_sorted), :99 (NaN observations at the query
 * times) and the time differencing of kernels/base.py:31-35: ts, ys [n] (sorted training times, observations),
 * q [K] (sorted query times), t0 = time before the first step (the reference's t0 = 0).
 * Outputs: t_all, y_all, dts [n+K] (merged times; observations with NaN at the queries; dts[k] = t_all[k] -
 * t_all[k-1], dts[0] = t_all[0] - t0) and q_idx [K] (int64: rows of the queries in the merged arrays — the
 * reference's boolean mask).  Ties between a training and a query time put the element of the SHORTER array first,
 * like the reference's scatter.
 */
int pssgp_merge_queries(pssgp_handle* h, int dtype, int64_t n, int64_t K, const void* ts, const void* ys, const void* q,
                        double t0, void* t_all, void* y_all, void* dts, void* q_idx, void* stream);

/* HOST routine: X [d,d] with F X + X F^T = G (float64, row-major), the d^2 x d^2 Kronecker solve inside
 * solve_lyap_vec (pssgp/kernels/math_utils.py:108-118; Pinf = -sym(X) for G = L Q L^T).  The same routine with F^T is
 * its adjoint (kernels/math_utils.py of this package). */
int pssgp_lyap_solve(const void* F, const void* G, int d, void* X);
int pssgp_sde_dim(const int32_t* spec, int spec_len, int* d_out, int* nparams_out);
int pssgp_sde_batch(const int32_t* spec, int spec_len, int64_t batch, const double* params, int64_t params_stride,
                    double* F, double* Pinf, double* H, int nthreads);
/* The same with the Jacobians w.r.t. the hyper-parameters (forward-mode dual numbers through the whole construction,
 * balancing vector held constant like the reference's tf.numpy_function, math_utils.py:68): dF, dPinf
 * [batch, nparams, d, d], dH [batch, nparams, d] = derivatives w.r.t. params[:, q] (the CONSTRAINED values: variance,
 * lengthscale, period).  Replaces TF autodiff through get_sde; at most 16 hyper-parameters. */
int pssgp_sde_batch_jac(const int32_t* spec, int spec_len, int64_t batch, const double* params, int64_t params_stride,
                        double* F, double* Pinf, double* H, double* dF, double* dPinf, double* dH, int nthreads);
/*
 * Log-likelihood of ONE series under `batch` hyper-parameter settings (BASELINE configs[4]b, the grid search the
 * reference runs as a Python loop over models): per setting discretise (kernels/base.py:29-47) + pkf with
 * log-likelihood (kalman/parallel.py:121-152), all enqueued on `stream` from this one call; the LGSSM of a setting
 * lives in the handle's workspace and is never returned.  F, Pinf [batch,d,d], H [batch,d], R [batch] (device, as
 * uploaded from pssgp_sde_batch); dts, y [n]; ll [batch].
 */
int pssgp_grid_loglik(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, const void* F, const void* Pinf,
                      const void* H, const void* R, const void* dts, const void* y, void* ll, void* stream);
/* The same with the gradient of every log-likelihood w.r.t. the SDE of its setting (per setting: discretise, fused
 * filter + adjoint = pssgp_pkfs_grad without smoother, pssgp_discretise_backward): dF, dPinf [batch,d,d] (through the
 * discretisation), dP0 [batch,d,d] (through the initial covariance; P0 = Pinf, so the total is dPinf + dP0), dH
 * [batch,d], dR [batch].  Contracted with the Jacobians of pssgp_sde_batch_jac this is the hyper-parameter gradient of
 * a whole batch of settings (gradient-based search, many MCMC chains) from one call. */
int pssgp_grid_loglik_grad(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, const void* F, const void* Pinf,
                           const void* H, const void* R, const void* dts, const void* y, void* ll, void* dF, void* dPinf,
                           void* dP0, void* dH, void* dR, void* stream);

/*
 * Sequential Kalman filter for `batch` independent series.  Replaces pssgp/kalman/sequential.py:11-47 (kf): per step
 * predict (mp = F m, Pp = sym(F P F^T + Q)), skip the update where y is NaN, else update with the scalar
 * observation and add log N(y; H mp, H Pp H^T + R) to the log-likelihood; m0 = 0 as in the reference.
 * y [batch,n]; lgssm_batched = 0: ONE LGSSM (P0 [d,d], Fs,Qs [n,d,d], H [d], R [1]) shared by all series,
 * != 0: one per series (P0 [batch,d,d], Fs,Qs [batch,n,d,d], H [batch,d], R [batch]).  d <= 32.
 * Outputs: fms [batch,n,d], fPs [batch,n,d,d]; mps, Pps (predicted moments, same shapes; both or neither: the
 * reference's return_predicted); ll [batch] or NULL.  The parallel axis is the batch: one thread per series for
 * d <= 4, one warp per series above.
 */
int pssgp_kf(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, int lgssm_batched,
             const void* P0, const void* Fs, const void* Qs, const void* H, const void* R, const void* y,
             void* fms, void* fPs, void* mps, void* Pps, void* ll, void* stream);
/*
 * Sequential RTS smoother.  Replaces pssgp/kalman/sequential.py:50-68 (ks): backwards from the last filtered state,
 * C = P_k F_{k+1}^T Pp_{k+1}^-1 (Cholesky), sm_k = m_k + C (sm_{k+1} - mp_{k+1}),
 * sP_k = sym(P_k + C (sP_{k+1} - Pp_{k+1}) C^T).  Inputs as produced by pssgp_kf; sms [batch,n,d], sPs [batch,n,d,d].
 * A predicted covariance that is not positive definite gives NaNs from that step on (the reference raises).
 */
int pssgp_ks(pssgp_handle* h, int dtype, int64_t batch, int64_t n, int d, int lgssm_batched,
             const void* Fs, const void* fms, const void* fPs, const void* mps, const void* Pps,
             void* sms, void* sPs, void* stream);

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
