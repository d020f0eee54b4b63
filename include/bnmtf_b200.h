/*
 * bnmtf_b200.h -- C ABI of the B200-native inference engine for the BNMF / NMF
 * model family of ThomasBrouwer/BNMTF_ARD.
 *
 * The reference is pure Python and defines no FFI of its own: its boundary is
 * the duck-typed class protocol of code/models (ctor, initialise, run, predict,
 * plus the per-parameter methods its tests call).  The entry points below are
 * what a ctypes binding of that protocol needs; each one cites the reference
 * method it replaces (paths relative to the reference checkout).  The Python
 * mirror of the class protocol lives in bnmtf_ard_b200/ and calls only these.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Pointers are HOST pointers unless the
 *    parameter name ends in `_dev`.  Matrices are C-contiguous (row-major)
 *    float64; masks are uint8 (0 = unobserved, non-zero = observed).
 *  - Every function returns 0 on success, non-zero on failure;
 *    bnmtf_last_error() returns a message for the calling thread.
 *  - Calls are synchronous on return.  One handle is used from one thread.
 *  - There is no CPU fallback: without a CUDA device every compute entry
 *    point fails.
 */
#ifndef BNMTF_B200_H
#define BNMTF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bnmtf_nmf_s* bnmtf_nmf_t;

/* inference method (which reference class the handle behaves as) */
enum {
  BNMTF_GIBBS = 0, /* code/models/bnmf_gibbs.py */
  BNMTF_ICM   = 1, /* code/models/nmf_icm.py    */
  BNMTF_VB    = 2, /* code/models/bnmf_vb.py    */
  BNMTF_NP    = 3  /* code/models/nmf_np.py     */
};

/* which factor: U (rows of R, I x K) or V (columns of R, J x K) */
enum { BNMTF_SIDE_U = 0, BNMTF_SIDE_V = 1 };

/* per-factor state arrays (n x K row-major).  VALUE is U/V for Gibbs, ICM, NP
 * and exp_U/exp_V for VB. */
enum { BNMTF_F_VALUE = 0, BNMTF_F_VAR = 1, BNMTF_F_MU = 2, BNMTF_F_TAU = 3 };

/* number of float64 values written per iteration into `iter_stats` of
 * bnmtf_nmf_run / per call of bnmtf_nmf_stats; see BNMTF_ST_* */
#define BNMTF_NSTATS 16
enum {
  BNMTF_ST_RSS      = 0,  /* sum_Omega (R - X Y^T)^2            bnmf_gibbs.py:193        */
  BNMTF_ST_SUM_E    = 1,  /* sum_Omega (R - X Y^T)                                      */
  BNMTF_ST_SUM_RCE  = 2,  /* sum_Omega (R - mean_Omega R) * E    -> R^2, Rp (:252-270)   */
  BNMTF_ST_TAU      = 3,  /* tau draw / mode, or exp_tau (VB)                           */
  BNMTF_ST_BETA_S   = 4,  /* beta* of tau                        bnmf_gibbs.py:191-193   */
  BNMTF_ST_ESD      = 5,  /* VB exp_square_diff                  bnmf_vb.py:241-244      */
  BNMTF_ST_ELBO     = 6,  /* VB elbo()                           bnmf_vb.py:191-232      */
  BNMTF_ST_IDIV     = 7,  /* NP compute_I_div()                  nmf_np.py:160-163       */
  BNMTF_ST_LOGTAU   = 8,  /* VB exp_logtau                                              */
  BNMTF_ST_ALPHA_S  = 9   /* alpha* of tau                                              */
};

const char* bnmtf_last_error(void);
int bnmtf_device_count(void);
/* ABI version of this library (bumped on any signature change). */
int bnmtf_abi_version(void);

/* ------------------------------------------------------------------------- *
 * Model handle.  Replaces the data side of the constructors
 * (bnmf_gibbs.py:61-101, bnmf_vb.py:55-96, nmf_icm.py:61-101, nmf_np.py:36-70):
 * copies R/M to the device, packs the observed entries (bit-exact values,
 * row-compressed for the U sweep and column-compressed for the V sweep) and
 * computes size_Omega and mean_Omega(R).
 *
 * Sharding (one process per GPU): this rank owns rows [row0,row1) of R for the
 * U sweep and columns [col0,col1) for the V sweep; `shard_rows`/`shard_cols`
 * are the equal shard sizes used by every rank (ceil(I/P), ceil(J/P)).  For a
 * single GPU pass row0=0,row1=I,col0=0,col1=J,shard_rows=I,shard_cols=J.
 * R and M are the FULL I x J matrices.
 * ------------------------------------------------------------------------- */
int bnmtf_nmf_create(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K,
                     const double* R, const uint8_t* M,
                     int64_t row0, int64_t row1, int64_t shard_rows,
                     int64_t col0, int64_t col1, int64_t shard_cols);

/* Same, from dense blocks already resident on `device`:
 *   Rrow_dev/Mrow_dev : rows [row0,row1) x all J columns, row-major
 *   Rcol_dev/Mcol_dev : all I rows x columns [col0,col1), row-major (ld = col1-col0)
 * The blocks are only read and may be freed after the call. */
int bnmtf_nmf_create_dev(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K,
                         const double* Rrow_dev, const uint8_t* Mrow_dev,
                         const double* Rcol_dev, const uint8_t* Mcol_dev,
                         int64_t row0, int64_t row1, int64_t shard_rows,
                         int64_t col0, int64_t col1, int64_t shard_cols);

/* Layout statistics of this rank's shards after create (out4: longest row of the U
 * shard, longest (row, column-tile) segment of it, and the same two for the V
 * shard), and the packing step.  The caller reduces the statistics with MAX over
 * ranks and passes the global values to bnmtf_nmf_pack, so that the packed
 * layout -- and every summation order -- is independent of the sharding (a
 * chain is bitwise identical on 1, 2, 4 or 8 GPUs).  global4 may be NULL for a
 * single rank.  The dense blocks given to bnmtf_nmf_create_dev must stay valid
 * until bnmtf_nmf_pack returns; every other entry point requires a packed
 * handle. */
int bnmtf_nmf_layout_stats(bnmtf_nmf_t h, int64_t* out4);
int bnmtf_nmf_pack(bnmtf_nmf_t h, const int64_t* global4);

int bnmtf_nmf_destroy(bnmtf_nmf_t h);

/* size_Omega, sum_Omega R, sum_Omega (R-mean)^2 of the LOCAL row shard
 * (bnmf_gibbs.py:74; compute_R2/compute_Rp :256-270).  out[3]. */
int bnmtf_nmf_data_stats(bnmtf_nmf_t h, double* out);
/* Global values after the wrapper has summed data_stats over ranks. */
int bnmtf_nmf_set_data_stats(bnmtf_nmf_t h, double size_Omega, double mean_R);

/* Join a multi-GPU job: `nccl_id` is the 128-byte ncclUniqueId created by rank
 * 0 (bnmtf_nccl_unique_id) and distributed by the caller (torch.distributed).
 * Loads libnccl lazily; not needed for world == 1. */
int bnmtf_nccl_unique_id(uint8_t* id128);
int bnmtf_nmf_set_comm(bnmtf_nmf_t h, int rank, int world, const uint8_t* nccl_id128);

/* Multi-GPU peer exchange (optional, after bnmtf_nmf_set_comm and before the first run; two-factor models): every rank
 * exports the CUDA IPC handles of its factor / statistics arrays (`blob`, BNMTF_PEER_BLOB_BYTES bytes), the caller
 * all-gathers the blobs (torch.distributed) and hands all of them -- world x blob_bytes, in rank order -- to every rank.
 * From then on the cluster sweeps write the rows they update straight into the other ranks' arrays over NVLink and a
 * device-side barrier replaces the NCCL all-gathers of the iteration.  Same chains, bit for bit.  No reference
 * counterpart (the reference is single-process numpy). */
#define BNMTF_PEER_BLOB_BYTES 448
int bnmtf_nmf_peer_export(bnmtf_nmf_t h, uint8_t* blob, int64_t blob_bytes);
int bnmtf_nmf_peer_import(bnmtf_nmf_t h, const uint8_t* blobs, int64_t blob_bytes);

/* Hyper-parameters (bnmf_gibbs.py:76-89).  lambdaU (I x K) / lambdaV (J x K)
 * are used when ARD == 0 and may be NULL when ARD != 0. */
int bnmtf_nmf_set_hyper(bnmtf_nmf_t h, int ARD, double alphatau, double betatau,
                        double alpha0, double beta0,
                        const double* lambdaU, const double* lambdaV);

/* Host-authoritative state: upload / download one factor array (n x K). */
int bnmtf_nmf_set_factor(bnmtf_nmf_t h, int side, int field, const double* src);
int bnmtf_nmf_get_factor(bnmtf_nmf_t h, int side, int field, double* dst);
/* tau (exp_tau for VB) and lambdak / exp_lambdak (K values; NULL when not ARD). */
int bnmtf_nmf_set_scalars(bnmtf_nmf_t h, double tau, const double* lambdak);
int bnmtf_nmf_get_scalars(bnmtf_nmf_t h, double* tau, double* lambdak);

/* VB variational parameters that are scalars / K-vectors (bnmf_vb.py:105-115,
 * 236-249,265-273): exp_tau, exp_logtau, alpha_s, beta_s and alphak_s, betak_s,
 * exp_lambdak, exp_loglambdak (each K values, NULL when not ARD).
 * get: scal4 = {exp_tau, exp_logtau, alpha_s, beta_s}; lamstate4K = the four
 * K-vectors back to back. */
int bnmtf_nmf_set_vb_scalars(bnmtf_nmf_t h, double exp_tau, double exp_logtau, double alpha_s, double beta_s,
                             const double* alphak_s, const double* betak_s,
                             const double* exp_lambdak, const double* exp_loglambdak);
int bnmtf_nmf_get_vb_scalars(bnmtf_nmf_t h, double* scal4, double* lamstate4K);

/* initialise() draws (bnmf_gibbs.py:110-132, bnmf_vb.py:105-142): fills VALUE
 * (MU for VB) of both factors with Exp(rate lambda) draws (random != 0) or
 * 1/lambda, lambda = lambdak (ARD) or lambdaU/V. */
int bnmtf_nmf_init_factors(bnmtf_nmf_t h, int method, int random, uint64_t seed);

/* Single-column conditionals from the CURRENT state, nothing is modified:
 *   Gibbs/ICM: tauU(k), muU(tauUk,k) / tauV(k), muV(tauVk,k)  bnmf_gibbs.py:203-219
 *   VB       : tau_U[:,k], mu_U[:,k] of update_U(k)/update_V(k) bnmf_vb.py:251-261
 *   NP       : out_mu = the updated column of update_U(k)/update_V(k)
 *              (nmf_np.py:120-126), out_tau = its denominator.
 * out_tau/out_mu have n = I (side U) or J (side V) entries.  Single-GPU only. */
int bnmtf_nmf_column_params(bnmtf_nmf_t h, int method, int side, int k,
                            double* out_tau, double* out_mu);

/* Residual statistics of the CURRENT state into out[BNMTF_NSTATS]:
 * RSS / sums for beta_s(), predict_while_running(), exp_square_diff(), elbo(),
 * compute_I_div().  Nothing is modified. */
int bnmtf_nmf_stats(bnmtf_nmf_t h, int method, double* out);

/* run(iterations)  (bnmf_gibbs.py:135-183, nmf_icm.py:135-186,
 * bnmf_vb.py:145-188, nmf_np.py:97-117), starting from the state on the
 * handle.  Trace pointers may be NULL:
 *   all_U   [iterations][I][K], all_V [iterations][J][K]   (Gibbs/ICM draws)
 *   all_lambdak [iterations][K]                             (ARD)
 *   iter_stats  [iterations][BNMTF_NSTATS]
 *   iter_seconds[iterations]  cumulative device-side seconds (all_times)
 * `seed` keys the Philox streams (Gibbs only). */
int bnmtf_nmf_run(bnmtf_nmf_t h, int method, int iterations, uint64_t seed,
                  double* all_U, double* all_V, double* all_lambdak,
                  double* iter_stats, double* iter_seconds);

/* number of kernels launched by this handle so far, and layout information, 16
 * values: [0]=threads per row team U, [1]=entries per thread U, [2],[3] same for
 * V, [4],[5]=padded entries U,V (row layout), [6],[7]=observed entries of the
 * local U,V shards, [8..11]=cluster-tile layout of the U side (in use, entries
 * per lane, rows per set, padded entries), [12..15] same for the V side. */
int64_t bnmtf_nmf_launch_count(bnmtf_nmf_t h);
int bnmtf_nmf_layout_info(bnmtf_nmf_t h, int64_t* out16);
/* on != 0: every sweep launch is bracketed by CUDA events (and synchronised) so that bnmtf_nmf_sweep_times can report the
 * kernel time per side -- measurement passes only, it serialises the run loop; resets the accumulated times.  Default: the
 * environment variable BNMTF_TIME_SWEEPS when the handle is created (bench.py's roofline leg; no reference counterpart). */
int bnmtf_nmf_time_sweeps(bnmtf_nmf_t h, int on);
/* device-event milliseconds spent in the U sweep / V sweep kernels of the last
 * bnmtf_nmf_run (sum over iterations), and their launch counts. out[4]. */
int bnmtf_nmf_sweep_times(bnmtf_nmf_t h, double* out4);

/* ------------------------------------------------------------------------- *
 * Tri-factorisation R ~ F S G^T (code/models/bnmtf_gibbs.py, nmtf_icm.py).
 * The handle type is shared with the NMF family: side BNMTF_SIDE_U is F (I x K),
 * side BNMTF_SIDE_V is G (J x L); layout_stats / pack / data_stats / set_comm /
 * set_factor / get_factor / destroy / launch_count / layout_info apply unchanged.
 * bnmtf_nmtf_run / column_params serve BNMTF_GIBBS (bnmtf_gibbs.py) and BNMTF_ICM (nmtf_icm.py);
 * the VB and NP variants have their own entry points below.
 * Limits (the reference has none; bnmtf_nmtf_create* fails with a message beyond them): L <= 48, K * L <= 3631.
 * ------------------------------------------------------------------------- */
int bnmtf_nmtf_create(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, int L,
                      const double* R, const uint8_t* M,
                      int64_t row0, int64_t row1, int64_t shard_rows,
                      int64_t col0, int64_t col1, int64_t shard_cols);
int bnmtf_nmtf_create_dev(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, int L,
                          const double* Rrow_dev, const uint8_t* Mrow_dev,
                          const double* Rcol_dev, const uint8_t* Mcol_dev,
                          int64_t row0, int64_t row1, int64_t shard_rows,
                          int64_t col0, int64_t col1, int64_t shard_cols);
/* bnmtf_gibbs.py:76-99: lambdaF (I x K) / lambdaG (J x L) when ARD == 0; lambdaS (K x L) always. */
int bnmtf_nmtf_set_hyper(bnmtf_nmf_t h, int ARD, double alphatau, double betatau, double alpha0, double beta0,
                         const double* lambdaF, const double* lambdaG, const double* lambdaS);
/* S (K x L row-major); field as BNMTF_F_*. */
int bnmtf_nmtf_set_S(bnmtf_nmf_t h, int field, const double* src);
int bnmtf_nmtf_get_S(bnmtf_nmf_t h, int field, double* dst);
/* tau, lambdaFk (K), lambdaGl (L); vectors may be NULL. */
int bnmtf_nmtf_set_scalars(bnmtf_nmf_t h, double tau, const double* lambdaFk, const double* lambdaGl);
int bnmtf_nmtf_get_scalars(bnmtf_nmf_t h, double* tau, double* lambdaFk, double* lambdaGl);
/* initialise(init_FG, init_S) draws (bnmtf_gibbs.py:121-164; 'random' -> Exp(rate lambda), 'exp' -> 1/lambda). */
int bnmtf_nmtf_init(bnmtf_nmf_t h, int method, int random_FG, int random_S, uint64_t seed_FG, uint64_t seed_S);
/* Single conditionals from the current state, nothing modified (bnmtf_gibbs.py:268-292):
 *   which 0: tauF(k), muF(.,k) (I values)   which 1: tauG(l), muG(.,l) (J values)
 *   which 2: tauS(k,l), muS(.,k,l) (one value each).  Single-GPU only. */
int bnmtf_nmtf_column_params(bnmtf_nmf_t h, int method, int which, int k, int l, double* out_tau, double* out_mu);
/* beta_s() / predict_while_running() sums of the current state into out[BNMTF_NSTATS]. */
int bnmtf_nmtf_stats(bnmtf_nmf_t h, int method, double* out);
/* run(iterations) (bnmtf_gibbs.py:170-229, nmtf_icm.py:172-232).  Traces may be NULL:
 * all_F [it][I][K], all_S [it][K][L], all_G [it][J][L], all_lambdaFk [it][K], all_lambdaGl [it][L]. */
int bnmtf_nmtf_run(bnmtf_nmf_t h, int method, int iterations, uint64_t seed,
                   double* all_F, double* all_S, double* all_G, double* all_lambdaFk, double* all_lambdaGl,
                   double* iter_stats, double* iter_seconds);

/* BNMTF VB (code/models/bnmtf_vb.py).  State: set_factor/get_factor (fields VALUE=exp, VAR,
 * MU, TAU of F and G), bnmtf_nmtf_set_S/get_S (same fields of S) and the scalars below:
 * scal4 = {exp_tau, exp_logtau, alpha_s, beta_s}; stateF4K = alphaFk_s, betaFk_s, exp_lambdaFk,
 * exp_loglambdaFk back to back (4 x K), stateG4L likewise (4 x L); both NULL when not ARD. */
int bnmtf_nmtf_set_vb_scalars(bnmtf_nmf_t h, double exp_tau, double exp_logtau, double alpha_s, double beta_s,
                              const double* stateF4K, const double* stateG4L);
int bnmtf_nmtf_get_vb_scalars(bnmtf_nmf_t h, double* scal4, double* stateF4K, double* stateG4L);
/* update_F(k) / update_G(l) / update_S(k,l) from the current state, nothing modified
 * (bnmtf_vb.py:336-375): which 0 -> tau_F[:,k], mu_F[:,k] (I values); 1 -> tau_G[:,l], mu_G[:,l]
 * (J values); 2 -> tau_S[k,l], mu_S[k,l] (one value each).  Single-GPU only. */
int bnmtf_nmtf_vb_update_params(bnmtf_nmf_t h, int which, int k, int l, double* out_tau, double* out_mu);
/* exp_square_diff() (:319-324), elbo() (:248-299) and the training-metric sums of the current state. */
int bnmtf_nmtf_vb_stats(bnmtf_nmf_t h, double* out);
/* run(iterations) (bnmtf_vb.py:193-245); read the final state back with the getters. */
int bnmtf_nmtf_vb_run(bnmtf_nmf_t h, int iterations, double* iter_stats, double* iter_seconds);

/* NMTF NP (code/models/nmtf_np.py): single multiplicative updates from the current state, nothing
 * modified (nmtf_np.py:165-188): which 0 -> new F[:,k] (I values), 1 -> new G[:,l] (J values),
 * 2 -> new S[k,l] (one value); and run(iterations) (:139-160).  bnmtf_nmtf_stats(h, BNMTF_NP, out)
 * gives compute_I_div() and the metric sums. */
int bnmtf_nmtf_np_update(bnmtf_nmf_t h, int which, int k, int l, double* out);
int bnmtf_nmtf_np_run(bnmtf_nmf_t h, int iterations, double* iter_stats, double* iter_seconds);

/* ------------------------------------------------------------------------- *
 * Masked K-means (code/models/kmeans/kmeans.py) on the packed layouts of a handle, used by
 * init_FG='kmeans' (bnmtf_gibbs.py:140-151).  side 0 clusters the rows of R (m = J coordinates),
 * side 1 the columns (KMeans(R.T, M.T, L), m = I).  centroids / mask_centroids are [Kc][m].
 *   assign : assignment() + closest_cluster()     kmeans.py:88-125  -> cluster index and MSE per point
 *   update : update_cluster() for every cluster with cluster_flag[c] != 0   kmeans.py:131-170
 *   minmax : per-coordinate min / max of the observed values               kmeans.py:50-52
 * Single-GPU only. */
int bnmtf_kmeans_assign(bnmtf_nmf_t h, int side, int Kc, const double* centroids, const uint8_t* mask_centroids,
                        int32_t* assign_out, double* dist_out);
int bnmtf_kmeans_update(bnmtf_nmf_t h, int side, int Kc, const int32_t* assign, const int32_t* cluster_flag,
                        double* centroids, uint8_t* mask_centroids);
int bnmtf_kmeans_minmax(bnmtf_nmf_t h, int side, double* mins, double* maxs);

/* ------------------------------------------------------------------------- *
 * On-device posterior summaries and held-out prediction (both handle kinds).
 *
 * bnmtf_nmf_set_summary: from the next bnmtf_nmf_run / bnmtf_nmtf_run on, add the draws of the
 *   iterations range(burn_in, iterations, thinning) into running sums on the device -- the
 *   quantities approx_expectation averages from all_U / all_V / all_tau / all_lambdak
 *   (bnmf_gibbs.py:222-230; F, S, G, lambdaFk, lambdaGl for bnmtf_gibbs.py:302-315), in the same
 *   summation order.  The run may then be called with NULL trace pointers.  thinning <= 0 turns
 *   the summary off and frees it.  Every run restarts the sums.
 * bnmtf_nmf_get_summary: mean (sum / count) of one field into a host array; `count` (optional)
 *   receives the number of accumulated draws.  dst may be NULL to query the count only.
 * bnmtf_nmf_predict: predict(M_pred, ...) (bnmf_gibbs.py:233-240, bnmf_vb.py:265-271,
 *   bnmtf_gibbs.py:318-326) on n test entries given as (row, column, value): R_pred from the
 *   current factors (BNMTF_PRED_CURRENT: U V^T, or F S G^T) or from the summary means
 *   (BNMTF_PRED_SUMMARY), then compute_MSE / compute_R2 / compute_Rp (bnmf_gibbs.py:252-270)
 *   with the two-pass (centred) sums of the reference.
 *   out8 = {n, MSE, R^2, Rp, SS_res, SS_total, mean_real, mean_pred}.
 * ------------------------------------------------------------------------- */
enum {
  BNMTF_SUM_FACTOR_U = 0, BNMTF_SUM_FACTOR_V = 1, BNMTF_SUM_S = 2,
  BNMTF_SUM_LAMBDA_U = 3, BNMTF_SUM_LAMBDA_V = 4, BNMTF_SUM_TAU = 5
};
enum { BNMTF_PRED_CURRENT = 0, BNMTF_PRED_SUMMARY = 1 };
int bnmtf_nmf_set_summary(bnmtf_nmf_t h, int burn_in, int thinning);
int bnmtf_nmf_get_summary(bnmtf_nmf_t h, int field, double* dst, int64_t* count);
int bnmtf_nmf_predict(bnmtf_nmf_t h, int source, int64_t n, const int32_t* rows, const int32_t* cols,
                      const double* vals, double* out8);

/* ------------------------------------------------------------------------- *
 * Batched run of many small NMF models: the fits of a cross-validation sweep, one per (fold,
 * parameter set), which the reference runs one after the other or in a multiprocessing.Pool
 * (code/cross_validation/matrix_cross_validation.py:83-101, parallel_matrix_cross_validation.py:53-66,
 * nested_matrix_cross_validation.py:47-80).
 *
 * handles[0..n): NMF handles (bnmtf_nmf_create) on ONE device, single-GPU, each with its hyper-
 *   parameters and state set exactly as before bnmtf_nmf_run, and all packed with the same global
 *   maxima (bnmtf_nmf_layout_stats of every model -> element-wise MAX -> bnmtf_nmf_pack) so that they
 *   share one row-team layout.  The models may differ in K, in their masks and in their shapes.
 * One launch of each update kernel (lambda_k, U sweep, V sweep, tau / ELBO) serves the whole batch and
 *   the iteration is replayed as one CUDA graph; model m follows bit for bit the chain
 *   bnmtf_nmf_run(handles[m], method, iterations, seeds[m], ...) gives it.
 * seeds[n] (Gibbs; may be NULL otherwise); iter_stats[m] -> [iterations][BNMTF_NSTATS] of model m (entries
 *   or the array may be NULL); iter_seconds[iterations]: cumulative device seconds of the batch.
 * Draw traces (all_U, all_V) are not produced: Gibbs / ICM models that need posterior means configure
 *   bnmtf_nmf_set_summary first.  Read the final states back with the per-handle getters.
 * launches (optional): kernels launched by this call.
 * ------------------------------------------------------------------------- */
int bnmtf_batch_run(int n, const bnmtf_nmf_t* handles, int method, int iterations, const uint64_t* seeds,
                    double* const* iter_stats, double* iter_seconds, int64_t* launches);
/* Optional, before creating the handles of a large batch: grow the device's stream-ordered memory pool (which
 * every allocation of this library comes from) to `bytes` of idle memory in one step instead of a few
 * megabytes per first-time allocation.  Best effort: returns 0 also when the device cannot spare that much. */
int bnmtf_pool_reserve(int device, int64_t bytes);

/* ------------------------------------------------------------------------- *
 * Distributions (code/models/distributions/*.py) on the device; n-element
 * host arrays in, host arrays out.  Draws use Philox4x32-10 keyed by `seed`.
 *   bnmtf_tn_draw        TN_vector_draw        truncated_normal_vector.py:37-50
 *   bnmtf_tn_moments     TN_vector_expectation/variance          :53-73
 *   bnmtf_gamma_draw     gamma_draw(alpha, beta) (shape, RATE)    gamma.py:11-14
 *   bnmtf_exp_draw       exponential_draw(lambda) (RATE)          exponential.py:7-9
 * ------------------------------------------------------------------------- */
int bnmtf_tn_draw(int device, uint64_t seed, int64_t n, const double* mu, const double* tau, double* out);
int bnmtf_tn_moments(int device, int64_t n, const double* mu, const double* tau, double* exp_out, double* var_out);
int bnmtf_gamma_draw(int device, uint64_t seed, int64_t n, const double* alpha, const double* beta, double* out);
int bnmtf_exp_draw(int device, uint64_t seed, int64_t n, const double* lambda, double* out);

#ifdef __cplusplus
}
#endif
#endif /* BNMTF_B200_H */
