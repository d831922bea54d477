/* bnmf.h -- C ABI of the B200-native Gibbs sampler behind jennalandy/bayesNMF.
 *
 * The reference is pure R and has no FFI; the seam this library replaces is the
 * private-method layer of the R6 class `bayesNMF_sampler`
 *     private$sample_prior_params()   R/bayesNMF_sampler.R:575-577 -> R/sample_priors.R:150-200
 *     private$sample_params()         R/bayesNMF_sampler.R:598-600 -> R/sample_params.R:51-89
 *     private$record_sample()         R/bayesNMF_sampler.R:651-672
 *     private$update_sample_metrics() R/bayesNMF_sampler.R:699-704 -> R/utils.R:339-348, :412-455
 * as called from initialize() (R/bayesNMF_sampler.R:232-257) and from the loop body
 * of run_gibbs_sampler() (R/bayesNMF_sampler.R:268-285, :337-348).
 *
 * Conventions (what an R `.Call` veneer, or ctypes, relies on):
 *   - plain C, no C++/torch types; every function returns 0 on success, non-zero on
 *     error, never throws or longjmps; bnmf_last_error() gives the message;
 *   - all host buffers are caller-owned, column-major, double (R's REALSXP) unless
 *     stated; the library copies in/out and never keeps a host pointer;
 *   - one host thread per handle; a handle is bound to one CUDA device;
 *   - indices are 0-based here; iteration numbers are the reference's (1 = the
 *     prior draw, R/bayesNMF_sampler.R:39-43).
 */
#ifndef BNMF_H
#define BNMF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bnmf_handle bnmf_handle;

enum { BNMF_POISSON = 0, BNMF_NORMAL = 1 };                         /* specs$likelihood */
enum { BNMF_TRUNCNORMAL = 0, BNMF_EXPONENTIAL = 1, BNMF_GAMMA = 2 };/* specs$prior      */
enum { BNMF_SBFI = 0, BNMF_BFI = 1, BNMF_BIC = 2 };                 /* specs$rank_method*/
enum { BNMF_F64 = 0, BNMF_F32 = 1 };                                /* state precision  */

/* columns of one sample_metrics row (R/utils.R:435-452, :341-342) */
enum {
  BNMF_MC_ITER = 0, BNMF_MC_RMSE, BNMF_MC_KL, BNMF_MC_LOGLIK, BNMF_MC_LOGPOST,
  BNMF_MC_NPARAMS, BNMF_MC_BIC, BNMF_MC_RANK, BNMF_MC_TEMP,
  BNMF_MC_PACC, BNMF_MC_EACC, BNMF_MC_COLS
};

/* bits of the `have` mask of bnmf_init_from_prior = names(init_params) of
 * R/bayesNMF_sampler.R:241 (sample_params(skip = names(init_params), from_prior = TRUE)) */
enum { BNMF_HAVE_P = 1, BNMF_HAVE_E = 2, BNMF_HAVE_A = 4, BNMF_HAVE_Z = 8, BNMF_HAVE_SIGMASQ = 16 };

typedef struct bnmf_config {
  int32_t K;             /* nrow(data)                        R/bayesNMF_sampler.R:142 */
  int32_t N;             /* max(rank)                         R/bayesNMF_sampler.R:143 */
  int64_t G;             /* ncol(data) held by THIS handle (its shard)                */
  int64_t G_total;       /* ncol(data) of the whole problem (== G when not sharded)   */
  int64_t g0;            /* global index of this shard's first column                 */
  int32_t likelihood;    /* BNMF_POISSON | BNMF_NORMAL                                */
  int32_t prior;         /* BNMF_TRUNCNORMAL | BNMF_EXPONENTIAL | BNMF_GAMMA          */
  int32_t MH;            /* specs$MH                                                  */
  int32_t learning_rank; /* specs$learning_rank  (length(rank) > 1)                   */
  int32_t rank_method;   /* BNMF_SBFI | BNMF_BFI   (BIC fans out on the host)         */
  int32_t precision;     /* BNMF_F64 (parity) | BNMF_F32                              */
  int32_t device;        /* CUDA device ordinal                                       */
  int32_t ring_cap;      /* samples kept on the device = MAP_over (0: keep none)      */
  uint64_t seed;         /* Philox key                                                */
} bnmf_config;

/* Check that (likelihood, prior, MH) is a model the reference accepts
 * (check_model, R/bayesNMF_sampler.R:623-645).  Needs no GPU.  0 = valid. */
int bnmf_check_model(int likelihood, int prior, int MH, char* msg, size_t msg_len);

/* Create a sampler: upload `data` (K x G column-major doubles; must be integer-valued
 * and >= 0 for the Poisson likelihood) and allocate all state in HBM.
 * Called where initialize() stores self$data/self$dims (R/bayesNMF_sampler.R:140-158). */
int bnmf_create(const bnmf_config* cfg, const double* data, bnmf_handle** out);
void bnmf_destroy(bnmf_handle* h);
/* message of the last failing call on this thread */
const char* bnmf_last_error(void);

/* Hyperprior parameters, scalar (rows = cols = 1) or full matrix, as filled by
 * fill_hyperprior_params_ (R/setup.R:15-88).  bnmf_create installs the reference's
 * data-dependent defaults (R/setup.R:123-181, from mean(data) of the handle's columns and
 * N); this call overrides them.  "alpha" / "beta" (scalars): sigmasq prior, default 3.  Names: "A_p","B_p","C_p","D_p","M_p",
 * "S_p" (K x N) and "A_e",...,"S_e" (N x G).  Matrices for *_e cover this shard. */
int bnmf_set_hyper(bnmf_handle* h, const char* name, const double* v, int64_t rows, int64_t cols);

/* Current state, by the reference's names.  Parameters: "P" (K x N), "E" (N x G),
 * "A" (1 x N), "R" (1), "sigmasq" (G).  Prior parameters: "Mu_p","Sigmasq_p",
 * "Lambda_p","Alpha_p","Beta_p" (K x N); "Mu_e",...,"Beta_e" (N x G); "Alpha","Beta"
 * (G, sigmasq prior).  Statistics replacing Z: "SP" = sum_g Z (K x N), "SE" = sum_k Z
 * (N x G).  "P_acceptance_rate" (K x N), "E_acceptance_rate" (N x G), "Mhat" (K x G).
 * Read-only: "rowsumE" (N; also settable: the fixed-point row sums of E), "data_sum" (1: the sum of this handle's data).
 * NaN entries in a prior parameter mean "draw this column from the hyperprior"
 * (R/sample_priors.R:33-60). */
int bnmf_set_state(bnmf_handle* h, const char* name, const double* v, int64_t len);
int bnmf_get_state(bnmf_handle* h, const char* name, double* out, int64_t len);

/* temperature_schedule (R/utils.R:307-332), indexed by iteration (entry 0 = iter 1). */
int bnmf_set_temperature_schedule(bnmf_handle* h, const double* temps, int64_t n);

/* Iteration 1: init_prior_params_ + sample_params_(from_prior = TRUE) + record_sample
 * + update_sample_metrics (R/bayesNMF_sampler.R:232-257).  States named in `have`
 * were supplied with bnmf_set_state and are kept verbatim; prior parameters that
 * were supplied are kept except for NaN columns.  metrics_row: BNMF_MC_COLS doubles. */
int bnmf_init_from_prior(bnmf_handle* h, uint32_t have, uint32_t have_prior, double* metrics_row);

/* Advance n_iters full Gibbs iterations on the device (prior parameters -> P -> E
 * -> R,A -> Z -> sigmasq -> record -> metrics; R/bayesNMF_sampler.R:273-285).
 *   converged   state$converged: real Metropolis-Hastings accept step on/off
 *               (R/sample_Pn.R:201-204, R/sample_En.R:198-201)
 *   metrics_out n_iters x BNMF_MC_COLS, row-major (one sample_metrics row per iter)
 *   P_out       NULL or K x N x n_iters   (samples$P of these iterations)
 *   A_out       NULL or N x n_iters       (samples$A)
 * Other samples stay in the device ring (bnmf_get_samples). */
int bnmf_step(bnmf_handle* h, int32_t n_iters, int32_t converged,
              double* metrics_out, double* P_out, double* A_out);

/* Sample ring (record_sample / update_list, R/bayesNMF_sampler.R:651-672,
 * R/helpers.R:111-119): the last `ring_cap` samples of "P", "E", "A".
 * `ago` = 0 is the newest sample. */
int bnmf_ring_count(bnmf_handle* h, int32_t* count);
int bnmf_get_sample(bnmf_handle* h, const char* name, int32_t ago, double* out, int64_t len);

/* get_MAP_ on the device (R/utils.R:194-288): over the newest `n_samples` ring
 * samples, keep those whose A equals the modal A, renormalise (R/helpers.R:35-49)
 * and average.  P_map K x N, E_map N x G, A_map N; n_match = samples averaged. */
int bnmf_get_map(bnmf_handle* h, int32_t n_samples, double* P_map, double* E_map,
                 double* A_map, int32_t* n_match);

/* The credible intervals of get_MAP_ (R/utils.R:264-287: apply(arr, c(1, 2), quantile, probs)) on
 * the device: over the same matching samples, element-wise quantiles (type 7) `lower_p` and
 * `upper_p` of the renormalised P (K x N) and E (N x G).  Any output may be NULL. */
int bnmf_get_credible_intervals(bnmf_handle* h, int32_t n_samples, double lower_p, double upper_p,
                                double* P_lower, double* P_upper, double* E_lower, double* E_upper,
                                int32_t* n_match);

/* assign_signatures_ensemble_ (R/postprocessing.R:175-341) on the retained samples: over the
 * newest `n_samples` ring samples that match the modal A, and over the included signatures
 * (keep_sigs: A == 1), every sample's P is assigned to the reference signatures by the Hungarian
 * algorithm on minus the cosine similarity (hungarian_assignment, R/helpers.R:287-362; the cosines
 * are computed on the device), an assignment votes with its cosine, the reference with the largest
 * share of an estimated signature's votes wins; MAP_cosine is the cosine of the MAP signature to its
 * winner, lower / upper the type-7 quantiles (1 +- credible_interval) / 2 of the samples' cosines to it.
 *   reference_P  K x n_ref, column-major (e.g. the 79 COSMIC SBS signatures)
 *   keep_sigs    N      0-based indices of the included signatures (first *n_keep entries valid)
 *   votes        N x n_ref, column-major with leading dimension N: votes[i + N*j] = share of the votes of
 *                estimated signature keep_sigs[i] that went to reference j (rows i < *n_keep)
 *   assignment   N      0-based reference index per included signature
 *   map_cosine, lower_cosine, upper_cosine   N each (first *n_keep entries valid) */
int bnmf_assign_signatures(bnmf_handle* h, int32_t n_samples, const double* reference_P, int32_t n_ref,
                           double credible_interval, int32_t* n_keep, int32_t* keep_sigs, double* votes,
                           int32_t* assignment, double* map_cosine, double* lower_cosine, double* upper_cosine,
                           int32_t* n_match);

/* run_gibbs_sampler (R/bayesNMF_sampler.R:265-408) with the convergence control of
 * R/convergence.R:60-154 and update_MAP_metrics_ (R/utils.R:356-397) behind the ABI: blocks of
 * iterations up to the next MAP check, the window mean of the metric, percent change, the
 * "no change" / "no best" / "max iters" rules (only once every temperature of the window is 1 and
 * iter >= miniters), then -- Metropolis-Hastings models -- `post_warmup` iterations with the real
 * accept step.  One call replaces a host round trip every MAP_every iterations.
 * The metric is a window mean of sample_metrics (logposterior, loglikelihood) or the BIC built
 * from it; RMSE / KL of the MAP reconstruction are not offered here (use bnmf_step + bnmf_get_map). */
enum { BNMF_METRIC_LOGPOSTERIOR = 0, BNMF_METRIC_LOGLIKELIHOOD = 1, BNMF_METRIC_BIC = 2 };
enum { BNMF_WHY_NONE = 0, BNMF_WHY_NO_CHANGE = 1, BNMF_WHY_NO_BEST = 2, BNMF_WHY_MAX_ITERS = 3 };
typedef struct bnmf_convergence_control {   /* new_convergence_control, R/convergence.R:16-45 */
  int32_t MAP_over, MAP_every;
  double tol;
  int32_t Ninarow_nochange, Ninarow_nobest, miniters, maxiters;
  int32_t metric;                           /* BNMF_METRIC_* */
} bnmf_convergence_control;
/* one row of state$MAP_metrics per check */
enum { BNMF_MM_ITER = 0, BNMF_MM_LOGLIK, BNMF_MM_LOGPOST, BNMF_MM_NPARAMS, BNMF_MM_BIC, BNMF_MM_RANK,
       BNMF_MM_A_COUNTS, BNMF_MM_MEAN_TEMP, BNMF_MM_COLS };
typedef struct bnmf_run_result {
  int32_t iter;             /* state$iter at return */
  int32_t converged, converged_iter, why;   /* BNMF_WHY_* */
  int32_t best_iter, n_checks, n_rows;
  int32_t inarow_no_change, inarow_no_best, inarow_na;
  double best_MAP_metric, prev_MAP_metric;
} bnmf_run_result;
/*   metrics_out      NULL or rows_cap x BNMF_MC_COLS: the sample_metrics rows of the iterations run
 *   map_metrics_out  NULL or checks_cap x BNMF_MM_COLS: the MAP_metrics rows of the checks made
 * The handle needs ring_cap >= MAP_over (the MAP of a check comes from the device ring). */
int bnmf_run(bnmf_handle* h, const bnmf_convergence_control* cc, int32_t post_warmup,
             double* metrics_out, int64_t rows_cap, double* map_metrics_out, int64_t checks_cap,
             bnmf_run_result* result);

/* Cross-shard reduction for genome-sharded runs (one process per GPU): each rank
 * creates its shard handle, rank 0 makes an id, every rank joins.  Afterwards
 * bnmf_step sums SP, rowSums(E) and the metric partials over ranks with NCCL. */
int bnmf_comm_unique_id(char* id128);
int bnmf_comm_init(bnmf_handle* h, const char* id128, int32_t rank, int32_t world);
/* Let `h` use the communicator `src` (same process, same device) already joined: further
 * chains or the fixed-rank samplers of a BIC fan-out need no second rendezvous. */
int bnmf_comm_share(bnmf_handle* h, bnmf_handle* src);

/* Device timing of the last bnmf_step call (CUDA events on the handle's stream), in
 * milliseconds: total = the whole call; iter = sum over iterations of the span from
 * the first to the last kernel of the iteration; zstat = share of the latent-count
 * kernel (events inside the iteration's chain: recorded when the L2 flush below is on
 * or the environment says BNMF_TIMING=z, else 0); launches = kernels launched. */
int bnmf_timing(bnmf_handle* h, double* total_ms, double* iter_ms, double* zstat_ms, int64_t* launches);

/* Benchmark hygiene: when bytes > 0, bnmf_step overwrites a scratch buffer of that size
 * before every iteration (outside the spans bnmf_timing reports as iter_ms) so that each
 * iteration starts with a cold L2.  0 switches it off (the default). */
int bnmf_set_l2_flush(bnmf_handle* h, size_t bytes);

/* Stand-alone entry for the fused latent-count + sufficient-statistic kernel on the
 * handle's current P, E, A at iteration `iter` (parity tests, roofline timing).
 * Overwrites SP / SE. */
int bnmf_sample_z(bnmf_handle* h, int32_t iter, double* kernel_ms);

/* Measurement: advance ONE iteration (as bnmf_step(h, 1, converged, ...) would, without the CUDA-graph replay
 * and without the side-stream overlap) with a CUDA event on the launching stream after every kernel, and
 * return the device time per kernel name: names = cap x 32 chars, ms = total of the kernel's launches,
 * counts = its launches; *n = distinct kernels (may exceed cap).  What bench.py's `roofline` object reads. */
int bnmf_profile_iteration(bnmf_handle* h, int32_t converged, char* names, double* ms, int32_t* counts,
                           int32_t cap, int32_t* n);

/* bnmf_destroy keeps a handle's device blocks in a process-wide cache (at most BNMF_CACHE_MB MiB,
 * default 4096) so that the next sampler -- bayesNMF() builds one per rank, R/bayesNMF.R -- starts
 * without allocator calls.  This hands the cached blocks of every device back to the driver. */
int bnmf_release_cached_memory(void);

#ifdef __cplusplus
}
#endif
#endif /* BNMF_H */
