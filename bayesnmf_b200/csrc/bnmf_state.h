// Device-resident sampler state shared by all kernels (POD, passed by value).
//
// Layout mirrors the reference's R6 fields (column-major everywhere, as in R):
//   self$data               -> M     K x G   (R/bayesNMF_sampler.R:140)
//   self$params$P / E / A   -> P K x N, E N x G, A N   (R/sample_params.R:16-41)
//   self$params$Z           -> never materialised; only its two margins
//                              SP = sum_g Z (K x N) and SE = sum_k Z (N x G) exist,
//                              the only things R/sample_Pn.R:101,109 and
//                              R/sample_En.R:100,108 ever read.
//   self$prior_params$*     -> Mu/Sigmasq/Lambda/Alpha/Beta _p (K x N), _e (N x G)
//   self$hyperprior_params$*-> scalar-or-matrix (R/setup.R:102-113 only fills a
//                              matrix when the user did not pass one)
// G here is the number of columns resident on THIS GPU (a contiguous shard
// [g0, g0+G) of G_total); P-side quantities are replicated on every shard.
#pragma once
#include <stdint.h>

namespace bnmf {

enum Likelihood : int { LIK_POISSON = 0, LIK_NORMAL = 1 };
enum Prior : int { PRIOR_TRUNCNORMAL = 0, PRIOR_EXPONENTIAL = 1, PRIOR_GAMMA = 2 };
enum RankMethod : int { RANK_SBFI = 0, RANK_BFI = 1, RANK_BIC = 2 };

// scalar-or-matrix hyperparameter
template <typename T> struct Hyper {
  const T* p;
  int is_matrix;
  __host__ __device__ __forceinline__ T at(long long i) const { return p[is_matrix ? i : 0]; }
};

// metric row columns (R/utils.R:435-452 + :341-342)
enum MetricCol : int {
  MC_ITER = 0, MC_RMSE, MC_KL, MC_LOGLIK, MC_LOGPOST, MC_NPARAMS, MC_BIC, MC_RANK, MC_TEMP,
  MC_PACC, MC_EACC, MC_COLS
};

// double partial sums reduced across tiles / shards each iteration
enum PartialCol : int {
  PC_SSE = 0,     // sum (Mhat - M)^2
  PC_KLV,         // variable part of padded KL:  - sum M' log Mhat'
  PC_LLV,         // variable part of the log-likelihood
  PC_LP_E,        // log prior of E
  PC_EACC,        // sum of E acceptance rates over active signatures
  PC_COLS
};

struct Ctrl {
  int iter;        // current iteration (1 = the prior draw, R/bayesNMF_sampler.R:39-43)
  int converged;   // state$converged: turns the real MH accept step on (R/sample_Pn.R:201)
  int row;         // row of `metrics` this iteration writes (reset per bnmf_step chunk)
  int ring_pos;    // next slot of the sample ring
  int ring_count;  // samples currently held
};

template <typename T> struct Dev {
  int K, N, G;              // G = local columns
  long long G_total, g0;    // global column count / offset of this shard
  int likelihood, prior, MH, learning_rank, rank_method;
  uint64_t seed;

  // data
  const int32_t* Mi;        // K x G counts (Poisson)
  const int32_t* Mt;        // G x K: the same, genome-major (k_zstat reads the counts of 32 genomes as one line)
  const int32_t* korder;    // K: mutation types by descending total count (the order k_zstat visits them in)
  const T* Mr;              // K x G reals  (Normal)
  double ll_const;          // - sum lgamma(M+1) over THIS shard's columns (Poisson)
  double ll_const_all;      // the same over all shards: what the rank learner adds to its two log-likelihoods,
                            // identical on every rank so that replicated draws of A_n cannot diverge
  double kl_const;          //   sum M' log M', M' = max(M, 1e-6)

  // parameters
  T* P; T* E; int32_t* A; int32_t* R; T* sigmasq;
  // prior parameters
  T* Mu_p; T* Sigmasq_p; T* Lambda_p; T* Alpha_p; T* Beta_p;
  T* Mu_e; T* Sigmasq_e; T* Lambda_e; T* Alpha_e; T* Beta_e;
  T* Alpha_g; T* Beta_g;    // sigmasq prior (R/sample_priors.R:133-140)
  // hyperprior parameters
  Hyper<T> A_p, B_p, C_p, D_p, M_p, S_p;
  Hyper<T> A_e, B_e, C_e, D_e, M_e, S_e;

  // latent-count margins
  unsigned long long* SP;   // K x N   (int64 so the cross-shard sum is exact)
  int32_t* SE;              // N x G
  // float reductions, replicated after the cross-shard sum
  long long* rowsumE_fx;    // N, fixed point 2^-24 (order-independent => bit-reproducible)
  T* colsumP;               // N
  double* lp_P;             // 1: log prior of P
  double* pacc_sum;         // 1: sum of P acceptance rates over active signatures

  // acceptance rates (MH)
  T* P_acc; T* E_acc;
  // running reconstruction Mhat = P diag(A) E (MH / Normal / rank-learning paths) and the
  // scratch of the sequential sweeps over signatures (bnmf_mh.cuh)
  T* Mhat;
  double* dvec;      // K: pending rank-1 update  Mhat[k,g] += dvec[k] * E[n_prev,g]
  double* prop;      // K: proposal for column n of P awaiting the accept step
  double* ppart;     // [n_gchunks][K][2] partial sums over genome chunks (P sweep)
  int n_gchunks, gchunk;
  double* apart;     // [n_eblocks][2] log-likelihood partials with A_n = 0 / 1
  int* nzE;          // [2][N] "row n of E has a non-zero" flags, by iteration parity
  int* nzP;          // [N]    "column n of P has a non-zero" flags
  double* paccpart;  // [N] sum_k P_acceptance_rate[k,n]

  // per-tile deterministic partials
  double* zpart;   int n_zitems;     // [n_zitems][PC_COLS] written by the column kernels
  double* epart;   int n_eblocks;    // [n_eblocks][PC_COLS]
  double* red;                       // [PC_COLS] tile-reduced (then shard-reduced) partials

  // control / outputs
  Ctrl* ctrl;
  const double* temps; int n_temps;  // temperature_schedule (R/utils.R:307-332)
  double* metrics; int metrics_cap;  // [metrics_cap][MC_COLS] rows of the current step chunk

  // sample ring (record_sample, R/bayesNMF_sampler.R:651-672)
  int ring_cap;
  T* ring_P; T* ring_E; int32_t* ring_A;
};

static const double RS_FX = 16777216.0;  // 2^24 fixed-point scale of rowsumE_fx

}  // namespace bnmf
