// Iteration-0 kernels: draw missing prior parameters from the hyperpriors
// (init_prior_params_, R/sample_priors.R:15-141) and data constants.
#pragma once
#include "bnmf_rng.cuh"
#include "bnmf_state.h"

namespace bnmf {

// flags[n] |= 1 if column n (p-side: a K-column; e-side: row n of an N x G matrix)
// of `v` contains a NaN  -- `any(is.na(...[, n]))`, R/sample_priors.R:33,40,47,54.
template <typename T>
__global__ void k_nan_cols(const T* v, long long len, int K, int N, int side, int* flags) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  T x = v[i];
  if (x != x) {
    int n = side == 0 ? (int)(i / K) : (int)(i % N);
    atomicOr(&flags[n], 1);
  }
}

// which: 0 = Mu, 1 = Sigmasq, 2 = Lambda, 3 = Alpha, 4 = Beta
template <typename T>
__global__ void k_init_prior(Dev<T> d, int side, const int* flags /* [5][N] */) {
  const int K = d.K, N = d.N;
  const long long cells = side == 0 ? (long long)K * N : (long long)N * d.G;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells) return;
  int n; long long c;
  if (side == 0) { n = (int)(idx / K); c = idx; }
  else { n = (int)(idx % N); c = (long long)n + (long long)N * (d.g0 + idx / N); }
  const uint32_t pur1 = side == 0 ? PUR_HYP_P1 : PUR_HYP_E1;
  const uint32_t pur2 = side == 0 ? PUR_HYP_P2 : PUR_HYP_E2;
  const Hyper<T>& hA = side == 0 ? d.A_p : d.A_e;
  const Hyper<T>& hB = side == 0 ? d.B_p : d.B_e;
  const Hyper<T>& hC = side == 0 ? d.C_p : d.C_e;
  const Hyper<T>& hD = side == 0 ? d.D_p : d.D_e;
  const Hyper<T>& hM = side == 0 ? d.M_p : d.M_e;
  const Hyper<T>& hS = side == 0 ? d.S_p : d.S_e;
  T* Mu = side == 0 ? d.Mu_p : d.Mu_e;
  T* Sg = side == 0 ? d.Sigmasq_p : d.Sigmasq_e;
  T* La = side == 0 ? d.Lambda_p : d.Lambda_e;
  T* Al = side == 0 ? d.Alpha_p : d.Alpha_e;
  T* Be = side == 0 ? d.Beta_p : d.Beta_e;
  if (d.prior == PRIOR_TRUNCNORMAL) {
    if (flags[0 * N + n])   // Mu ~ N(M, sd = sqrt(S))               R/sample_priors.R:34-38
      Mu[idx] = (T)normal_draw<double>(make_stream(d.seed, 0, pur1, c), (double)hM.at(idx),
                                       sqrt((double)hS.at(idx)));
    if (flags[1 * N + n])   // Sigmasq ~ InvGamma(A, rate = B)       R/sample_priors.R:41-45
      Sg[idx] = (T)(1.0 / gamma_draw<double>(make_stream(d.seed, 0, pur2, c), (double)hA.at(idx),
                                             (double)hB.at(idx)));
  } else if (d.prior == PRIOR_EXPONENTIAL) {
    if (flags[2 * N + n])   // Lambda ~ Gamma(A, rate = B)           R/sample_priors.R:72-76
      La[idx] = (T)gamma_draw<double>(make_stream(d.seed, 0, pur1, c), (double)hA.at(idx),
                                      (double)hB.at(idx));
  } else {
    if (flags[4 * N + n])   // Beta ~ Gamma(A, B)                    R/sample_priors.R:103-107
      Be[idx] = (T)gamma_draw<double>(make_stream(d.seed, 0, pur1, c), (double)hA.at(idx),
                                      (double)hB.at(idx));
    if (flags[3 * N + n])   // Alpha ~ Gamma(C, D)                   R/sample_priors.R:110-114
      Al[idx] = (T)gamma_draw<double>(make_stream(d.seed, 0, pur2, c), (double)hC.at(idx),
                                      (double)hD.at(idx));
  }
}

// R (expected rank) ~ Unif{0..N} and A_n ~ Bernoulli(clip(R/N)) when the rank is
// learned (R/sample_params.R:103-106, :222-226, :239); A = 1 otherwise (:75-77).
template <typename T>
__global__ void k_init_rank(Dev<T> d, int keepA) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int N = d.N;
  if (keepA) return;
  if (!d.learning_rank) {
    for (int n = 0; n < N; ++n) d.A[n] = 1;
    *d.R = N;
    return;
  }
  const int iter = d.ctrl->iter;
  // sample(0:N, 1, prob = uniform): inverse CDF on N+1 equal cells
  U4 w = make_stream(d.seed, iter, PUR_R, 0).at(0);
  int r = (int)(u01<double>(w.x) * (double)(N + 1));
  if (r > N) r = N;
  *d.R = r;
  double q = (double)r / (double)N;
  const double clipv = 0.4 / (double)N;
  if (q < clipv) q = clipv;
  if (q > 1.0 - clipv) q = 1.0 - clipv;
  for (int n = 0; n < N; ++n) {
    U4 wa = make_stream(d.seed, iter, PUR_A, (uint64_t)n).at(0);
    d.A[n] = u01<double>(wa.x) < q ? 1 : 0;
  }
}

}  // namespace bnmf
