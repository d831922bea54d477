// Kernels of the sweep-based Gibbs iteration: the Normal likelihood, the Poisson
// likelihood with Metropolis-Hastings proposals, and rank learning (SBFI / BFI).
//
//   prior parameters   R/sample_priors.R:214-308          k_hyper
//   P sweep            R/sample_Pn.R:54-87,132-248        k_p_rows (one launch, a cluster per mutation type)  |  k_p_pass1 -> k_p_draw [-> k_p_pass2 -> k_p_accept]   per signature
//   E sweep            R/sample_En.R:54-86,131-241        k_e_sweep                                            one launch
//   R, A sweep         R/sample_params.R:101-241          k_r -> (k_a_pass -> k_a_draw)                        per signature
//   sigmasq + metrics  R/sample_params.R:275-286, R/utils.R:412-455     k_final, k_pprior
//
// All of them work on a running reconstruction Mhat = P diag(A) E (K x G, resident in
// HBM) that is corrected by rank-1 updates instead of being recomputed by a dgemm per
// conditional as the reference does (2 to 8 get_Mhat calls per signature,
// R/sample_Pn.R:136,152,209-231).  A changed column of P (or a flipped A_n) is applied
// lazily: the next pass that streams Mhat anyway adds dvec[k] * E[n_prev, g] on the fly.
// The sweeps over n are truly sequential (each conditional sees the columns already
// updated), which is why P needs one streaming pass per signature while E -- whose
// conditionals are independent across genomes -- runs all N updates of a column
// on-chip in one launch.
#pragma once
#include "bnmf_poisson.cuh"
#include "bnmf_rng.cuh"
#include "bnmf_state.h"

namespace bnmf {

constexpr double LOG_SQRT_2PI = 0.9189385332046727;

// log(pnorm(t)), stable in both tails
__device__ __forceinline__ double log_ndtr(double t) {
  if (t >= 0.0) return log1p(-0.5 * erfc(t * 0.7071067811865476));
  return log(0.5 * erfcx(-t * 0.7071067811865476)) - 0.5 * t * t;
}
// log(truncnorm::dtruncnorm(x, a = 0, b = Inf, mean, sd))   (R/utils.R:134-145)
__device__ __forceinline__ double dtruncnorm0_log(double x, double mean, double sd) {
  const double z = (x - mean) / sd;
  return -LOG_SQRT_2PI - log(sd) - 0.5 * z * z - log_ndtr(mean / sd);
}
__device__ __forceinline__ double dnorm_log(double x, double mean, double var) {
  const double r = x - mean;
  return -LOG_SQRT_2PI - 0.5 * log(var) - 0.5 * (r * r) / var;
}
template <typename T> __device__ __forceinline__ double Mat(const Dev<T>& d, long long i) {
  return d.likelihood == LIK_POISSON ? (double)d.Mi[i] : (double)d.Mr[i];
}
// log Metropolis-Hastings ratio contribution of one cell (R/sample_Pn.R:213-238):
// dpois(M; new) - dpois(M; old) + dnorm(M; old, var = max(new, 1)) - dnorm(M; new, var = max(old, 1));
// the lgamma(M + 1) and -log sqrt(2 pi) terms cancel inside the cell.
__device__ __forceinline__ double mh_cell(double m, double mh_old, double mh_new) {
  const double lo = mh_old > 1e-6 ? mh_old : 1e-6, ln = mh_new > 1e-6 ? mh_new : 1e-6;
  const double Lo = log(lo), Ln = log(ln);
  const double dp = (m * Ln - ln) - (m * Lo - lo);
  const double vo = mh_new > 1.0 ? mh_new : 1.0, vn = mh_old > 1.0 ? mh_old : 1.0;
  // log(max(x, 1)) = max(log(max(x, 1e-6)), 0): two logarithms serve all four densities
  const double Lvo = Ln > 0.0 ? Ln : 0.0, Lvn = Lo > 0.0 ? Lo : 0.0;
  const double ro = m - mh_old, rn = m - mh_new;
  const double n_old = -0.5 * Lvo - 0.5 * (ro * ro) / vo;
  const double n_new = -0.5 * Lvn - 0.5 * (rn * rn) / vn;
  return dp + (n_old - n_new);
}
__device__ __forceinline__ double mh_ratio(double D) {
  if (!(D == D)) return D;           // NaN stays NaN: the comparison u < NaN rejects
  const double r = exp(D);
  return r < 1.0 ? r : 1.0;
}

// draw of one element from its prior (R/sample_Pn.R:12-30,56-74; R/sample_En.R:12-30,56-73)
template <typename T>
__device__ __forceinline__ double prior_draw(const Dev<T>& d, const Stream& st, int side, long long idx) {
  if (d.prior == PRIOR_TRUNCNORMAL) {
    const double mu = (double)(side == 0 ? d.Mu_p : d.Mu_e)[idx];
    const double sg = (double)(side == 0 ? d.Sigmasq_p : d.Sigmasq_e)[idx];
    return truncnorm0_draw<double>(st, mu, sqrt(sg));
  }
  if (d.prior == PRIOR_EXPONENTIAL)
    return gamma_draw<double>(st, 1.0, (double)(side == 0 ? d.Lambda_p : d.Lambda_e)[idx]);
  return gamma_draw<double>(st, (double)(side == 0 ? d.Alpha_p : d.Alpha_e)[idx],
                            (double)(side == 0 ? d.Beta_p : d.Beta_e)[idx]);
}
template <typename T>
__device__ __forceinline__ double prior_logdens(const Dev<T>& d, int side, long long idx, double x) {
  if (d.prior == PRIOR_TRUNCNORMAL)
    return dtruncnorm0_log(x, (double)(side == 0 ? d.Mu_p : d.Mu_e)[idx],
                           sqrt((double)(side == 0 ? d.Sigmasq_p : d.Sigmasq_e)[idx]));
  if (d.prior == PRIOR_EXPONENTIAL) return dexp_log(x, (double)(side == 0 ? d.Lambda_p : d.Lambda_e)[idx]);
  return dgamma_log(x, (double)(side == 0 ? d.Alpha_p : d.Alpha_e)[idx], (double)(side == 0 ? d.Beta_p : d.Beta_e)[idx]);
}

// ---- truncated-normal draws with the first attempts' variates computed up front ------------------
// What attempt t of truncnorm0_draw (bnmf_rng.cuh) consumes -- a standard normal, or an exponential
// and the logarithm of its accept uniform -- does not depend on the moments of the conditional.  The
// sweeps draw N conditionals one after the other per row / column; evaluating Philox, logarithm,
// square root and cosine inside that chain costs ~2.5 us per draw with everything else waiting, so
// the variates of attempts 0 .. P_PRE-1 of all N draws are evaluated beforehand, in parallel, and
// the chain only selects.  Same expressions, same results as truncnorm0_draw.
constexpr int P_PRE = 4;
__device__ __forceinline__ void tn_variates(const Stream& st, int t, double& z, double& e, double& u) {
  const U4 w = st.at((uint32_t)t);
  z = normal_from<double>(w.x, w.y);
  e = -tlog<double>(u01<double>(w.x));
  u = tlog<double>(u01<double>(w.y));
}
// vZ, vE, vU: the P_PRE variates of this draw
__device__ __forceinline__ double truncnorm0_staged(const Stream& st, double mu, double sd,
                                                    const double* vZ, const double* vE, const double* vU) {
  const double alpha = -mu / sd;
  bool found = false;
  double x;
  if (alpha <= 0.45) {
    double z = alpha;
    for (int t = 0; t < P_PRE && !found; ++t) { const double zz = vZ[t]; if (zz >= alpha) { z = zz; found = true; } }
    x = mu + sd * z;
    x = x < 0.0 ? 0.0 : x;
  } else {
    const double lam = 0.5 * (alpha + sqrt(alpha * alpha + 4.0));
    double e = 0.0;
    for (int t = 0; t < P_PRE && !found; ++t) {
      const double ee = vE[t] / lam;
      const double dz = (alpha + ee) - lam;
      if (vU[t] <= -0.5 * (dz * dz)) { e = ee; found = true; }
    }
    x = sd * e;
  }
  if (!found) x = truncnorm0_draw<double>(st, mu, sd, (uint32_t)P_PRE);
  return x;
}

// ------------------------------------------------------------------------------
// k_hyper: prior-parameter updates of the truncated-normal and exponential priors,
// element-wise (they read only the previous P / E; R/sample_priors.R:214-308).
//   Mu ~ N(num/den, sd = 1/den)      -- the reference passes the variance as sd (:219,:235)
//   Sigmasq ~ InvGamma(A + 1/2, B + (X - Mu)^2 / 2), the E side with A_e as base (:267)
//   Lambda ~ Gamma(A + 1, B + X)
// ------------------------------------------------------------------------------
template <typename T>
__global__ void k_hyper(Dev<T> d, int side) {
  const int K = d.K, N = d.N;
  const long long cells = side == 0 ? (long long)K * N : (long long)N * d.G;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells) return;
  const long long c = side == 0 ? idx : (idx % N) + (long long)N * (d.g0 + idx / N);
  const int iter = d.ctrl->iter;
  const uint32_t pur1 = side == 0 ? PUR_HYP_P1 : PUR_HYP_E1, pur2 = side == 0 ? PUR_HYP_P2 : PUR_HYP_E2;
  const Hyper<T>& hA = side == 0 ? d.A_p : d.A_e;
  const Hyper<T>& hB = side == 0 ? d.B_p : d.B_e;
  const double X = (double)(side == 0 ? d.P : d.E)[idx];
  if (d.prior == PRIOR_TRUNCNORMAL) {
    const Hyper<T>& hM = side == 0 ? d.M_p : d.M_e;
    const Hyper<T>& hS = side == 0 ? d.S_p : d.S_e;
    T* Mu = side == 0 ? d.Mu_p : d.Mu_e;
    T* Sg = side == 0 ? d.Sigmasq_p : d.Sigmasq_e;
    const double S = (double)hS.at(idx), sg = (double)Sg[idx];
    const double num = (double)hM.at(idx) / S + X / sg;
    const double den = 1.0 / S + 1.0 / sg;
    const double mu = (double)(T)normal_draw<double>(make_stream(d.seed, iter, pur1, c), num / den, 1.0 / den);
    Mu[idx] = (T)mu;
    const double base = side == 0 ? (double)hB.at(idx) : (double)hA.at(idx);
    const double r = X - mu;
    Sg[idx] = (T)(1.0 / gamma_draw<double>(make_stream(d.seed, iter, pur2, c), (double)hA.at(idx) + 0.5, base + (r * r) / 2.0));
  } else if (d.prior == PRIOR_EXPONENTIAL) {
    T* La = side == 0 ? d.Lambda_p : d.Lambda_e;
    La[idx] = (T)gamma_draw<double>(make_stream(d.seed, iter, pur1, c), (double)hA.at(idx) + 1.0, (double)hB.at(idx) + X);
  }
}

// k_prior_fill: P or E entirely from the prior (iteration 1, R/bayesNMF_sampler.R:241).
template <typename T>
__global__ void k_prior_fill(Dev<T> d, int side) {
  const int K = d.K, N = d.N;
  const long long cells = side == 0 ? (long long)K * N : (long long)N * d.G;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells) return;
  const long long c = side == 0 ? idx : (idx % N) + (long long)N * (d.g0 + idx / N);
  const Stream st = make_stream(d.seed, d.ctrl->iter, side == 0 ? PUR_P : PUR_E, c);
  (side == 0 ? d.P : d.E)[idx] = (T)prior_draw(d, st, side, idx);
}

// non-zero flags of the columns of P / rows of E, and NaN acceptance rates (iteration 1)
template <typename T>
__global__ void k_nzflags(Dev<T> d) {
  const int K = d.K, N = d.N;
  const int iter = d.ctrl->iter;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)K * N && d.P[i] != (T)0) atomicOr(&d.nzP[i / K], 1);
  if (i < (long long)N * d.G && d.E[i] != (T)0) atomicOr(&d.nzE[(iter & 1) * N + (int)(i % N)], 1);
}
template <typename T> __global__ void k_fill(T* p, long long n, T v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// k_mhat_full: Mhat = P diag(A) E from scratch (get_Mhat_, R/utils.R:29-49); run once per
// iteration so that rounding of the rank-1 corrections cannot accumulate.
template <typename T>
__global__ void k_mhat_full(Dev<T> d) {
  const int K = d.K, N = d.N;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K * d.G) return;
  const int k = (int)(i % K);
  const long long g = i / K;
  double acc = 0.0;
  for (int n = 0; n < N; ++n)
    if (d.A[n]) acc += (double)d.P[k + (long long)K * n] * (double)d.E[n + (long long)N * g];
  d.Mhat[i] = (T)acc;
}

// ------------------------------------------------------------------------------
// P sweep, signature n.  Thread (kx, gy) owns mutation type k and every GY-th genome of
// the block's chunk; a warp reads 32 consecutive k of one genome (coalesced, K is the
// fast axis of M and Mhat).  The reduction over genomes is a private running sum per
// thread, combined over gy in shared memory and over chunks by k_p_draw, both in a
// fixed order (bit-reproducible).
//   pass 1: apply the pending rank-1 update, then
//           num1[k] = sum_g E[n,g] (M - Mhat_{-n})[k,g] / s[k,g],  den[k] = sum_g A_n E[n,g]^2 / s[k,g]
//           with s = Mhat (MH proposal, R/sample_Pn.R:137-139) or s = sigmasq_g (Normal)
//   pass 2: log MH ratio of the proposal (R/sample_Pn.R:206-238), only after convergence
// ------------------------------------------------------------------------------
template <typename T>
__global__ void k_p_pass1(Dev<T> d, int n, int n_prev) {
  extern __shared__ double sm[];
  const int K = d.K, N = d.N;
  const int kx = threadIdx.x, gy = threadIdx.y, GY = blockDim.y, KX = blockDim.x;
  const int k = blockIdx.y * KX + kx;
  const long long gbeg = (long long)blockIdx.x * d.gchunk;
  const long long gend = gbeg + d.gchunk < d.G ? gbeg + d.gchunk : d.G;
  const int An = d.A[n];
  const bool normal = d.likelihood == LIK_NORMAL;
  double num1 = 0.0, den = 0.0;
  if (k < K) {
    const double dv = n_prev >= 0 ? d.dvec[k] : 0.0;
    const double pkn = (double)d.P[k + (long long)K * n];
    // restrict-qualified views: lets the loads of several genomes be in flight at once
    T* __restrict__ Mh = d.Mhat;
    const T* __restrict__ Ev = d.E;
    const T* __restrict__ Sg = d.sigmasq;
    const int32_t* __restrict__ Mi = d.Mi;
    const T* __restrict__ Mr = d.Mr;
#pragma unroll 4
    for (long long g = gbeg + gy; g < gend; g += GY) {
      const long long i = k + (long long)K * g;
      double mh = (double)Mh[i];
      if (n_prev >= 0) {
        mh = (double)(T)(mh + dv * (double)Ev[n_prev + (long long)N * g]);
        Mh[i] = (T)mh;
      }
      if (An) {
        const double e = (double)Ev[n + (long long)N * g];
        const double inv = 1.0 / (normal ? (double)Sg[g] : mh);
        const double mh_no = mh - pkn * e;
        const double m = normal ? (double)Mr[i] : (double)Mi[i];
        num1 += e * ((m - mh_no) * inv);
        den += (e * e) * inv;
      }
    }
  }
  sm[(gy * KX + kx) * 2 + 0] = num1;
  sm[(gy * KX + kx) * 2 + 1] = den;
  __syncthreads();
  if (gy == 0 && k < K) {
    for (int y = 1; y < GY; ++y) { num1 += sm[(y * KX + kx) * 2 + 0]; den += sm[(y * KX + kx) * 2 + 1]; }
    double* pp = d.ppart + ((long long)blockIdx.x * K + k) * 2;
    pp[0] = num1; pp[1] = den;
  }
}

template <typename T>
__global__ void k_p_pass2(Dev<T> d, int n) {
  extern __shared__ double sm[];
  const int K = d.K, N = d.N;
  const int kx = threadIdx.x, gy = threadIdx.y, GY = blockDim.y, KX = blockDim.x;
  const int k = blockIdx.y * KX + kx;
  const long long gbeg = (long long)blockIdx.x * d.gchunk;
  const long long gend = gbeg + d.gchunk < d.G ? gbeg + d.gchunk : d.G;
  double D = 0.0;
  if (k < K && d.A[n]) {
    const double dp = d.prop[k] - (double)d.P[k + (long long)K * n];
    const T* __restrict__ Mh = d.Mhat;
    const T* __restrict__ Ev = d.E;
    const bool normal = d.likelihood == LIK_NORMAL;
#pragma unroll 4
    for (long long g = gbeg + gy; g < gend; g += GY) {
      const long long i = k + (long long)K * g;
      const double mh = (double)Mh[i];
      const double m = normal ? (double)d.Mr[i] : (double)d.Mi[i];
      D += mh_cell(m, mh, mh + dp * (double)Ev[n + (long long)N * g]);
    }
  }
  sm[gy * KX + kx] = D;
  __syncthreads();
  if (gy == 0 && k < K) {
    for (int y = 1; y < GY; ++y) D += sm[y * KX + kx];
    d.ppart[((long long)blockIdx.x * K + k) * 2] = D;
  }
}

// k_p_draw: finish the reduction over genome chunks (a warp per mutation type, fixed
// butterfly order), form the conditional (or proposal) moments, draw.
template <typename T>
__global__ void k_p_draw(Dev<T> d, int n) {
  const int K = d.K, N = d.N;
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K) return;
  const int iter = d.ctrl->iter;
  const int An = d.A[n];
  const long long c = k + (long long)K * n;
  const bool zero_row = d.nzE[((iter - 1) & 1) * N + n] == 0;   // all(E[n, ] == 0), R/sample_Pn.R:56
  double num1 = 0.0, den = 0.0;
  if (An != 0 && !zero_row) {
    for (int ch = lane; ch < d.n_gchunks; ch += 32) {
      const double* pp = d.ppart + ((long long)ch * K + k) * 2;
      num1 += pp[0]; den += pp[1];
    }
    num1 = warp_sum(num1); den = warp_sum(den);
  }
  if (lane != 0) return;
  const double Pold = (double)d.P[c];
  const Stream st = make_stream(d.seed, iter, PUR_P, c);
  double x;
  if (An == 0 || zero_row) {
    x = prior_draw(d, st, 0, c);
  } else {
    double mu, v;
    if (d.prior == PRIOR_EXPONENTIAL) {
      mu = (num1 - (double)d.Lambda_p[c]) / den; v = 1.0 / den;
    } else {
      const double sg = (double)d.Sigmasq_p[c];
      den = den + 1.0 / sg;
      mu = (num1 + (double)d.Mu_p[c] / sg) / den; v = 1.0 / den;
    }
    x = truncnorm0_draw<double>(st, mu, sqrt(v));
  }
  x = (double)(T)x;
  const bool mh_step = d.MH && d.ctrl->converged && An != 0;
  if (mh_step) {                      // k_p_pass2 / k_p_accept decide
    d.prop[k] = x;
    d.dvec[k] = 0.0;
    return;
  }
  if (d.MH && An != 0) d.P_acc[c] = (T)1;                        // R/sample_Pn.R:201-204
  d.P[c] = (T)x;
  d.dvec[k] = An ? x - Pold : 0.0;
  if (x != 0.0) atomicOr(&d.nzP[n], 1);
}

template <typename T>
__global__ void k_p_accept(Dev<T> d, int n) {
  const int K = d.K;
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K || d.A[n] == 0) return;
  const long long c = k + (long long)K * n;
  double D = 0.0;
  for (int ch = lane; ch < d.n_gchunks; ch += 32) D += d.ppart[((long long)ch * K + k) * 2];
  D = warp_sum(D);
  if (lane != 0) return;
  const double ratio = mh_ratio(D);
  d.P_acc[c] = (T)ratio;
  const double u = u01<double>(make_stream(d.seed, d.ctrl->iter, PUR_MH_P, c).at(0).x);
  const double Pold = (double)d.P[c];
  const double x = u < ratio ? d.prop[k] : Pold;
  d.P[c] = (T)x;
  d.dvec[k] = x - Pold;
  if (x != 0.0) atomicOr(&d.nzP[n], 1);
}

// ------------------------------------------------------------------------------
// k_p_rows: the whole P sweep in one launch.  The conditionals of P[k, .] touch only row k of
// M and of the running Mhat, so a thread-block cluster owns one mutation type: its CS blocks
// keep disjoint slices of the row in shared memory for all N signatures (M and Mhat are read
// once per iteration instead of once or twice per signature).  Row n of E -- read from a
// transposed copy made once per iteration, genomes contiguous -- arrives by a bulk asynchronous
// copy (cp.async.bulk + mbarrier) into one of two buffers while signature n-1 is being drawn.
// The sums over genomes are reduced inside the block and then across the cluster through
// distributed shared memory; block 0 draws (and, after convergence, runs the Metropolis-Hastings
// accept step on a second cluster-wide sum) and hands the change of P[k,n] back to every
// block, which applies the rank-1 correction to its slice.
// Same arithmetic as k_p_pass1 / k_p_draw / k_p_pass2 / k_p_accept, a different (fixed)
// summation order.
// ------------------------------------------------------------------------------
template <typename T>
__global__ void k_transpose_E(Dev<T> d, T* __restrict__ Et, long long Gp) {
  // Et[n * Gp + g] = E[n + N * g]   (row stride Gp = G rounded up to an even number of elements)
  const long long NG = (long long)d.N * d.G;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NG) return;
  const long long n = i / d.G, g = i - n * d.G;
  Et[n * Gp + g] = d.E[n + (long long)d.N * g];
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a double into the shared memory of block `rank` of this cluster (same offset as `local`)
__device__ __forceinline__ void dsmem_store(double* local, uint32_t rank, double v) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sa), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(ra), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tBNMF_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra BNMF_DONE;\n\tbra BNMF_WAIT;\n\tBNMF_DONE:\n\t}"
      :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// bulk asynchronous copy global -> shared memory of this block; bytes, src and dst multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// sum of (a, b) over the block in a fixed order; valid in thread 0
template <int THREADS>
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch /*2 * THREADS/32*/) {
  a = warp_sum(a); b = warp_sum(b);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { scratch[2 * wid] = a; scratch[2 * wid + 1] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; b = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) { a += scratch[2 * w]; b += scratch[2 * w + 1]; }
  }
}

// bytes of shared memory of k_p_rows for a slice of gslice genomes (gslice a multiple of 16 / sizeof(T))
template <typename T> __host__ __device__ inline size_t p_rows_smem(int gslice, int threads, bool normal) {
  return (size_t)gslice * (3 * sizeof(T) + (normal ? sizeof(T) + sizeof(double) : sizeof(int32_t))) + 16 +
         (size_t)(2 * (threads / 32) + 18 + 3 * 64 + 3 * 64 * P_PRE) * sizeof(double) + 2 * sizeof(uint64_t) + 2 * 64 * sizeof(int);
}

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_p_rows(Dev<T> d, const T* __restrict__ Et, long long Gp, int CS, int gslice) {
  extern __shared__ __align__(16) unsigned char prow_raw[];
  const int K = d.K, N = d.N, G = d.G;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int k = blockIdx.x / CS;
  const int g0 = (int)rank * gslice;
  const int ng = max(0, min(G, g0 + gslice) - g0);
  constexpr int EA = 16 / (int)sizeof(T);                                    // elements per 16 bytes: bulk copies move whole 16-byte units
  const uint32_t e_bytes = (uint32_t)(((ng + EA - 1) / EA) * 16);
  const bool normal = d.likelihood == LIK_NORMAL;
  // shared: E buffers (2 x T) | Mhat slice (T) | M slice (T or int32) | reduction scratch | mailbox | broadcast slot | 2 mbarriers
  T* Eb0 = smem_ptr<T>((uint32_t)__cvta_generic_to_shared(prow_raw));
  T* Mh = smem_ptr<T>((uint32_t)__cvta_generic_to_shared(prow_raw) + 2u * (uint32_t)gslice * (uint32_t)sizeof(T));
  unsigned char* after = reinterpret_cast<unsigned char*>(Mh + gslice);
  T* Mrs = smem_ptr<T>((uint32_t)__cvta_generic_to_shared(after));
  int32_t* Mis = smem_ptr<int32_t>((uint32_t)__cvta_generic_to_shared(after));
  after += (size_t)gslice * (normal ? sizeof(T) : sizeof(int32_t));
  after = reinterpret_cast<unsigned char*>(((uintptr_t)after + 15) & ~(uintptr_t)15);
  // (offsets that depend on the likelihood hide the address space from the compiler: going through the
  //  32-bit shared-window address keeps the accesses LDS / STS instead of generic loads)
  double* Sinv = smem_ptr<double>((uint32_t)__cvta_generic_to_shared(after));   // [gslice] 1 / sigmasq_g (Normal likelihood only)
  if (normal) after += (size_t)gslice * sizeof(double);
  double* scratch = smem_ptr<double>((uint32_t)__cvta_generic_to_shared(after));   // [2 * THREADS / 32]
  double* mail = scratch + 2 * (THREADS / 32);              // [8][2]  partial sums of every block (used in block 0)
  double* bc = mail + 16;                                   // [2]     {value, flag} from block 0
  double* sP = bc + 2;                                      // [64] row k of P and of its prior parameters, A, the
  double* sQ1 = sP + 64;                                    //      zero-row flags: read once, not once per signature
  double* sQ2 = sQ1 + 64;
  double* vZ = sQ2 + 64;                                    // [64][P_PRE] base variates of the first attempts of every draw:
  double* vE = vZ + 64 * P_PRE;                             //   standard normal | -log u (exponential) | log u' (its accept test)
  double* vU = vE + 64 * P_PRE;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(vU + 64 * P_PRE);   // [2]     "row n of E has landed in buffer n & 1"
  int* sA = reinterpret_cast<int*>(mbar + 2);               // [64]
  int* sZ = sA + 64;                                        // [64]

  if (tid == 0) {
    mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && ng > 0) {                                  // row 0 of E on its way while the slices are gathered
    mbar_expect_tx(&mbar[0], e_bytes);
    bulk_g2s(Eb0, Et + g0, e_bytes, &mbar[0]);
  }
  const int iter = d.ctrl->iter;
  for (int j = tid; j < ng; j += THREADS) {
    const long long i = k + (long long)K * (g0 + j);
    Mh[j] = d.Mhat[i];
    if (normal) { Mrs[j] = d.Mr[i]; Sinv[j] = 1.0 / (double)d.sigmasq[g0 + j]; } else Mis[j] = d.Mi[i];
  }
  for (int n = tid; n < N; n += THREADS) {
    const long long c = k + (long long)K * n;
    sA[n] = d.A[n];
    sZ[n] = d.nzE[((iter - 1) & 1) * N + n] == 0;           // all(E[n, ] == 0), R/sample_Pn.R:56
    sP[n] = (double)d.P[c];
    if (d.prior == PRIOR_EXPONENTIAL) { sQ1[n] = (double)d.Lambda_p[c]; sQ2[n] = 0.0; }
    else if (d.prior == PRIOR_TRUNCNORMAL) { sQ1[n] = (double)d.Mu_p[c]; sQ2[n] = (double)d.Sigmasq_p[c]; }
  }
  // variates of the first attempts of the row's N draws, N x P_PRE threads in parallel (tn_variates)
  for (int i = tid; i < N * P_PRE; i += THREADS) {
    const int n = i / P_PRE, t = i - n * P_PRE;
    tn_variates(make_stream(d.seed, iter, PUR_P, k + (long long)K * n), t, vZ[i], vE[i], vU[i]);
  }
  const bool mh_on = d.MH && d.ctrl->converged;
  __syncthreads();

  for (int n = 0; n < N; ++n) {
    const int An = sA[n];
    const long long c = k + (long long)K * n;
    const double pkn = sP[n];
    const bool active = An != 0 && !sZ[n];
    const T* En = Eb0 + (size_t)(n & 1) * gslice;      // (one base pointer: the loads stay LDS)
    // the other buffer is free (its last readers passed the barrier that ended signature n-1): prefetch row n+1
    if (tid == 0 && n + 1 < N && ng > 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&mbar[(n + 1) & 1], e_bytes);
      bulk_g2s(Eb0 + (size_t)((n + 1) & 1) * gslice, Et + (long long)(n + 1) * Gp + g0, e_bytes, &mbar[(n + 1) & 1]);
    }
    if (ng > 0) mbar_wait(&mbar[n & 1], (uint32_t)((n >> 1) & 1));
    // ---- pass 1: the two sums of the conditional ----
    double num1 = 0.0, den = 0.0;
    if (active) {
#pragma unroll 4
      for (int j = tid; j < ng; j += THREADS) {
        const double mh = (double)Mh[j];
        const double e = (double)En[j];
        const double inv = normal ? Sinv[j] : 1.0 / mh;
        const double mh_no = mh - pkn * e;
        const double m = normal ? (double)Mrs[j] : (double)Mis[j];
        num1 += e * ((m - mh_no) * inv);
        den += (e * e) * inv;
      }
      block_sum2<THREADS>(num1, den, scratch);
      if (tid == 0) { dsmem_store(&mail[2 * rank], 0, num1); dsmem_store(&mail[2 * rank + 1], 0, den); }
    }
    cluster_sync_all();
    // ---- block 0 draws ----
    if (rank == 0 && tid == 0) {
      const Stream st = make_stream(d.seed, iter, PUR_P, c);
      double x;
      if (!active) {
        x = prior_draw(d, st, 0, c);
      } else {
        double s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < CS; ++r) { s1 += mail[2 * r]; s2 += mail[2 * r + 1]; }
        double mu, v;
        if (d.prior == PRIOR_EXPONENTIAL) {
          mu = (s1 - sQ1[n]) / s2; v = 1.0 / s2;
        } else {
          const double sg = sQ2[n];
          s2 = s2 + 1.0 / sg;
          mu = (s1 + sQ1[n] / sg) / s2; v = 1.0 / s2;
        }
        x = truncnorm0_staged(st, mu, sqrt(v), vZ + n * P_PRE, vE + n * P_PRE, vU + n * P_PRE);
      }
      x = (double)(T)x;
      if (!(mh_on && An != 0)) {
        if (d.MH && An != 0) d.P_acc[c] = (T)1;                        // R/sample_Pn.R:201-204
        d.P[c] = (T)x;
        if (x != 0.0) atomicOr(&d.nzP[n], 1);
        const double dv = An ? x - pkn : 0.0;
        for (int r = 0; r < CS; ++r) { dsmem_store(&bc[0], r, dv); dsmem_store(&bc[1], r, 0.0); }
      } else {
        for (int r = 0; r < CS; ++r) { dsmem_store(&bc[0], r, x); dsmem_store(&bc[1], r, 1.0); }
      }
    }
    cluster_sync_all();
    double dv = bc[0];
    if (bc[1] != 0.0) {
      // ---- Metropolis-Hastings accept step (R/sample_Pn.R:199-248): log ratio over the row ----
      const double dp = dv - pkn;          // bc[0] carries the proposal
      double D = 0.0, zero = 0.0;
#pragma unroll 2
      for (int j = tid; j < ng; j += THREADS) {
        const double mh = (double)Mh[j];
        const double m = normal ? (double)Mrs[j] : (double)Mis[j];
        D += mh_cell(m, mh, mh + dp * (double)En[j]);
      }
      block_sum2<THREADS>(D, zero, scratch);
      if (tid == 0) dsmem_store(&mail[2 * rank], 0, D);
      cluster_sync_all();
      if (rank == 0 && tid == 0) {
        double Dt = 0.0;
        for (int r = 0; r < CS; ++r) Dt += mail[2 * r];
        const double ratio = mh_ratio(Dt);
        d.P_acc[c] = (T)ratio;
        const double u = u01<double>(make_stream(d.seed, iter, PUR_MH_P, c).at(0).x);
        const double xn = u < ratio ? dv : pkn;
        d.P[c] = (T)xn;
        if (xn != 0.0) atomicOr(&d.nzP[n], 1);
        for (int r = 0; r < CS; ++r) dsmem_store(&bc[0], r, xn - pkn);
      }
      cluster_sync_all();
      dv = bc[0];
    }
    // ---- rank-1 correction of the slice (every later conditional sees the new column) ----
    if (dv != 0.0) {
#pragma unroll 4
      for (int j = tid; j < ng; j += THREADS) Mh[j] = (T)((double)Mh[j] + dv * (double)En[j]);
    }
    __syncthreads();   // buffer n & 1 and Mh are settled before the next prefetch / pass
    // (bc / mail are rewritten only after the next cluster barrier, which every block reaches
    //  after it has read them)
  }
  for (int j = tid; j < ng; j += THREADS) d.Mhat[k + (long long)K * (g0 + j)] = Mh[j];
  if (rank == 0 && tid == 0) d.dvec[k] = 0.0;
}

// ------------------------------------------------------------------------------
// P sweep of the Normal likelihood through two Gram matrices.  There the variance of a cell does
// not depend on the mutation type (s[k,g] = sigmasq_g, R/sample_Pn.R:140-143), so with
// W = diag(1 / sigmasq) the sums of get_mu_sigmasq_Pn_normal (R/sample_Pn.R:132-187) factor:
//     num1[k] = (M W E')[k,n] - sum_{m != n} P[k,m] A_m (E W E')[m,n],     den = A_n (E W E')[n,n]
// (P[k,m] the current value: already redrawn for m < n).  One pass over M and E per iteration
// (k_gram_part + k_gram_fold, fixed summation order) replaces a pass per signature; the N
// conditionals of a mutation type are then a chain over two small matrices (k_p_gram, a warp per
// mutation type), and Mhat is rebuilt once (k_mhat_full).  Algebraically the reference's sums,
// not their summation order.
// ------------------------------------------------------------------------------
constexpr int GRAM_GC = 32;      // genomes per block of k_gram_part
constexpr int GRAM_KT = 128;     // mutation types per block (one per thread)
// part[chunk][K*N + N*N]: A1 = M W E' (K x N, k fastest) then A2 = E W E' (N x N)
template <typename T>
__global__ void __launch_bounds__(GRAM_KT) k_gram_part(Dev<T> d, double* __restrict__ part) {
  extern __shared__ double gsm[];
  const int K = d.K, N = d.N;
  double* Ew = gsm;                    // [N][GC]  E[n,g] / sigmasq_g
  double* Es = gsm + N * GRAM_GC;      // [N][GC]  E[n,g]
  const long long g0 = (long long)blockIdx.x * GRAM_GC;
  const int ng = (int)min((long long)GRAM_GC, (long long)d.G - g0);
  for (int i = threadIdx.x; i < N * GRAM_GC; i += GRAM_KT) {
    const int j = i / N, n = i - j * N;            // n fastest: coalesced reads of E
    double e = 0.0, w = 0.0;
    if (j < ng) { e = (double)d.E[n + (long long)N * (g0 + j)]; w = 1.0 / (double)d.sigmasq[g0 + j]; }
    Es[n * GRAM_GC + j] = e;
    Ew[n * GRAM_GC + j] = e * w;
  }
  __syncthreads();
  double* out = part + (long long)blockIdx.x * ((long long)K * N + (long long)N * N);
  const int k = blockIdx.y * GRAM_KT + threadIdx.x;
  if (k < K) {
    double mrow[GRAM_GC];
#pragma unroll
    for (int j = 0; j < GRAM_GC; ++j) mrow[j] = j < ng ? (double)d.Mr[k + (long long)K * (g0 + j)] : 0.0;
    for (int n = 0; n < N; ++n) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < GRAM_GC; ++j) acc += mrow[j] * Ew[n * GRAM_GC + j];
      out[k + (long long)K * n] = acc;
    }
  }
  if (blockIdx.y == 0) {
    for (int i = threadIdx.x; i < N * N; i += GRAM_KT) {
      const int m = i % N, n = i / N;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < GRAM_GC; ++j) acc += Ew[m * GRAM_GC + j] * Es[n * GRAM_GC + j];
      out[(long long)K * N + i] = acc;
    }
  }
}
// gram[i] = sum over chunks of part[chunk][i]: a warp per element, lanes stride over the chunks, fixed butterfly
__global__ void k_gram_fold(const double* __restrict__ part, double* __restrict__ gram, long long len, int n_chunks) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= len) return;
  double a = 0.0;
  for (int c = lane; c < n_chunks; c += 32) a += part[(long long)c * len + i];
  a = warp_sum(a);
  if (lane == 0) gram[i] = a;
}
// the N conditionals of mutation type k, one warp per k
template <typename T, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_p_gram(Dev<T> d, const double* __restrict__ gram) {
  extern __shared__ double psm[];
  const int K = d.K, N = d.N;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int k = blockIdx.x * WARPS + wid;
  if (k >= K) return;
  double* wbase = psm + (size_t)wid * (N * (1 + 3 * P_PRE));
  double* Prow = wbase;                // [N] current row of P (A folded in below)
  double* vZ = Prow + N;               // [N][P_PRE] variates of the first attempts of the N draws
  double* vE = vZ + N * P_PRE;
  double* vU = vE + N * P_PRE;
  const double* A1 = gram;
  const double* A2 = gram + (long long)K * N;
  const int iter = d.ctrl->iter;
  for (int n = lane; n < N; n += 32) Prow[n] = (double)d.P[k + (long long)K * n];
  for (int i = lane; i < N * P_PRE; i += 32) {
    const int n = i / P_PRE, t = i - n * P_PRE;
    tn_variates(make_stream(d.seed, iter, PUR_P, k + (long long)K * n), t, vZ[i], vE[i], vU[i]);
  }
  __syncwarp();
  for (int n = 0; n < N; ++n) {
    const int An = d.A[n];
    const long long c = k + (long long)K * n;
    const bool zero_row = d.nzE[((iter - 1) & 1) * N + n] == 0;   // all(E[n, ] == 0), R/sample_Pn.R:56
    const Stream st = make_stream(d.seed, iter, PUR_P, c);
    double x;
    if (An == 0 || zero_row) {
      x = prior_draw(d, st, 0, c);
    } else {
      double s = 0.0;
      for (int m = lane; m < N; m += 32)
        if (m != n && d.A[m]) s += Prow[m] * A2[m + (long long)N * n];
      s = warp_sum(s);
      const double num1 = A1[c] - s;
      double den = A2[n + (long long)N * n];
      double mu, v;
      if (d.prior == PRIOR_EXPONENTIAL) {
        mu = (num1 - (double)d.Lambda_p[c]) / den; v = 1.0 / den;
      } else {
        const double sg = (double)d.Sigmasq_p[c];
        den = den + 1.0 / sg;
        mu = (num1 + (double)d.Mu_p[c] / sg) / den; v = 1.0 / den;
      }
      x = truncnorm0_staged(st, mu, sqrt(v), vZ + n * P_PRE, vE + n * P_PRE, vU + n * P_PRE);
    }
    x = (double)(T)x;
    __syncwarp();
    if (lane == 0) {
      Prow[n] = x;
      d.P[c] = (T)x;
      if (x != 0.0) atomicOr(&d.nzP[n], 1);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------
// k_e_sweep: all N updates of E[., g] for one genome per warp.  The genome's column of M
// and of the running Mhat stay in shared memory for the whole sweep; column n of P is
// staged once per block and shared by its warps; the reductions over k are warp
// butterflies (fixed order).  One streaming pass over M and Mhat per iteration, whatever N.
// ------------------------------------------------------------------------------
// doubles of shared memory per warp of k_e_sweep besides its two columns: the genome's column of E
// and of the prior parameters, and the variates of the first attempts of its N draws (tn_variates)
__host__ __device__ inline int e_sweep_extra(int N, int stage) { return (3 + (stage ? 3 * P_PRE : 0)) * N; }
// doubles per warp for its two columns: Mhat (double) and the data (double for the Normal likelihood,
// int32 counts for the Poisson one)
__host__ __device__ inline int e_sweep_cols(int K, bool normal) { return K + (normal ? K : (K + 1) / 2); }

// sum over the LPG consecutive lanes of a sub-group (LPG = 8, 16 or 32), fixed butterfly order
template <int LPG> __device__ __forceinline__ double seg_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = LPG >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// LPG lanes per genome: with short columns (K <= 128) a warp carries 4 (2) genomes, each on its own
// 8 (16) lanes, so that the scalar part of a conditional is not repeated by 32 lanes.
template <typename T, int LPG>
__global__ void k_e_sweep(Dev<T> d, int n_prev, int stage) {
  extern __shared__ double sm[];
  const int K = d.K, N = d.N;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int GPW = 32 / LPG;                               // genomes per warp
  const int sub = lane / LPG, l = lane - sub * LPG;
  const int SPB = (blockDim.x >> 5) * GPW;                    // genome slots per block
  const int slot = wid * GPW + sub;
  const bool normal = d.likelihood == LIK_NORMAL;
  const int cols = e_sweep_cols(K, normal);
  double* Pn = sm;
  double* Mh = sm + K + (size_t)slot * cols;
  double* MvD = Mh + K;                                       // data column: doubles (Normal) ...
  int32_t* MvI = reinterpret_cast<int32_t*>(Mh + K);         // ... or counts (Poisson)
  double* wx = sm + K + (size_t)SPB * cols + (size_t)slot * e_sweep_extra(N, stage);
  double* sE = wx;                       // [N] E[., g] before the sweep
  double* sQ1 = sE + N;                  // [N] Lambda_e | Mu_e
  double* sQ2 = sQ1 + N;                 // [N] Sigmasq_e
  double* vZ = sQ2 + N;                  // [N][P_PRE] variates of the first attempts of every draw
  double* vE = vZ + N * P_PRE;
  double* vU = vE + N * P_PRE;
  const long long g = (long long)blockIdx.x * SPB + slot;
  const bool valid = g < d.G;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);   // whole sub-groups are valid or not
  const int iter = d.ctrl->iter, converged = d.ctrl->converged;
  if (valid) {
    const double eprev = n_prev >= 0 ? (double)d.E[n_prev + (long long)N * g] : 0.0;
    for (int k = l; k < K; k += LPG) {
      const long long i = k + (long long)K * g;
      double mh = (double)d.Mhat[i];
      if (n_prev >= 0) mh = (double)(T)(mh + d.dvec[k] * eprev);
      Mh[k] = mh;
      if (normal) MvD[k] = (double)d.Mr[i]; else MvI[k] = d.Mi[i];
    }
    // everything the N sequential conditionals of this genome read besides Mhat: once, in parallel
    for (int n = l; n < N; n += LPG) {
      const long long idx = n + (long long)N * g;
      sE[n] = (double)d.E[idx];
      if (d.prior == PRIOR_EXPONENTIAL) { sQ1[n] = (double)d.Lambda_e[idx]; sQ2[n] = 0.0; }
      else if (d.prior == PRIOR_TRUNCNORMAL) { sQ1[n] = (double)d.Mu_e[idx]; sQ2[n] = (double)d.Sigmasq_e[idx]; }
    }
    if (stage) for (int i = l; i < N * P_PRE; i += LPG) {
      const int n = i / P_PRE, t = i - n * P_PRE;
      tn_variates(make_stream(d.seed, iter, PUR_E, n + (long long)N * (d.g0 + g)), t, vZ[i], vE[i], vU[i]);
    }
  }
  const double inv_sg = (valid && normal) ? 1.0 / (double)d.sigmasq[g] : 0.0;
  for (int n = 0; n < N; ++n) {
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) Pn[k] = (double)d.P[k + (long long)K * n];
    __syncthreads();
    if (valid) {
      const int An = d.A[n];
      const long long idx = n + (long long)N * g;
      const long long c = n + (long long)N * (d.g0 + g);
      const double Eold = sE[n];
      const Stream st = make_stream(d.seed, iter, PUR_E, c);
      double x;
      if (An == 0 || d.nzP[n] == 0) {                               // R/sample_En.R:12,56
        x = prior_draw(d, st, 1, idx);
      } else {
        double num1 = 0.0, den = 0.0;
        for (int k = l; k < K; k += LPG) {
          const double mh = Mh[k], p = Pn[k];
          const double inv = normal ? inv_sg : 1.0 / mh;
          const double mh_no = mh - p * Eold;
          num1 += p * (((normal ? MvD[k] : (double)MvI[k]) - mh_no) * inv);
          den += (p * p) * inv;
        }
        num1 = seg_sum<LPG>(num1, vmask); den = seg_sum<LPG>(den, vmask);
        double mu, v;
        if (d.prior == PRIOR_EXPONENTIAL) {
          mu = (num1 - sQ1[n]) / den; v = 1.0 / den;
        } else {
          const double s2 = sQ2[n];
          den = den + 1.0 / s2;
          mu = (num1 + sQ1[n] / s2) / den; v = 1.0 / den;
        }
        x = stage ? truncnorm0_staged(st, mu, sqrt(v), vZ + n * P_PRE, vE + n * P_PRE, vU + n * P_PRE)
                  : truncnorm0_draw<double>(st, mu, sqrt(v));
      }
      x = (double)(T)x;
      double Enew = x;
      if (d.MH && An != 0) {
        if (!converged) {
          if (l == 0) d.E_acc[idx] = (T)1;                          // R/sample_En.R:198-201
        } else {
          double D = 0.0;
          const double de = x - Eold;
          for (int k = l; k < K; k += LPG) { const double mh = Mh[k]; D += mh_cell(normal ? MvD[k] : (double)MvI[k], mh, mh + Pn[k] * de); }
          D = seg_sum<LPG>(D, vmask);
          const double ratio = mh_ratio(D);
          if (l == 0) d.E_acc[idx] = (T)ratio;
          const double u = u01<double>(make_stream(d.seed, iter, PUR_MH_E, c).at(0).x);
          Enew = u < ratio ? x : Eold;
        }
      }
      if (An != 0 && Enew != Eold) {
        const double de = Enew - Eold;
        for (int k = l; k < K; k += LPG) Mh[k] = (double)(T)(Mh[k] + Pn[k] * de);
      }
      if (l == 0) {
        d.E[idx] = (T)Enew;
        if (Enew != 0.0) atomicOr(&d.nzE[(iter & 1) * N + n], 1);
      }
    }
    __syncwarp();
  }
  if (valid)
    for (int k = l; k < K; k += LPG) d.Mhat[k + (long long)K * g] = (T)Mh[k];
}

// ------------------------------------------------------------------------------
// k_e_gram: the E sweep of the NORMAL likelihood in Gram-matrix form (get_mu_sigmasq_En_normal +
// sample_En_normal, R/sample_En.R:131-184, :54-86).  With sigmasq[k,g] = sigmasq_g the sums over the
// mutation types of the reference's conditional of E[n,g] collapse onto two small objects,
//     num1 = ( (P' M)[n,g] - sum_{m != n} (P' P)[n,m] A_m E[m,g] ) / sigmasq_g ,  den = A_n (P' P)[n,n] / sigmasq_g
// (E[m,g] the current value: already redrawn for m < n) -- the transpose of what k_p_gram does for P.
// One thread per genome: its column of the data goes through shared memory once (coalesced), c = P' M[., g]
// is accumulated eight signatures at a time in registers, then the N conditionals are a chain of N-term
// dot products with P' P (shared memory, broadcast) and N truncated-normal draws.  Mhat is neither read nor
// kept up to date: it is rebuilt once after the sweep (mhat_rebuild -> the tensor cores) for sigmasq and the
// metrics.  Algebraically the reference's sums, not their summation order.
// ------------------------------------------------------------------------------
constexpr int EG_T = 192, EG_KC = 16, EG_PRE = 2;    // threads per block, rows of the data staged at a time, staged attempts per draw
constexpr int EG_ARR = 4 + 3 * EG_PRE;               // per-genome arrays of N doubles: c, E, two prior parameters, the variates
__host__ __device__ inline size_t e_gram_smem(int K, int N, int GB) {
  const int NPAD = (N + 7) & ~7;
  return ((size_t)K * NPAD + (size_t)N * N + (size_t)GB * (EG_KC + 1) + (size_t)EG_ARR * N * GB) * sizeof(double) + 64 * sizeof(int);
}
// Work layout: 192 threads per GB genomes (GB <= 96, chosen by the host so that the whole shard is resident in one
// wave of two blocks per SM when it fits).  The parallel phases use every thread (staging, P' P, the variates of the
// first EG_PRE attempts of every draw, P' M: thread (genome, half) accumulates eight signatures at a time in
// registers while the block streams the data through shared memory 16 mutation types at a time); the chain of the
// N conditionals is scalar work per genome and runs one thread per genome, with the later attempts of a rejected
// draw (rare after two staged ones) tried warp-uniformly.
template <typename T>
__global__ void __launch_bounds__(EG_T) k_e_gram(Dev<T> d, int GB) {
  extern __shared__ double egs[];
  const int K = d.K, N = d.N, NPAD = (N + 7) & ~7;
  double* Ps = egs;                                   // [K][NPAD]   P, zero padded
  double* Q = Ps + (size_t)K * NPAD;                  // [N][N]      (P' P)[n,m] A_m
  double* Ms = Q + (size_t)N * N;                     // [GB][EG_KC+1] rows kc .. kc+15 of the block's columns of the data
  double* cs = Ms + (size_t)GB * (EG_KC + 1);         // [N][GB]     P' M[., g]
  double* es = cs + (size_t)N * GB;                   // [N][GB]     E[., g], current
  double* q1 = es + (size_t)N * GB;                   // [N][GB]     Lambda_e | Mu_e
  double* q2 = q1 + (size_t)N * GB;                   // [N][GB]     Sigmasq_e
  double* vZ = q2 + (size_t)N * GB;                   // [EG_PRE][N][GB]  the variates of the first attempts of every draw (tn_variates)
  double* vE = vZ + (size_t)EG_PRE * N * GB;
  double* vU = vE + (size_t)EG_PRE * N * GB;
  int* nzs = reinterpret_cast<int*>(vU + (size_t)EG_PRE * N * GB);   // [N] "row n of E got a non-zero"
  const int t = threadIdx.x;
  const long long g0 = (long long)blockIdx.x * GB;
  const int ng = (int)(d.G - g0 < GB ? d.G - g0 : GB);
  const int iter = d.ctrl->iter;
  for (int i = t; i < K * NPAD; i += EG_T) { const int k = i / NPAD, n = i - k * NPAD; Ps[i] = n < N ? (double)d.P[k + (long long)K * n] : 0.0; }
  if (t < N) nzs[t] = 0;
  // everything the chain reads besides the two Gram objects, in parallel: prior parameters, E, the variates
  for (int i = t; i < N * ng; i += EG_T) {            // i = n + N * genome: E and the prior parameters are read coalesced
    const int gl = i / N, n = i - gl * N;
    const long long idx = (long long)N * g0 + i;
    es[n * GB + gl] = (double)d.E[idx];
    // the prior's two contributions to the conditional, with their divisions done here, outside the chain:
    //   q1 = what is added to num1 (-Lambda | Mu / Sigmasq),  q2 = what is added to den (0 | 1 / Sigmasq)
    if (d.prior == PRIOR_EXPONENTIAL) { q1[n * GB + gl] = -(double)d.Lambda_e[idx]; q2[n * GB + gl] = 0.0; }
    else { const double s2 = (double)d.Sigmasq_e[idx]; q1[n * GB + gl] = (double)d.Mu_e[idx] / s2; q2[n * GB + gl] = 1.0 / s2; }
  }
  for (int i = t; i < EG_PRE * N * ng; i += EG_T) {
    const int a = i / (N * ng), r = i - a * (N * ng), gl = r / N, n = r - gl * N;
    const size_t o = ((size_t)a * N + n) * GB + gl;
    tn_variates(make_stream(d.seed, iter, PUR_E, (long long)N * (d.g0 + g0) + r), a, vZ[o], vE[o], vU[o]);
  }
  __syncthreads();
  for (int i = t; i < N * N; i += EG_T) {
    const int a = i / N, b = i - a * N;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += Ps[k * NPAD + a] * Ps[k * NPAD + b];
    Q[i] = (a == b || d.A[b]) ? s : 0.0;              // the diagonal is den; off the diagonal excluded signatures drop out
  }
  // c = P' M[., g]: thread u takes genome u mod GB... two "halves" of threads split the signature groups of 8
  {
    const int HT = EG_T / 2;                          // 96 threads per half >= GB: a half covers the genomes in one pass
    const int h = t / HT, tl = t - h * HT;
    const T* src = d.Mr + (long long)K * g0;
    const int NG8 = NPAD / 8;
    const int rounds = (NG8 + 1) / 2;                 // both halves run the same number of rounds (an odd group count: the second half idles in the last)
    for (int r = 0; r < rounds; ++r) {
      const int n0 = 8 * (2 * r + h);
      const bool live = n0 < NPAD;
      for (int gb = 0; gb < GB; gb += HT) {           // genomes gb + tl
        const int gl = gb + tl;
        double acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0;
        for (int kc = 0; kc < K; kc += EG_KC) {
          __syncthreads();                            // (the chunk before has been consumed)
          const int rows = K - kc < EG_KC ? K - kc : EG_KC;
          for (int i = t; i < ng * EG_KC; i += EG_T) {  // 16 consecutive doubles of a column = one 128-byte line
            const int c = i / EG_KC, j = i - c * EG_KC;
            Ms[c * (EG_KC + 1) + j] = j < rows ? (double)src[(long long)c * K + kc + j] : 0.0;
          }
          __syncthreads();
          if (live && gl < ng) {
            const double* mrow = Ms + gl * (EG_KC + 1);
            for (int j = 0; j < rows; ++j) {
              const double m = mrow[j];
              const double* pr = Ps + (kc + j) * NPAD + n0;
#pragma unroll
              for (int u = 0; u < 8; ++u) acc[u] += pr[u] * m;
            }
          }
        }
        if (live && gl < ng) {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (n0 + j < N) cs[(n0 + j) * GB + gl] = acc[j];
        }
      }
    }
  }
  __syncthreads();
  // ---- the chain: one thread per genome (GB <= 96 < 128 threads) ----
  {
    const int gl = t;
    const bool valid = gl < ng;
    const unsigned gmask = __ballot_sync(0xffffffffu, valid);
    const long long g = g0 + gl;
    if (valid) {
      const double inv_sg = 1.0 / (double)d.sigmasq[g];
      for (int n = 0; n < N; ++n) {
        const int An = d.A[n];
        const long long idx = n + (long long)N * g;
        const Stream st = make_stream(d.seed, iter, PUR_E, n + (long long)N * (d.g0 + g));
        double x;
        if (An == 0 || d.nzP[n] == 0) {                               // R/sample_En.R:12,56
          x = prior_draw(d, st, 1, idx);
        } else {
          const double* qr = Q + n * N;
          double dot0 = 0.0, dot1 = 0.0;
          int m = 0;
          for (; m + 1 < N; m += 2) {
            if (m != n) dot0 += qr[m] * es[m * GB + gl];
            if (m + 1 != n) dot1 += qr[m + 1] * es[(m + 1) * GB + gl];
          }
          if (m < N && m != n) dot0 += qr[m] * es[m * GB + gl];
          // mu = (num1 + q1) / den, v = 1 / den (R/sample_En.R:150-184) with one division and one square root on the
          // chain: sd = sqrt(v), mu = (num1 + q1) v, alpha = -mu / sd = -(num1 + q1) sd
          const double num = (cs[n * GB + gl] - (dot0 + dot1)) * inv_sg + q1[n * GB + gl];
          const double den = qr[n] * inv_sg + q2[n * GB + gl];
          const double v = 1.0 / den;
          const double mu = num * v;
          const double sd = sqrt(v), alpha = -(num * sd);
          const double lam = alpha > 0.45 ? 0.5 * (alpha + sqrt(alpha * alpha + 4.0)) : 1.0;
          bool found = false;
          x = 0.0;
#pragma unroll
          for (int a = 0; a < EG_PRE; ++a) {
            const size_t o = ((size_t)a * N + n) * GB + gl;
            if (!found) {
              if (alpha <= 0.45) { const double zz = vZ[o]; if (zz >= alpha) { x = mu + sd * zz; x = x < 0.0 ? 0.0 : x; found = true; } }
              else { const double ee = vE[o] / lam; const double dz = (alpha + ee) - lam; if (vU[o] <= -0.5 * (dz * dz)) { x = sd * ee; found = true; } }
            }
          }
          // later attempts (same attempts, same order, same expressions as truncnorm0_draw).  This branch is uniform
          // over the warp's valid lanes (A_n and nzP[n] are); the loop runs until no genome of the warp needs another
          // attempt, so that the rejected lanes share their passes through the variate code
          for (uint32_t at = EG_PRE; __any_sync(gmask, !found) && at < 4096u; ++at) {
            if (!found) {
              double z, e, u;
              tn_variates(st, (int)at, z, e, u);
              if (alpha <= 0.45) { if (z >= alpha) { x = mu + sd * z; x = x < 0.0 ? 0.0 : x; found = true; } }
              else { const double ee = e / lam; const double dz = (alpha + ee) - lam; if (u <= -0.5 * (dz * dz)) { x = sd * ee; found = true; } }
            }
          }
          if (!found) x = alpha <= 0.45 ? fmax(mu + sd * alpha, 0.0) : 0.0;       // (truncnorm0_draw's value when every attempt fails)
        }
        x = (double)(T)x;
        es[n * GB + gl] = x;
        d.E[idx] = (T)x;
        if (x != 0.0) nzs[n] = 1;
      }
    }
  }
  __syncthreads();
  if (t < N && nzs[t]) atomicOr(&d.nzE[(iter & 1) * N + t], 1);
}

// ------------------------------------------------------------------------------
// Rank learning (R/sample_params.R:101-241).
// k_r: R | A  ~ categorical over 0..N with weight (q_r^{sum A} (1 - q_r)^{N - sum A})^T.
// ------------------------------------------------------------------------------
__device__ __forceinline__ double prior_prob_1(double R, int N) {   // compute_prior_prob_1, :178-187
  double q = R / (double)N;
  const double c = 0.4 / (double)N;
  if (q < c) q = c;
  if (q > 1.0 - c) q = 1.0 - c;
  return q;
}
template <typename T> __device__ __forceinline__ double temperature(const Dev<T>& d, int iter) {
  return (iter >= 1 && iter <= d.n_temps) ? d.temps[iter - 1] : 1.0;
}
template <typename T>
__global__ void k_r(Dev<T> d) {
  // one thread per candidate rank r = 0..N (three pow() each: serial they are 20 us of an iteration at N = 10),
  // then thread 0 walks the N + 1 weights in order -- the sums are those of the serial loop
  __shared__ double probs[65];
  const int N = d.N, iter = d.ctrl->iter;
  const double Tm = temperature(d, iter);
  int sA = 0;
  for (int n = 0; n < N; ++n) sA += d.A[n];
  for (int r = threadIdx.x; r <= N; r += blockDim.x) {
    const double q = prior_prob_1((double)r, N);
    probs[r] = (1.0 / (double)(N + 1)) * pow(pow(q, (double)sA) * pow(1.0 - q, (double)(N - sA)), Tm);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double tot = 0.0;
  for (int r = 0; r <= N; ++r) tot += probs[r];
  const double u = u01<double>(make_stream(d.seed, iter, PUR_R, 0).at(0).x);
  double cdf = 0.0;
  int r = 0;
  for (int j = 0; j <= N; ++j) { cdf += probs[j] / tot; if (cdf <= u) ++r; }
  *d.R = r < N ? r : N;
}

template <typename T, int THREADS>
__device__ __forceinline__ void a_draw_block(const Dev<T>& d, int n, double l0, double l1);

// k_a_pass: one streaming pass giving the log-likelihood with A_n = 0 and with A_n = 1
// (the two get_loglik calls of R/sample_params.R:115-116); a warp per genome column.
// FUSE = 1 (unsharded runs): the block that finishes last folds the per-block partials in a fixed
// order and draws A_n itself -- one launch per signature instead of three (k_a_reduce, k_a_draw).
template <typename T, int FUSE>
__global__ void __launch_bounds__(256) k_a_pass(Dev<T> d, int n, int n_prev, unsigned* ticket) {
  __shared__ double s0[32], s1[32];
  const int K = d.K, N = d.N;
  const int WPB = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long g = (long long)blockIdx.x * WPB + wid;
  const bool normal = d.likelihood == LIK_NORMAL;
  double l0 = 0.0, l1 = 0.0;
  if (g < d.G) {
    const int An = d.A[n];
    const double e = (double)d.E[n + (long long)N * g];
    const double eprev = n_prev >= 0 ? (double)d.E[n_prev + (long long)N * g] : 0.0;
    const double sg = normal ? (double)d.sigmasq[g] : 0.0;
    for (int k = lane; k < K; k += 32) {
      const long long i = k + (long long)K * g;
      double mh = (double)d.Mhat[i];
      if (n_prev >= 0) { mh = (double)(T)(mh + d.dvec[k] * eprev); d.Mhat[i] = (T)mh; }
      const double pe = (double)d.P[k + (long long)K * n] * e;
      const double mh0 = An ? mh - pe : mh, mh1 = An ? mh : mh + pe;
      const double m = Mat(d, i);
      if (normal) { l0 += dnorm_log(m, mh0, sg); l1 += dnorm_log(m, mh1, sg); }
      else {
        const double a = mh0 > 1e-6 ? mh0 : 1e-6, b = mh1 > 1e-6 ? mh1 : 1e-6;
        l0 += m * log(a) - a; l1 += m * log(b) - b;
      }
    }
  }
  l0 = warp_sum(l0); l1 = warp_sum(l1);
  if (lane == 0) { s0[wid] = l0; s1[wid] = l1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < WPB; ++w) { a += s0[w]; b += s1[w]; }
    d.apart[2 * (long long)blockIdx.x] = a; d.apart[2 * (long long)blockIdx.x + 1] = b;
  }
  if (FUSE) {
    __shared__ bool last;
    __shared__ double scratch[8];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    double a = 0.0, b = 0.0;
    const volatile double* ap = d.apart;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) { a += ap[2 * (long long)i]; b += ap[2 * (long long)i + 1]; }
    a = block_sum<256>(a, scratch);
    b = block_sum<256>(b, scratch);
    __shared__ double tot[2];
    if (threadIdx.x == 0) { tot[0] = a; tot[1] = b; *ticket = 0u; }
    __syncthreads();
    a_draw_block<T, 256>(d, n, tot[0], tot[1]);
  }
}

// k_a_draw: A_n ~ Bernoulli(p), SBFI / BFI (R/sample_params.R:118-165); one block.
// (a device function: also the tail of k_a_pass<.., FUSE = 1>)
template <typename T, int THREADS>
__device__ __forceinline__ void a_draw_block(const Dev<T>& d, int n, double l0, double l1) {
  __shared__ int s_delta;
  const int K = d.K, N = d.N;
  if (threadIdx.x == 0) {
    const int iter = d.ctrl->iter;
    if (d.likelihood == LIK_POISSON) { l0 += d.ll_const_all; l1 += d.ll_const_all; }
    const double q = prior_prob_1((double)*d.R, N);
    const double Tm = temperature(d, iter);
    const int Aold = d.A[n];
    int sA = 0;
    for (int j = 0; j < N; ++j) sA += d.A[j];
    const double sA0 = (double)(sA - Aold), sA1 = sA0 + 1.0;
    double lp0, lp1;
    if (d.rank_method == RANK_SBFI) {
      const double GK = (double)d.G_total + (double)K, lg = log((double)d.G_total);
      const double b0 = l0 - sA0 * GK * lg / 2.0, b1 = l1 - sA1 * GK * lg / 2.0;
      lp0 = log(1.0 - q) + Tm * b0; lp1 = log(q) + Tm * b1;
    } else {
      lp0 = log(1.0 - q) + Tm * l0; lp1 = log(q) + Tm * l1;
    }
    const double hi = lp0 > lp1 ? lp0 : lp1, lo = lp0 > lp1 ? lp1 : lp0;
    const double s = hi + log(1.0 + exp(lo - hi));                 // sumLog, :199-206
    double p = exp(lp1 - s);
    if (!(p == p)) {                                               // overflow branch, :143-163
      const bool n1 = !(lp1 == lp1), n0 = !(lp0 == lp0);
      if (n1 && n0) p = 0.5; else if (n1) p = 0.0; else if (n0) p = 1.0;
      else if (lp1 > lp0) p = 1.0; else if (lp1 < lp0) p = 0.0; else p = 0.5;
    }
    const double u = u01<double>(make_stream(d.seed, iter, PUR_A, (uint64_t)n).at(0).x);
    const int Anew = u < p ? 1 : 0;
    d.A[n] = Anew;
    s_delta = Anew - Aold;
  }
  __syncthreads();
  const double dl = (double)s_delta;
  for (int k = threadIdx.x; k < K; k += THREADS) d.dvec[k] = dl * (double)d.P[k + (long long)K * n];
}
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_a_draw(Dev<T> d, int n, const double* lsum) {
  a_draw_block<T, THREADS>(d, n, lsum[0], lsum[1]);
}

// ------------------------------------------------------------------------------
// k_a_sweep: the N passes of the rank learner (k_a_pass + its draw, above) as ONE cooperative launch for problems
// whose column blocks are all resident at once (C2: 63 blocks): a pass, a grid barrier, block 0 folds the partials
// in block order and draws A_n (a_draw_block), a grid barrier, the next signature.  Same sums in the same order as
// the per-signature launches; what goes is 2/3 of their 12 us apiece (launch, ticket, tail).
// `bar` counts arrivals (zeroed by the host before the launch); co-residency comes from the cooperative launch.
// ------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
  }
  __syncthreads();
}
template <typename T>
__global__ void __launch_bounds__(256) k_a_sweep(Dev<T> d, unsigned* bar) {
  __shared__ double s0[32], s1[32];
  __shared__ double scratch[8];
  __shared__ double tot[2];
  const int K = d.K, N = d.N;
  const int WPB = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long g = (long long)blockIdx.x * WPB + wid;
  const bool normal = d.likelihood == LIK_NORMAL;
  unsigned phase = 0;
  for (int n = 0; n < N; ++n) {
    const int n_prev = n - 1;
    double l0 = 0.0, l1 = 0.0;
    if (g < d.G) {
      const int An = d.A[n];
      const double e = (double)d.E[n + (long long)N * g];
      const double eprev = n_prev >= 0 ? (double)d.E[n_prev + (long long)N * g] : 0.0;
      const double sg = normal ? (double)d.sigmasq[g] : 0.0;
      for (int k = lane; k < K; k += 32) {
        const long long i = k + (long long)K * g;
        double mh = (double)d.Mhat[i];
        if (n_prev >= 0) { mh = (double)(T)(mh + __ldcg(&d.dvec[k]) * eprev); d.Mhat[i] = (T)mh; }
        const double pe = (double)d.P[k + (long long)K * n] * e;
        const double mh0 = An ? mh - pe : mh, mh1 = An ? mh : mh + pe;
        const double m = Mat(d, i);
        if (normal) { l0 += dnorm_log(m, mh0, sg); l1 += dnorm_log(m, mh1, sg); }
        else {
          const double a = mh0 > 1e-6 ? mh0 : 1e-6, b = mh1 > 1e-6 ? mh1 : 1e-6;
          l0 += m * log(a) - a; l1 += m * log(b) - b;
        }
      }
    }
    l0 = warp_sum(l0); l1 = warp_sum(l1);
    if (lane == 0) { s0[wid] = l0; s1[wid] = l1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < WPB; ++w) { a += s0[w]; b += s1[w]; }
      d.apart[2 * (long long)blockIdx.x] = a; d.apart[2 * (long long)blockIdx.x + 1] = b;
    }
    grid_barrier(bar, ++phase * gridDim.x);
    if (blockIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) { a += __ldcg(&d.apart[2 * (long long)i]); b += __ldcg(&d.apart[2 * (long long)i + 1]); }
      a = block_sum<256>(a, scratch);
      b = block_sum<256>(b, scratch);
      if (threadIdx.x == 0) { tot[0] = a; tot[1] = b; }
      __syncthreads();
      a_draw_block<T, 256>(d, n, tot[0], tot[1]);       // A[n], dvec
    }
    grid_barrier(bar, ++phase * gridDim.x);
  }
}

// ------------------------------------------------------------------------------
// k_final: last streaming pass of the iteration, a warp per genome column:
//   apply the pending update; sigmasq_g ~ InvGamma(Alpha_g + K/2, Beta_g + SSE_g / 2)
//   (R/sample_params.R:275-286); metric partials (log-likelihood, SSE, padded KL;
//   R/utils.R:412-455, :467-471); log prior of E[., g] and the E acceptance sum.
// ------------------------------------------------------------------------------
template <typename T>
__global__ void k_final(Dev<T> d, int n_prev, int keep_sigmasq) {
  __shared__ double sp[32][PC_COLS];
  const int K = d.K, N = d.N;
  const int WPB = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long g = (long long)blockIdx.x * WPB + wid;
  const bool normal = d.likelihood == LIK_NORMAL;
  double sse = 0.0, ll = 0.0, kl = 0.0, lpe = 0.0, eacc = 0.0;
  if (g < d.G) {
    const double eprev = n_prev >= 0 ? (double)d.E[n_prev + (long long)N * g] : 0.0;
    for (int k = lane; k < K; k += 32) {
      const long long i = k + (long long)K * g;
      double mh = (double)d.Mhat[i];
      if (n_prev >= 0) { mh = (double)(T)(mh + d.dvec[k] * eprev); d.Mhat[i] = (T)mh; }
      const double r = Mat(d, i) - mh;
      sse += r * r;
    }
    sse = warp_sum(sse);
    double sg = 0.0;
    if (normal) {
      if (!keep_sigmasq) {
        sg = (double)(T)(1.0 / gamma_draw<double>(make_stream(d.seed, d.ctrl->iter, PUR_SIGMASQ, (uint64_t)(d.g0 + g)),
                                                  (double)d.Alpha_g[g] + (double)K / 2.0, (double)d.Beta_g[g] + 0.5 * sse));
        if (lane == 0) d.sigmasq[g] = (T)sg;
      } else sg = (double)d.sigmasq[g];
    }
    for (int k = lane; k < K; k += 32) {
      const long long i = k + (long long)K * g;
      const double mh = (double)d.Mhat[i], m = Mat(d, i);
      const double lam = mh > 1e-6 ? mh : 1e-6;
      const double L = log(lam);
      kl -= (m > 1e-6 ? m : 1e-6) * L;
      ll += normal ? dnorm_log(m, mh, sg) : m * L - lam;
    }
    for (int n = lane; n < N; n += 32) {
      const long long idx = n + (long long)N * g;
      lpe += prior_logdens(d, 1, idx, (double)d.E[idx]);
      if (d.MH && d.A[n]) eacc += (double)d.E_acc[idx];
    }
    ll = warp_sum(ll); kl = warp_sum(kl); lpe = warp_sum(lpe); eacc = warp_sum(eacc);
  }
  if (lane == 0) { sp[wid][PC_SSE] = sse; sp[wid][PC_KLV] = kl; sp[wid][PC_LLV] = ll; sp[wid][PC_LP_E] = lpe; sp[wid][PC_EACC] = eacc; }
  __syncthreads();
  if (threadIdx.x < PC_COLS) {
    double s = 0.0;
    for (int w = 0; w < WPB; ++w) s += sp[w][threadIdx.x];
    d.epart[(long long)blockIdx.x * PC_COLS + threadIdx.x] = s;
  }
}

// k_pprior: log prior of column n of P and the sum of its acceptance rates (block n).
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_pprior(Dev<T> d) {
  __shared__ double scratch[THREADS / 32];
  const int n = blockIdx.x, K = d.K;
  double lp = 0.0, pa = 0.0;
  for (int k = threadIdx.x; k < K; k += THREADS) {
    const long long c = k + (long long)K * n;
    lp += prior_logdens(d, 0, c, (double)d.P[c]);
    if (d.MH) pa += (double)d.P_acc[c];
  }
  const double a = block_sum<THREADS>(lp, scratch);
  const double b = block_sum<THREADS>(pa, scratch);
  if (threadIdx.x == 0) { d.zpart[(long long)d.n_zitems * PC_COLS + n] = a; d.paccpart[n] = b; }
}

}  // namespace bnmf
