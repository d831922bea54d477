// Kernels of the Poisson (non-MH) Gibbs iteration: the data-augmented sampler with
// latent counts Z (R/sample_params.R:79-85, :253-265) and conjugate Gamma updates
// (R/sample_Pn.R:98-120, R/sample_En.R:97-119, R/sample_priors.R:284-397).
//
// One iteration = k_begin_iter -> k_pside -> k_eside -> [rank kernels] -> k_zstat
//                 -> k_reduce_partials -> [cross-shard sum] -> k_metrics
#pragma once
#include "bnmf_rng.cuh"
#include "bnmf_state.h"

namespace bnmf {

// ------------------------------------------------------------------------------
// warp helpers (fixed butterfly order => deterministic sums)
// ------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide deterministic sum of one double per thread; result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch /*THREADS/32*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) r += scratch[w];
  }
  return r;
}

// log densities used by get_logpost_ (R/utils.R:147-174)
__device__ __forceinline__ double dgamma_log(double x, double shape, double rate) {
  return shape * log(rate) - lgamma(shape) + (shape - 1.0) * log(x) - rate * x;
}
__device__ __forceinline__ double dexp_log(double x, double rate) { return log(rate) - rate * x; }

// ------------------------------------------------------------------------------
// k_begin_iter: advance the iteration counter, reset per-iteration scratch.
// ------------------------------------------------------------------------------
template <typename T>
__global__ void k_begin_iter(Dev<T> d, int* work_ctr, int n_ctr) {
  const int t = threadIdx.x;
  const int iter = d.ctrl->iter + 1;
  __syncthreads();
  if (t == 0) { d.ctrl->iter = iter; d.ctrl->row += 1; }
  for (int i = t; i < n_ctr; i += blockDim.x) work_ctr[i] = 0;
  if (t == 0) { *d.lp_P = 0.0; *d.pacc_sum = 0.0; }
  // non-zero flags written during this iteration (bnmf_mh.cuh)
  for (int n = t; n < d.N; n += blockDim.x) { d.nzP[n] = 0; d.nzE[(iter & 1) * d.N + n] = 0; }
}

// ------------------------------------------------------------------------------
// k_pside: P-side of one sweep for the Poisson / non-MH model.  Block n owns
// signature n (column n of P and of every *_p prior parameter):
//   prior parameters  (R/sample_priors.R:170-198; element-wise in (k,n), they read
//                      only P[k,n] of the previous iteration)
//   P[k,n] ~ Gamma(shape + SP[k,n], rate + A_n * rowSums(E)[n])   (R/sample_Pn.R:98-120)
//   colSums(P)[n], sum_k log prior(P[k,n])                         (R/utils.R:162-167)
// It consumes (and clears) SP and rowsumE_fx produced by the previous iteration.
// from_prior = 1 draws P from its prior (R/sample_Pn.R:12-30) and skips the
// hyper-updates; keepP = 1 leaves a user-supplied P untouched (skip = names(init_params)).
// ------------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_pside(Dev<T> d, int from_prior, int keepP) {
  __shared__ double scratch[THREADS / 32];
  const int n = blockIdx.x;
  const int K = d.K, N = d.N;
  const int iter = d.ctrl->iter;
  const int An = d.A[n];
  const double rsE = (double)d.rowsumE_fx[n] / RS_FX;
  double csum = 0.0, lp = 0.0;
  for (int k = threadIdx.x; k < K; k += THREADS) {
    const long long c = (long long)k + (long long)K * n;
    double Pold = (double)d.P[c];
    double Pnew = Pold;
    double lpc = 0.0;
    if (d.prior == PRIOR_GAMMA) {
      double al = (double)d.Alpha_p[c], be = (double)d.Beta_p[c];
      if (!from_prior) {
        be = gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_P1, c),
                                (double)d.A_p.at(c) + al, (double)d.B_p.at(c) + Pold);
        al = alpha_draw(make_stream(d.seed, iter, PUR_HYP_P2, c),
                        (double)d.C_p.at(c), (double)d.D_p.at(c), be, Pold);
        d.Beta_p[c] = (T)be; d.Alpha_p[c] = (T)al;
      }
      if (!keepP) {
        double shape = al, rate = be;
        if (!from_prior) { shape += (double)d.SP[c]; rate += An ? rsE : 0.0; }
        Pnew = gamma_draw<double>(make_stream(d.seed, iter, PUR_P, c), shape, rate);
      }
      lpc = dgamma_log((double)(T)Pnew, (double)(T)al, (double)(T)be);
    } else {  // PRIOR_EXPONENTIAL
      double la = (double)d.Lambda_p[c];
      if (!from_prior) {
        la = gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_P1, c),
                                (double)d.A_p.at(c) + 1.0, (double)d.B_p.at(c) + Pold);
        d.Lambda_p[c] = (T)la;
      }
      if (!keepP) {
        double shape = 1.0, rate = la;
        if (!from_prior) { shape += (double)d.SP[c]; rate += An ? rsE : 0.0; }
        Pnew = gamma_draw<double>(make_stream(d.seed, iter, PUR_P, c), shape, rate);
      }
      lpc = dexp_log((double)(T)Pnew, (double)(T)la);
    }
    d.P[c] = (T)Pnew;
    d.SP[c] = 0ull;
    if (d.ring_cap > 0) d.ring_P[(long long)d.ctrl->ring_pos * K * N + c] = (T)Pnew;
    csum += (double)(T)Pnew;
    lp += lpc;
  }
  double cs = block_sum<THREADS>(csum, scratch);
  double lps = block_sum<THREADS>(lp, scratch);
  if (threadIdx.x == 0) {
    d.colsumP[n] = (T)cs;
    d.rowsumE_fx[n] = 0ll;
    // deterministic: each block owns slot n of a small array folded by k_reduce_partials
    d.zpart[(long long)(d.n_zitems) * PC_COLS + n] = lps;
  }
}

// ------------------------------------------------------------------------------
// k_eside: E-side of one sweep, one thread per cell (n,g), idx = n + N*g (coalesced):
//   prior parameters (R/sample_priors.R:175-177,190-197; they read E of the
//                     previous iteration only)
//   E[n,g] ~ Gamma(shape + SE[n,g], rate + A_n * colSums(P)[n])    (R/sample_En.R:97-119)
//   rowSums(E) (fixed point, order-independent), sum log prior(E)   (R/utils.R:168-173)
// Consumes and clears SE.  Block b writes its partial to epart[b].
// ------------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_eside(Dev<T> d, int from_prior, int keepE) {
  __shared__ double scratch[THREADS / 32];
  __shared__ long long fx[THREADS];
  const int N = d.N;
  const long long cells = (long long)N * d.G;
  const long long base = (long long)blockIdx.x * THREADS;
  const long long idx = base + threadIdx.x;
  const int iter = d.ctrl->iter;
  double lp = 0.0;
  long long myfx = 0;
  if (idx < cells) {
    const int n = (int)(idx % N);
    const long long gl = idx / N;
    const long long c = (long long)n + (long long)N * (d.g0 + gl);  // global cell id
    const int An = d.A[n];
    const double csP = An ? (double)d.colsumP[n] : 0.0;
    double Eold = (double)d.E[idx], Enew = Eold;
    if (d.prior == PRIOR_GAMMA) {
      double al = (double)d.Alpha_e[idx], be = (double)d.Beta_e[idx];
      if (!from_prior) {
        be = gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_E1, c),
                                (double)d.A_e.at(idx) + al, (double)d.B_e.at(idx) + Eold);
        al = alpha_draw(make_stream(d.seed, iter, PUR_HYP_E2, c),
                        (double)d.C_e.at(idx), (double)d.D_e.at(idx), be, Eold);
        d.Beta_e[idx] = (T)be; d.Alpha_e[idx] = (T)al;
      }
      if (!keepE) {
        double shape = al, rate = be;
        if (!from_prior) { shape += (double)d.SE[idx]; rate += csP; }
        Enew = gamma_draw<double>(make_stream(d.seed, iter, PUR_E, c), shape, rate);
      }
      lp = dgamma_log((double)(T)Enew, (double)(T)al, (double)(T)be);
    } else {
      double la = (double)d.Lambda_e[idx];
      if (!from_prior) {
        la = gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_E1, c),
                                (double)d.A_e.at(idx) + 1.0, (double)d.B_e.at(idx) + Eold);
        d.Lambda_e[idx] = (T)la;
      }
      if (!keepE) {
        double shape = 1.0, rate = la;
        if (!from_prior) { shape += (double)d.SE[idx]; rate += csP; }
        Enew = gamma_draw<double>(make_stream(d.seed, iter, PUR_E, c), shape, rate);
      }
      lp = dexp_log((double)(T)Enew, (double)(T)la);
    }
    d.E[idx] = (T)Enew;
    d.SE[idx] = 0;
    if (d.ring_cap > 0) d.ring_E[(long long)d.ctrl->ring_pos * cells + idx] = (T)Enew;
    myfx = llrint((double)(T)Enew * RS_FX);
  }
  fx[threadIdx.x] = myfx;
  double lps = block_sum<THREADS>(lp, scratch);   // contains __syncthreads => fx visible
  if (threadIdx.x < N) {
    // thread j sums the cells of this block whose n == (base + j) % N ... i.e. offsets j, j+N, ...
    long long s = 0;
    for (int o = threadIdx.x; o < THREADS; o += N) s += fx[o];
    const int n = (int)((base + threadIdx.x) % N);
    atomicAdd((unsigned long long*)&d.rowsumE_fx[n], (unsigned long long)s);
  }
  if (threadIdx.x == 0) {
    double* ep = d.epart + (long long)blockIdx.x * PC_COLS;
    ep[PC_SSE] = 0.0; ep[PC_KLV] = 0.0; ep[PC_LLV] = 0.0; ep[PC_LP_E] = lps; ep[PC_EACC] = 0.0;
  }
}

// ------------------------------------------------------------------------------
// k_zstat: THE fused latent-count + sufficient-statistic kernel.
//
// For every cell (k,g) of the count matrix it samples
//     Z[k, . ,g] ~ Multinomial(M[k,g], p_n / sum p),  p_n = P[k,n] A_n E[n,g]
// (R/sample_params.R:253-265) as M[k,g] independent categorical picks by inverse
// CDF, and reduces them on the fly into SP[k,n] = sum_g Z and SE[n,g] = sum_k Z --
// the K x N x G tensor never exists.  Because sum_n p_n is Mhat[k,g] it also yields,
// for free, the per-iteration metrics of R/utils.R:412-455 (log-likelihood, RMSE,
// padded KL).
//
// Mapping: a warp owns a work item = 32 consecutive genomes (one per lane) x 32
// consecutive mutation types; lanes walk down the 32 rows together.  Per lane the
// column E[.,g] and the SE accumulators live in registers, the running CDF and the
// per-cell pick histogram in a lane-private, bank-conflict-free slice of shared
// memory ([n][thread]).  After each row the histogram is summed across the warp
// with REDUX and added to a block-level SP table in shared memory; SE leaves the
// SM once per item, SP once per block.  Items are handed out dynamically.
//
// Arithmetic contract (what oracle/ restates bit-for-bit):
//   p_n   = Pa[k,n] * E[n,g]            (Pa = P with excluded signatures zeroed)
//   cdf_n = cdf_{n-1} + p_n              (sequential, no FMA)
//   pick  = #{ n : cdf_n <= u * cdf_{N-1} },  u = (w + 0.5) 2^-32,
//           w = word (j mod 4) of Philox block j/4 of stream (iter, PUR_Z, k + K*g)
// ------------------------------------------------------------------------------
constexpr int ZT = 256;  // threads per block of k_zstat

template <int NP> struct NextPow2 {
  static constexpr int value = NP <= 4 ? 4 : NP <= 8 ? 8 : NP <= 16 ? 16 : NP <= 32 ? 32 : 64;
};

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }

template <typename T, int NP>
__global__ void __launch_bounds__(ZT, (NP <= 32 ? 2 : 1))
k_zstat(Dev<T> d, int KT, int* work_ctr) {
  constexpr int NP2 = NextPow2<NP>::value;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = d.K, N = d.N, G = d.G;
  T* Psm = reinterpret_cast<T*>(smem_raw);                 // [KT][NP]
  T* cdf = Psm + (size_t)KT * NP;                          // [NP2][ZT]
  int* hist = reinterpret_cast<int*>(cdf + (size_t)NP2 * ZT);  // [NP][ZT]
  int* spacc = hist + (size_t)NP * ZT;                     // [KT][N]
  __shared__ int s_item[ZT / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ky = blockIdx.y;            // k-tile
  const int k0 = ky * KT;
  const int krows = min(KT, K - k0);
  const int iter = d.ctrl->iter;

  // stage P (with A folded in) and clear block accumulators
  for (int i = tid; i < KT * NP; i += ZT) {
    int kk = i / NP, n = i - kk * NP;
    T v = (T)0;
    if (kk < krows && n < N && d.A[n]) v = d.P[(long long)(k0 + kk) + (long long)K * n];
    Psm[i] = v;
  }
  for (int i = tid; i < KT * N; i += ZT) spacc[i] = 0;
  for (int n = 0; n < NP; ++n) hist[n * ZT + tid] = 0;
  for (int n = NP; n < NP2; ++n) cdf[n * ZT + tid] = (T)INFINITY;
  __syncthreads();

  const int rts = (krows + 31) / 32;                 // row sub-tiles in this k-tile
  const int cts = (G + 31) / 32;                     // column tiles
  const int n_items = rts * cts;
  const int rts_all = ((d.K + KT - 1) / KT) * ((KT + 31) / 32);  // item id stride (for zpart)

  for (;;) {
    if (lane == 0) s_item[wid] = atomicAdd(&work_ctr[ky], 1);
    __syncwarp();
    const int item = s_item[wid];
    __syncwarp();
    if (item >= n_items) break;
    const int ct = item / rts, rt = item - ct * rts;
    const int g = ct * 32 + lane;
    const bool valid = g < G;
    const long long gg = d.g0 + g;

    T Ereg[NP];
    int se[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      Ereg[n] = (valid && n < N) ? d.E[(long long)n + (long long)N * g] : (T)0;
      se[n] = 0;
    }
    double a_sse = 0.0, a_kl = 0.0, a_ll = 0.0;

    const int kk_end = min(32, krows - rt * 32);
    for (int r = 0; r < kk_end; ++r) {
      const int kk = rt * 32 + r;
      const int k = k0 + kk;
      const int m = valid ? d.Mi[(long long)k + (long long)K * g] : 0;
      // running CDF
      T acc = (T)0;
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        acc = add_rn<T>(acc, mul_rn<T>(Psm[kk * NP + n], Ereg[n]));
        cdf[n * ZT + tid] = acc;
      }
      const T total = acc;
      if (valid) {
        const double mh = (double)total;
        const double lam = mh > 1e-6 ? mh : 1e-6;
        const double L = log(lam);
        const double md = (double)m;
        a_ll += md * L - lam;
        a_kl -= (m > 0 ? md : 1e-6) * L;
        const double diff = mh - md;
        a_sse += diff * diff;
      }
      const bool work = (m > 0) && (total > (T)0);
      if (__any_sync(0xffffffffu, work)) {
        if (work) {
          const Stream s = make_stream(d.seed, iter, PUR_Z, (unsigned long long)k + (unsigned long long)K * gg);
          for (int j = 0; j < m; j += 4) {
            const U4 w = s.at((uint32_t)(j >> 2));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            const int lim = min(4, m - j);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < lim) {
                const T t = mul_rn<T>(u01<T>(ww[q]), total);
                int pos = 0;
#pragma unroll
                for (int st = NP2 / 2; st > 0; st >>= 1)
                  if (cdf[(pos + st - 1) * ZT + tid] <= t) pos += st;
                if (pos > NP - 1) pos = NP - 1;   // unreachable for finite totals
                hist[pos * ZT + tid] += 1;
              }
            }
          }
        }
        __syncwarp();
        int mytot0 = 0, mytot1 = 0;
#pragma unroll
        for (int n = 0; n < NP; ++n) {
          const int v = hist[n * ZT + tid];
          hist[n * ZT + tid] = 0;
          se[n] += v;
          const int tot = __reduce_add_sync(0xffffffffu, v);
          if (n < 32) { if (lane == n) mytot0 = tot; }
          else        { if (lane == n - 32) mytot1 = tot; }
        }
        if (lane < N && mytot0) atomicAdd(&spacc[kk * N + lane], mytot0);
        if (NP > 32 && lane + 32 < N && mytot1) atomicAdd(&spacc[kk * N + lane + 32], mytot1);
      }
    }
    // SE leaves the SM once per item
    if (valid) {
#pragma unroll
      for (int n = 0; n < NP; ++n)
        if (n < N && se[n]) atomicAdd(&d.SE[(long long)n + (long long)N * g], se[n]);
    }
    // per-item metric partials, fixed reduction order
    a_sse = warp_sum(a_sse); a_kl = warp_sum(a_kl); a_ll = warp_sum(a_ll);
    if (lane == 0) {
      const long long gi = (long long)ct * rts_all + (long long)ky * ((KT + 31) / 32) + rt;
      double* zp = d.zpart + gi * PC_COLS;
      zp[PC_SSE] = a_sse; zp[PC_KLV] = a_kl; zp[PC_LLV] = a_ll; zp[PC_LP_E] = 0.0; zp[PC_EACC] = 0.0;
    }
  }
  __syncthreads();
  for (int i = tid; i < krows * N; i += ZT) {
    const int v = spacc[i];
    if (v) {
      const int kk = i / N, n = i - kk * N;
      atomicAdd(&d.SP[(long long)(k0 + kk) + (long long)K * n], (unsigned long long)v);
    }
  }
}

// ------------------------------------------------------------------------------
// k_reduce_partials: fold per-item / per-block partials in a fixed order (one block).
// red[c] = sum_items zpart[.][c] + sum_blocks epart[.][c]  (+ data constants)
// lp_P   = sum_n zpart[n_zitems + n]   (slots written by k_pside)
// ------------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_reduce_partials(Dev<T> d) {
  __shared__ double scratch[THREADS / 32];
  for (int c = 0; c < PC_COLS; ++c) {
    double v = 0.0;
    for (int i = threadIdx.x; i < d.n_zitems; i += THREADS) v += d.zpart[(long long)i * PC_COLS + c];
    for (int i = threadIdx.x; i < d.n_eblocks; i += THREADS) v += d.epart[(long long)i * PC_COLS + c];
    double s = block_sum<THREADS>(v, scratch);
    if (threadIdx.x == 0) {
      if (c == PC_LLV) s += d.ll_const;
      if (c == PC_KLV) s += d.kl_const;
      d.red[c] = s;
    }
  }
  double v = 0.0;
  for (int i = threadIdx.x; i < d.N; i += THREADS) v += d.zpart[(long long)d.n_zitems * PC_COLS + i];
  double s = block_sum<THREADS>(v, scratch);
  if (threadIdx.x == 0) *d.lp_P = s;
}

// ------------------------------------------------------------------------------
// k_metrics: compose the sample_metrics row (R/utils.R:339-348, :412-455), record A
// into the ring and advance it (record_sample, R/bayesNMF_sampler.R:651-672).
// ------------------------------------------------------------------------------
template <typename T>
__global__ void k_metrics(Dev<T> d) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Ctrl* c = d.ctrl;
  const int row = c->row;
  int rank = 0;
  for (int n = 0; n < d.N; ++n) rank += d.A[n];
  const double cellsKG = (double)d.K * (double)d.G_total;
  const double loglik = d.red[PC_LLV];
  const double nparams = (double)rank * ((double)d.G_total + (double)d.K);
  double* m = d.metrics + (long long)row * MC_COLS;
  m[MC_ITER] = (double)c->iter;
  m[MC_RMSE] = sqrt(d.red[PC_SSE] / cellsKG);
  m[MC_KL] = d.red[PC_KLV];
  m[MC_LOGLIK] = loglik;
  m[MC_LOGPOST] = loglik + *d.lp_P + d.red[PC_LP_E];
  m[MC_NPARAMS] = nparams;
  m[MC_BIC] = -2.0 * loglik + nparams * log((double)d.G_total);
  m[MC_RANK] = (double)rank;
  m[MC_TEMP] = (c->iter >= 1 && c->iter <= d.n_temps) ? d.temps[c->iter - 1] : 1.0;
  if (d.MH) {
    // mean over the active signatures (R/utils.R:444-452); NaN when none is active
    double pa = 0.0;
    for (int n = 0; n < d.N; ++n) if (d.A[n]) pa += d.paccpart[n];
    const double na = rank > 0 ? (double)rank : nan("");
    m[MC_PACC] = pa / (na * (double)d.K);
    m[MC_EACC] = d.red[PC_EACC] / (na * (double)d.G_total);
  } else {
    m[MC_PACC] = 1.0; m[MC_EACC] = 1.0;
  }
  if (d.ring_cap > 0) {
    for (int n = 0; n < d.N; ++n) d.ring_A[(long long)c->ring_pos * d.N + n] = d.A[n];
    c->ring_pos = (c->ring_pos + 1) % d.ring_cap;
    if (c->ring_count < d.ring_cap) c->ring_count += 1;
  }
}

}  // namespace bnmf
