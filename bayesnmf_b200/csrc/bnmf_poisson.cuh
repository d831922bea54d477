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

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
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
// k_pside_hyper: Beta_p / Alpha_p of iteration `iter` from P and the prior parameters of the
// iteration before (R/sample_priors.R:284-324, :356-390); like k_eside_hyper it runs on the side
// stream under the previous k_zstat, which takes two of the three samplers of k_pside off the
// critical path of an iteration (one block per signature: pure latency).
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_pside_hyper(Dev<T> d, int iter_arg) {
  const int iter = iter_arg < 0 ? d.ctrl->iter : iter_arg;      // < 0: the iteration under way (launches replayed from a graph)
  const int n = blockIdx.x, K = d.K;
  for (int k = threadIdx.x; k < K; k += THREADS) {
    const long long c = (long long)k + (long long)K * n;
    const double Pold = (double)d.P[c], al0 = (double)d.Alpha_p[c];
    const double be = (double)(T)gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_P1, c),
                                                    (double)d.A_p.at(c) + al0, (double)d.B_p.at(c) + Pold);
    const double al = (double)(T)alpha_draw(make_stream(d.seed, iter, PUR_HYP_P2, c),
                                            (double)d.C_p.at(c), (double)d.D_p.at(c), be, Pold, al0);
    d.Beta_p[c] = (T)be; d.Alpha_p[c] = (T)al;
  }
}

template <typename T, int THREADS, int PRIOR, int FROM_PRIOR, int HYPER_DONE = 0>
__global__ void __launch_bounds__(THREADS) k_pside(Dev<T> d, int keepP, int* begin_ctr = nullptr, int n_ctr = 0, unsigned* begin_ticket = nullptr) {
  constexpr int from_prior = FROM_PRIOR || HYPER_DONE;      // (hyper-draws already made: take them as stored)
  constexpr int prior_draw_only = FROM_PRIOR;
  __shared__ double scratch[THREADS / 32];
  const int n = blockIdx.x;
  const int K = d.K, N = d.N;
  // begin_ctr != nullptr: this launch is also k_begin_iter (one launch less on the chain of an iteration that is
  // not replayed from a graph): every block works on iteration ctrl->iter + 1, the last one to finish does what
  // k_begin_iter does -- by then every block has read the counter
  const int iter = d.ctrl->iter + (begin_ctr ? 1 : 0);
  const int An = d.A[n];
  const double rsE = (double)d.rowsumE_fx[n] / RS_FX;
  double csum = 0.0, lp = 0.0;
  for (int k = threadIdx.x; k < K; k += THREADS) {
    const long long c = (long long)k + (long long)K * n;
    double Pold = (double)d.P[c];
    double Pnew = Pold;
    double lpc = 0.0;
    if (PRIOR == PRIOR_GAMMA) {
      double al = (double)d.Alpha_p[c], be = (double)d.Beta_p[c];
      if (!from_prior) {
        // (double)(T): every later use sees the value as stored in the state precision
        be = (double)(T)gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_P1, c),
                                           (double)d.A_p.at(c) + al, (double)d.B_p.at(c) + Pold);
        al = (double)(T)alpha_draw(make_stream(d.seed, iter, PUR_HYP_P2, c),
                                   (double)d.C_p.at(c), (double)d.D_p.at(c), be, Pold, al);
        d.Beta_p[c] = (T)be; d.Alpha_p[c] = (T)al;
      }
      if (!keepP) {
        double shape = al, rate = be;
        if (!prior_draw_only) { shape += (double)d.SP[c]; rate += An ? rsE : 0.0; }
        Pnew = gamma_draw<double>(make_stream(d.seed, iter, PUR_P, c), shape, rate);
      }
      lpc = dgamma_log((double)(T)Pnew, (double)(T)al, (double)(T)be);
    } else {  // PRIOR_EXPONENTIAL
      double la = (double)d.Lambda_p[c];
      if (!from_prior) {
        la = (double)(T)gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_P1, c),
                                           (double)d.A_p.at(c) + 1.0, (double)d.B_p.at(c) + Pold);
        d.Lambda_p[c] = (T)la;
      }
      if (!keepP) {
        double shape = 1.0, rate = la;
        if (!prior_draw_only) { shape += (double)d.SP[c]; rate += An ? rsE : 0.0; }
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
  if (begin_ctr) {
    __shared__ bool last;
    if (threadIdx.x == 0) { __threadfence(); last = atomicAdd(begin_ticket, 1u) == gridDim.x - 1; }
    __syncthreads();
    if (!last) return;
    for (int i = threadIdx.x; i < n_ctr; i += THREADS) begin_ctr[i] = 0;
    for (int j = threadIdx.x; j < N; j += THREADS) { d.nzP[j] = 0; d.nzE[(iter & 1) * N + j] = 0; }
    if (threadIdx.x == 0) { d.ctrl->iter = iter; d.ctrl->row += 1; *d.lp_P = 0.0; *d.pacc_sum = 0.0; *begin_ticket = 0u; }
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
// k_eside_hyper: the two hyper-draws of the gamma prior on the E side for iteration `iter`,
//   Beta_e[n,g]  ~ Gamma(A_e + Alpha_e, B_e + E)                      (R/sample_priors.R:325-344)
//   Alpha_e[n,g] ~ its log-concave conditional given the new Beta_e    (R/sample_priors.R:356-397)
// as a kernel of their own.  They read nothing but E and the prior parameters of the previous
// iteration -- not P, not the latent counts -- so the host launches them for iteration t+1 on a
// low-priority side stream while k_zstat of iteration t is running: the fp64 pipe that k_zstat
// leaves idle does 3/4 of the E side's work for free (64 registers x 256 threads fit next to the
// two k_zstat blocks of an SM).  k_eside<..., HYPER_DONE = 1> then takes the values as stored.
// `iter` comes by value: the device's counter may already have moved on when a block starts.
// The Alpha draw is a rejection sampler (1.19 attempts per cell): a warp that loops until its slowest
// lane accepts runs ~2.9 passes with most lanes idle.  Here every cell makes attempt 0 in place; the
// rejected ones (16 %) park their envelope in a global list (`retry`: 21 columns of `cap` doubles --
// shared memory belongs to k_zstat, which runs on the same SMs) and k_alpha_retry finishes them with
// full warps.  Attempt a of a cell reads block a of the cell's Philox stream wherever it is made, so
// the draws are those of alpha_draw() (bnmf_rng.cuh), which the fused k_eside and the oracle use.
#define BNMF_ALPHA_ENV_COLS 21
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_eside_hyper(Dev<T> d, int iter_arg, double* __restrict__ retry, int* __restrict__ n_retry, int cap) {
  const int iter = iter_arg < 0 ? d.ctrl->iter : iter_arg;
  const long long cells = (long long)d.N * d.G;
  const long long idx = (long long)blockIdx.x * THREADS + threadIdx.x;
  const long long ii = idx < cells ? idx : cells - 1;     // the whole block runs the staged sampler (its barriers)
  const long long c = (ii % d.N) + (long long)d.N * (d.g0 + ii / d.N);
  const double al0 = (double)d.Alpha_e[ii], Eold = (double)d.E[ii];
  const double be = (double)(T)gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_E1, c),
                                                  (double)d.A_e.at(ii) + al0, (double)d.B_e.at(ii) + Eold);
  __syncthreads();
  AlphaEnv e;
  alpha_setup<true>(e, (double)d.C_e.at(ii), (double)d.D_e.at(ii), be, Eold, al0);
  __syncthreads();
  const Stream st = make_stream(d.seed, iter, PUR_HYP_E2, c);
  double x = e.m;
  bool store = idx < cells;
  bool done = alpha_attempt(e, st, 0u, x) || idx >= cells;
  // rejected cells take consecutive slots of the list (one atomic per warp)
  const unsigned rej = __ballot_sync(0xffffffffu, !done);
  if (rej) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs((int)rej) - 1) base = atomicAdd(n_retry, __popc(rej));
    base = __shfl_sync(0xffffffffu, base, __ffs((int)rej) - 1);
    if (!done) {
      const int slot = base + __popc(rej & ((1u << lane) - 1u));
      if (slot < cap) {
        double* r = retry + slot;
        const size_t cs = (size_t)cap;
        r[0 * cs] = e.t.cm1; r[1 * cs] = e.t.b;
#pragma unroll
        for (int q = 0; q < 3; ++q) { r[(2 + q) * cs] = e.hv[q]; r[(5 + q) * cs] = e.sl[q]; r[(8 + q) * cs] = e.xs[q]; r[(15 + q) * cs] = e.mass[q]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) r[(11 + q) * cs] = e.z[q];
        r[18 * cs] = e.tot; r[19 * cs] = e.m;
        r[20 * cs] = __longlong_as_double(ii);
        store = false;                                    // Alpha_e[ii] is written by k_alpha_retry
      } else {                                            // list full: finish here
        for (uint32_t a = 1; a < BNMF_ALPHA_MAX_ATTEMPTS; ++a) if (alpha_attempt(e, st, a, x)) break;
      }
    }
  }
  if (idx < cells) d.Beta_e[ii] = (T)be;
  if (store) d.Alpha_e[ii] = (T)x;
}

// Attempts 1, 2, ... of the cells k_eside_hyper parked.  A warp owns ALPHA_RETRY_PER_WARP consecutive
// list entries; a lane whose cell has accepted takes the warp's next entry instead of idling until the
// slowest lane is through (a 16 % rejection rate per attempt would otherwise cost ~3 passes per warp).
#define ALPHA_RETRY_PER_WARP 128
template <typename T>
__device__ __forceinline__ void alpha_env_load(const Dev<T>& d, int iter, const double* __restrict__ retry, size_t cs, int slot,
                                               AlphaEnv& e, long long& ii, Stream& st) {
  const double* r = retry + slot;
  e.t.cm1 = r[0 * cs]; e.t.b = r[1 * cs];
#pragma unroll
  for (int q = 0; q < 3; ++q) { e.hv[q] = r[(2 + q) * cs]; e.sl[q] = r[(5 + q) * cs]; e.xs[q] = r[(8 + q) * cs]; e.mass[q] = r[(15 + q) * cs]; }
#pragma unroll
  for (int q = 0; q < 4; ++q) e.z[q] = r[(11 + q) * cs];
  e.tot = r[18 * cs]; e.m = r[19 * cs];
  ii = __double_as_longlong(r[20 * cs]);
  const long long c = (ii % d.N) + (long long)d.N * (d.g0 + ii / d.N);
  st = make_stream(d.seed, iter, PUR_HYP_E2, c);
}
template <typename T>
__global__ void __launch_bounds__(256) k_alpha_retry(Dev<T> d, int iter_arg, const double* __restrict__ retry, const int* __restrict__ n_retry, int cap) {
  const int iter = iter_arg < 0 ? d.ctrl->iter : iter_arg;
  const int n = min(*n_retry, cap);
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long w0 = (long long)warp * ALPHA_RETRY_PER_WARP;
  if (w0 >= n) return;
  const int w1 = (int)min((long long)n, w0 + ALPHA_RETRY_PER_WARP);
  const size_t cs = (size_t)cap;
  int next = (int)w0 + 32;                    // first entry nobody has taken yet (warp-uniform)
  int slot = (int)w0 + lane;
  bool busy = slot < w1;
  AlphaEnv e; long long ii = 0; Stream st;
  uint32_t a = 1;
  double x = 0.0;
  if (busy) { alpha_env_load<T>(d, iter, retry, cs, slot, e, ii, st); x = e.m; }
  while (__any_sync(0xffffffffu, busy)) {
    bool acc = false;
    if (busy) {
      acc = alpha_attempt(e, st, a, x) || a + 1 >= BNMF_ALPHA_MAX_ATTEMPTS;
      ++a;
      if (acc) d.Alpha_e[ii] = (T)x;
    }
    const unsigned freed = __ballot_sync(0xffffffffu, busy && acc);
    if (busy && acc) {
      slot = next + __popc(freed & ((1u << lane) - 1u));
      busy = slot < w1;
      if (busy) { alpha_env_load<T>(d, iter, retry, cs, slot, e, ii, st); x = e.m; a = 1; }
    }
    next += __popc(freed);
  }
}

template <typename T, int THREADS, int PRIOR, int FROM_PRIOR, int HYPER_DONE = 0>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_eside(Dev<T> d, int keepE) {
  constexpr int from_prior = FROM_PRIOR || HYPER_DONE;      // (hyper-draws already made: take them as stored)
  constexpr int prior_draw_only = FROM_PRIOR;
  __shared__ double scratch[THREADS / 32];
  __shared__ long long fx[THREADS];
  const int N = d.N;
  const long long cells = (long long)N * d.G;
  const long long base = (long long)blockIdx.x * THREADS;
  const long long idx = base + threadIdx.x;
  const int iter = d.ctrl->iter;
  double lp = 0.0;
  long long myfx = 0;
  // Threads past the end redo the last cell (and store nothing) so that the whole block runs the
  // same stages; the barriers between the stages keep the eight warps of a block inside the
  // same stretch of code -- the samplers are long, and a block whose warps drift apart
  // stalls on instruction fetch.
  const bool act = idx < cells;
  const long long ii = act ? idx : cells - 1;
  {
    const int n = (int)(ii % N);
    const long long gl = ii / N;
    const long long c = (long long)n + (long long)N * (d.g0 + gl);  // global cell id
    const int An = d.A[n];
    const double csP = An ? (double)d.colsumP[n] : 0.0;
    double Eold = (double)d.E[ii], Enew = Eold;
    if (PRIOR == PRIOR_GAMMA) {
      double al = (double)d.Alpha_e[ii], be = (double)d.Beta_e[ii];
      if (!from_prior) {
        be = (double)(T)gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_E1, c),
                                           (double)d.A_e.at(ii) + al, (double)d.B_e.at(ii) + Eold);
        __syncthreads();
        al = (double)(T)alpha_draw<true>(make_stream(d.seed, iter, PUR_HYP_E2, c),
                                         (double)d.C_e.at(ii), (double)d.D_e.at(ii), be, Eold, al);
        __syncthreads();
        if (act) { d.Beta_e[ii] = (T)be; d.Alpha_e[ii] = (T)al; }
      }
      if (!keepE) {
        double shape = al, rate = be;
        if (!prior_draw_only) { shape += (double)d.SE[ii]; rate += csP; }
        Enew = gamma_draw<double>(make_stream(d.seed, iter, PUR_E, c), shape, rate);
      }
      __syncthreads();
      lp = dgamma_log((double)(T)Enew, (double)(T)al, (double)(T)be);
    } else {
      double la = (double)d.Lambda_e[ii];
      if (!from_prior) {
        la = (double)(T)gamma_draw<double>(make_stream(d.seed, iter, PUR_HYP_E1, c),
                                           (double)d.A_e.at(ii) + 1.0, (double)d.B_e.at(ii) + Eold);
        __syncthreads();
        if (act) d.Lambda_e[ii] = (T)la;
      }
      if (!keepE) {
        double shape = 1.0, rate = la;
        if (!prior_draw_only) { shape += (double)d.SE[ii]; rate += csP; }
        Enew = gamma_draw<double>(make_stream(d.seed, iter, PUR_E, c), shape, rate);
      }
      lp = dexp_log((double)(T)Enew, (double)(T)la);
    }
    if (act) {
      d.E[ii] = (T)Enew;
      d.SE[ii] = 0;
      if (d.ring_cap > 0) d.ring_E[(long long)d.ctrl->ring_pos * cells + ii] = (T)Enew;
      myfx = llrint((double)(T)Enew * RS_FX);
    } else lp = 0.0;
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
// k_sides: k_pside and k_eside of a steady-state iteration of the gamma-prior model (hyper-draws already made on
// the side stream) as ONE launch.  Blocks [0, N) are the P side (block n = signature n, as k_pside); the others
// are the E side (as k_eside): they form everything of their draw that does not need the column sums of the new
// P -- the loads, the unit-rate Gamma variate (its shape is Alpha_e + SE) -- and only then wait for the N P-side
// blocks to have published colSums(P) (a counter that grows by N per launch: nothing to reset; blocks are handed
// out in index order, the P-side blocks are resident before any block that waits for them).  One launch and the
// P side's latency less on the chain exchange -> P -> E -> Z of an iteration (an 8-GPU shard: ~8 us of 270).
// `iter` comes by value: the launch is also k_begin_iter (k_pside's begin_ctr), whose update of the device's counter
// the E-side blocks must not race with.  Draw for draw the two kernels it replaces.
// ------------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_sides(Dev<T> d, int iter, int* begin_ctr, int n_ctr, unsigned* begin_ticket, unsigned long long* pub, unsigned long long pub_target) {
  __shared__ double scratch[THREADS / 32];
  __shared__ long long fx[THREADS];
  const int K = d.K, N = d.N;
  if ((int)blockIdx.x < N) {
    // ---- P side of signature n (k_pside<T, THREADS, PRIOR_GAMMA, 0, 1>) ----
    const int n = blockIdx.x;
    const int An = d.A[n];
    const double rsE = (double)d.rowsumE_fx[n] / RS_FX;
    double csum = 0.0, lp = 0.0;
    // the first 128 threads, 128 apart -- k_pside's assignment of mutation types to threads, so that the two sums
    // below are formed in its order for any K (the other warps add zeros)
    for (int k = threadIdx.x; k < K && threadIdx.x < 128; k += 128) {
      const long long c = (long long)k + (long long)K * n;
      const double al = (double)d.Alpha_p[c], be = (double)d.Beta_p[c];
      const double Pnew = gamma_draw<double>(make_stream(d.seed, iter, PUR_P, c), al + (double)d.SP[c], be + (An ? rsE : 0.0));
      lp += dgamma_log((double)(T)Pnew, (double)(T)al, (double)(T)be);
      d.P[c] = (T)Pnew;
      d.SP[c] = 0ull;
      if (d.ring_cap > 0) d.ring_P[(long long)d.ctrl->ring_pos * K * N + c] = (T)Pnew;
      csum += (double)(T)Pnew;
    }
    const double cs = block_sum<THREADS>(csum, scratch);
    const double lps = block_sum<THREADS>(lp, scratch);
    __shared__ bool last;
    if (threadIdx.x == 0) {
      d.colsumP[n] = (T)cs;
      d.rowsumE_fx[n] = 0ll;
      d.zpart[(long long)(d.n_zitems) * PC_COLS + n] = lps;
      __threadfence();
      atomicAdd(pub, 1ull);                                   // colSums(P)[n] is out
      last = atomicAdd(begin_ticket, 1u) == (unsigned)N - 1u;
    }
    __syncthreads();
    if (!last) return;
    // the last P-side block is k_begin_iter (no block of this launch reads the device's counter)
    for (int i = threadIdx.x; i < n_ctr; i += THREADS) begin_ctr[i] = 0;
    for (int j = threadIdx.x; j < N; j += THREADS) { d.nzP[j] = 0; d.nzE[(iter & 1) * N + j] = 0; }
    if (threadIdx.x == 0) { d.ctrl->iter = iter; d.ctrl->row += 1; *d.lp_P = 0.0; *d.pacc_sum = 0.0; *begin_ticket = 0u; }
    return;
  }
  // ---- E side (k_eside<T, THREADS, PRIOR_GAMMA, 0, 1>) ----
  const int eb = (int)blockIdx.x - N;
  const long long cells = (long long)N * d.G;
  const long long base = (long long)eb * THREADS;
  const long long idx = base + threadIdx.x;
  const bool act = idx < cells;
  const long long ii = act ? idx : cells - 1;
  const int n = (int)(ii % N);
  const long long c = (long long)n + (long long)N * (d.g0 + ii / N);  // global cell id
  const int An = d.A[n];
  const double al = (double)d.Alpha_e[ii], be = (double)d.Beta_e[ii];
  const double g0 = gamma_unit<double>(make_stream(d.seed, iter, PUR_E, c), al + (double)d.SE[ii]);
  if (threadIdx.x == 0) {
    while (ld_acquire_gpu_u64(pub) < pub_target) __nanosleep(20);
  }
  __syncthreads();
  const double csP = An ? (double)__ldcg(&d.colsumP[n]) : 0.0;
  const double Enew = gamma_scale<double>(g0, be + csP);
  double lp = dgamma_log((double)(T)Enew, (double)(T)al, (double)(T)be);
  long long myfx = 0;
  if (act) {
    d.E[ii] = (T)Enew;
    d.SE[ii] = 0;
    if (d.ring_cap > 0) d.ring_E[(long long)d.ctrl->ring_pos * cells + ii] = (T)Enew;
    myfx = llrint((double)(T)Enew * RS_FX);
  } else lp = 0.0;
  fx[threadIdx.x] = myfx;
  const double lps = block_sum<THREADS>(lp, scratch);   // contains __syncthreads => fx visible
  if (threadIdx.x < N) {
    long long s = 0;
    for (int o = threadIdx.x; o < THREADS; o += N) s += fx[o];
    const int nn = (int)((base + threadIdx.x) % N);
    atomicAdd((unsigned long long*)&d.rowsumE_fx[nn], (unsigned long long)s);
  }
  if (threadIdx.x == 0) {
    double* ep = d.epart + (long long)eb * PC_COLS;
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
// Mapping: a warp owns a work item = 32 consecutive genomes (one per lane) x ZR
// consecutive mutation types and walks down the rows.  Per row:
//   phase 1 (lane = cell)   the lane accumulates the running CDF of its cell and turns it into
//                           32-bit integer pick thresholds in a warp-private shared-memory
//                           table ([n][cell]: bank = cell); metric partials; E[.,g] and the
//                           SE accumulators stay in registers.
//   phase 2 (lane = share)  counts per cell are heavy-tailed (max/mean ~ 4 over a warp), so
//                           the picks of the 32 cells are cut into quads (4 picks = one
//                           Philox block), laid end to end, and every lane takes an equal
//                           contiguous share of the row's quads, whatever cells they belong
//                           to.  A pick is a branch-free binary search of the cell's threshold
//                           column (the two top levels sit in registers) that ends on the
//                           address of its histogram counter.  A cell that starts inside a
//                           lane's share is counted straight into the cell's histogram column
//                           (exclusive, no atomics); the part of a cell that spills into
//                           following lanes is counted in those lanes' private columns and
//                           folded in by a short, conflict-free fix-up (lanes continuing the
//                           same cell take turns).
//   reduce                  REDUX sums the histogram columns across the warp into the
//                           block-level SP table; each lane adds its own column to SE.
// SE leaves the SM once per item, SP once per block.  Items are handed out dynamically.
//
// Arithmetic contract (what oracle/ restates bit-for-bit; T = state precision):
//   p_n   = Pa[k,n] * E[n,g]            (Pa = P with excluded signatures zeroed)
//   cdf_n = cdf_{n-1} + p_n              (sequential, one rounding per operation, no FMA)
//   thr_n = sat_u32(floor(cdf_n * (2^32 / cdf_{N-1})))   for n < N-1   (NaN -> 0)
//   pick  = #{ n < N-1 : thr_n <= min(w, 2^32 - 2) },
//           w = word (j mod 4) of Philox block j/4 of stream (iter, PUR_Z, k + K*g)
// i.e. the fp64 inverse CDF evaluated at the 32-bit uniform w: P(pick <= n) = thr_n 2^-32.
// ------------------------------------------------------------------------------
// warps per block of k_zstat: 8 up to 32 signatures, 4 beyond (shared memory per warp doubles)
template <int NP> struct ZWarps { static constexpr int value = NP <= 32 ? 8 : 4; };
// threshold columns are padded to a power of two (entries >= N-1 hold "never")
template <int NP> struct ZPad { static constexpr int value = NP <= 4 ? 4 : NP <= 8 ? 8 : NP <= 16 ? 16 : NP <= 32 ? 32 : 64; };
// rows of a threshold table that a search can touch: NP rounded up to a quarter of the padded width
__host__ __device__ constexpr int zthr_rows(int NP) {
  const int NPAD = NP <= 4 ? 4 : NP <= 8 ? 8 : NP <= 16 ? 16 : NP <= 32 ? 32 : 64;
  const int q = NPAD / 4;
  return ((NP + q - 1) / q) * q;
}

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
// 2^32 / total, IEEE division
__device__ __forceinline__ double pick_scale(double total) { return __ddiv_rn(4294967296.0, total); }
__device__ __forceinline__ float pick_scale(float total) { return __fdiv_rn(4294967296.0f, total); }
// saturating floor to u32 (NaN -> 0)
__device__ __forceinline__ uint32_t pick_thr(double x) { return __double2uint_rd(x); }
__device__ __forceinline__ uint32_t pick_thr(float x) { return __float2uint_rd(x); }

// The ten round keys of Philox4x32-10 for the run seed, computed once on the host: as a kernel
// parameter they are constant-bank operands of the round's XOR instead of twenty additions per
// block of random words.
struct ZKeys { uint32_t k0[10], k1[10]; };
__host__ inline ZKeys make_zkeys(uint64_t seed) {
  ZKeys r; uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) { r.k0[i] = a; r.k1[i] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  return r;
}
__device__ __forceinline__ U4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const ZKeys& rk) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ rk.k0[r];
    const uint32_t n2 = hi0 ^ c3 ^ rk.k1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// The pick loop is issue-bound.  Its shared-memory tables are addressed by 32-bit shared-window
// addresses (LDS/STS with register + immediate addressing, no 64-bit pointer arithmetic), and a
// level of the search is a load, a compare and one predicated add.
extern __shared__ __align__(16) unsigned char smem_raw[];
template <typename V> __device__ __forceinline__ V* smem_ptr(uint32_t sa) {
  return reinterpret_cast<V*>(__cvta_shared_to_generic((size_t)sa));
}
template <int OFF> __device__ __forceinline__ uint32_t lds_u32_off(uint32_t sa) { return *smem_ptr<const uint32_t>(sa + OFF); }
// a += INC if v <= w
template <int INC> __device__ __forceinline__ void add_if_le(uint32_t& a, uint32_t v, uint32_t w) {
  asm("{ .reg .pred p; setp.le.u32 p, %1, %2; @p add.u32 %0, %0, %3; }" : "+r"(a) : "r"(v), "r"(w), "n"(INC));
}

// levels S, S/2, .. 1 of four concurrent binary searches (S = remaining half-width in entries)
template <int S, bool ALL4>
__device__ __forceinline__ void search_levels(uint32_t (&a)[4], const uint32_t (&ww)[4], int lim) {
  if constexpr (S >= 1) {
    uint32_t v[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) if (ALL4 || p < lim) v[p] = lds_u32_off<(S - 1) * 128>(a[p]);
#pragma unroll
    for (int p = 0; p < 4; ++p) if (ALL4 || p < lim) add_if_le<S * 128>(a[p], v[p], ww[p]);
    search_levels<S / 2, ALL4>(a, ww, lim);
  }
}

// One quad: four words of one Philox block, four searches, four counter increments.
// A search returns the address of thr[#{ n : thr_n <= w }][cell] in the cell's
// non-decreasing threshold column (row stride 128 bytes, last entry "never"): branch-free binary
// search, the pivots of its first two levels in registers (loaded once per cell), the four
// searches advanced level by level so that their loads are in flight together.  `hdb` = byte
// offset from a threshold entry to the counter of the same (n, cell).  `lim` = picks of the cell
// still to draw (>= 1); with ALL4 the four searches are straight-line code whatever `lim`.
struct ZPivots { uint32_t mid, lo, hi, q0, q1, q2, q3; };   // entries NPAD/2-1 | NPAD/4-1, 3NPAD/4-1 | (2j+1)NPAD/8-1
template <int NPAD, bool ALL4>
__device__ __forceinline__ void zstat_quad(const U4& w, uint32_t col, const ZPivots& pv, int lim, uint32_t hdb) {
  uint32_t ww[4] = {w.x, w.y, w.z, w.w};
  uint32_t a[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    ww[p] = min(ww[p], 0xfffffffeu);
    const bool up = pv.mid <= ww[p];
    a[p] = up ? col + (NPAD / 2) * 128 : col;
    const bool up2 = (up ? pv.hi : pv.lo) <= ww[p];
    if (up2) a[p] += (NPAD / 4) * 128;
    if (NPAD >= 8) {
      const uint32_t lo3 = up ? pv.q2 : pv.q0, hi3 = up ? pv.q3 : pv.q1;
      add_if_le<(NPAD / 8) * 128>(a[p], up2 ? hi3 : lo3, ww[p]);
    }
  }
  search_levels<NPAD / 16, ALL4>(a, ww, lim);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p == 0 || p < lim) {
      int* h = smem_ptr<int>(a[p] + hdb);
      *h = *h + 1;
    }
  }
}

// Fix-up pass of a dense row (see k_zstat): the counts a lane drew for a cell that began in an earlier lane's share
// (cont[.][lane]) are summed over the consecutive lanes that continue the same cell -- a segmented shuffle
// reduction of LEVELS doubling steps, `run` = lanes behind this one on the same cell -- and added to the cell's
// histogram column by the first of them.
template <int NP, int LEVELS>
__device__ __forceinline__ void zstat_fixup(int* cont, int* hist, int lane, bool F, bool head, int run, int c_first) {
#pragma unroll
  for (int n = 0; n < NP; ++n) {
    int v = cont[n * 32 + lane];
    if (F) cont[n * 32 + lane] = 0;
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) {
      const int t = __shfl_down_sync(0xffffffffu, v, 1 << l);
      if ((1 << l) <= run) v += t;
    }
    if (head && v) hist[n * 32 + c_first] += v;
  }
}

// shared memory of k_zstat in bytes (host and device agree on it): per warp
// { E tile (T; a genome's column padded to NP + 1), thr, hist, cont (int), cell descriptors, the row of P (T) }; nothing is
// shared between the warps
template <typename T> __host__ __device__ inline size_t zstat_warp_bytes(int NP) {
  return (size_t)32 * (zthr_rows(NP) + 2 * NP) * sizeof(int) + (size_t)32 * (NP + 1) * sizeof(T) + 136 * sizeof(int) + (size_t)NP * sizeof(T);
}
template <typename T> __host__ __device__ inline size_t zstat_smem_bytes(int NP, int W) {
  return (size_t)W * zstat_warp_bytes<T>(NP);
}

// A warp is on its own from start to end: it takes work items -- ZR mutation types x 32 genomes -- from ONE global
// queue (the ticket of the next item is drawn while the current one is processed), stages the row of P it is
// working on in 160 bytes of its own shared memory, and adds every finished row of SP straight to global memory
// (integer atomics, one 64-bit RED per signature and row: 6 M per launch at C3, spread over K N addresses).  No
// block-level barrier, no K tiling: the fixed cost of a launch is one item's tail instead of one per K tile
// (0.084 -> 0.03 ms; it is what a 12,500-genome shard of an 8-GPU run pays 8 times as dearly).  Mutation types
// are visited in the order `korder` (descending total count, fixed at bnmf_create), so that the last items of a
// column tile -- the tail of the launch -- are its lightest.
// counts of genome g (local), mutation type k: the genome-major copy (a row of 32 genomes is one line)
#ifdef BNMF_Z_NO_MT
#define ZM(g, k) ZLD(&d.Mi[(long long)(k) + (long long)K * (g)])
#else
#define ZM(g, k) ZLD(&d.Mt[(long long)(g) + (long long)G * (k)])
#endif
// streamed operands (an element of E or M is read once per item): with 2 x 112 KB of shared memory the SM's L1 is
// a few KB -- it is left to the rows of P, which every warp rereads
#ifdef BNMF_Z_LDCG
#define ZLD(p) __ldcg(p)
#else
#define ZLD(p) (*(p))
#endif
template <typename T, int NP, int MODE /* 0 dense, 1 sparse rows drawn lane-per-cell, 2 dense with several warps per item */>
#ifndef ZV_REGS
#define ZV_REGS 96
#endif
// (96 registers: two of its blocks and a 256-thread block of the side stream's hyper-draw kernels share an SM's
//  register file -- the overlap of section 4.2 of DESIGN.md needs all three resident)
__global__ void __launch_bounds__(32 * ZWarps<NP>::value) __maxnreg__(NP <= 32 ? ZV_REGS : 168)
k_zstat(Dev<T> d, const __grid_constant__ ZKeys rk, int ZR_A, int ZR_B, int ctA, int lgS_arg /* log2 of the warps per item (MODE 2) */, int* work_ctr) {
  constexpr bool SPARSE = MODE == 1;
  const int lgS = MODE == 2 ? lgS_arg : 0;
  constexpr int NPAD = ZPad<NP>::value;
  constexpr int TR = zthr_rows(NP);
  const int K = d.K, N = d.N, G = d.G;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned char* wtabs = smem_raw;
  // E tile of the item as it lies in global memory -- genome by genome --, a genome's column padded to the odd length
  // ES: the tile is copied with coalesced loads (it is N * 32 contiguous elements; a load per lane and signature
  // touched 32 sectors and held up the shared-memory traffic of every other warp of the SM behind it) and lane
  // c reads its column Esm[c * ES + n] conflict-free
  constexpr int ES = NP + 1;
  T* Esm = reinterpret_cast<T*>(wtabs + (size_t)wid * zstat_warp_bytes<T>(NP));   // [32 cells][ES]
  uint32_t* thr = reinterpret_cast<uint32_t*>(Esm + ES * 32);       // [TR][32 cells]  pick thresholds of the row
  int* hist = reinterpret_cast<int*>(thr) + TR * 32;                // [NP][32 cells]  counts of the item so far (its SE)
  int* cont = hist + NP * 32;                                       // [NP][32 lanes]  counts of a share's spilled first cell
  // cell descriptors of the current row: {first quad, picks, Philox counter words 0,1}; entry 32 = {all quads}
  int4* desc = reinterpret_cast<int4*>(cont + NP * 32);
  T* Prow = reinterpret_cast<T*>(reinterpret_cast<int*>(desc) + 136);   // [NP]  row k of P with A folded in
  // shared-window addresses of the two tables; opaque so that they stay in registers instead of
  // being recomputed inside the (rarely taken, hence sunk-into) cell-switch branch of the pick loop
  uint32_t thr_sa = (uint32_t)__cvta_generic_to_shared(thr), desc_sa = (uint32_t)__cvta_generic_to_shared(desc);
  asm volatile("mov.u32 %0, %0;" : "+r"(thr_sa));
  asm volatile("mov.u32 %0, %0;" : "+r"(desc_sa));
  const uint32_t c3 = ((uint32_t)d.ctrl->iter << 8) | (uint32_t)PUR_Z;
#pragma unroll
  for (int n = 0; n < ES; ++n) Esm[n * 32 + lane] = (T)0;          // signatures >= N of a column stay zero
  const uint32_t ediv = ((1u << 20) + (unsigned)N - 1u) / (unsigned)N;   // i / N = (i * ediv) >> 20 for i < 2048
#pragma unroll
  for (int n = 0; n < TR; ++n) thr[n * 32 + lane] = 0xffffffffu;   // entries >= N-1 stay "never"
#pragma unroll
  for (int n = 0; n < NP; ++n) { hist[n * 32 + lane] = 0; cont[n * 32 + lane] = 0; }
  // signatures this lane stages of a row of P: lane, lane + 32 (excluded ones -- A_n = 0 -- as zero)
  const bool use0 = lane < N && d.A[lane] != 0, use1 = NP > 32 && lane + 32 < N && d.A[lane + 32] != 0;

  // column tiles [0, ctA) are cut into chunks of ZR_A mutation types, the last ones into finer chunks of ZR_B:
  // the items handed out at the end of a launch are small, so is the time the last warp works alone
  const int rtsA = (K + ZR_A - 1) / ZR_A, rtsB = (K + ZR_B - 1) / ZR_B;
  const int cts = (G + 31) / 32;                     // column tiles
  const long long nA = (long long)rtsA * ctA;
  const long long n_items = nA + (long long)rtsB * (cts - ctA);
  {
    // the first item of a warp is its own number (no 2,368 tickets drawn from one counter at once); later ones
    // come from the queue, which starts behind them
    const int n_warps = (int)gridDim.x * ZWarps<NP>::value;
    int next = (int)blockIdx.x * ZWarps<NP>::value + wid;
  for (;;) {
    // Small problems (fewer items than resident warps; the heaviest row-tile alone is 40 us of one warp at 96 x 100):
    // S warps share an item -- each forms the row's thresholds itself and draws the S-th part of its quads into
    // its own tables; the margins are integer atomics, so the result does not change.  Split 0 writes the metric
    // partials of the item.
    const int unit = next;
    if (unit >= (n_items << lgS)) break;
    if (lane == 0) next = n_warps + atomicAdd(work_ctr, 1);      // the next ticket travels while this item is processed
    const int item = unit >> lgS, split = unit - (item << lgS);
    // items in chunk-major order: chunk 0 (it holds the heaviest mutation types) of every column tile first, the
    // lightest chunk last -- longest processing time first, the warps that finish last are on the cheapest items
    // (a single row of the heaviest type over 32 genomes is ~8 rows' worth of picks: taken late it IS the tail)
    const bool fine = item >= nA;
    const int ZR = fine ? ZR_B : ZR_A, rts = fine ? rtsB : rtsA;
    const int ctn = fine ? cts - ctA : ctA;
    const int j = fine ? (int)(item - nA) : item;
    const int rt = j / ctn;
    const int ct = (fine ? ctA : 0) + (j - rt * ctn);
    const int g = ct * 32 + lane;
    const bool valid = g < G;
    const unsigned long long cell0 = (unsigned long long)K * (unsigned long long)(d.g0 + (long long)ct * 32);

    // E tile of the item: N * 32 contiguous elements, copied as they lie (coalesced), a genome's column padded to ES
    {
      const T* __restrict__ Eg = d.E + (long long)N * ((long long)ct * 32);
      const int n_el = N * min(32, G - ct * 32);              // elements of the tile that exist
#ifndef BNMF_Z_EUNROLL
#define BNMF_Z_EUNROLL 4
#endif
      constexpr int EU = BNMF_Z_EUNROLL;     // loads in flight per lane (all NP at once was measured: 6 % slower overall)
#pragma unroll EU
      for (int jj = 0; jj < NP; ++jj) {
        const int i = jj * 32 + lane;
        if (i < N * 32) {
          const uint32_t c = ((uint32_t)i * ediv) >> 20;      // genome of element i; its signature is i - c N
          Esm[i + (int)c * (ES - N)] = i < n_el ? ZLD(Eg + i) : (T)0;
        }
      }
    }
    // chunk rt = the mutation types of rank rt, rt + rts, rt + 2 rts, ... in `korder` (descending total count):
    // every chunk holds a like share of heavy and light types, so items weigh alike; heaviest row first
    int k = d.korder[rt];
    T p0 = use0 ? d.P[(long long)k + (long long)K * lane] : (T)0, p1 = (T)0;
    if (NP > 32) p1 = use1 ? d.P[(long long)k + (long long)K * (lane + 32)] : (T)0;
#ifndef BNMF_Z_NO_MNEXT
    int m_next = valid ? ZM(g, k) : 0;
#endif
    double a_sse = 0.0, a_kl = 0.0, a_ll = 0.0;
    int sp_prev0 = 0, sp_prev1 = 0;     // lane n: sum over the tile's cells of hist[n][.] after the previous row

    for (int r = 0; r < ZR && r * rts + rt < K; ++r) {
#ifdef BNMF_Z_NO_MNEXT
      const int m = valid ? ZM(g, k) : 0;
#else
      const int m = m_next;
#endif
      const int k_this = k;
      {   // row k of P into the warp's shared memory (read by every lane in phase 1); the next row's (and the next
          // row's counts) are on their way
        if (lane < NP) Prow[lane] = p0;
        if (NP > 32) { if (lane + 32 < NP) Prow[lane + 32] = p1; }
        __syncwarp();
        if ((r + 1) < ZR && (r + 1) * rts + rt < K) {
          k = d.korder[(r + 1) * rts + rt];
          p0 = use0 ? d.P[(long long)k + (long long)K * lane] : (T)0;
          if (NP > 32) p1 = use1 ? d.P[(long long)k + (long long)K * (lane + 32)] : (T)0;
#ifndef BNMF_Z_NO_MNEXT
          m_next = valid ? ZM(g, k) : 0;
#endif
        }
      }
      // ---- phase 1: total of this lane's cell, metric partials ----
      T total = (T)0;
#pragma unroll
      for (int n = 0; n < NP; ++n) total = add_rn<T>(total, mul_rn<T>(Prow[n], Esm[lane * ES + n]));
      if (valid && split == 0) {
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
      if (!__any_sync(0xffffffffu, work)) continue;
      // integer pick thresholds of this lane's cell (second pass over the products: reloading the
      // two tiles is cheaper than 2 NP live registers, which would serialise the pick loop below)
      {
        asm volatile("" ::: "memory");
        const T scale = pick_scale(total);
        T acc = (T)0;
#pragma unroll
        for (int n = 0; n < NP - 1; ++n) {
          acc = add_rn<T>(acc, mul_rn<T>(Prow[n], Esm[lane * ES + n]));
          thr[n * 32 + lane] = n < N - 1 ? pick_thr(mul_rn<T>(acc, scale)) : 0xffffffffu;
        }
      }
      // quads of this lane's cell
      const int q = work ? (int)(((unsigned)m + 3u) >> 2) : 0;
      // Sparse rows (exome-like counts: a quad or two per cell): sharing out the row's quads costs more -- scan,
      // cell table, spilled-cell fix-up: ~300 instructions -- than the idle lanes it saves.  Every lane then draws
      // the picks of its own cell: its own threshold column, its own histogram column, nothing to fix up.
      // (SPARSE: a kernel variant of its own, chosen at bnmf_create for data with few counts per cell -- the extra
      //  branch costs the dense kernel registers it needs to share an SM with the side stream's kernels)
      if (SPARSE && 32 * __reduce_max_sync(0xffffffffu, q) <= __reduce_add_sync(0xffffffffu, q) + 80) {
        if (q > 0) {
          const unsigned long long cell = cell0 + (unsigned long long)k_this + (unsigned long long)((unsigned)K * (unsigned)lane);
          const uint32_t col = thr_sa + 4u * (unsigned)lane;
          ZPivots pv;
          pv.mid = lds_u32_off<(NPAD / 2 - 1) * 128>(col);
          pv.lo = lds_u32_off<(NPAD / 4 - 1) * 128>(col); pv.hi = lds_u32_off<(3 * NPAD / 4 - 1) * 128>(col);
          pv.q0 = pv.q1 = pv.q2 = pv.q3 = 0xffffffffu;
          if (NPAD >= 8) {
            pv.q0 = lds_u32_off<(NPAD / 8 - 1) * 128>(col);
            pv.q1 = lds_u32_off<(3 * NPAD / 8 - 1) * 128>(col);
            pv.q2 = (5 * NPAD / 8 - 1 < TR) ? lds_u32_off<(5 * NPAD / 8 - 1) * 128>(col) : 0xffffffffu;
            pv.q3 = (7 * NPAD / 8 - 1 < TR) ? lds_u32_off<(7 * NPAD / 8 - 1) * 128>(col) : 0xffffffffu;
          }
          for (int sub = 0; sub < q; ++sub) {
            const U4 w = philox_rk((uint32_t)cell, (uint32_t)(cell >> 32), (uint32_t)sub, c3, rk);
            zstat_quad<NPAD, false>(w, col, pv, m - 4 * sub, 4u * TR * 32);
          }
        }
        __syncwarp();
      } else {
      // ---- dense rows: the quads of the 32 cells laid end to end, an equal share per lane ----
      int incl = q;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int excl = incl - q;
      const int Tq = __shfl_sync(0xffffffffu, incl, 31);
      {
        const unsigned long long cell = cell0 + (unsigned long long)k_this + (unsigned long long)((unsigned)K * (unsigned)lane);
        desc[lane] = make_int4(excl, m, (int)(uint32_t)cell, (int)(uint32_t)(cell >> 32));
        if (lane == 31) desc[32].x = Tq;
      }
      __syncwarp();   // threshold columns and the cell table visible to the whole warp
      // ---- phase 2: an equal contiguous share [qd, qhi) of the row's quads per lane ----
      // this warp's part of the row (a row holds fewer than 2^27 quads: cells <= 2^24 counts)
      const int q_lo = (int)(((unsigned long long)(unsigned)Tq * (unsigned)split) >> lgS);
      const int q_n = (int)(((unsigned long long)(unsigned)Tq * (unsigned)(split + 1)) >> lgS) - q_lo;
      int qd = q_lo + (int)(((unsigned long long)lane * (unsigned)q_n) >> 5);
      const int qhi = q_lo + (int)(((unsigned long long)(lane + 1) * (unsigned)q_n) >> 5);
      // owner of the first quad: the cell c with excl[c] <= qd < excl[c] + q[c]
      int c_first = 0, ec_first = 0;
#pragma unroll
      for (int st = 16; st > 0; st >>= 1) {
        const int e = __shfl_sync(0xffffffffu, excl, c_first + st);
        if (e <= qd) { c_first += st; ec_first = e; }
      }
      const bool F = ec_first < qd;                  // the first cell started in an earlier lane's share
      {
        const bool dense = q_n >= 96;                // rows of mostly full quads: branch-free searches
        uint32_t da = desc_sa + 16u * (unsigned)(c_first - 1);
        int ec = 0, nxt = qd, mc = 0;                // qd >= nxt: the first pass loads the first cell
        bool spill = F;                              // first cell of a spilled share -> cont[.][lane]
        uint32_t col = thr_sa, hdb = 0;
        ZPivots pv = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        uint32_t cc0 = 0, cc1 = 0;
        for (; qd < qhi; ++qd) {
          if (qd >= nxt) {          // next cell that has picks (the first one may have begun earlier)
            do {
              da += 16u;
              const int4 dd = *smem_ptr<const int4>(da);
              ec = dd.x; mc = dd.y; cc0 = (uint32_t)dd.z; cc1 = (uint32_t)dd.w;
              nxt = (int)lds_u32_off<16>(da);
            } while (nxt == ec);
            const uint32_t c4 = (da - desc_sa) >> 2;           // 4 * cell
            col = thr_sa + c4;
            // pivots of the first three levels of the search (rows >= TR do not exist: "never")
            pv.mid = lds_u32_off<(NPAD / 2 - 1) * 128>(col);
            pv.lo = lds_u32_off<(NPAD / 4 - 1) * 128>(col); pv.hi = lds_u32_off<(3 * NPAD / 4 - 1) * 128>(col);
            if (NPAD >= 8) {
              pv.q0 = lds_u32_off<(NPAD / 8 - 1) * 128>(col);
              pv.q1 = lds_u32_off<(3 * NPAD / 8 - 1) * 128>(col);
              pv.q2 = (5 * NPAD / 8 - 1 < TR) ? lds_u32_off<(5 * NPAD / 8 - 1) * 128>(col) : 0xffffffffu;
              pv.q3 = (7 * NPAD / 8 - 1 < TR) ? lds_u32_off<(7 * NPAD / 8 - 1) * 128>(col) : 0xffffffffu;
            }
            hdb = spill ? 4u * (unsigned)((TR + NP) * 32 + lane) - c4 : 4u * TR * 32;   // counter column: cont[.][lane] | hist[.][c]
            spill = false;
          }
          const int sub = qd - ec;
          const int lim = mc - 4 * sub;              // picks in this quad: min(4, lim) >= 1
          const U4 w = philox_rk(cc0, cc1, (uint32_t)sub, c3, rk);
          if (dense) zstat_quad<NPAD, true>(w, col, pv, lim, hdb);
          else       zstat_quad<NPAD, false>(w, col, pv, lim, hdb);
        }
      }
      // ---- fix-up: lanes whose share began inside a cell hand their counts to its column.  Lanes
      //      continuing the same cell are consecutive: their columns are first summed into the
      //      first of them by a segmented shuffle reduction, so one pass serves every cell ----
      {
        const unsigned fm = __ballot_sync(0xffffffffu, F);
        if (fm) {
          const int cprev = __shfl_up_sync(0xffffffffu, c_first, 1);
          const bool chain = F && lane > 0 && ((fm >> (lane - 1)) & 1u) && cprev == c_first;
          const unsigned bm = __ballot_sync(0xffffffffu, chain);
          const unsigned above = lane < 31 ? (bm >> (lane + 1)) : 0u;
          const int run = F ? __ffs((int)~above) - 1 : 0;          // lanes lane+1 .. lane+run continue my first cell
          const int maxrun = __reduce_max_sync(0xffffffffu, run);
          const bool head = F && !chain;
#ifdef BNMF_Z_OLD_FIXUP
#pragma unroll
          for (int n = 0; n < NP; ++n) {
            int v = cont[n * 32 + lane];
            if (F) cont[n * 32 + lane] = 0;
            for (int dd = 1; dd <= maxrun; dd <<= 1) {
              const int t = __shfl_down_sync(0xffffffffu, v, dd);
              if (dd <= run) v += t;
            }
            if (head && v) hist[n * 32 + c_first] += v;
          }
#else
          // straight-line code per number of doubling steps the longest chain of the row needs
          if (maxrun == 0) zstat_fixup<NP, 0>(cont, hist, lane, F, head, run, c_first);
          else if (maxrun == 1) zstat_fixup<NP, 1>(cont, hist, lane, F, head, run, c_first);
          else if (maxrun < 4) zstat_fixup<NP, 2>(cont, hist, lane, F, head, run, c_first);
          else if (maxrun < 8) zstat_fixup<NP, 3>(cont, hist, lane, F, head, run, c_first);
          else zstat_fixup<NP, 5>(cont, hist, lane, F, head, run, c_first);
#endif
          __syncwarp();
        }
      }
      }   // (dense rows)
      // the histogram keeps counting through the rows of the item (its columns are SE); the sum of
      // a row of it over the tile's cells, minus the same sum one row earlier, is this row's SP
      int mytot0 = 0, mytot1 = 0;
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const int tot = __reduce_add_sync(0xffffffffu, hist[n * 32 + lane]);
        if (n < 32) { if (lane == n) mytot0 = tot; }
        else        { if (lane == n - 32) mytot1 = tot; }
      }
      if (lane < N && mytot0 != sp_prev0) atomicAdd(&d.SP[(long long)k_this + (long long)K * lane], (unsigned long long)(mytot0 - sp_prev0));
      if (NP > 32 && lane + 32 < N && mytot1 != sp_prev1) atomicAdd(&d.SP[(long long)k_this + (long long)K * (lane + 32)], (unsigned long long)(mytot1 - sp_prev1));
      sp_prev0 = mytot0; sp_prev1 = mytot1;
      __syncwarp();   // the threshold columns are rewritten by the next row
    }
    // SE leaves the SM once per item
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      const int v = hist[n * 32 + lane];
      hist[n * 32 + lane] = 0;
      if (valid && n < N && v) atomicAdd(&d.SE[(long long)n + (long long)N * g], v);
    }
    // per-item metric partials, fixed reduction order
    a_sse = warp_sum(a_sse); a_kl = warp_sum(a_kl); a_ll = warp_sum(a_ll);
    if (lane == 0 && split == 0) {
      double* zp = d.zpart + (long long)item * PC_COLS;
      zp[PC_SSE] = a_sse; zp[PC_KLV] = a_kl; zp[PC_LLV] = a_ll; zp[PC_LP_E] = 0.0; zp[PC_EACC] = 0.0;
    }
    next = __shfl_sync(0xffffffffu, next, 0);
  }
  }
}

// ------------------------------------------------------------------------------
// End of an iteration, ONE kernel:
//   1. fold the per-item / per-block partials in a fixed order:
//        red[c] = sum_items zpart[.][c] + sum_blocks epart[.][c]  (+ data constants)
//        lp_P   = sum_n zpart[n_zitems + n]   (slots written by k_pside / k_pprior)
//      RED_BLOCKS blocks each fold a fixed contiguous slice; the last one to finish (ticket) folds the
//      slices in index order, so the result does not depend on scheduling;
//   2. genome-sharded runs: that last block sums SP, rowSums(E) (integers: exact) and the metric partials
//      (doubles, in rank order: the same bits on every rank) over the GPUs of the node -- a one-shot
//      all-reduce over NVLink peer memory (below) instead of two NCCL launches;
//   3. it composes the sample_metrics row (R/utils.R:339-348, :412-455), records A into the ring and
//      advances it (record_sample, R/bayesNMF_sampler.R:651-672).
//
// One-shot all-reduce ("push, flag, sum locally"): every rank owns an exchange buffer that its peers have mapped
// (CUDA IPC over NVLink / NVSwitch, bnmf_comm_init); it holds, per slot, one region and one flag per source rank.
// Exchange number `seq` uses slot seq % XCHG_SLOTS:
//   push   the ~16 KB payload is stored into region (slot, own rank) of EVERY rank's buffer (remote stores are
//          fire-and-forget: no round trip), then, after a block barrier, thread q publishes to rank q:
//          flag(slot, own rank) = seq with st.release.sys (the release is cumulative over the barrier);
//   wait   thread r spins on the LOCAL flag (slot, r) >= seq (ld.acquire.sys);
//   sum    every thread sums its elements over the world regions of its own buffer, in rank order: integers
//          exactly, doubles to the same bits on every rank (L1 is bypassed: the lines were written by peers).
// A rank is at most one exchange ahead of any peer (it needs the peer's flag to finish), so a region is rewritten
// only after its owner has summed it.  Cost: one NVLink store latency + one fence; no remote loads.
// ------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 64;
constexpr int XCHG_SLOTS = 4;
constexpr int XCHG_MAX_WORLD = 8;
struct Xchg {
  int world, rank;                              // world <= 1: no exchange
  unsigned long long seq;                       // number of this exchange (1, 2, ...), the same on every rank
  unsigned long long* buf[XCHG_MAX_WORLD];      // rank r's exchange buffer as mapped into this process
  int slot_words;                               // 8-byte words per (slot, source) region; the flags follow the regions
  int* err;                                     // set when a peer's flag does not arrive (bnmf_step reports it)
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

template <typename T> __device__ void metrics_row(const Dev<T>& d);

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_reduce_partials(Dev<T> d, double* slices /*[RED_BLOCKS][PC_COLS]*/, unsigned* ticket, const Xchg x) {
  __shared__ double scratch[THREADS / 32];
  __shared__ bool last;
  const long long total = (long long)d.n_zitems + d.n_eblocks;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = lo + per < total ? lo + per : total;
  {   // one pass over the slice for all PC_COLS columns (the sums and their order are those of a pass per column)
    __shared__ double wsum[THREADS / 32][PC_COLS];
    double v[PC_COLS];
#pragma unroll
    for (int c = 0; c < PC_COLS; ++c) v[c] = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += THREADS) {
      const double* row = i < d.n_zitems ? d.zpart + i * PC_COLS : d.epart + (i - d.n_zitems) * PC_COLS;
#pragma unroll
      for (int c = 0; c < PC_COLS; ++c) v[c] += row[c];
    }
#pragma unroll
    for (int c = 0; c < PC_COLS; ++c) v[c] = warp_sum(v[c]);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int c = 0; c < PC_COLS; ++c) wsum[threadIdx.x >> 5][c] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < PC_COLS) {
      double r = 0.0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) r += wsum[w][threadIdx.x];
      slices[blockIdx.x * PC_COLS + threadIdx.x] = r;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < PC_COLS) {
    const int c = threadIdx.x;
    double s = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) s += ((volatile double*)slices)[b * PC_COLS + c];
    if (c == PC_LLV) s += d.ll_const;
    if (c == PC_KLV) s += d.kl_const;
    d.red[c] = s;
  }
  if (threadIdx.x == 32) {
    double s = 0.0;
    for (int i = 0; i < d.N; ++i) s += d.zpart[(long long)d.n_zitems * PC_COLS + i];
    *d.lp_P = s;
    *ticket = 0u;
  }
  __syncthreads();
  if (x.world > 1) {
    // ---- one-shot all-reduce of [SP | rowSums(E)] (K N + N integers) and red (PC_COLS doubles) ----
    const int n_i64 = d.K * d.N + d.N, nw = n_i64 + PC_COLS;
    const int slot = (int)(x.seq % XCHG_SLOTS);
    const size_t region = ((size_t)slot * x.world + x.rank) * x.slot_words;            // where this rank's payload goes, in every buffer
    const size_t flags = (size_t)XCHG_SLOTS * x.world * x.slot_words + (size_t)slot * x.world;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(d.SP);     // SP and rowsumE_fx are one array
    for (int i = threadIdx.x; i < nw; i += THREADS) {
      const unsigned long long v = i < n_i64 ? src[i] : (unsigned long long)__double_as_longlong(d.red[i - n_i64]);
#pragma unroll
      for (int q = 0; q < XCHG_MAX_WORLD; ++q) if (q < x.world) x.buf[q][region + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < x.world) st_release_sys(x.buf[threadIdx.x] + flags + x.rank, x.seq);
    if (threadIdx.x < x.world) {
      const unsigned long long* f = x.buf[x.rank] + flags + threadIdx.x;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) < x.seq) {
        __nanosleep(32);
        if (clock64() - t0 > 8000000000LL) { atomicExch(x.err, 1 + (int)threadIdx.x); break; }    // ~4 s: a peer is gone
      }
    }
    __syncthreads();
    const unsigned long long* mine = x.buf[x.rank] + (size_t)slot * x.world * x.slot_words;
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(d.SP);
    for (int i = threadIdx.x; i < nw; i += THREADS) {
      unsigned long long v[XCHG_MAX_WORLD];
#pragma unroll
      for (int r = 0; r < XCHG_MAX_WORLD; ++r) v[r] = r < x.world ? __ldcg(mine + (size_t)r * x.slot_words + i) : 0ull;
      if (i < n_i64) {
        unsigned long long s = 0ull;
#pragma unroll
        for (int r = 0; r < XCHG_MAX_WORLD; ++r) s += v[r];
        dst[i] = s;
      } else {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) s += __longlong_as_double((long long)v[r]);
        d.red[i - n_i64] = s;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && x.world >= 0) metrics_row<T>(d);
}

// ------------------------------------------------------------------------------
// metrics_row: compose the sample_metrics row (R/utils.R:339-348, :412-455), record A
// into the ring and advance it (record_sample, R/bayesNMF_sampler.R:651-672).  One thread.
// (k_metrics: the same as a kernel of its own, after an NCCL exchange.)
// ------------------------------------------------------------------------------
template <typename T>
__device__ void metrics_row(const Dev<T>& d) {
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
template <typename T>
__global__ void k_metrics(Dev<T> d) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  metrics_row<T>(d);
}

}  // namespace bnmf
