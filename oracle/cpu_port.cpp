// TEST / BASELINE INFRASTRUCTURE ONLY -- never loaded by the product (bayesnmf_b200/).
//
// CPU port of the Poisson latent-count Gibbs iteration (fixed rank, gamma or exponential prior)
// in C++ with OpenMP: the `cpu_baseline` / `--impl reference` arm of bench.py ("kind": "port-c++")
// and a second, independent executor of the oracle's arithmetic (tests/test_cpu_port.py compares
// it with oracle/gibbs.py: latent-count margins bit for bit, draws to 1e-12).
//
// It restates, function by function, what oracle/gibbs.py restates of the reference:
//   init_prior_params_        R/sample_priors.R:15-141
//   sample_prior_params_      R/sample_priors.R:150-200, :284-397
//   sample_Pn_poisson         R/sample_Pn.R:98-120        (and the prior draw :12-30)
//   sample_En_poisson         R/sample_En.R:97-119
//   sample_Zkg                R/sample_params.R:253-265   (M[k,g] categorical picks by inverse CDF)
//   compute_metrics_          R/utils.R:412-455, get_loglik_ :62-112, get_logpost_ :123-183,
//                             padded_KL_ :467-471
// The random variates are the shared-source __host__ __device__ draw code of the kernels
// (bayesnmf_b200/csrc/bnmf_rng.cuh: Philox4x32-10, Marsaglia-Tsang gamma, the exact rejection
// sampler of Alpha), compiled here by g++ -- SURVEY.md section 8(d)(ii).  R itself is not installed
// in this image, so this is a port, not the reference.
//
// Build: g++ -O3 -fopenmp -ffp-contract=off -shared -fPIC (oracle/cpu_port.py).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../bayesnmf_b200/csrc/bnmf_rng.cuh"

using namespace bnmf;

namespace {

enum { CP_EXPONENTIAL = 1, CP_GAMMA = 2 };
const double FX = 16777216.0;   // 2^24: rowSums(E) is an exact fixed-point sum (order independent)

struct Port {
  int K, N; long long G, G_total, g0; int prior; uint64_t seed; int iter;
  std::vector<int32_t> M;                        // K x G, column-major
  std::vector<double> P, E;                      // K x N, N x G
  std::vector<double> Alpha_p, Beta_p, Lambda_p, Alpha_e, Beta_e, Lambda_e;
  std::vector<long long> SP; std::vector<int32_t> SE;
  double A_p, B_p, C_p, D_p, A_e, B_e, C_e, D_e;
  double ll_const, kl_const;
  double row[9];
};

inline double dgamma_log(double x, double shape, double rate) { return shape * log(rate) - lgamma(shape) + (shape - 1.0) * log(x) - rate * x; }
inline double dexp_log(double x, double rate) { return log(rate) - rate * x; }

// sample_Zkg for every cell + the likelihood metrics that fall out of sum_n p_n = Mhat[k,g]
void latent_counts(Port& s, double* sse_out, double* klv_out, double* llv_out) {
  const int K = s.K, N = s.N; const long long G = s.G;
  std::fill(s.SP.begin(), s.SP.end(), 0LL);
  double sse = 0.0, klv = 0.0, llv = 0.0;
#pragma omp parallel reduction(+ : sse, klv, llv)
  {
    std::vector<long long> sp((size_t)K * N, 0LL);
    std::vector<uint32_t> thr((size_t)N);
    std::vector<double> cdf((size_t)N);
#pragma omp for schedule(dynamic, 64)
    for (long long g = 0; g < G; ++g) {
      const double* e = &s.E[(size_t)N * g];
      int32_t* se = &s.SE[(size_t)N * g];
      for (int n = 0; n < N; ++n) se[n] = 0;
      for (int k = 0; k < K; ++k) {
        const int m = s.M[(size_t)k + (size_t)K * g];
        double acc = 0.0;
        for (int n = 0; n < N; ++n) { acc = acc + s.P[(size_t)k + (size_t)K * n] * e[n]; cdf[n] = acc; }   // A = 1 (fixed rank)
        const double total = acc;
        const double lam = total > 1e-6 ? total : 1e-6, L = log(lam), md = (double)m;
        llv += md * L - lam;
        klv -= (m > 0 ? md : 1e-6) * L;
        sse += (total - md) * (total - md);
        if (!(m > 0 && total > 0.0)) continue;
        const double scale = 4294967296.0 / total;
        for (int n = 0; n < N - 1; ++n) {
          const double x = floor(cdf[n] * scale);
          thr[n] = !(x == x) ? 0u : x >= 4294967295.0 ? 4294967295u : x <= 0.0 ? 0u : (uint32_t)x;
        }
        const Stream st = make_stream(s.seed, (uint32_t)s.iter, PUR_Z, (uint64_t)k + (uint64_t)K * (uint64_t)(g + s.g0));
        for (int j = 0; j < m; j += 4) {
          const U4 w4 = st.at((uint32_t)(j >> 2));
          const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
          const int lim = m - j < 4 ? m - j : 4;
          for (int q = 0; q < lim; ++q) {
            const uint32_t w = ww[q] < 0xfffffffeu ? ww[q] : 0xfffffffeu;
            int lo = 0, hi = N - 1;                      // pick = #{ n < N-1 : thr_n <= w } (thr is non-decreasing)
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (thr[mid] <= w) lo = mid + 1; else hi = mid; }
            se[lo] += 1; sp[(size_t)k + (size_t)K * lo] += 1;
          }
        }
      }
    }
#pragma omp critical
    for (size_t i = 0; i < sp.size(); ++i) s.SP[i] += sp[i];
  }
  *sse_out = sse; *klv_out = klv; *llv_out = llv;
}

// the P side of one iteration: prior parameters, then P[k,n] | Z, E
double p_side(Port& s, bool from_prior) {
  const int K = s.K, N = s.N;
  std::vector<long long> rs((size_t)N, 0LL);
  for (int n = 0; n < N; ++n) {
    long long a = 0;
#pragma omp parallel for reduction(+ : a)
    for (long long g = 0; g < s.G; ++g) a += llrint(s.E[(size_t)n + (size_t)N * g] * FX);
    rs[n] = a;
  }
  double lp = 0.0;
#pragma omp parallel for reduction(+ : lp) schedule(dynamic, 16)
  for (long long c = 0; c < (long long)K * N; ++c) {
    const int n = (int)(c / K);
    const double Pold = s.P[c];
    const double rsE = (double)rs[n] / FX;
    double Pnew;
    if (s.prior == CP_GAMMA) {
      double al = s.Alpha_p[c], be = s.Beta_p[c];
      if (!from_prior) {
        be = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_HYP_P1, c), s.A_p + al, s.B_p + Pold);
        al = alpha_draw(make_stream(s.seed, s.iter, PUR_HYP_P2, c), s.C_p, s.D_p, be, Pold, al);
        s.Beta_p[c] = be; s.Alpha_p[c] = al;
      }
      double shape = al, rate = be;
      if (!from_prior) { shape += (double)s.SP[c]; rate += rsE; }
      Pnew = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_P, c), shape, rate);
      lp += dgamma_log(Pnew, al, be);
    } else {
      double la = s.Lambda_p[c];
      if (!from_prior) { la = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_HYP_P1, c), s.A_p + 1.0, s.B_p + Pold); s.Lambda_p[c] = la; }
      double shape = 1.0, rate = la;
      if (!from_prior) { shape += (double)s.SP[c]; rate += rsE; }
      Pnew = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_P, c), shape, rate);
      lp += dexp_log(Pnew, la);
    }
    s.P[c] = Pnew;
  }
  return lp;
}

double e_side(Port& s, bool from_prior) {
  const int K = s.K, N = s.N;
  std::vector<double> cs((size_t)N, 0.0);
  for (int n = 0; n < N; ++n) { double a = 0.0; for (int k = 0; k < K; ++k) a += s.P[(size_t)k + (size_t)K * n]; cs[n] = a; }
  double lp = 0.0;
#pragma omp parallel for reduction(+ : lp) schedule(dynamic, 256)
  for (long long ii = 0; ii < (long long)N * s.G; ++ii) {
    const int n = (int)(ii % N);
    const uint64_t c = (uint64_t)n + (uint64_t)N * (uint64_t)(s.g0 + ii / N);
    const double Eold = s.E[ii];
    double Enew;
    if (s.prior == CP_GAMMA) {
      double al = s.Alpha_e[ii], be = s.Beta_e[ii];
      if (!from_prior) {
        be = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_HYP_E1, c), s.A_e + al, s.B_e + Eold);
        al = alpha_draw(make_stream(s.seed, s.iter, PUR_HYP_E2, c), s.C_e, s.D_e, be, Eold, al);
        s.Beta_e[ii] = be; s.Alpha_e[ii] = al;
      }
      double shape = al, rate = be;
      if (!from_prior) { shape += (double)s.SE[ii]; rate += cs[n]; }
      Enew = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_E, c), shape, rate);
      lp += dgamma_log(Enew, al, be);
    } else {
      double la = s.Lambda_e[ii];
      if (!from_prior) { la = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_HYP_E1, c), s.A_e + 1.0, s.B_e + Eold); s.Lambda_e[ii] = la; }
      double shape = 1.0, rate = la;
      if (!from_prior) { shape += (double)s.SE[ii]; rate += cs[n]; }
      Enew = gamma_draw<double>(make_stream(s.seed, s.iter, PUR_E, c), shape, rate);
      lp += dexp_log(Enew, la);
    }
    s.E[ii] = Enew;
  }
  return lp;
}

void iteration(Port& s, bool from_prior) {
  const double lpP = p_side(s, from_prior);
  const double lpE = e_side(s, from_prior);
  double sse, klv, llv;
  latent_counts(s, &sse, &klv, &llv);
  const double loglik = llv + s.ll_const;
  const double nparams = (double)s.N * ((double)s.G_total + (double)s.K);
  s.row[0] = (double)s.iter;
  s.row[1] = sqrt(sse / ((double)s.K * (double)s.G_total));
  s.row[2] = klv + s.kl_const;
  s.row[3] = loglik;
  s.row[4] = loglik + lpP + lpE;
  s.row[5] = nparams;
  s.row[6] = -2.0 * loglik + nparams * log((double)s.G_total);
  s.row[7] = (double)s.N;
  s.row[8] = 1.0;
}

}  // namespace

extern "C" {

// hyper = {A_p, B_p, C_p, D_p, A_e, B_e, C_e, D_e} (scalars, fill_hyperprior_params_ R/setup.R:15-88)
void* cp_create(int K, int N, long long G, long long G_total, long long g0, int prior, uint64_t seed,
                const double* data /*K x G column-major*/, const double* hyper) {
  if (prior != CP_GAMMA && prior != CP_EXPONENTIAL) return nullptr;
  Port* s = new Port();
  s->K = K; s->N = N; s->G = G; s->G_total = G_total; s->g0 = g0; s->prior = prior; s->seed = seed; s->iter = 1;
  const size_t KG = (size_t)K * G, KN = (size_t)K * N, NG = (size_t)N * G;
  s->M.resize(KG);
  double ll = 0.0, kl = 0.0;
#pragma omp parallel for reduction(+ : ll, kl)
  for (long long i = 0; i < (long long)KG; ++i) {
    const double m = data[i];
    s->M[i] = (int32_t)m;
    ll -= lgamma(m + 1.0);
    const double mp = m > 1e-6 ? m : 1e-6;
    kl += mp * log(mp);
  }
  s->ll_const = ll; s->kl_const = kl;
  s->P.assign(KN, 0.0); s->E.assign(NG, 0.0);
  s->Alpha_p.assign(KN, 0.0); s->Beta_p.assign(KN, 0.0); s->Lambda_p.assign(KN, 0.0);
  s->Alpha_e.assign(NG, 0.0); s->Beta_e.assign(NG, 0.0); s->Lambda_e.assign(NG, 0.0);
  s->SP.assign(KN, 0); s->SE.assign(NG, 0);
  s->A_p = hyper[0]; s->B_p = hyper[1]; s->C_p = hyper[2]; s->D_p = hyper[3];
  s->A_e = hyper[4]; s->B_e = hyper[5]; s->C_e = hyper[6]; s->D_e = hyper[7];
  return s;
}
void cp_destroy(void* h) { delete static_cast<Port*>(h); }

// iteration 1 = the prior draw (R/bayesNMF_sampler.R:232-257): prior parameters from the hyperpriors
// (iteration-0 streams), P and E from their priors, Z from its full conditional, the metrics row
void cp_init(void* h, double* row9) {
  Port& s = *static_cast<Port*>(h);
  const long long KN = (long long)s.K * s.N, NG = (long long)s.N * s.G;
#pragma omp parallel for
  for (long long c = 0; c < KN; ++c) {
    if (s.prior == CP_GAMMA) {
      s.Beta_p[c] = gamma_draw<double>(make_stream(s.seed, 0, PUR_HYP_P1, c), s.A_p, s.B_p);
      s.Alpha_p[c] = gamma_draw<double>(make_stream(s.seed, 0, PUR_HYP_P2, c), s.C_p, s.D_p);
    } else s.Lambda_p[c] = gamma_draw<double>(make_stream(s.seed, 0, PUR_HYP_P1, c), s.A_p, s.B_p);
  }
#pragma omp parallel for
  for (long long ii = 0; ii < NG; ++ii) {
    const uint64_t c = (uint64_t)(ii % s.N) + (uint64_t)s.N * (uint64_t)(s.g0 + ii / s.N);
    if (s.prior == CP_GAMMA) {
      s.Beta_e[ii] = gamma_draw<double>(make_stream(s.seed, 0, PUR_HYP_E1, c), s.A_e, s.B_e);
      s.Alpha_e[ii] = gamma_draw<double>(make_stream(s.seed, 0, PUR_HYP_E2, c), s.C_e, s.D_e);
    } else s.Lambda_e[ii] = gamma_draw<double>(make_stream(s.seed, 0, PUR_HYP_E1, c), s.A_e, s.B_e);
  }
  s.iter = 1;
  iteration(s, true);
  if (row9) memcpy(row9, s.row, sizeof(s.row));
}

// n full Gibbs iterations (R/bayesNMF_sampler.R:273-285); rows = n x 9 (iter, RMSE, KL, loglik, logpost,
// n_params, BIC, rank, temp)
void cp_step(void* h, int n, double* rows) {
  Port& s = *static_cast<Port*>(h);
  for (int i = 0; i < n; ++i) {
    s.iter += 1;
    iteration(s, false);
    if (rows) memcpy(rows + 9 * (size_t)i, s.row, sizeof(s.row));
  }
}

// name: 0 P, 1 E, 2 SP, 3 SE, 4 Alpha_p, 5 Beta_p, 6 Alpha_e, 7 Beta_e, 8 Lambda_p, 9 Lambda_e
long long cp_get(void* h, int name, double* out) {
  Port& s = *static_cast<Port*>(h);
  const std::vector<double>* v = nullptr;
  switch (name) {
    case 0: v = &s.P; break; case 1: v = &s.E; break;
    case 2: for (size_t i = 0; i < s.SP.size(); ++i) out[i] = (double)s.SP[i]; return (long long)s.SP.size();
    case 3: for (size_t i = 0; i < s.SE.size(); ++i) out[i] = (double)s.SE[i]; return (long long)s.SE.size();
    case 4: v = &s.Alpha_p; break; case 5: v = &s.Beta_p; break; case 6: v = &s.Alpha_e; break; case 7: v = &s.Beta_e; break;
    case 8: v = &s.Lambda_p; break; case 9: v = &s.Lambda_e; break;
    default: return -1;
  }
  memcpy(out, v->data(), v->size() * sizeof(double));
  return (long long)v->size();
}

int cp_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
