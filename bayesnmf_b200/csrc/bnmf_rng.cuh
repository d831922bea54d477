// Counter-based random draws shared by every bayesnmf_b200 kernel.
//
// All randomness in the sampler comes from Philox4x32-10 (Salmon et al., SC'11)
// keyed by the run seed and addressed by (iteration, purpose, cell, sub-draw), so a
// draw is a pure function of *what* is being drawn, never of which thread, block,
// GPU or shard happens to draw it.  The reference (R) uses the global Mersenne
// Twister stream (stats::rgamma / rexp / rnorm / runif / rmultinom,
// truncnorm::rtruncnorm, invgamma::rinvgamma, armspp::arms -- see
// /root/reference/R/sample_Pn.R:14-27,79-85,116-118,243, R/sample_priors.R:219-397,
// R/sample_params.R:104,164,239,263,279); parity with it is distributional, parity
// with oracle/ (which restates exactly the arithmetic below in numpy) is
// draw-for-draw.
//
// Everything here is __host__ __device__ so that tests/hostcheck can execute the
// very same code on the CPU; the product never does.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define BNMF_HD __host__ __device__ __forceinline__
// the big samplers are called, not inlined: a kernel that inlines three gamma draws and an
// adaptive-rejection draw is ~140 KB of straight-line code and stalls on instruction fetch
#define BNMF_HD_CALL __host__ __device__ __noinline__
#else
#define BNMF_HD inline
#define BNMF_HD_CALL inline
#endif

namespace bnmf {

// ---- draw purposes (low 8 bits of counter word 3) -------------------------------
enum Purpose : uint32_t {
  PUR_Z        = 0,   // latent-count picks, cell = k + K*g, sub = pick/4
  PUR_P        = 1,   // P[k,n] draw (gamma / truncated normal / prior), cell = k + K*n
  PUR_E        = 2,   // E[n,g] draw, cell = n + N*g
  PUR_HYP_P1   = 3,   // Beta_p | Lambda_p | Mu_p
  PUR_HYP_P2   = 4,   // Alpha_p | Sigmasq_p
  PUR_HYP_E1   = 5,   // Beta_e | Lambda_e | Mu_e
  PUR_HYP_E2   = 6,   // Alpha_e | Sigmasq_e
  PUR_MH_P     = 7,   // Metropolis-Hastings accept uniform for P[k,n]
  PUR_MH_E     = 8,   // Metropolis-Hastings accept uniform for E[n,g]
  PUR_A        = 9,   // inclusion indicator A_n, cell = n
  PUR_R        = 10,  // expected rank R, cell = 0
  PUR_SIGMASQ  = 11,  // sigmasq_g, cell = g
};

struct U4 { uint32_t x, y, z, w; };

BNMF_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// Philox4x32-10.  Counter (c0..c3), key (k0,k1).
BNMF_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                         uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// The sampler's addressing convention: key = 64-bit seed; counter =
// (cell_lo, cell_hi, sub, iter<<8 | purpose).
struct Stream {
  uint32_t k0, k1;   // seed
  uint32_t c3;       // iter<<8 | purpose
  uint32_t c0, c1;   // cell
  BNMF_HD U4 at(uint32_t sub) const { return philox4x32_10(c0, c1, sub, c3, k0, k1); }
};

BNMF_HD Stream make_stream(uint64_t seed, uint32_t iter, uint32_t purpose, uint64_t cell) {
  Stream s;
  s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32);
  s.c3 = (iter << 8) | purpose;
  s.c0 = (uint32_t)cell; s.c1 = (uint32_t)(cell >> 32);
  return s;
}

// ---- uniforms -------------------------------------------------------------------
// Open interval (0,1): (w + 0.5) * 2^-32 is exact in double.  In float only 24 bits
// survive, so the float path drops the low byte first (oracle: uniform_bits=24).
template <typename T> BNMF_HD T u01(uint32_t w);
template <> BNMF_HD double u01<double>(uint32_t w) {
  return ((double)w + 0.5) * 2.3283064365386963e-10;  // 2^-32
}
template <> BNMF_HD float u01<float>(uint32_t w) {
  return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-8f;  // 2^-24
}

// ---- small math helpers kept explicit so that no a*b+c is contracted -----------
template <typename T> BNMF_HD T tlog(T x) { return log(x); }
template <typename T> BNMF_HD T texp(T x) { return exp(x); }
template <typename T> BNMF_HD T tsqrt(T x) { return sqrt(x); }
template <typename T> BNMF_HD T tlgamma(T x) { return lgamma(x); }
#if defined(__CUDACC__)
template <> BNMF_HD float tlog<float>(float x) { return logf(x); }
template <> BNMF_HD float texp<float>(float x) { return expf(x); }
template <> BNMF_HD float tsqrt<float>(float x) { return sqrtf(x); }
template <> BNMF_HD float tlgamma<float>(float x) { return lgammaf(x); }
#endif

// Standard normal by Box-Muller (cosine branch) from two words.
template <typename T> BNMF_HD T normal_from(uint32_t w0, uint32_t w1) {
  T u1 = u01<T>(w0), u2 = u01<T>(w1);
  T r = tsqrt<T>((T)-2 * tlog<T>(u1));
  return r * (T)cos((T)6.283185307179586 * u2);
}

// Exponential(rate): -log(u)/rate.  (stats::rexp, R/sample_Pn.R:21)
template <typename T> BNMF_HD T exponential_draw(const Stream& s, T rate) {
  U4 w = s.at(0);
  return -tlog<T>(u01<T>(w.x)) / rate;
}

// Normal(mean, sd).  (stats::rnorm, R/sample_priors.R:34,219)
template <typename T> BNMF_HD T normal_draw(const Stream& s, T mean, T sd) {
  U4 w = s.at(0);
  return mean + sd * normal_from<T>(w.x, w.y);
}

// Gamma(shape, rate) by Marsaglia & Tsang (2000); shape < 1 boosted by U^(1/shape).
// Attempt t consumes Philox block `sub0 + t`: words (x,y) -> normal, z -> accept
// uniform, w -> boost uniform.  Result floored at the smallest normal so that a
// later log() stays finite (R's rgamma can return exactly 0 for tiny shapes).
// (stats::rgamma(n, shape, rate): R/sample_Pn.R:23-27,116-118, R/sample_priors.R:285-344)
// gamma_unit: the variate before the division by the rate (everything that needs the shape only; k_sides draws it
// while the rate is still being formed); gamma_scale: the division and the floor.  gamma_draw = the two in a row.
template <typename T> BNMF_HD_CALL T gamma_unit(const Stream s, T shape, uint32_t sub0 = 0) {
  const bool boost = shape < (T)1;
  const T a = boost ? shape + (T)1 : shape;
  const T d = a - (T)(1.0 / 3.0);
  const T c = (T)1 / tsqrt<T>((T)9 * d);
  T g = d;  // fallback if the attempt cap is ever hit (probability ~ 0)
  for (uint32_t t = 0; t < 4096u; ++t) {
    U4 w = s.at(sub0 + t);
    T x = normal_from<T>(w.x, w.y);
    T v = (T)1 + c * x;
    if (v <= (T)0) continue;
    v = v * v * v;
    T u = u01<T>(w.z);
    T x2 = x * x;
    bool acc = u < (T)1 - (T)0.0331 * (x2 * x2);
    if (!acc) acc = tlog<T>(u) < (T)0.5 * x2 + d * ((T)1 - v + tlog<T>(v));
    if (acc) {
      g = d * v;
      if (boost) g = g * (T)pow(u01<T>(w.w), (T)1 / shape);
      break;
    }
  }
  return g;
}
template <typename T> BNMF_HD T gamma_scale(T g, T rate) {
  g = g / rate;
  const T tiny = sizeof(T) == 8 ? (T)DBL_MIN : (T)FLT_MIN;
  return g < tiny ? tiny : g;
}
template <typename T> BNMF_HD T gamma_draw(const Stream s, T shape, T rate, uint32_t sub0 = 0) {
  return gamma_scale<T>(gamma_unit<T>(s, shape, sub0), rate);
}

// Normal(mean, sd) truncated to [0, inf).  Every call site of truncnorm::rtruncnorm
// in the reference has a = 0, b = Inf (R/sample_Pn.R:14-19,59-64,79-85,
// R/sample_En.R:14-19,59-64,78-84).  With alpha = -mean/sd:
//   alpha <= 0.45 : plain normal rejection z >= alpha (acceptance >= 0.326)
//   alpha >  0.45 : Robert (1995) translated-exponential rejection; the result is
//                   formed as sd*(z-alpha) directly so a mean far below zero does
//                   not cancel catastrophically.
// Attempt t consumes Philox block t: (x,y) -> normal or (x -> exp, y -> accept).
// t0 = first attempt to try (callers that have already evaluated attempts 0 .. t0-1 themselves).
template <typename T> BNMF_HD_CALL T truncnorm0_draw(const Stream s, T mean, T sd, uint32_t t0 = 0) {
  const T alpha = -mean / sd;
  if (alpha <= (T)0.45) {
    T z = alpha;
    for (uint32_t t = t0; t < 4096u; ++t) {
      U4 w = s.at(t);
      T zz = normal_from<T>(w.x, w.y);
      if (zz >= alpha) { z = zz; break; }
    }
    T x = mean + sd * z;
    return x < (T)0 ? (T)0 : x;
  }
  const T lam = (T)0.5 * (alpha + tsqrt<T>(alpha * alpha + (T)4));
  T e = (T)0;
  for (uint32_t t = t0; t < 4096u; ++t) {
    U4 w = s.at(t);
    T ee = -tlog<T>(u01<T>(w.x)) / lam;   // z - alpha
    T dz = (alpha + ee) - lam;
    if (tlog<T>(u01<T>(w.y)) <= (T)-0.5 * (dz * dz)) { e = ee; break; }
  }
  return sd * e;
}

// ---- digamma / trigamma (recurrence to x >= 6, then asymptotic series) ----------
// Evaluated together: the Newton step on h' needs both and they share every reciprocal.
template <typename T> BNMF_HD_CALL void digamma_trigamma(T x, T& psi, T& tri) {
  T r1 = (T)0, r2 = (T)0;
  while (x < (T)6) {
    const T inv = (T)1 / x;
    r1 = r1 - inv;
    r2 = r2 + inv * inv;
    x = x + (T)1;
  }
  const T inv = (T)1 / x;
  const T f = inv * inv;
  const T t1 = f * ((T)(-1.0 / 12.0) + f * ((T)(1.0 / 120.0) + f * ((T)(-1.0 / 252.0) +
               f * ((T)(1.0 / 240.0) + f * (T)(-1.0 / 132.0)))));
  psi = r1 + tlog<T>(x) - (T)0.5 * inv + t1;
  const T t2 = inv + (T)0.5 * f +
               (f * inv) * ((T)(1.0 / 6.0) + f * ((T)(-1.0 / 30.0) + f * ((T)(1.0 / 42.0) +
               f * (T)(-1.0 / 30.0))));
  tri = r2 + t2;
}
template <typename T> BNMF_HD T digamma(T x) { T p, t; digamma_trigamma<T>(x, p, t); return p; }
template <typename T> BNMF_HD T trigamma(T x) { T p, t; digamma_trigamma<T>(x, p, t); return t; }

// ---- shape parameter of the Gamma prior --------------------------------------
// The reference draws Alpha[.] with armspp::arms(n_samples = 1, log_pdf, 1e-3, 1e4)
// (R/sample_priors.R:356-397) from
//   log f(x) = (C-1) log x - D x + x log(Beta) + (x-1) log(X) - lgamma(x)
//            = cm1 log x - b x - lgamma(x) + const,   b = D - log(Beta) - log(X).
// f'' = -(C-1)/x^2 - trigamma(x) < -C/x^2 < 0, so the target is log-concave for
// every C > 0 and ARMS degenerates to exact adaptive rejection sampling.  We draw
// exactly from the same density with a fixed three-tangent envelope (tangents at
// mode-s, mode, mode+s, s = Laplace sd), i.e. non-adaptive ARS.  Any choice of
// tangent points gives a valid envelope; only the acceptance rate depends on them.
struct AlphaTarget { double cm1, b; };
BNMF_HD_CALL double alpha_h(const AlphaTarget t, double x) { return t.cm1 * log(x) - t.b * x - lgamma(x); }
BNMF_HD double alpha_hp(const AlphaTarget& t, double x) { return t.cm1 / x - t.b - digamma<double>(x); }
BNMF_HD double alpha_hpp(const AlphaTarget& t, double x) { return -t.cm1 / (x * x) - trigamma<double>(x); }
BNMF_HD void alpha_hp_hpp(const AlphaTarget& t, double x, double& f, double& fp) {
  double psi, tri;
  digamma_trigamma<double>(x, psi, tri);
  f = t.cm1 / x - t.b - psi;
  fp = -t.cm1 / (x * x) - tri;
}

// integral of exp(-|s| y) over [0, w]: the mass of a linear-exponent segment of width w
// relative to its higher end (never overflows).
BNMF_HD double seg_unit(double s, double w) {
  const double sw = fabs(s) * w;
  if (sw < 1e-8) return w * (1.0 - 0.5 * sw);
  return -expm1(-sw) / fabs(s);
}
// x in [a,b] with partial mass fraction q in (0,1) of the segment exp(s x) over [a,b].
BNMF_HD double seg_inv(double s, double a, double b, double q) {
  double w = b - a;
  double sw = s * w;
  if (fabs(sw) < 1e-8) return a + q * w;
  if (sw > 0.0) return b + log1p((1.0 - q) * expm1(-sw)) / s;
  return a + log1p(q * expm1(sw)) / s;
}

// digamma(1e-3), digamma(1e4): the two end-point tests of the mode search, as literals so that
// the oracle, the host check and the kernels decide them identically
#define BNMF_PSI_LO (-1000.5755719318103)
#define BNMF_PSI_HI (9.21029037114285)
#define BNMF_ALPHA_NEWTON 16
#define BNMF_ALPHA_TOL 1e-2

// x0 = where the search for the mode starts: the current value of Alpha (a draw from this
// very conditional one sweep ago, hence within about one standard deviation of the mode).
// The mode is only needed approximately: any three increasing tangent points give a valid
// envelope of a concave log-density, their position only moves the acceptance rate.
// STAGES = true (device, every thread of the block calls it): block barriers between the mode
// search, the envelope and the rejection loop keep the warps of a block in the same code.
#if defined(__CUDA_ARCH__)
#define BNMF_STAGE() do { if (STAGES) __syncthreads(); } while (0)
#else
#define BNMF_STAGE() do { } while (0)
#endif
// Envelope of one Alpha conditional: everything an attempt of the rejection loop reads.
struct AlphaEnv {
  AlphaTarget t;
  double hv[3], sl[3], xs[3], z[4], mass[3], tot, m;
};
template <bool STAGES = false>
BNMF_HD void alpha_setup(AlphaEnv& e, double C, double D, double beta, double X, double x0) {
  const double LO = 1e-3, HI = 1e4;
  AlphaTarget t; t.cm1 = C - 1.0; t.b = D - log(beta) - log(X);
  // --- approximate mode: safeguarded Newton on h' (strictly decreasing) to BNMF_ALPHA_TOL ---
  double m;
  if (t.cm1 / LO - t.b - BNMF_PSI_LO <= 0.0) m = LO;
  else if (t.cm1 / HI - t.b - BNMF_PSI_HI >= 0.0) m = HI;
  else {
    double a = LO, b = HI;
    double x = x0;
    if (!(x > LO && x < HI)) x = C > 1.0 ? (C < HI ? C : 0.5 * HI) : 1.0;
    for (int it = 0; it < BNMF_ALPHA_NEWTON; ++it) {
      double f, fp;
      alpha_hp_hpp(t, x, f, fp);
      if (f > 0.0) a = x; else b = x;
      double xn = x * exp(-f / (x * fp));               // Newton step in log x (h' is close to linear in it)
      if (!(xn > a && xn < b)) xn = sqrt(a * b);        // safeguard: geometric bisection
      const bool conv = fabs(xn - x) <= BNMF_ALPHA_TOL * x;
      x = xn;
      if (conv) break;
    }
    m = x;
  }
  BNMF_STAGE();
  double hp_m, hpp_m;
  alpha_hp_hpp(t, m, hp_m, hpp_m);
  double s = 1.0 / sqrt(-hpp_m);
  // --- three tangent points inside (0, inf), strictly increasing ---
  double xs[3];
  xs[1] = m;
  xs[0] = (m - s > 0.5 * m) ? m - s : 0.5 * m;
  xs[2] = m + s;
  if (m <= LO) { xs[0] = LO; xs[1] = LO + s; xs[2] = LO + 2.0 * s; }
  if (m >= HI) { xs[2] = HI; xs[1] = HI - s; xs[0] = HI - 2.0 * s; if (xs[0] < 0.5 * HI) { xs[0] = 0.5 * HI; xs[1] = 0.75 * HI; } }
  double hv[3], sl[3];
  for (int j = 0; j < 3; ++j) {
    hv[j] = alpha_h(t, xs[j]);
    sl[j] = (xs[j] == m) ? hp_m : alpha_hp(t, xs[j]);   // the centre point is m unless the mode sits on a bound
  }
  // --- segment boundaries: tangent intersections, clipped to [LO, HI] ---
  double z[4];
  z[0] = LO; z[3] = HI;
  for (int j = 0; j < 2; ++j) {
    double den = sl[j] - sl[j + 1];
    double zz = (hv[j + 1] - hv[j] + sl[j] * xs[j] - sl[j + 1] * xs[j + 1]) / den;
    if (!(zz >= xs[j])) zz = xs[j];
    if (!(zz <= xs[j + 1])) zz = xs[j + 1];
    if (zz < LO) zz = LO;
    if (zz > HI) zz = HI;
    z[j + 1] = zz;
  }
  // masses of the three envelope segments relative to the envelope's overall maximum (a
  // piecewise-linear exponent peaks at a segment end): nothing overflows even when the
  // tangent points do not bracket the mode
  double top[3];
  for (int j = 0; j < 3; ++j) {
    const double el = hv[j] + sl[j] * (z[j] - xs[j]), er = hv[j] + sl[j] * (z[j + 1] - xs[j]);
    top[j] = el > er ? el : er;
  }
  double hmax = top[0];
  if (top[1] > hmax) hmax = top[1];
  if (top[2] > hmax) hmax = top[2];
  double tot = 0.0;
  for (int j = 0; j < 3; ++j) {
    e.mass[j] = (z[j + 1] > z[j]) ? exp(top[j] - hmax) * seg_unit(sl[j], z[j + 1] - z[j]) : 0.0;
    tot += e.mass[j];
  }
  e.t = t; e.tot = tot; e.m = m;
  for (int j = 0; j < 3; ++j) { e.hv[j] = hv[j]; e.sl[j] = sl[j]; e.xs[j] = xs[j]; }
  for (int j = 0; j < 4; ++j) e.z[j] = z[j];
}
// attempt number `a` of the rejection loop: true = accepted, the draw is in x
BNMF_HD bool alpha_attempt(const AlphaEnv& e, const Stream st, uint32_t a, double& x) {
  U4 w = st.at(a);
  double r = u01<double>(w.x) * e.tot;
  int j = 0;
  if (r >= e.mass[0]) { r -= e.mass[0]; j = 1; if (r >= e.mass[1]) { r -= e.mass[1]; j = 2; } }
  if (!(e.mass[j] > 0.0)) return false;
  double q = r / e.mass[j];
  if (q >= 1.0) q = 1.0 - 1e-16;
  double xc = seg_inv(e.sl[j], e.z[j], e.z[j + 1], q);
  if (xc < e.z[j]) xc = e.z[j];
  if (xc > e.z[j + 1]) xc = e.z[j + 1];
  double env = e.hv[j] + e.sl[j] * (xc - e.xs[j]);
  if (log(u01<double>(w.y)) <= alpha_h(e.t, xc) - env) { x = xc; return true; }
  return false;
}
#define BNMF_ALPHA_MAX_ATTEMPTS 4096u
template <bool STAGES = false>
BNMF_HD_CALL double alpha_draw(const Stream st, double C, double D, double beta, double X, double x0) {
  AlphaEnv e;
  alpha_setup<STAGES>(e, C, D, beta, X, x0);
  BNMF_STAGE();
  double x = e.m;
  for (uint32_t a = 0; a < BNMF_ALPHA_MAX_ATTEMPTS; ++a)
    if (alpha_attempt(e, st, a, x)) break;
  return x;
}

}  // namespace bnmf
