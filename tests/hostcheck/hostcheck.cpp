// TEST INFRASTRUCTURE ONLY.  Compiles the __host__ __device__ draw code of
// bayesnmf_b200/csrc/bnmf_rng.cuh with g++ so the very arithmetic the kernels run can
// be compared with oracle/ on a machine without a GPU.  Never used by the product.
#include "../../bayesnmf_b200/csrc/bnmf_rng.cuh"
using namespace bnmf;
extern "C" {
void hc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  U4 o = philox4x32_10(c0, c1, c2, c3, k0, k1); out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}
void hc_gamma(uint64_t seed, uint32_t it, uint32_t pur, const uint64_t* cell, const double* shape, const double* rate, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = gamma_draw<double>(make_stream(seed, it, pur, cell[i]), shape[i], rate[i]);
}
void hc_truncnorm(uint64_t seed, uint32_t it, uint32_t pur, const uint64_t* cell, const double* mean, const double* sd, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = truncnorm0_draw<double>(make_stream(seed, it, pur, cell[i]), mean[i], sd[i]);
}
void hc_normal(uint64_t seed, uint32_t it, uint32_t pur, const uint64_t* cell, const double* mean, const double* sd, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = normal_draw<double>(make_stream(seed, it, pur, cell[i]), mean[i], sd[i]);
}
void hc_exp(uint64_t seed, uint32_t it, uint32_t pur, const uint64_t* cell, const double* rate, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = exponential_draw<double>(make_stream(seed, it, pur, cell[i]), rate[i]);
}
void hc_alpha(uint64_t seed, uint32_t it, uint32_t pur, const uint64_t* cell, const double* C, const double* D, const double* beta, const double* X, const double* x0, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = alpha_draw(make_stream(seed, it, pur, cell[i]), C[i], D[i], beta[i], X[i], x0[i]);
}
void hc_digamma(const double* x, double* d, double* t, long n) { for (long i = 0; i < n; ++i) { d[i] = digamma<double>(x[i]); t[i] = trigamma<double>(x[i]); } }
}
