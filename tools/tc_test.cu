// Stand-alone check of k_mhat_tc (bayesnmf_b200/csrc/bnmf_tc.cuh) against an fp64 host product.  dev tool:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_test tools/tc_test.cu && gpurun -- ./tools/tc_test
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../bayesnmf_b200/csrc/bnmf_tc.cuh"
using namespace bnmf;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
static double urand() { return (double)rand() / RAND_MAX; }
int run(int K, int N, long long G) {
  std::vector<double> P((size_t)K * N), E((size_t)N * G), M((size_t)K * G), ref((size_t)K * G);
  std::vector<int> A(N, 1);
  if (N > 2) A[1] = 0;
  for (auto& v : P) v = urand() < 0.2 ? 1e-5 * urand() : urand() * 0.3;
  for (auto& v : E) v = -log(urand() + 1e-12) * 200.0;
  for (long long g = 0; g < G; ++g) if (g % 17 == 3) for (int n = 0; n < N; ++n) E[n + N * g] = 0.0;     // an all-zero column
  for (int k = 0; k < K; ++k) for (long long g = 0; g < G; ++g) { double s = 0; for (int n = 0; n < N; ++n) s += (A[n] ? P[k + (size_t)K * n] : 0.0) * E[n + N * g]; ref[k + (size_t)K * g] = s; }
  double *dP, *dE, *dM; int* dA;
  CK(cudaMalloc(&dP, P.size() * 8)); CK(cudaMalloc(&dE, E.size() * 8)); CK(cudaMalloc(&dM, M.size() * 8)); CK(cudaMalloc(&dA, N * 4));
  CK(cudaMemcpy(dP, P.data(), P.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dE, E.data(), E.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA, A.data(), N * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dM, 0xff, M.size() * 8));
  const size_t smem = tc_smem_bytes<double>(N);
  CK(cudaFuncSetAttribute(k_mhat_tc<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((G + TC_N - 1) / TC_N), (unsigned)((K + TC_M - 1) / TC_M));
  unsigned char* dPd; int* deP;
  CK(cudaMalloc(&dPd, tc_planes_bytes(K, N))); CK(cudaMalloc(&deP, K * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_tc_prep_P<double><<<grid.y, 128>>>(dP, dA, dPd, deP, K, N);
  k_mhat_tc<double><<<grid, 128, smem>>>(dPd, deP, dE, dM, K, N, G);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) { k_tc_prep_P<double><<<grid.y, 128>>>(dP, dA, dPd, deP, K, N); k_mhat_tc<double><<<grid, 128, smem>>>(dPd, deP, dE, dM, K, N, G); }
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  CK(cudaMemcpy(M.data(), dM, M.size() * 8, cudaMemcpyDeviceToHost));
  double max_abs = 0, max_rel = 0, max_scaled = 0;
  for (int k = 0; k < K; ++k) {
    double pm = 0; for (int n = 0; n < N; ++n) if (A[n]) pm = fmax(pm, P[k + (size_t)K * n]);
    for (long long g = 0; g < G; ++g) {
      double em = 0; for (int n = 0; n < N; ++n) em = fmax(em, E[n + N * g]);
      const double d = fabs(M[k + (size_t)K * g] - ref[k + (size_t)K * g]);
      if (!(d == d)) { printf("NaN at %d %lld\n", k, g); return 1; }
      max_abs = fmax(max_abs, d);
      if (ref[k + (size_t)K * g] > 0) max_rel = fmax(max_rel, d / ref[k + (size_t)K * g]);
      if (pm * em > 0) max_scaled = fmax(max_scaled, d / (pm * em));
    }
  }
  printf("K=%d N=%d G=%lld: max abs %.3e, max rel %.3e, max |err|/(rowmax*colmax) %.3e (bound N*2^-40 = %.3e), %.3f ms/launch, %.1f GB/s written\n",
         K, N, G, max_abs, max_rel, max_scaled, N * 9.094947017729282e-13, ms / 5, 8.0 * K * G / (ms / 5 * 1e-3) / 1e9);
  cudaFree(dP); cudaFree(dE); cudaFree(dM); cudaFree(dA);
  return max_scaled < N * 4e-12 ? 0 : 2;
}
int main() {
  int rc = 0;
  rc |= run(96, 15, 200);
  rc |= run(96, 5, 64);
  rc |= run(300, 40, 130);
  rc |= run(96, 15, 20000);
  rc |= run(1536, 40, 50000);
  rc |= run(96, 48, 1000);
  printf(rc ? "FAILED\n" : "tc_test ok\n");
  return rc;
}
