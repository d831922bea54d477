// get_MAP_ on the device sample ring (R/utils.R:194-288): mode of A over the newest
// n_samples samples (get_mode, R/helpers.R:63-79), then over the samples that match it the
// element-wise mean of the renormalised P and E (renormalize, R/helpers.R:35-49:
// P[,n] / colSums(P)[n], E[n,] * colSums(P)[n]).  Only K x N and N x G results cross PCIe;
// the E samples (N x G each) never leave HBM.  Included by bnmf_api.cu.

// colsum[s][n] = sum_k P_s[k,n] for the matching ring slots; one block per (n, s)
template <typename T> static __global__ void k_map_colsum(Dev<T> d, const int* slots, double* colsum) {
  __shared__ double scratch[4];
  const int n = blockIdx.x, s = blockIdx.y;
  const T* P = d.ring_P + (long long)slots[s] * d.K * d.N + (long long)d.K * n;
  double v = 0.0;
  for (int k = threadIdx.x; k < d.K; k += 128) v += (double)P[k];
  const double r = block_sum<128>(v, scratch);
  if (threadIdx.x == 0) colsum[(long long)s * d.N + n] = r;
}
// Reduce("+", samples) / length, in sample order (oldest first) like the reference
template <typename T> static __global__ void k_map_mean_P(Dev<T> d, const int* slots, int n_match, const double* colsum, double* out) {
  const long long KN = (long long)d.K * d.N;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KN) return;
  const int n = (int)(i / d.K);
  double acc = 0.0;
  for (int s = 0; s < n_match; ++s) acc += (double)d.ring_P[(long long)slots[s] * KN + i] / colsum[(long long)s * d.N + n];
  out[i] = acc / (double)n_match;
}
template <typename T> static __global__ void k_map_mean_E(Dev<T> d, const int* slots, int n_match, const double* colsum, double* out) {
  const long long NG = (long long)d.N * d.G;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NG) return;
  const int n = (int)(i % d.N);
  double acc = 0.0;
  for (int s = 0; s < n_match; ++s) acc += (double)d.ring_E[(long long)slots[s] * NG + i] * colsum[(long long)s * d.N + n];
  out[i] = acc / (double)n_match;
}

template <typename T> int Sampler<T>::get_map(int n_samples, double* P_map, double* E_map, double* A_map, int* n_match_out) {
  CK(cudaSetDevice(cfg.device));
  if (d.ring_cap <= 0) return fail("bnmf_get_map: the handle was created with ring_cap = 0");
  Ctrl hc; CK(cudaMemcpyAsync(&hc, d.ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  if (n_samples < 1 || n_samples > hc.ring_count)
    return fail("bnmf_get_map: n_samples = %d outside the %d samples held", n_samples, hc.ring_count);
  const int N = cfg.N, K = cfg.K; const long long KN = (long long)K * N, NG = (long long)N * cfg.G;
  std::vector<int32_t> ringA((size_t)d.ring_cap * N);
  CK(cudaMemcpyAsync(ringA.data(), d.ring_A, sizeof(int32_t) * ringA.size(), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  // samples oldest -> newest, as considered_idx is increasing in the reference
  std::vector<int> slot(n_samples);
  std::vector<std::string> key(n_samples);
  for (int j = 0; j < n_samples; ++j) {
    const int ago = n_samples - 1 - j;
    slot[j] = ((hc.ring_pos - 1 - ago) % d.ring_cap + d.ring_cap) % d.ring_cap;
    std::string s(N, '0');
    for (int n = 0; n < N; ++n) s[n] = ringA[(size_t)slot[j] * N + n] ? '1' : '0';
    key[j] = s;
  }
  // get_mode: table() orders the patterns alphabetically and the stable decreasing sort keeps
  // that order among ties, so the modal pattern is the most frequent, smallest string
  std::map<std::string, int> cnt;
  for (auto& s : key) cnt[s] += 1;
  std::string mode; int best = -1;
  for (auto& kv : cnt) if (kv.second > best) { best = kv.second; mode = kv.first; }
  std::vector<int> match;
  for (int j = 0; j < n_samples; ++j) if (key[j] == mode) match.push_back(slot[j]);
  const int nm = (int)match.size();
  int* dslots; double* colsum; double* outP; double* outE;
  CK(cudaMalloc((void**)&dslots, sizeof(int) * nm));
  CK(cudaMalloc((void**)&colsum, sizeof(double) * (size_t)nm * N));
  CK(cudaMalloc((void**)&outP, sizeof(double) * KN));
  CK(cudaMalloc((void**)&outE, sizeof(double) * NG));
  CK(cudaMemcpyAsync(dslots, match.data(), sizeof(int) * nm, cudaMemcpyHostToDevice, stream));
  k_map_colsum<T><<<dim3(N, nm), 128, 0, stream>>>(d, dslots, colsum);
  k_map_mean_P<T><<<blocks(KN, 128), 128, 0, stream>>>(d, dslots, nm, colsum, outP);
  k_map_mean_E<T><<<blocks(NG, 256), 256, 0, stream>>>(d, dslots, nm, colsum, outE);
  if (P_map) CK(cudaMemcpyAsync(P_map, outP, sizeof(double) * KN, cudaMemcpyDeviceToHost, stream));
  if (E_map) CK(cudaMemcpyAsync(E_map, outE, sizeof(double) * NG, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  CK(cudaGetLastError());
  cudaFree(dslots); cudaFree(colsum); cudaFree(outP); cudaFree(outE);
  if (A_map) for (int n = 0; n < N; ++n) A_map[n] = mode[n] == '1' ? 1.0 : 0.0;
  if (n_match_out) *n_match_out = nm;
  return 0;
}
