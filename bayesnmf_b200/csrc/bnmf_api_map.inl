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

// slots of the newest n_samples ring samples whose A equals the modal A (oldest first), and that A
template <typename T> int Sampler<T>::map_slots(int n_samples, std::vector<int>& match, std::string& mode) {
  if (d.ring_cap <= 0) return fail("bnmf_get_map: the handle was created with ring_cap = 0");
  Ctrl hc; CK(cudaMemcpyAsync(&hc, d.ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  if (n_samples < 1 || n_samples > hc.ring_count)
    return fail("bnmf_get_map: n_samples = %d outside the %d samples held", n_samples, hc.ring_count);
  const int N = cfg.N;
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
  int best = -1;
  for (auto& kv : cnt) if (kv.second > best) { best = kv.second; mode = kv.first; }
  match.clear();
  for (int j = 0; j < n_samples; ++j) if (key[j] == mode) match.push_back(slot[j]);
  return 0;
}

template <typename T> int Sampler<T>::get_map(int n_samples, double* P_map, double* E_map, double* A_map, int* n_match_out) {
  CK(cudaSetDevice(cfg.device));
  std::vector<int> match; std::string mode;
  if (map_slots(n_samples, match, mode)) return 1;
  const int N = cfg.N, K = cfg.K; const long long KN = (long long)K * N, NG = (long long)N * cfg.G;
  const int nm = (int)match.size();
  Scratch sc(stream, cfg.device);
  int* dslots; double* colsum; double* outP; double* outE;
  CK(sc.get(&dslots, sizeof(int) * nm));
  CK(sc.get(&colsum, sizeof(double) * (size_t)nm * N));
  CK(sc.get(&outP, sizeof(double) * KN));
  CK(sc.get(&outE, sizeof(double) * NG));
  CK(cudaMemcpyAsync(dslots, match.data(), sizeof(int) * nm, cudaMemcpyHostToDevice, stream));
  k_map_colsum<T><<<dim3(N, nm), 128, 0, stream>>>(d, dslots, colsum);
  k_map_mean_P<T><<<blocks(KN, 128), 128, 0, stream>>>(d, dslots, nm, colsum, outP);
  k_map_mean_E<T><<<blocks(NG, 256), 256, 0, stream>>>(d, dslots, nm, colsum, outE);
  if (P_map) CK(cudaMemcpyAsync(P_map, outP, sizeof(double) * KN, cudaMemcpyDeviceToHost, stream));
  if (E_map) CK(cudaMemcpyAsync(E_map, outE, sizeof(double) * NG, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  CK(cudaGetLastError());
  if (A_map) for (int n = 0; n < N; ++n) A_map[n] = mode[n] == '1' ? 1.0 : 0.0;
  if (n_match_out) *n_match_out = nm;
  return 0;
}

// Credible intervals of get_MAP_ (R/utils.R:264-287): element-wise quantiles (R's default, type 7:
// x[j] + (h - j)(x[j+1] - x[j]), h = (n - 1) p) of the renormalised P and E over the samples that
// match the modal A.  A block takes CI_EPB consecutive elements: every matching sample contributes
// one contiguous segment (coalesced), a warp sorts the values of one element in shared memory
// (bitonic, padded with +inf to a power of two) and picks the two quantiles.  The E samples
// (N x G each, up to MAP_over of them) never leave HBM.
constexpr int CI_EPB = 8;
template <typename T> static __global__ void __launch_bounds__(32 * CI_EPB)
k_ci(Dev<T> d, const int* slots, int nm, int NS, const double* colsum, int side, double plo, double phi,
     double* out_lo, double* out_hi) {
  extern __shared__ double cv[];                       // [CI_EPB][NS]
  const long long len = side == 0 ? (long long)d.K * d.N : (long long)d.N * d.G;
  const T* ring = side == 0 ? d.ring_P : d.ring_E;
  const long long i0 = (long long)blockIdx.x * CI_EPB;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < CI_EPB * NS; t += blockDim.x) {
    const int e = t % CI_EPB, s = t / CI_EPB;
    const long long i = i0 + e;
    double v = INFINITY;
    if (s < nm && i < len) {
      const int n = side == 0 ? (int)(i / d.K) : (int)(i % d.N);
      const double cs = colsum[(long long)s * d.N + n];
      const double x = (double)ring[(long long)slots[s] * len + i];
      v = side == 0 ? x / cs : x * cs;                 // renormalize, R/helpers.R:35-49
      if (v != v) v = INFINITY;                        // (0 / 0 columns sort last, as NaN would in R's quantile with na.rm)
    }
    cv[e * NS + s] = v;
  }
  __syncthreads();
  double* x = cv + wid * NS;
  for (int k = 2; k <= NS; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int idx = lane; idx < NS; idx += 32) {
        const int partner = idx ^ j;
        if (partner > idx) {
          const double a = x[idx], b = x[partner];
          const bool up = (idx & k) == 0;
          if ((a > b) == up) { x[idx] = b; x[partner] = a; }
        }
      }
      __syncwarp();
    }
  }
  const long long i = i0 + wid;
  if (lane < 2 && i < len) {
    const double p = lane == 0 ? plo : phi;
    const double h = (double)(nm - 1) * p;
    int j = (int)floor(h);
    if (j > nm - 1) j = nm - 1;
    const double g = h - (double)j;
    const double a = x[j], b = x[j + 1 < nm ? j + 1 : nm - 1];
    const double q = a + g * (b - a);
    (lane == 0 ? out_lo : out_hi)[i] = q;
  }
}

template <typename T> int Sampler<T>::get_ci(int n_samples, double plo, double phi, double* P_lo, double* P_hi,
                                             double* E_lo, double* E_hi, int* n_match_out) {
  CK(cudaSetDevice(cfg.device));
  if (!(plo >= 0.0 && phi <= 1.0 && plo <= phi)) return fail("bnmf_get_credible_intervals: need 0 <= lower <= upper <= 1");
  std::vector<int> match; std::string mode;
  if (map_slots(n_samples, match, mode)) return 1;
  const int N = cfg.N, K = cfg.K; const long long KN = (long long)K * N, NG = (long long)N * cfg.G;
  const int nm = (int)match.size();
  int NS = 2; while (NS < nm) NS <<= 1;
  const size_t smem = (size_t)CI_EPB * NS * sizeof(double);
  if (smem > (size_t)200 * 1024) return fail("bnmf_get_credible_intervals: %d matching samples do not fit the sort buffer", nm);
  CK(cudaFuncSetAttribute(k_ci<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  Scratch sc(stream, cfg.device);
  int* dslots; double* colsum; double* lo; double* hi;
  CK(sc.get(&dslots, sizeof(int) * nm));
  CK(sc.get(&colsum, sizeof(double) * (size_t)nm * N));
  CK(sc.get(&lo, sizeof(double) * std::max(KN, NG)));
  CK(sc.get(&hi, sizeof(double) * std::max(KN, NG)));
  CK(cudaMemcpyAsync(dslots, match.data(), sizeof(int) * nm, cudaMemcpyHostToDevice, stream));
  k_map_colsum<T><<<dim3(N, nm), 128, 0, stream>>>(d, dslots, colsum);
  for (int side = 0; side < 2; ++side) {
    double* o_lo = side == 0 ? P_lo : E_lo; double* o_hi = side == 0 ? P_hi : E_hi;
    if (!o_lo && !o_hi) continue;
    const long long len = side == 0 ? KN : NG;
    k_ci<T><<<(unsigned)((len + CI_EPB - 1) / CI_EPB), 32 * CI_EPB, smem, stream>>>(d, dslots, nm, NS, colsum, side, plo, phi, lo, hi);
    if (o_lo) CK(cudaMemcpyAsync(o_lo, lo, sizeof(double) * len, cudaMemcpyDeviceToHost, stream));
    if (o_hi) CK(cudaMemcpyAsync(o_hi, hi, sizeof(double) * len, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }
  CK(cudaGetLastError());
  if (n_match_out) *n_match_out = nm;
  return 0;
}

// ---- assign_signatures_ensemble_ (R/postprocessing.R:175-341) ---------------------------------------
// cos[s][i][j] = cosine(P_s[, keep[i]], ref[, j]) for the matching ring slots (s = n_match: the MAP
// signatures, mean of the renormalised samples); a block per (i, s), threads over j
template <typename T> static __global__ void k_assign_cos(Dev<T> d, const int* slots, int nm, const int* keep, int n_keep,
                                                           const double* ref, int n_ref, const double* Pmap, double* cosv) {
  extern __shared__ double col[];                 // [K] the estimated signature
  __shared__ double scratch[4];
  const int i = blockIdx.x, s = blockIdx.y, K = d.K;
  const int n = keep[i];
  double ss = 0.0;
  for (int k = threadIdx.x; k < K; k += 128) {
    const double v = s < nm ? (double)d.ring_P[(long long)slots[s] * K * d.N + (long long)K * n + k] : Pmap[(long long)K * n + k];
    col[k] = v; ss += v * v;
  }
  __shared__ double norm_e;
  const double tot = block_sum<128>(ss, scratch);
  if (threadIdx.x == 0) norm_e = sqrt(tot);
  __syncthreads();
  for (int j = threadIdx.x; j < n_ref; j += 128) {
    double dot = 0.0, rr = 0.0;
    for (int k = 0; k < K; ++k) { const double r = ref[(long long)K * j + k]; dot += col[k] * r; rr += r * r; }
    cosv[((long long)s * n_keep + i) * n_ref + j] = dot / (norm_e * sqrt(rr));     // lsa::cosine
  }
}

// minimum-cost assignment of n rows to m >= n columns (Hungarian algorithm with potentials, O(n^2 m));
// cost row-major n x m; asg[i] = column of row i
static void hungarian(const std::vector<double>& cost, int n, int m, std::vector<int>& asg) {
  const double INF = 1e300;
  std::vector<double> u(n + 1, 0.0), v(m + 1, 0.0);
  std::vector<int> p(m + 1, 0), way(m + 1, 0);
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    std::vector<double> minv(m + 1, INF);
    std::vector<char> used(m + 1, 0);
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = INF; int j1 = 0;
      for (int j = 1; j <= m; ++j) if (!used[j]) {
        const double cur = cost[(size_t)(i0 - 1) * m + (j - 1)] - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      for (int j = 0; j <= m; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; } else minv[j] -= delta;
      }
      j0 = j1;
    } while (p[j0] != 0);
    do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
  }
  asg.assign(n, -1);
  for (int j = 1; j <= m; ++j) if (p[j]) asg[p[j] - 1] = j - 1;
}

template <typename T> int Sampler<T>::assign(int n_samples, const double* ref, int n_ref, double ci, int* n_keep_out, int* keep_out,
                                             double* votes, int* asg_out, double* mapc, double* lo_out, double* hi_out, int* n_match_out) {
  CK(cudaSetDevice(cfg.device));
  if (!ref || n_ref < 1) return fail("bnmf_assign_signatures: a reference matrix (K x n_ref) is needed");
  if (!(ci > 0.0 && ci < 1.0)) return fail("bnmf_assign_signatures: credible_interval must be in (0, 1)");
  std::vector<int> match; std::string mode;
  if (map_slots(n_samples, match, mode)) return 1;
  const int N = cfg.N, K = cfg.K, nm = (int)match.size();
  std::vector<int> keep;
  for (int n = 0; n < N; ++n) if (mode[n] == '1') keep.push_back(n);
  const int nk = (int)keep.size();
  if (n_keep_out) *n_keep_out = nk;
  if (n_match_out) *n_match_out = nm;
  if (keep_out) for (int i = 0; i < nk; ++i) keep_out[i] = keep[i];
  if (nk == 0) return 0;
  if (nk > n_ref) return fail("bnmf_assign_signatures: %d included signatures but only %d reference signatures", nk, n_ref);
  // MAP signatures (mean of the renormalised matching samples), then all cosines on the device
  const long long KN = (long long)K * N;
  Scratch sc(stream, cfg.device);
  int* dslots; int* dkeep; double* colsum; double* dPmap; double* dref; double* dcos;
  CK(sc.get(&dslots, sizeof(int) * nm)); CK(sc.get(&dkeep, sizeof(int) * nk));
  CK(sc.get(&colsum, sizeof(double) * (size_t)nm * N)); CK(sc.get(&dPmap, sizeof(double) * KN));
  CK(sc.get(&dref, sizeof(double) * (size_t)K * n_ref));
  CK(sc.get(&dcos, sizeof(double) * (size_t)(nm + 1) * nk * n_ref));
  CK(cudaMemcpyAsync(dslots, match.data(), sizeof(int) * nm, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(dkeep, keep.data(), sizeof(int) * nk, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(dref, ref, sizeof(double) * (size_t)K * n_ref, cudaMemcpyHostToDevice, stream));
  k_map_colsum<T><<<dim3(N, nm), 128, 0, stream>>>(d, dslots, colsum);
  k_map_mean_P<T><<<blocks(KN, 128), 128, 0, stream>>>(d, dslots, nm, colsum, dPmap);
  k_assign_cos<T><<<dim3(nk, nm + 1), 128, sizeof(double) * K, stream>>>(d, dslots, nm, dkeep, nk, dref, n_ref, dPmap, dcos);
  std::vector<double> cosv((size_t)(nm + 1) * nk * n_ref);
  CK(cudaMemcpyAsync(cosv.data(), dcos, sizeof(double) * cosv.size(), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  CK(cudaGetLastError());
  // votes: every sample's Hungarian assignment votes with its cosine (R/postprocessing.R:277-301)
  std::vector<double> V((size_t)nk * n_ref, 0.0), cost((size_t)nk * n_ref);
  std::vector<int> a;
  for (int s = 0; s < nm; ++s) {
    const double* cs = cosv.data() + (size_t)s * nk * n_ref;
    for (size_t t = 0; t < cost.size(); ++t) cost[t] = -cs[t];
    hungarian(cost, nk, n_ref, a);
    for (int i = 0; i < nk; ++i) V[(size_t)i * n_ref + a[i]] += cs[(size_t)i * n_ref + a[i]];
  }
  std::vector<int> win(nk, 0);
  for (int i = 0; i < nk; ++i) {
    double tot = 0.0;
    for (int j = 0; j < n_ref; ++j) tot += V[(size_t)i * n_ref + j];
    int best = 0; double bp = -1.0;
    for (int j = 0; j < n_ref; ++j) {
      const double pv = V[(size_t)i * n_ref + j] / tot;
      if (votes) votes[(size_t)i + (size_t)N * j] = pv;
      if (pv > bp) { bp = pv; best = j; }          // which.max: the first maximum
    }
    win[i] = best;
    if (asg_out) asg_out[i] = best;
    if (mapc) mapc[i] = cosv[((size_t)nm * nk + i) * n_ref + best];
    // credible interval of the samples' cosines to the winner, quantile type 7
    std::vector<double> x(nm);
    for (int s = 0; s < nm; ++s) x[s] = cosv[((size_t)s * nk + i) * n_ref + best];
    std::sort(x.begin(), x.end());
    const double probs[2] = {(1.0 - ci) / 2.0, 1.0 - (1.0 - ci) / 2.0};
    for (int q = 0; q < 2; ++q) {
      const double hh = (double)(nm - 1) * probs[q];
      int j = (int)std::floor(hh); if (j > nm - 1) j = nm - 1;
      const double g = hh - (double)j;
      const double val = x[j] + g * (x[j + 1 < nm ? j + 1 : nm - 1] - x[j]);
      if (q == 0 && lo_out) lo_out[i] = val;
      if (q == 1 && hi_out) hi_out[i] = val;
    }
  }
  return 0;
}
