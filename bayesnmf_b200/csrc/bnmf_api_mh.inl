// Host orchestration of the sweep-based models (Normal likelihood, Poisson + MH) and of
// rank learning; kernels in bnmf_mh.cuh.  Included by bnmf_api.cu.

template <typename T> static __global__ void k_ring_copy(Dev<T> d) {
  // record_sample for P and E (R/bayesNMF_sampler.R:651-672); ring_A is written by k_metrics
  const long long KN = (long long)d.K * d.N, NG = (long long)d.N * d.G;
  const long long pos = d.ctrl->ring_pos;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < KN + NG; i += (long long)gridDim.x * blockDim.x) {
    if (i < KN) d.ring_P[pos * KN + i] = d.P[i];
    else d.ring_E[pos * NG + (i - KN)] = d.E[i - KN];
  }
}
template <typename T> static __global__ void k_a_reduce(Dev<T> d, int nblocks, double* out) {
  __shared__ double scratch[8];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) { a += d.apart[2 * (long long)i]; b += d.apart[2 * (long long)i + 1]; }
  const double l0 = block_sum<256>(a, scratch);
  const double l1 = block_sum<256>(b, scratch);
  if (threadIdx.x == 0) { out[0] = l0; out[1] = l1; }
}
template <typename T> int Sampler<T>::mh_setup() {
  const int K = cfg.K, N = cfg.N; const long long G = cfg.G;
  if (dalloc(&d.dvec, K) || dalloc(&d.prop, K) || dalloc(&d.nzE, 2 * N) || dalloc(&d.nzP, N) || dalloc(&d.paccpart, N) ||
      dalloc(&asum, 2)) return 1;
  sweep_model = cfg.MH || cfg.likelihood == BNMF_NORMAL;
  if (!(sweep_model || cfg.learning_rank)) return 0;
  // P sweep decomposition: KX mutation types x GY genome lanes per block, genomes in chunks
  p_kx = ((K + 31) / 32) * 32; if (p_kx > 128) p_kx = 128;
  p_gy = 256 / p_kx; if (p_gy < 1) p_gy = 1;
  p_ktiles = (K + p_kx - 1) / p_kx;
  long long chunks = (G + 63) / 64;                       // at least 64 genomes per block
  const long long cap = std::max(1, 2368 / p_ktiles);     // up to 16 blocks per SM: the passes are latency-bound
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  d.gchunk = (int)((G + chunks - 1) / chunks);
  d.n_gchunks = (int)((G + d.gchunk - 1) / d.gchunk);
  col_blocks = (int)((G + 7) / 8);
  if (dalloc(&d.ppart, (long long)d.n_gchunks * K * 2) || dalloc(&d.apart, 2LL * col_blocks)) return 1;
  // E sweep: lanes per genome by the column length, genome slots per block by the shared memory
  e_lpg = K <= 128 ? 8 : K <= 256 ? 16 : 32;
  if (G < 148LL * 32) e_lpg = 32;                              // few genomes: every one its own warp, the GPU has room
  {
    const int gpw = 32 / e_lpg;
    const size_t budget = 220 * 1024;
    const long long ecols = e_sweep_cols(K, cfg.likelihood == BNMF_NORMAL);
    // (the variates of the draws are staged only where that does not cost a resident warp: with long
    //  columns the sweep is bound by the reductions over K, not by the latency of a draw)
    const long long s0 = ((long long)(budget / sizeof(double)) - K) / (ecols + e_sweep_extra(N, 0));
    const long long s1 = ((long long)(budget / sizeof(double)) - K) / (ecols + e_sweep_extra(N, 1));
    if (s0 < 1) return fail("k_e_sweep: K = %d does not fit one genome column in shared memory", K);
    const long long wmax = s0 / gpw <= 16 ? 16 : 8;            // long columns: one block per SM with every warp that fits; short: 8-warp blocks
    const long long w0 = std::max<long long>(1, std::min<long long>(wmax, s0 / gpw)), w1 = std::min<long long>(wmax, s1 / gpw);
    e_stage = w1 >= w0 ? 1 : 0;
    e_wpb = (int)(e_stage ? w1 : w0);
    while (e_wpb > 1 && (G + (long long)e_wpb * gpw - 1) / ((long long)e_wpb * gpw) < 2 * 148) e_wpb >>= 1;   // small problems: more, smaller blocks
    if (s0 < gpw) { e_lpg = 32; e_wpb = (int)std::min<long long>(16, s0); e_stage = 0; }   // (not even one warp of short-column slots fits)
    e_slots = e_wpb * (32 / e_lpg);
    e_smem = ((size_t)K + (size_t)e_slots * (ecols + e_sweep_extra(N, e_stage))) * sizeof(double);
    CK(cudaFuncSetAttribute(k_e_sweep<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e_smem));
    CK(cudaFuncSetAttribute(k_e_sweep<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e_smem));
    CK(cudaFuncSetAttribute(k_e_sweep<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e_smem));
  }
  // Normal likelihood: the P sweep through two Gram matrices (k_gram_part ...); BNMF_GRAM=0 turns it off
  if (cfg.likelihood == BNMF_NORMAL && !(getenv("BNMF_GRAM") && atoi(getenv("BNMF_GRAM")) == 0)) {
    gram_chunks = (int)((G + GRAM_GC - 1) / GRAM_GC);
    const long long len = (long long)K * N + (long long)N * N;
    if (dalloc(&gram_part, (long long)gram_chunks * len) || dalloc(&gram_buf, len)) return 1;
  }
  // rank learner: one cooperative launch when all its column blocks fit on the device at once (BNMF_A_COOP=0: a launch per signature)
  if (cfg.learning_rank && !(getenv("BNMF_A_COOP") && atoi(getenv("BNMF_A_COOP")) == 0)) {
    int per_sm = 0, sms = 148, coop = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg.device);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_a_sweep<T>, 256, 0) == cudaSuccess)
      a_coop = (long long)col_blocks <= (long long)per_sm * sms;
    else cudaGetLastError();
  }
  // Normal likelihood: the E sweep in Gram-matrix form (BNMF_EGRAM=0: k_e_sweep), when P fits in shared memory
  if (gram_buf && !(getenv("BNMF_EGRAM") && atoi(getenv("BNMF_EGRAM")) == 0)) {
    // genomes per block: the whole shard in one wave of two blocks per SM when that fits the shared memory (a
    // genome's chain of N conditionals is latency; a second wave doubles it), else the most that fits
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
    int gb = (int)((G + 2LL * sms - 1) / (2LL * sms));
    gb = std::max(32, std::min(96, (gb + 3) & ~3));
    while (gb > 32 && e_gram_smem(K, N, gb) > (size_t)110 * 1024) gb -= 4;
    if (e_gram_smem(K, N, gb) <= (size_t)110 * 1024) {
      eg_gb = gb; eg_smem = e_gram_smem(K, N, gb);
      CK(cudaFuncSetAttribute(k_e_gram<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eg_smem));
    }
  }
  // Normal likelihood: Mhat on the tensor cores
  if (cfg.likelihood == BNMF_NORMAL && N <= TC_MAX_N && !(getenv("BNMF_TC") && atoi(getenv("BNMF_TC")) == 0)) {
    tc_smem = tc_smem_bytes<T>(N);
    CK(cudaFuncSetAttribute(k_mhat_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
    if (dalloc(&tc_Pd, (long long)tc_planes_bytes(K, N)) || dalloc(&tc_eP, K)) return 1;
  }
  // Row-resident P sweep: a cluster of CS blocks per mutation type keeps the row of M and Mhat in
  // shared memory.  CS is the cluster size that needs the fewest waves of clusters over the K rows
  // (ties: the smaller slice per block); BNMF_P_ROWS=0 keeps the pass-per-signature kernels.
  const char* pr_env = getenv("BNMF_P_ROWS");
  if (sweep_model && !(pr_env && atoi(pr_env) == 0)) {
    const bool normal = cfg.likelihood == BNMF_NORMAL;
    const long long EA = 16 / (long long)sizeof(T);
    pr_Gp = ((G + EA - 1) / EA) * EA;
    double best = 0.0;
    for (int cs = 1; cs <= 8; ++cs) {
      long long gs = (G + cs - 1) / cs;
      gs = ((gs + EA - 1) / EA) * EA;
      const int threads = gs >= 2048 ? 512 : 256;
      const size_t smem = p_rows_smem<T>((int)gs, threads, normal);
      if (smem > (size_t)220 * 1024) continue;
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3((unsigned)(cs * K)); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem; lc.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      int nclusters = 0;
      cudaError_t e1, e2;
      if (threads == 512) {
        e1 = cudaFuncSetAttribute(k_p_rows<T, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e2 = cudaOccupancyMaxActiveClusters(&nclusters, k_p_rows<T, 512>, &lc);
      } else {
        e1 = cudaFuncSetAttribute(k_p_rows<T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e2 = cudaOccupancyMaxActiveClusters(&nclusters, k_p_rows<T, 256>, &lc);
      }
      if (e1 != cudaSuccess || e2 != cudaSuccess || nclusters < 1) { cudaGetLastError(); continue; }
      const long long waves = (K + nclusters - 1) / nclusters;
      const double cost = (double)waves * (1.0 + (double)gs / 16384.0);   // fixed latency per signature + the slice
      if (getenv("BNMF_TRACE")) fprintf(stderr, "[k_p_rows] cs %d slice %lld threads %d smem %zu clusters %d waves %lld cost %.2f\n", cs, gs, threads, smem, nclusters, waves, cost);
      if (pr_cs == 0 || cost < best) { best = cost; pr_cs = cs; pr_gslice = (int)gs; pr_threads = threads; pr_smem = smem; }
    }
    if (pr_cs) {
      if (pr_threads == 512) CK(cudaFuncSetAttribute(k_p_rows<T, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pr_smem));
      else CK(cudaFuncSetAttribute(k_p_rows<T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pr_smem));
      if (dalloc(&Et, (long long)N * pr_Gp)) return 1;
    }
  }
  return 0;
}

// Mhat = P diag(A) E from scratch.  Normal likelihood: on the tensor cores (k_mhat_tc, bnmf_tc.cuh: tcgen05.mma on
// exact 8-bit digit planes, error bounded absolutely -- what the residuals M - Mhat of the Normal conditionals
// need); the Poisson models take log(Mhat) and divide by it cell by cell and keep the fp64 kernel.  BNMF_TC=0: fp64.
template <typename T> int Sampler<T>::mhat_rebuild() {
  const long long KG = (long long)cfg.K * cfg.G;
  if (tc_Pd) {
    const dim3 grid((unsigned)((cfg.G + TC_N - 1) / TC_N), (unsigned)((cfg.K + TC_M - 1) / TC_M));
    k_tc_prep_P<T><<<grid.y, 128, 0, stream>>>(d.P, d.A, tc_Pd, tc_eP, cfg.K, cfg.N); mark("k_tc_prep_P");
    k_mhat_tc<T><<<grid, 128, tc_smem, stream>>>(tc_Pd, tc_eP, d.E, d.Mhat, cfg.K, cfg.N, (long long)cfg.G); mark("k_mhat_tc");
    launches += 2;
    return 0;
  }
  k_mhat_full<T><<<blocks(KG, 256), 256, 0, stream>>>(d); mark("k_mhat_full"); ++launches;
  return 0;
}

// P sweep of the Normal likelihood: Gram matrices, the chain per mutation type, Mhat rebuilt
template <typename T> int Sampler<T>::p_gram_launch() {
  const int K = cfg.K, N = cfg.N;
  const long long len = (long long)K * N + (long long)N * N, KG = (long long)K * cfg.G;
  const dim3 grid((unsigned)gram_chunks, (unsigned)((K + GRAM_KT - 1) / GRAM_KT));
  k_gram_part<T><<<grid, GRAM_KT, (size_t)2 * N * GRAM_GC * sizeof(double), stream>>>(d, gram_part); mark("k_gram_part");
  k_gram_fold<<<blocks(len * 32, 256), 256, 0, stream>>>(gram_part, gram_buf, len, gram_chunks); mark("k_gram_fold");
  constexpr int PW = 4;
  k_p_gram<T, PW><<<(K + PW - 1) / PW, 32 * PW, (size_t)PW * N * (1 + 3 * P_PRE) * sizeof(double), stream>>>(d, gram_buf); mark("k_p_gram");
  launches += 3;
  return eg_smem ? 0 : mhat_rebuild();       // (the Gram-matrix E sweep reads no Mhat: it is rebuilt once, after that sweep)
}

template <typename T> int Sampler<T>::p_rows_launch() {
  const long long NG = (long long)cfg.N * cfg.G;
  k_transpose_E<T><<<blocks(NG, 256), 256, 0, stream>>>(d, Et, pr_Gp); ++launches; mark("k_transpose_E");
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(pr_cs * cfg.K)); lc.blockDim = dim3(pr_threads); lc.dynamicSmemBytes = pr_smem; lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = pr_cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  if (pr_threads == 512) CK(cudaLaunchKernelEx(&lc, k_p_rows<T, 512>, d, (const T*)Et, pr_Gp, pr_cs, pr_gslice));
  else CK(cudaLaunchKernelEx(&lc, k_p_rows<T, 256>, d, (const T*)Et, pr_Gp, pr_cs, pr_gslice));
  ++launches; mark("k_p_rows");
  return 0;
}

template <typename T> int Sampler<T>::rank_sweep_kernels(int* pending) {
  const int N = cfg.N;
  k_r<T><<<1, 96, 0, stream>>>(d); mark("k_r"); ++launches;
  if (world <= 1 && a_coop) {          // every column block resident at once: the N passes as one cooperative launch
    CK(cudaMemsetAsync(red_ticket + 1, 0, sizeof(unsigned), stream));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)col_blocks); lc.blockDim = dim3(256); lc.dynamicSmemBytes = 0; lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CK(cudaLaunchKernelEx(&lc, k_a_sweep<T>, d, red_ticket + 1));
    ++launches; mark("k_a_sweep");
    *pending = N - 1;
    return 0;
  }
  for (int n = 0; n < N; ++n) {
    if (world <= 1) {     // nothing to exchange: the last block of the pass reduces and draws
      k_a_pass<T, 1><<<col_blocks, 256, 0, stream>>>(d, n, n ? n - 1 : -1, red_ticket + 1); mark("k_a_pass"); ++launches;
      continue;
    }
    k_a_pass<T, 0><<<col_blocks, 256, 0, stream>>>(d, n, n ? n - 1 : -1, nullptr); mark("k_a_pass"); ++launches;
    k_a_reduce<T><<<1, 256, 0, stream>>>(d, col_blocks, asum); mark("k_a_reduce"); ++launches;
    if (allreduce_buf(asum, 2, NCCL_FLOAT64, NCCL_SUM)) return 1;
    k_a_draw<T, 256><<<1, 256, 0, stream>>>(d, n, asum); mark("k_a_draw"); ++launches;
  }
  *pending = N - 1;
  return 0;
}

// rank learning inside the Poisson / latent-count iteration (R/sample_params.R:67-74)
template <typename T> int Sampler<T>::rank_sweep() {
  const long long KG = (long long)cfg.K * cfg.G;
  k_mhat_full<T><<<blocks(KG, 256), 256, 0, stream>>>(d); mark("k_mhat_full"); ++launches;
  int pending = -1;
  return rank_sweep_kernels(&pending);   // the pending update is dropped: Mhat is rebuilt next iteration
}

template <typename T> int Sampler<T>::mh_iteration(int from_prior, uint32_t have) {
  const int K = cfg.K, N = cfg.N; const long long G = cfg.G;
  const long long KN = (long long)K * N, NG = (long long)N * G, KG = (long long)K * G;
  if (from_prior) {
    if (!(have & BNMF_HAVE_P)) { k_prior_fill<T><<<blocks(KN, 128), 128, 0, stream>>>(d, 0); mark("k_prior_fill"); ++launches; }
    if (!(have & BNMF_HAVE_E)) { k_prior_fill<T><<<blocks(NG, 128), 128, 0, stream>>>(d, 1); mark("k_prior_fill"); ++launches; }
    k_init_rank<T><<<1, 32, 0, stream>>>(d, (have & BNMF_HAVE_A) ? 1 : 0); mark("k_init_rank"); ++launches;
    CK(cudaMemsetAsync(d.nzP, 0, sizeof(int) * N, stream));
    CK(cudaMemsetAsync(d.nzE, 0, sizeof(int) * 2 * N, stream));
    k_nzflags<T><<<blocks(std::max(KN, NG), 256), 256, 0, stream>>>(d); mark("k_nzflags"); ++launches;
    if (cfg.MH) {   // matrix(nrow, ncol) is NA-filled (R/bayesNMF_sampler.R:236-237)
      k_fill<T><<<blocks(KN, 256), 256, 0, stream>>>(d.P_acc, KN, (T)NAN);
      k_fill<T><<<blocks(NG, 256), 256, 0, stream>>>(d.E_acc, NG, (T)NAN);
      launches += 2;
    }
    if (mhat_rebuild()) return 1;
    k_final<T><<<col_blocks, 256, 0, stream>>>(d, -1, (have & BNMF_HAVE_SIGMASQ) ? 1 : 0); mark("k_final"); ++launches;
  } else {
    // the E side's prior parameters (they read E only) are not needed before the E sweep: a parallel branch, on the
    // side stream, under the P sweep
    const bool fork_e = side != nullptr && !prof_on;
    if (fork_e) {
      CK(cudaEventRecord(ev_fork, stream));
      CK(cudaStreamWaitEvent(side, ev_fork, 0));
      k_hyper<T><<<blocks(NG, 128), 128, 0, side>>>(d, 1);
      CK(cudaEventRecord(ev_join, side));
    }
    k_hyper<T><<<blocks(KN, 128), 128, 0, stream>>>(d, 0); mark("k_hyper");
    if (!fork_e) { k_hyper<T><<<blocks(NG, 128), 128, 0, stream>>>(d, 1); mark("k_hyper"); }
    launches += 2;
    if (!gram_buf) { if (mhat_rebuild()) return 1; }   // (the Gram-matrix P sweep rebuilds Mhat after itself)
    const dim3 pgrid(d.n_gchunks, p_ktiles), pblock(p_kx, p_gy);
    const size_t psm = (size_t)p_kx * p_gy * 2 * sizeof(double);
    const int dblocks = (K + 3) / 4;       // k_p_draw / k_p_accept: a warp per mutation type
    if (gram_buf) { if (p_gram_launch()) return 1; }
    else if (pr_cs) { if (p_rows_launch()) return 1; }
    else for (int n = 0; n < N; ++n) {
      k_p_pass1<T><<<pgrid, pblock, psm, stream>>>(d, n, n ? n - 1 : -1); mark("k_p_pass1");
      k_p_draw<T><<<dblocks, 128, 0, stream>>>(d, n); mark("k_p_draw");
      launches += 2;
      if (cfg.MH && h_converged) {
        k_p_pass2<T><<<pgrid, pblock, psm, stream>>>(d, n); mark("k_p_pass2");
        k_p_accept<T><<<dblocks, 128, 0, stream>>>(d, n); mark("k_p_accept");
        launches += 2;
      }
    }
    if (fork_e) CK(cudaStreamWaitEvent(stream, ev_join, 0));
    if (gram_buf && eg_smem) {       // Normal likelihood: both sweeps through Gram matrices, Mhat from the tensor cores afterwards
      k_e_gram<T><<<(unsigned)((G + eg_gb - 1) / eg_gb), EG_T, eg_smem, stream>>>(d, eg_gb); mark("k_e_gram"); ++launches;
      if (mhat_rebuild()) return 1;
    } else {
      const unsigned eg = (unsigned)((G + e_slots - 1) / e_slots);
      const int np = (pr_cs || gram_buf) ? -1 : N - 1;
      if (e_lpg == 8) k_e_sweep<T, 8><<<eg, 32 * e_wpb, e_smem, stream>>>(d, np, e_stage);
      else if (e_lpg == 16) k_e_sweep<T, 16><<<eg, 32 * e_wpb, e_smem, stream>>>(d, np, e_stage);
      else k_e_sweep<T, 32><<<eg, 32 * e_wpb, e_smem, stream>>>(d, np, e_stage);
      mark("k_e_sweep");
      ++launches;
    }
    int pending = -1;
    if (cfg.learning_rank) { if (rank_sweep_kernels(&pending)) return 1; }
    k_final<T><<<col_blocks, 256, 0, stream>>>(d, pending, 0); mark("k_final"); ++launches;
  }
  k_pprior<T, 128><<<N, 128, 0, stream>>>(d); mark("k_pprior"); ++launches;
  if (d.ring_cap > 0) { k_ring_copy<T><<<std::min(1024, blocks(KN + NG, 256)), 256, 0, stream>>>(d); mark("k_ring_copy"); ++launches; }
  CK(cudaGetLastError());
  return 0;
}

template <typename T> int Sampler<T>::refresh_metrics_only() {
  return fail("bnmf_init_from_prior: a user-supplied Z cannot be honoured (the latent counts are never materialised; supply SP and SE instead)");
}

// Alpha_g = alpha, Beta_g = beta for every genome (R/sample_priors.R:133-140, the `%in%`
// there tests values, so the vectors are always rebuilt from the scalars; defaults 3, 3:
// R/bayesNMF_sampler.R:222-230)
template <typename T> int Sampler<T>::init_sigmasq_prior() {
  k_fill<T><<<blocks(cfg.G, 256), 256, 0, stream>>>(d.Alpha_g, cfg.G, (T)sig_alpha);
  k_fill<T><<<blocks(cfg.G, 256), 256, 0, stream>>>(d.Beta_g, cfg.G, (T)sig_beta);
  return 0;
}
