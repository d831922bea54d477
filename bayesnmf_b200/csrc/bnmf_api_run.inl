// bnmf_run: the driver loop of run_gibbs_sampler (R/bayesNMF_sampler.R:265-408) with
// check_convergence_ (R/convergence.R:60-154) and the part of update_MAP_metrics_ (R/utils.R:356-397)
// that the convergence rules read, on the host side of the library.  Included by bnmf_api.cu.

template <typename T> int Sampler<T>::run(const bnmf_convergence_control* cc, int post_warmup, double* metrics_out,
                                          int64_t rows_cap, double* map_out, int64_t checks_cap, bnmf_run_result* res) {
  CK(cudaSetDevice(cfg.device));
  if (!cc || !res) return fail("bnmf_run: convergence control and result must not be NULL");
  if (cc->MAP_over < 1 || cc->MAP_every < 1 || cc->maxiters < 1) return fail("bnmf_run: MAP_over, MAP_every, maxiters must be >= 1");
  if (cc->metric < BNMF_METRIC_LOGPOSTERIOR || cc->metric > BNMF_METRIC_BIC)
    return fail("bnmf_run: metric must be logposterior, loglikelihood or BIC (RMSE / KL of the MAP: use bnmf_step + bnmf_get_map)");
  if (d.ring_cap < cc->MAP_over) return fail("bnmf_run: ring_cap = %d < MAP_over = %d", d.ring_cap, cc->MAP_over);
  if (h_rows.empty()) return fail("bnmf_run: call bnmf_init_from_prior first");
  if (post_warmup < 0) post_warmup = 0;
  Ctrl hc; CK(cudaMemcpyAsync(&hc, d.ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  int iter = hc.iter;
  memset(res, 0, sizeof(*res));
  // state$prev_MAP_metric / best_MAP_metric / inarow_* / converged live in the sampler, as in the reference
  // (self$state, R/convergence.R:84-141): a second call on the same handle (resume, raised maxiters) carries on
  // from the last check instead of starting at the 'first check' branch; bnmf_init_from_prior resets them
  bool& have_prev = rs.have_prev;
  double& prev = rs.prev; double& best = rs.best;
  int& inarow_no_change = rs.inarow_no_change; int& inarow_no_best = rs.inarow_no_best; int& inarow_na = rs.inarow_na;
  int& best_iter = rs.best_iter;
  bool& converged = rs.converged; int& why = rs.why; int& converged_iter = rs.converged_iter;
  int64_t n_rows = 0, n_checks = 0;
  std::vector<double> buf;
  const double logG = std::log((double)cfg.G_total);

  auto advance = [&](int n, int conv) -> int {
    buf.resize((size_t)n * MC_COLS);
    if (step(n, conv, buf.data(), nullptr, nullptr)) return 1;
    if (metrics_out) {
      const int64_t room = rows_cap - n_rows;
      const int64_t take = room < n ? (room < 0 ? 0 : room) : n;
      if (take > 0) memcpy(metrics_out + n_rows * MC_COLS, buf.data(), sizeof(double) * MC_COLS * (size_t)take);
    }
    n_rows += n;
    iter += n;
    return 0;
  };
  // window mean over the rows whose iteration lies in (iter - MAP_over, iter]  (R/utils.R:372-379)
  auto window_mean = [&](int col) -> double {
    const long long rows = (long long)(h_rows.size() / MC_COLS);
    double s = 0.0; long long c = 0;
    for (long long r = rows - 1; r >= 0; --r) {
      const double it = h_rows[(size_t)r * MC_COLS + MC_ITER];
      if (it <= (double)(iter - cc->MAP_over)) break;
      if (it <= (double)iter) { s += h_rows[(size_t)r * MC_COLS + col]; ++c; }
    }
    return c ? s / (double)c : NAN;
  };
  auto check = [&]() -> int {
    // get_MAP: the modal A of the newest min(MAP_over, held) samples (rank, A_counts)
    Ctrl h2; CK(cudaMemcpyAsync(&h2, d.ctrl, sizeof(h2), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    std::vector<int> match; std::string mode;
    if (map_slots(std::min(cc->MAP_over, h2.ring_count), match, mode)) return 1;
    int rank = 0; for (char ch : mode) rank += ch == '1';
    const double loglik = window_mean(MC_LOGLIK), logpost = window_mean(MC_LOGPOST);
    const double n_params = (double)rank * ((double)cfg.G_total + (double)cfg.K);
    const double bic = -2.0 * loglik + n_params * logG;
    double mt = 0.0; int lo0 = std::max(iter - cc->MAP_over, 0), cnt = 0;
    for (int i = lo0; i < iter; ++i) { mt += i < (int)h_temps.size() ? h_temps[i] : 1.0; ++cnt; }
    if (map_out && n_checks < checks_cap) {
      double* m = map_out + n_checks * BNMF_MM_COLS;
      m[BNMF_MM_ITER] = iter; m[BNMF_MM_LOGLIK] = loglik; m[BNMF_MM_LOGPOST] = logpost; m[BNMF_MM_NPARAMS] = n_params;
      m[BNMF_MM_BIC] = bic; m[BNMF_MM_RANK] = rank; m[BNMF_MM_A_COUNTS] = (double)match.size();
      m[BNMF_MM_MEAN_TEMP] = cnt ? mt / cnt : NAN;
    }
    ++n_checks;
    // check_convergence_, R/convergence.R:66-141
    double metric = cc->metric == BNMF_METRIC_LOGPOSTERIOR ? -logpost : cc->metric == BNMF_METRIC_LOGLIKELIHOOD ? -loglik : bic;
    if (!have_prev) { prev = metric + 1.0; best = metric + 1.0; have_prev = true; }
    const double pc = (metric - prev) / prev;
    prev = metric;
    if (pc != pc) { inarow_no_change = 0; inarow_no_best = 0; ++inarow_na; }
    else if (std::fabs(pc) < cc->tol) { ++inarow_no_change; inarow_na = 0; }
    else { inarow_no_change = 0; inarow_na = 0; }
    bool all_one = true;                       // R's 1-based window [iter - MAP_over, iter]
    for (int i = std::max(iter - cc->MAP_over, 1); i <= iter && all_one; ++i)
      if (i - 1 < (int)h_temps.size() && h_temps[i - 1] != 1.0) all_one = false;
    if (all_one && iter >= cc->miniters) {
      if (metric < best) { best = metric; best_iter = iter; inarow_no_best = 0; } else ++inarow_no_best;
      if (inarow_no_change >= cc->Ninarow_nochange) { converged = true; why = BNMF_WHY_NO_CHANGE; }
      else if (inarow_no_best >= cc->Ninarow_nobest) { converged = true; why = BNMF_WHY_NO_BEST; }
      else if (iter >= cc->maxiters) { converged = true; why = BNMF_WHY_MAX_ITERS; }
    }
    return 0;
  };

  while (!converged && iter < cc->maxiters) {
    const int n = std::min(cc->MAP_every - iter % cc->MAP_every, cc->maxiters - iter);
    if (advance(n, 0)) return 1;
    if ((iter % cc->MAP_every == 0 && iter >= std::max(cc->MAP_over, cc->MAP_every)) || iter >= cc->maxiters) {
      if (check()) return 1;
      if (converged && !converged_iter) converged_iter = iter;
    }
  }
  if (cfg.MH) {                                 // R/bayesNMF_sampler.R:337-348
    int& done = rs.post_done;                   // (a handle that has made its post-warm-up iterations makes no more)
    while (done < post_warmup) {
      const int n = std::min(cc->MAP_every - iter % cc->MAP_every, post_warmup - done);
      if (advance(n, 1)) return 1;
      done += n;
      if (check()) return 1;
    }
  }
  res->iter = iter; res->converged = converged ? 1 : 0; res->converged_iter = converged_iter; res->why = why;
  res->best_iter = best_iter; res->n_checks = (int32_t)n_checks; res->n_rows = (int32_t)n_rows;
  res->inarow_no_change = inarow_no_change; res->inarow_no_best = inarow_no_best; res->inarow_na = inarow_na;
  res->best_MAP_metric = best; res->prev_MAP_metric = prev;
  return 0;
}
