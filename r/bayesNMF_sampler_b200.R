# Patch of the R6 class `bayesNMF_sampler` (jennalandy/bayesNMF, R/bayesNMF_sampler.R) that routes the
# Gibbs hot path through libbnmf_b200.so via the .Call veneer r/rcall.c.  Everything else of the
# package (convergence.R, get_MAP_, summary(), plot(), postprocessing*.R) keeps reading
# self$params, self$samples and self$state$sample_metrics exactly as before.
#
# NOT RUN in this repository (no R in the build image): the same entry points are exercised from
# Python (bayesnmf_b200/_lib.py, bayesnmf_b200/sampler.py mirrors the control flow below).
#
# Usage on a machine with R and a B200:
#   R CMD SHLIB r/rcall.c -Iinclude -Lbayesnmf_b200 -lbnmf_b200 -o bayesNMFb200.so
#   dyn.load("bayesNMFb200.so"); source("r/bayesNMF_sampler_b200.R")
#   bayesNMF_sampler$set("private", "h", NULL)            # the handle (external pointer)
#   bayesNMF_sampler$set("private", "b200_initialize", b200_initialize)
#   bayesNMF_sampler$set("private", "b200_block", b200_block)
#   ... and the two call sites marked below replace R/bayesNMF_sampler.R:241-257 and :273-285 / :340-348.

.lik_id    <- c(poisson = 0L, normal = 1L)
.prior_id  <- c(truncnormal = 0L, exponential = 1L, gamma = 2L)
.method_id <- c(SBFI = 0L, BFI = 1L)
.have_bits <- c(P = 1L, E = 2L, A = 4L, Z = 8L, sigmasq = 16L)      # BNMF_HAVE_* of include/bnmf.h
# BNMF_HAVE_PRIOR_* of include/bnmf.h: one bit per prior-parameter matrix the user supplied
# (bits 0..4 = Mu, Sigmasq, Lambda, Alpha, Beta on the P side, bits 8..12 on the E side); a matrix
# whose bit is clear is drawn from its hyperprior, a set bit keeps it (NA columns are still drawn:
# init_prior_params_, R/sample_priors.R:15-141)
.have_prior_bits <- c(Mu_p = 1L, Sigmasq_p = 2L, Lambda_p = 4L, Alpha_p = 8L, Beta_p = 16L,
                      Mu_e = 256L, Sigmasq_e = 512L, Lambda_e = 1024L, Alpha_e = 2048L, Beta_e = 4096L)
.state_dims <- function(self, name) {
  K <- self$dims$K; N <- self$dims$N; G <- self$dims$G
  if (name %in% c("P", "Mu_p", "Sigmasq_p", "Lambda_p", "Alpha_p", "Beta_p", "P_acceptance_rate")) c(K, N)
  else if (name %in% c("E", "Mu_e", "Sigmasq_e", "Lambda_e", "Alpha_e", "Beta_e", "E_acceptance_rate")) c(N, G)
  else if (name == "A") c(1L, N) else if (name == "R") c(1L, 1L)
  else if (name %in% c("sigmasq", "Alpha", "Beta")) c(1L, G) else stop("unknown state ", name)
}

# replaces R/bayesNMF_sampler.R:241-257 (the prior draw that is iteration 1); called at the end of
# initialize(), after check_model() (:217), fill_hyperprior_params_ (:214) and the temperature
# schedule (:136) exist
b200_initialize <- function(self, private, init_params, init_prior_params, seed = 0, device = 0L) {
  cc <- self$specs$convergence_control
  private$h <- .Call("R_bnmf_create", self$data, self$dims$N, .lik_id[[self$specs$likelihood]],
                     .prior_id[[self$specs$prior]], self$specs$MH, self$specs$learning_rank,
                     .method_id[[if (self$specs$rank_method == "BFI") "BFI" else "SBFI"]], as.numeric(seed),
                     as.integer(cc$MAP_over), as.integer(device), NULL)
  # bnmf_create already installed the defaults of R/setup.R:123-181; user-supplied values override
  for (nm in names(self$hyperprior_params)) .Call("R_bnmf_set_hyper", private$h, nm, self$hyperprior_params[[nm]])
  .Call("R_bnmf_set_temps", private$h, as.numeric(self$temperature_schedule))
  for (nm in names(init_params))       .Call("R_bnmf_set_state", private$h, nm, init_params[[nm]])
  # Normal likelihood: the reference adds the scalars alpha / beta of the sigmasq prior to init_prior_params
  # (R/bayesNMF_sampler.R:222-230); they are hyperparameters of the handle, not state matrices
  for (nm in intersect(names(init_prior_params), c("alpha", "beta")))
    .Call("R_bnmf_set_hyper", private$h, nm, as.numeric(init_prior_params[[nm]]))
  prior_mats <- intersect(names(init_prior_params), names(.have_prior_bits))
  for (nm in prior_mats) .Call("R_bnmf_set_state", private$h, nm, init_prior_params[[nm]])  # NA columns are drawn
  have <- sum(.have_bits[intersect(names(init_params), names(.have_bits))])
  have_prior <- sum(.have_prior_bits[prior_mats])
  row1 <- .Call("R_bnmf_init", private$h, as.integer(have), as.integer(have_prior))
  b200_pull_state(self, private)
  private$record_sample()                                   # samples$P / E / A [[1]] from self$params
  b200_append_metrics(self, matrix(row1, nrow = 1))
  invisible(self)
}

# self$params / prior_params / acceptance_rates <- device state (what save_object, get_Mhat and
# compute_metrics_ read, SURVEY.md Appendix D)
b200_pull_state <- function(self, private) {
  get <- function(nm) { d <- .state_dims(self, nm); .Call("R_bnmf_get_state", private$h, nm, d[1], d[2]) }
  for (nm in c("P", "E", "A", if (self$specs$learning_rank) "R", if (self$specs$likelihood == "normal") "sigmasq"))
    self$params[[nm]] <- get(nm)
  for (nm in names(self$prior_params)) if (is.matrix(self$prior_params[[nm]])) self$prior_params[[nm]] <- get(nm)
  if (self$specs$MH) {
    self$acceptance_rates$P_acceptance_rate <- get("P_acceptance_rate")
    self$acceptance_rates$E_acceptance_rate <- get("E_acceptance_rate")
  }
}

.metric_cols <- c("iter", "RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "temp",
                  "mean_P_acceptance_rate", "mean_E_acceptance_rate")   # R/utils.R:435-452, :341-342
b200_append_metrics <- function(self, rows) {
  df <- as.data.frame(rows); names(df) <- .metric_cols
  if (!self$specs$MH) df <- df[, 1:9]
  self$state$sample_metrics <- rbind(self$state$sample_metrics, df)     # update_sample_metrics_, R/utils.R:339-348
}

# replaces the body of the while loop of run_gibbs_sampler() (R/bayesNMF_sampler.R:273-285, and
# :340-348 with converged = TRUE): one call advances to the next MAP / convergence check
b200_block <- function(self, private, converged = FALSE, done = 0L) {
  cc <- self$specs$convergence_control
  # warm-up: up to the next multiple of MAP_every, never past maxiters (:268-271, :288-296); post-warm-up
  # MH phase (:337): `done` of the post_warmup iterations have been made, state$iter is already past maxiters
  left <- if (converged) self$specs$post_warmup - done else cc$maxiters - self$state$iter
  n <- min(cc$MAP_every - self$state$iter %% cc$MAP_every, left)
  if (n <= 0) return(invisible(0L))
  out <- .Call("R_bnmf_step", private$h, as.integer(n), converged, self$dims$K, self$dims$N)
  b200_append_metrics(self, t(out[[1]]))
  if (isTRUE(self$specs$save_all_samples)) for (i in seq_len(n)) {      # samples$P[[iter]], samples$A[[iter]]
    self$samples$P[[self$state$iter + i]] <- out[[2]][, , i]
    self$samples$A[[self$state$iter + i]] <- matrix(out[[3]][, i], nrow = 1)
  }
  self$state$iter <- self$state$iter + n
  b200_pull_state(self, private)
  invisible(n)
}

# the whole of run_gibbs_sampler() in one call (convergence control behind the ABI); fills
# state$sample_metrics, state$MAP_metrics (loglikelihood, logposterior, n_params, BIC, rank, A counts,
# mean temperature per check), state$iter / converged / converged_iter / why, then the MAP
b200_run <- function(self, private) {
  cc <- self$specs$convergence_control
  metric_id <- c(logposterior = 0, loglikelihood = 1, BIC = 2)[[cc$metric]]
  out <- .Call("R_bnmf_run", private$h,
               as.numeric(c(cc$MAP_over, cc$MAP_every, cc$tol, cc$Ninarow_nochange, cc$Ninarow_nobest, cc$miniters, cc$maxiters, metric_id)),
               as.integer(if (self$specs$MH) self$specs$post_warmup else 0L))
  r <- out[[1]]
  b200_append_metrics(self, t(out[[2]][, seq_len(r[6]), drop = FALSE]))
  mm <- as.data.frame(t(out[[3]][, seq_len(r[7]), drop = FALSE]))
  names(mm) <- c("iter", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "MAP_A_counts", "mean_temp")
  self$state$MAP_metrics <- mm
  self$state$iter <- r[1]; self$state$converged <- r[2] == 1; self$state$converged_iter <- r[3]
  self$state$why <- c("no change", "no best", "max iters")[r[4]]
  b200_pull_state(self, private)
  invisible(self)
}

# get_MAP_ (R/utils.R:194-288) on the device ring: mode of A, renormalise, mean over the matching
# samples; samples$E never crosses PCIe unless credible intervals over E are wanted on the host
b200_get_MAP <- function(self, private, n_samples) {
  r <- .Call("R_bnmf_get_map", private$h, as.integer(n_samples), self$dims$K, self$dims$N, self$dims$G)
  list(P = r[[1]], E = r[[2]], A = matrix(r[[3]], nrow = 1), n_match = r[[4]])
}
# credible_intervals of get_MAP_ (R/utils.R:264-287): quantile(type 7) over the same samples, on the device
b200_credible_intervals <- function(self, private, n_samples, credible_interval = 0.95) {
  probs <- c(0.5 - credible_interval / 2, 0.5 + credible_interval / 2)
  r <- .Call("R_bnmf_get_ci", private$h, as.integer(n_samples), probs, self$dims$K, self$dims$N, self$dims$G)
  list(P = list(lower = r[[1]], upper = r[[2]]), E = list(lower = r[[3]], upper = r[[4]]))
}
# assign_signatures_ensemble_ (R/postprocessing.R:175-341) over the retained samples; returns the
# `assignments` and `votes` data frames the reference stores in self$reference_comparison
b200_assign_signatures <- function(self, private, reference_P, n_samples, credible_interval = 0.95) {
  r <- .Call("R_bnmf_assign_signatures", private$h, as.integer(n_samples), reference_P, credible_interval, self$dims$N)
  nk <- r[[7]]; i <- seq_len(nk)
  ref_names <- if (is.null(colnames(reference_P))) paste0("Ref", seq_len(ncol(reference_P))) else colnames(reference_P)
  assignments <- data.frame(sig_est = r[[1]][i], sig_ref = ref_names[r[[2]][i]], MAP_cosine = r[[4]][i],
                            lower_cosine = r[[5]][i], upper_cosine = r[[6]][i])
  v <- r[[3]][i, , drop = FALSE]
  w <- which(v > 0, arr.ind = TRUE)
  votes <- data.frame(sig_est = r[[1]][w[, 1]], sig_ref = ref_names[w[, 2]], prop_votes = v[w])
  votes <- votes[order(votes$sig_est, -votes$prop_votes), ]
  list(assignments = assignments, votes = votes)
}
b200_sample_E <- function(self, private, ago = 0L)
  .Call("R_bnmf_get_sample", private$h, "E", as.integer(ago), self$dims$N, self$dims$G)
