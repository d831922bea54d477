/* .Call veneer between the R6 class `bayesNMF_sampler` of jennalandy/bayesNMF and the C ABI of
 * libbnmf_b200.so (include/bnmf.h).  It replaces the four private-method calls of the Gibbs loop
 * (private$sample_prior_params / sample_params / record_sample / update_sample_metrics,
 * R/bayesNMF_sampler.R:273-285, :340-348) and the prior draw of initialize() (:232-257).
 *
 * Build on a machine that has R:   R CMD SHLIB rcall.c -I../include -L../bayesnmf_b200 -lbnmf_b200
 * R is not part of this repository's build image: tests/test_abi.py only checks that this file
 * compiles against a stub of R's C API (tests/stubs) and names every bnmf_* entry it calls.
 *
 * Conventions: R owns every SEXP; the library copies in and out before returning and never keeps a
 * host pointer; a non-zero status becomes an R condition after temporaries are unprotected; the
 * handle is an external pointer whose finalizer calls bnmf_destroy (finalize(),
 * R/bayesNMF_sampler.R:740-745).  All calls come from the main R thread. */
#include <stdint.h>
#include <string.h>
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "bnmf.h"

static void fin(SEXP p) {
  bnmf_handle* h = (bnmf_handle*)R_ExternalPtrAddr(p);
  if (h) { bnmf_destroy(h); R_ClearExternalPtr(p); }
}
static bnmf_handle* H(SEXP p) {
  bnmf_handle* h = (bnmf_handle*)R_ExternalPtrAddr(p);
  if (!h) Rf_error("bnmf: closed handle");
  return h;
}
#define CK(rc) do { if (rc) Rf_error("%s", bnmf_last_error()); } while (0)

/* check_model(), R/bayesNMF_sampler.R:623-645 */
SEXP R_bnmf_check_model(SEXP lik, SEXP prior, SEXP MH) {
  char msg[256];
  if (bnmf_check_model(Rf_asInteger(lik), Rf_asInteger(prior), Rf_asLogical(MH), msg, sizeof msg)) Rf_error("%s", msg);
  return R_NilValue;
}

/* the allocation part of initialize(), R/bayesNMF_sampler.R:140-158; data = self$data (K x G),
 * shard = c(G_total, g0) for genome-sharded runs or NULL */
SEXP R_bnmf_create(SEXP data, SEXP N, SEXP lik, SEXP prior, SEXP MH, SEXP learn, SEXP method, SEXP seed,
                   SEXP ring, SEXP device, SEXP shard) {
  bnmf_config c;
  memset(&c, 0, sizeof c);
  c.K = Rf_nrows(data); c.G = Rf_ncols(data); c.G_total = c.G; c.g0 = 0; c.N = Rf_asInteger(N);
  if (!Rf_isNull(shard)) { c.G_total = (int64_t)REAL(shard)[0]; c.g0 = (int64_t)REAL(shard)[1]; }
  c.likelihood = Rf_asInteger(lik); c.prior = Rf_asInteger(prior); c.MH = Rf_asLogical(MH);
  c.learning_rank = Rf_asLogical(learn); c.rank_method = Rf_asInteger(method);
  c.precision = BNMF_F64; c.device = Rf_asInteger(device); c.ring_cap = Rf_asInteger(ring);
  c.seed = (uint64_t)Rf_asReal(seed);
  SEXP d = PROTECT(Rf_coerceVector(data, REALSXP));          /* INTSXP counts -> double, column-major */
  bnmf_handle* h = NULL;
  int rc = bnmf_create(&c, REAL(d), &h);
  UNPROTECT(1);
  CK(rc);
  SEXP p = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, fin, TRUE);
  UNPROTECT(1);
  return p;
}
SEXP R_bnmf_destroy(SEXP p) { fin(p); return R_NilValue; }

/* fill_hyperprior_params_ results (R/setup.R:15-88): scalar or full matrix */
SEXP R_bnmf_set_hyper(SEXP p, SEXP name, SEXP v) {
  SEXP d = PROTECT(Rf_coerceVector(v, REALSXP));
  int64_t r = Rf_isMatrix(v) ? Rf_nrows(v) : 1, c = Rf_isMatrix(v) ? Rf_ncols(v) : XLENGTH(v);
  int rc = bnmf_set_hyper(H(p), CHAR(STRING_ELT(name, 0)), REAL(d), r, c);
  UNPROTECT(1);
  CK(rc);
  return R_NilValue;
}
/* init_params / init_prior_params supplied by the user, R/bayesNMF_sampler.R:232-233 */
SEXP R_bnmf_set_state(SEXP p, SEXP name, SEXP v) {
  SEXP d = PROTECT(Rf_coerceVector(v, REALSXP));
  int rc = bnmf_set_state(H(p), CHAR(STRING_ELT(name, 0)), REAL(d), (int64_t)XLENGTH(d));
  UNPROTECT(1);
  CK(rc);
  return R_NilValue;
}
/* self$params$P / E / A / R / sigmasq, prior_params, acceptance_rates after a block of iterations */
SEXP R_bnmf_get_state(SEXP p, SEXP name, SEXP nrow, SEXP ncol) {
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(nrow), Rf_asInteger(ncol)));
  int rc = bnmf_get_state(H(p), CHAR(STRING_ELT(name, 0)), REAL(out), (int64_t)XLENGTH(out));
  UNPROTECT(1);
  CK(rc);
  return out;
}
/* self$temperature_schedule, R/utils.R:307-332 */
SEXP R_bnmf_set_temps(SEXP p, SEXP t) {
  CK(bnmf_set_temperature_schedule(H(p), REAL(t), (int64_t)XLENGTH(t)));
  return R_NilValue;
}
/* sample_params(skip = names(init_params), from_prior = TRUE) + record_sample + metrics row 1,
 * R/bayesNMF_sampler.R:241-257 */
SEXP R_bnmf_init(SEXP p, SEXP have, SEXP have_prior) {
  SEXP row = PROTECT(Rf_allocVector(REALSXP, BNMF_MC_COLS));
  int rc = bnmf_init_from_prior(H(p), (uint32_t)Rf_asInteger(have), (uint32_t)Rf_asInteger(have_prior), REAL(row));
  UNPROTECT(1);
  CK(rc);
  return row;
}
/* the loop body :273-285 (converged = FALSE) or :340-348 (TRUE) for n iterations;
 * list(metrics = 11 x n (a column per iteration), P = K x N x n, A = N x n) */
SEXP R_bnmf_step(SEXP p, SEXP n, SEXP converged, SEXP K, SEXP N) {
  int ni = Rf_asInteger(n), k = Rf_asInteger(K), nn = Rf_asInteger(N);
  SEXP met = PROTECT(Rf_allocMatrix(REALSXP, BNMF_MC_COLS, ni));
  SEXP P = PROTECT(Rf_alloc3DArray(REALSXP, k, nn, ni));
  SEXP A = PROTECT(Rf_allocMatrix(REALSXP, nn, ni));
  int rc = bnmf_step(H(p), ni, Rf_asLogical(converged), REAL(met), REAL(P), REAL(A));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, met); SET_VECTOR_ELT(out, 1, P); SET_VECTOR_ELT(out, 2, A);
  UNPROTECT(4);
  CK(rc);
  return out;
}
/* run_gibbs_sampler with the convergence control behind the ABI (R/bayesNMF_sampler.R:265-408,
 * R/convergence.R:60-154): cc = c(MAP_over, MAP_every, tol, Ninarow_nochange, Ninarow_nobest, miniters,
 * maxiters, metric id); list(result = c(iter, converged, converged_iter, why, best_iter), metrics, MAP_metrics) */
SEXP R_bnmf_run(SEXP p, SEXP cc, SEXP post_warmup) {
  bnmf_convergence_control c;
  const double* v = REAL(cc);
  c.MAP_over = (int32_t)v[0]; c.MAP_every = (int32_t)v[1]; c.tol = v[2]; c.Ninarow_nochange = (int32_t)v[3];
  c.Ninarow_nobest = (int32_t)v[4]; c.miniters = (int32_t)v[5]; c.maxiters = (int32_t)v[6]; c.metric = (int32_t)v[7];
  const int pw = Rf_asInteger(post_warmup);
  const int rows = c.maxiters + pw + 8, checks = rows / (c.MAP_every > 0 ? c.MAP_every : 1) + 8;
  SEXP met = PROTECT(Rf_allocMatrix(REALSXP, BNMF_MC_COLS, rows));
  SEXP mm = PROTECT(Rf_allocMatrix(REALSXP, BNMF_MM_COLS, checks));
  bnmf_run_result r;
  memset(&r, 0, sizeof r);
  int rc = bnmf_run(H(p), &c, pw, REAL(met), rows, REAL(mm), checks, &r);
  SEXP res = PROTECT(Rf_allocVector(REALSXP, 7));
  REAL(res)[0] = r.iter; REAL(res)[1] = r.converged; REAL(res)[2] = r.converged_iter; REAL(res)[3] = r.why;
  REAL(res)[4] = r.best_iter; REAL(res)[5] = r.n_rows; REAL(res)[6] = r.n_checks;
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, res); SET_VECTOR_ELT(out, 1, met); SET_VECTOR_ELT(out, 2, mm);
  UNPROTECT(4);
  CK(rc);
  return out;
}
/* get_MAP_, R/utils.R:194-288, on the device ring: list(P, E, A, n_match) */
SEXP R_bnmf_get_map(SEXP p, SEXP n_samples, SEXP K, SEXP N, SEXP G) {
  int k = Rf_asInteger(K), nn = Rf_asInteger(N), g = Rf_asInteger(G);
  SEXP P = PROTECT(Rf_allocMatrix(REALSXP, k, nn));
  SEXP E = PROTECT(Rf_allocMatrix(REALSXP, nn, g));
  SEXP A = PROTECT(Rf_allocVector(REALSXP, nn));
  int32_t n_match = 0;
  int rc = bnmf_get_map(H(p), Rf_asInteger(n_samples), REAL(P), REAL(E), REAL(A), &n_match);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, P); SET_VECTOR_ELT(out, 1, E); SET_VECTOR_ELT(out, 2, A);
  SET_VECTOR_ELT(out, 3, Rf_ScalarInteger(n_match));
  UNPROTECT(4);
  CK(rc);
  return out;
}
/* credible_intervals of get_MAP_, R/utils.R:264-287, on the device ring:
 * list(P_lower, P_upper, E_lower, E_upper, n_match) */
SEXP R_bnmf_get_ci(SEXP p, SEXP n_samples, SEXP probs, SEXP K, SEXP N, SEXP G) {
  int k = Rf_asInteger(K), nn = Rf_asInteger(N), g = Rf_asInteger(G);
  SEXP Pl = PROTECT(Rf_allocMatrix(REALSXP, k, nn)), Ph = PROTECT(Rf_allocMatrix(REALSXP, k, nn));
  SEXP El = PROTECT(Rf_allocMatrix(REALSXP, nn, g)), Eh = PROTECT(Rf_allocMatrix(REALSXP, nn, g));
  int32_t n_match = 0;
  int rc = bnmf_get_credible_intervals(H(p), Rf_asInteger(n_samples), REAL(probs)[0], REAL(probs)[1],
                                       REAL(Pl), REAL(Ph), REAL(El), REAL(Eh), &n_match);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
  SET_VECTOR_ELT(out, 0, Pl); SET_VECTOR_ELT(out, 1, Ph); SET_VECTOR_ELT(out, 2, El); SET_VECTOR_ELT(out, 3, Eh);
  SET_VECTOR_ELT(out, 4, Rf_ScalarInteger(n_match));
  UNPROTECT(5);
  CK(rc);
  return out;
}
/* assign_signatures_ensemble_, R/postprocessing.R:175-341, over the retained samples:
 * list(keep_sigs (1-based), sig_ref (1-based column of reference_P), votes (n_keep x n_ref shares),
 *      MAP_cosine, lower_cosine, upper_cosine, n_match) */
SEXP R_bnmf_assign_signatures(SEXP p, SEXP n_samples, SEXP reference_P, SEXP credible_interval, SEXP N) {
  const int nn = Rf_asInteger(N), n_ref = Rf_ncols(reference_P);
  SEXP ref = PROTECT(Rf_coerceVector(reference_P, REALSXP));
  SEXP votes = PROTECT(Rf_allocMatrix(REALSXP, nn, n_ref));
  SEXP mc = PROTECT(Rf_allocVector(REALSXP, nn)), lo = PROTECT(Rf_allocVector(REALSXP, nn)), hi = PROTECT(Rf_allocVector(REALSXP, nn));
  SEXP keep = PROTECT(Rf_allocVector(REALSXP, nn)), sig = PROTECT(Rf_allocVector(REALSXP, nn));
  int32_t nk = 0, nm = 0, keep_i[64], asg_i[64];
  memset(REAL(votes), 0, sizeof(double) * (size_t)nn * n_ref);
  int rc = nn > 64 ? 1 : bnmf_assign_signatures(H(p), Rf_asInteger(n_samples), REAL(ref), n_ref, Rf_asReal(credible_interval), &nk,
                                                keep_i, REAL(votes), asg_i, REAL(mc), REAL(lo), REAL(hi), &nm);
  for (int i = 0; i < nn; ++i) { REAL(keep)[i] = i < nk ? keep_i[i] + 1 : 0; REAL(sig)[i] = i < nk ? asg_i[i] + 1 : 0; }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 8));
  SET_VECTOR_ELT(out, 0, keep); SET_VECTOR_ELT(out, 1, sig); SET_VECTOR_ELT(out, 2, votes); SET_VECTOR_ELT(out, 3, mc);
  SET_VECTOR_ELT(out, 4, lo); SET_VECTOR_ELT(out, 5, hi); SET_VECTOR_ELT(out, 6, Rf_ScalarInteger(nk)); SET_VECTOR_ELT(out, 7, Rf_ScalarInteger(nm));
  UNPROTECT(8);
  CK(rc);
  return out;
}
/* samples$E[[i]] on demand (update_list ring, R/helpers.R:111-119); ago = 0 is the newest */
SEXP R_bnmf_get_sample(SEXP p, SEXP name, SEXP ago, SEXP nrow, SEXP ncol) {
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(nrow), Rf_asInteger(ncol)));
  int rc = bnmf_get_sample(H(p), CHAR(STRING_ELT(name, 0)), Rf_asInteger(ago), REAL(out), (int64_t)XLENGTH(out));
  UNPROTECT(1);
  CK(rc);
  return out;
}
SEXP R_bnmf_ring_count(SEXP p) {
  int32_t n = 0;
  CK(bnmf_ring_count(H(p), &n));
  return Rf_ScalarInteger(n);
}
/* genome-sharded runs, one R worker per GPU: rank 0 makes the id (a raw vector of 128 bytes),
 * the workers exchange it by their own means (MPI, sockets) and every one joins */
SEXP R_bnmf_comm_unique_id(void) {
  SEXP id = PROTECT(Rf_allocVector(RAWSXP, 128));
  int rc = bnmf_comm_unique_id((char*)RAW(id));
  UNPROTECT(1);
  CK(rc);
  return id;
}
SEXP R_bnmf_comm_init(SEXP p, SEXP id, SEXP rank, SEXP world) {
  if (XLENGTH(id) != 128) Rf_error("bnmf: the communicator id has 128 bytes");
  CK(bnmf_comm_init(H(p), (const char*)RAW(id), Rf_asInteger(rank), Rf_asInteger(world)));
  return R_NilValue;
}
SEXP R_bnmf_comm_share(SEXP p, SEXP src) { CK(bnmf_comm_share(H(p), H(src))); return R_NilValue; }
/* device time of the last bnmf_step: c(total_ms, iter_ms, zstat_ms, launches) -> time$per_iter */
SEXP R_bnmf_timing(SEXP p) {
  double t = 0, i = 0, z = 0; int64_t l = 0;
  CK(bnmf_timing(H(p), &t, &i, &z, &l));
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4));
  REAL(out)[0] = t; REAL(out)[1] = i; REAL(out)[2] = z; REAL(out)[3] = (double)l;
  UNPROTECT(1);
  return out;
}

SEXP R_bnmf_release_cached_memory(void) { bnmf_release_cached_memory(); return R_NilValue; }

static const R_CallMethodDef calls[] = {
  {"R_bnmf_check_model", (DL_FUNC)&R_bnmf_check_model, 3}, {"R_bnmf_create", (DL_FUNC)&R_bnmf_create, 11},
  {"R_bnmf_destroy", (DL_FUNC)&R_bnmf_destroy, 1},         {"R_bnmf_set_hyper", (DL_FUNC)&R_bnmf_set_hyper, 3},
  {"R_bnmf_set_state", (DL_FUNC)&R_bnmf_set_state, 3},     {"R_bnmf_get_state", (DL_FUNC)&R_bnmf_get_state, 4},
  {"R_bnmf_set_temps", (DL_FUNC)&R_bnmf_set_temps, 2},     {"R_bnmf_init", (DL_FUNC)&R_bnmf_init, 3},
  {"R_bnmf_step", (DL_FUNC)&R_bnmf_step, 5},               {"R_bnmf_get_map", (DL_FUNC)&R_bnmf_get_map, 5},
  {"R_bnmf_get_ci", (DL_FUNC)&R_bnmf_get_ci, 6},           {"R_bnmf_run", (DL_FUNC)&R_bnmf_run, 3},
  {"R_bnmf_assign_signatures", (DL_FUNC)&R_bnmf_assign_signatures, 5},
  {"R_bnmf_get_sample", (DL_FUNC)&R_bnmf_get_sample, 5},   {"R_bnmf_ring_count", (DL_FUNC)&R_bnmf_ring_count, 1},
  {"R_bnmf_comm_unique_id", (DL_FUNC)&R_bnmf_comm_unique_id, 0}, {"R_bnmf_comm_init", (DL_FUNC)&R_bnmf_comm_init, 4},
  {"R_bnmf_comm_share", (DL_FUNC)&R_bnmf_comm_share, 2},   {"R_bnmf_timing", (DL_FUNC)&R_bnmf_timing, 1},
  {"R_bnmf_release_cached_memory", (DL_FUNC)&R_bnmf_release_cached_memory, 0},
  {NULL, NULL, 0}};
void R_init_bayesNMFb200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
