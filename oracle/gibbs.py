"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

fp64 numpy restatement of the Gibbs hot path of jennalandy/bayesNMF, function by
function, following the R sources under /root/reference (cited per function).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker / the reported CPU baseline.

Parity status: **parity unpinned**.  The reference ships no tests, golden vectors or
fixtures for this path and R is not installed here, so the
restatement is pinned to (a) the formulas in the R sources, deterministic given the
draws, (b) the bundled example_data.rds / COSMIC csv (tests/golden) for an end-to-end
Monte-Carlo-level known answer, and (c) scipy CDFs for every random-variate
generator.  Random draws come from the Philox streams of oracle/philox.py (the R code
uses the global Mersenne-Twister stream), so equality with R is distributional, while
equality with the CUDA path is draw-for-draw.

State layout follows the reference: data K x G, P K x N, E N x G, A length N
(the 1 x N matrix of R/sample_params.R:36), column-major semantics are irrelevant in
numpy -- cell ids used for RNG addressing are k + K*n (P side), n + N*g (E side),
k + K*g (Z), with g the GLOBAL genome index (g0 + local).

Reference quirks that are reproduced on purpose:
  1. rnorm(., num/denom, 1/denom): posterior variance used as sd (R/sample_priors.R:219,235)
  2. sample_Sigmasq_En uses A_e as the base of the rate            (R/sample_priors.R:267)
  3. hyper S is a variance in the Mu update, sqrt(S) is the sd at init (R/sample_priors.R:37,215-218)
  4. MH proposal variance = unclipped Mhat; accept ratio uses pmax(Mhat, 1)   (R/sample_Pn.R:137-139,221,228)
  5. every proposal accepted until state$converged                 (R/sample_Pn.R:201-204)
  8. log prior counts excluded signatures; no A, R, sigmasq terms    (R/utils.R:131-182)
 13. Z = 0 for excluded signatures                                  (R/sample_params.R:254-261)
 14. P and E use the Z of the previous iteration                    (R/sample_params.R:54-85)
"""
import numpy as np
from scipy.special import gammaln, log_ndtr, xlogy

from . import draws as dr
from . import philox as px

LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)


# --------------------------------------------------------------------------------------
# defaults and setup  (R/setup.R)
# --------------------------------------------------------------------------------------
def default_hyperprior_params(prior, mean_data, N):
    """get_default_*_hyperprior_params_  (R/setup.R:123-181)."""
    if prior == "truncnormal":
        v = dict(m=0.0, s=np.sqrt(mean_data / N), a=N + 1.0, b=np.sqrt(N))
        return {f"{k}_{e}": x for e in ("p", "e") for k, x in v.items()}
    if prior == "exponential":
        v = dict(a=10.0 * np.sqrt(N), b=10.0 * np.sqrt(mean_data))
        return {f"{k}_{e}": x for e in ("p", "e") for k, x in v.items()}
    if prior == "gamma":
        v = dict(a=10.0 * np.sqrt(N), b=10.0, c=10.0 * np.sqrt(mean_data), d=10.0)
        return {f"{k}_{e}": x for e in ("p", "e") for k, x in v.items()}
    raise ValueError(prior)


def check_model(likelihood, prior, MH):
    """check_model  (R/bayesNMF_sampler.R:623-645)."""
    if likelihood not in ("normal", "poisson"):
        raise ValueError("likelihood must be one of normal, poisson")
    if likelihood == "normal":
        if prior not in ("truncnormal", "exponential"):
            raise ValueError("prior must be one of c('truncnormal','exponential') with `likelihood = 'normal'`")
    else:
        if prior not in ("gamma", "exponential", "truncnormal"):
            raise ValueError("prior must be one of c('gamma','exponential','truncnormal') with `likelihood = 'poisson'`")
        if prior == "gamma" and MH:
            raise ValueError("gamma prior cannot be used in a MH-within-gibbs sampler")
        if prior == "truncnormal" and not MH:
            raise ValueError("truncnormal prior can only be used in a MH-within-gibbs sampler")


def get_temp_sched(length, n_temp, rng=None):
    """get_temp_sched_  (R/utils.R:307-332).  The random sub-sample branch consumes the
    host RNG (`sort(sample(temp_sched, n_temp))`); numpy's Generator stands in for R's."""
    nX = max(int(round(n_temp / 374.0)), 1)
    sched = [0.0] * nX
    for x in range(9, 4, -1):
        sched += [10.0 ** (-x)] * nX
    sched += [10.0 ** (-4)] * int(round(8 * nX))
    for y in range(4, 0, -1):
        for xi in range(90):
            x = xi * 0.1
            sched += [(1 + x) * 10.0 ** (-y)] * nX
    sched = np.asarray(sched)
    if len(sched) > n_temp:
        rng = rng or np.random.default_rng(0)
        sched = np.sort(rng.choice(sched, size=n_temp, replace=False))
    return np.concatenate([sched, np.ones(max(length - len(sched), 0))])


# --------------------------------------------------------------------------------------
# latent counts  (R/sample_params.R:253-265)
# --------------------------------------------------------------------------------------
def z_cdf(P, A, E, f32=False):
    """Running CDF of p_n = P[k,n] A_n E[n,g], accumulated sequentially over n (in float32
    arithmetic, one rounding per multiply and per add, for the float state of BNMF_F32)."""
    dt = np.float32 if f32 else np.float64
    Pa = np.where(np.asarray(A).reshape(1, -1) != 0, P, 0.0).astype(dt)
    prob = Pa[:, :, None] * np.asarray(E).astype(dt)[None, :, :]     # K x N x G
    return np.add.accumulate(prob, axis=1, dtype=dt)                # sequential adds, no pairwise


def z_thresholds(cdf, N, f32=False):
    """32-bit integer pick thresholds of every cell: thr_n = sat_u32(floor(cdf_n * (2^32 / cdf_{N-1})))
    for n < N-1 (NaN -> 0, as the float -> unsigned conversion of the kernel does), evaluated in
    the state precision with one rounding per operation.  P(pick <= n) = thr_n * 2^-32."""
    dt = np.float32 if f32 else np.float64
    total = cdf[:, -1, :]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        scale = dt(4294967296.0) / total                       # IEEE division
        x = (cdf[:, : N - 1, :] * scale[:, None, :]).astype(dt)
    x = np.where(np.isnan(x), 0.0, x)
    return np.clip(np.floor(x).astype(np.float64), 0.0, 4294967295.0).astype(np.uint64)   # K x (N-1) x G


def sample_Z_stats(M, P, A, E, seed, it, g0=0, bits=32, return_Z=False, chunk=2_000_000, f32=False):
    """sample_Zkg for every cell, reduced to SP = sum_g Z (K x N), SE = sum_k Z (N x G).

    Reference: probs_n = P[k,n]*A[1,n]*E[n,g]; all-zero -> Z = 0; else
    rmultinom(1, size = M[k,g], prob = probs/sum(probs))  (R/sample_params.R:253-265).
    rmultinom is a chain of conditional binomials; a multinomial is equally the
    histogram of M[k,g] independent categorical draws, which is what is used here: the
    inverse of the running CDF (accumulated in the state precision) at the 32-bit uniform
    w of the pick, pick = #{n < N-1 : thr_n <= min(w, 2^32 - 2)} with the integer thresholds
    of z_thresholds (the arithmetic contract of k_zstat, csrc/bnmf_poisson.cuh).
    `bits` is accepted for symmetry with the draw functions and not used."""
    M = np.asarray(M)
    K, G = M.shape
    N = P.shape[1]
    cdf = z_cdf(P, A, E, f32)                      # K x N x G
    total = cdf[:, -1, :]                          # K x G  == Mhat
    thr = z_thresholds(cdf, N, f32)                # K x (N-1) x G
    Mi = M.astype(np.int64)
    work = (Mi > 0) & (total > 0.0)
    kk, gg = np.nonzero(work)
    cnt = Mi[kk, gg]
    SP = np.zeros((K, N), dtype=np.int64)
    SE = np.zeros((N, G), dtype=np.int64)
    Z = np.zeros((K, N, G), dtype=np.int64) if return_Z else None
    if len(kk) == 0:
        return (SP, SE, Z) if return_Z else (SP, SE)
    cell_id = kk.astype(np.uint64) + np.uint64(K) * (gg.astype(np.uint64) + np.uint64(g0))
    # expand to picks, chunked to bound memory
    starts = np.concatenate([[0], np.cumsum(cnt)])
    c0 = 0
    ncell = len(kk)
    while c0 < ncell:
        c1 = int(np.searchsorted(starts, starts[c0] + chunk, side="right")) - 1
        c1 = max(c1, c0 + 1)
        c1 = min(c1, ncell)
        cs = cnt[c0:c1]
        owner = np.repeat(np.arange(c0, c1), cs)
        j = np.arange(len(owner)) - np.repeat(starts[c0:c1] - starts[c0], cs)
        w = px.words(seed, it, px.PUR_Z, cell_id[owner], j >> 2)
        W = np.stack(w, axis=0)
        word = W[j & 3, np.arange(len(owner))].astype(np.uint64)
        word = np.minimum(word, np.uint64(0xFFFFFFFE))
        th = thr[kk[owner], :, gg[owner]]          # picks x (N-1)
        pick = (th <= word[:, None]).sum(axis=1)
        np.add.at(SP, (kk[owner], pick), 1)
        np.add.at(SE, (pick, gg[owner]), 1)
        if return_Z:
            np.add.at(Z, (kk[owner], pick, gg[owner]), 1)
        c0 = c1
    return (SP, SE, Z) if return_Z else (SP, SE)


# --------------------------------------------------------------------------------------
# densities
# --------------------------------------------------------------------------------------
def dpois_log(x, lam):
    """stats::dpois(x, lambda, log = TRUE); direct form x log(lam) - lam - lgamma(x+1)
    (R's saddle-point dpois_raw agrees to ~1e-13 relative)."""
    return xlogy(x, lam) - lam - gammaln(x + 1.0)


def dnorm_log(x, mean, var):
    return -LOG_SQRT_2PI - 0.5 * np.log(var) - 0.5 * (x - mean) ** 2 / var


def dtruncnorm0_log(x, mean, sd):
    """log(truncnorm::dtruncnorm(x, a = 0, b = Inf, mean, sd)), in log space
    (R/utils.R:134-145 takes log() of the density and can underflow to -Inf)."""
    z = (x - mean) / sd
    return -LOG_SQRT_2PI - np.log(sd) - 0.5 * z * z - log_ndtr(mean / sd)


def dgamma_log(x, shape, rate):
    return shape * np.log(rate) - gammaln(shape) + (shape - 1.0) * np.log(x) - rate * x


def dexp_log(x, rate):
    return np.log(rate) - rate * x


# --------------------------------------------------------------------------------------
# MAP over retained samples  (R/utils.R:194-288, R/helpers.R:35-79)
# --------------------------------------------------------------------------------------
def get_mode(A_list):
    """get_mode (R/helpers.R:63-79): most frequent pattern; table() orders the pattern strings
    alphabetically and the stable decreasing sort keeps that order among ties."""
    keys = ["".join("1" if a else "0" for a in np.asarray(A).reshape(-1)) for A in A_list]
    counts = {}
    for k in keys:
        counts[k] = counts.get(k, 0) + 1
    mode = sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))[0][0]
    idx = [i for i, k in enumerate(keys) if k == mode]
    return np.array([1.0 if c == "1" else 0.0 for c in mode]), idx


def renormalize(P, E):
    """renormalize (R/helpers.R:35-49): columns of P sum to 1, product P E unchanged."""
    cs = P.sum(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        return P / cs[None, :], E * cs[:, None]


def get_MAP(P_list, E_list, A_list):
    """get_MAP_ with final = FALSE (R/utils.R:194-261): samples oldest first; returns
    (P_MAP, E_MAP, A_MAP, idx of the samples averaged)."""
    A_map, idx = get_mode(A_list)
    acc_P = acc_E = None
    for i in idx:
        P, E = renormalize(np.asarray(P_list[i], dtype=np.float64), np.asarray(E_list[i], dtype=np.float64))
        acc_P = P if acc_P is None else acc_P + P
        acc_E = E if acc_E is None else acc_E + E
    return acc_P / len(idx), acc_E / len(idx), A_map, idx


class _Store(dict):
    """State container that rounds every stored float array to float32 when emulating the
    BNMF_F32 state precision (the kernels compute in double and store T)."""

    def __init__(self, f32, *a, **kw):
        self.f32 = f32
        super().__init__()
        for k, v in dict(*a, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if self.f32 and isinstance(v, np.ndarray) and v.dtype == np.float64:
            v = v.astype(np.float32).astype(np.float64)
        super().__setitem__(k, v)


class OracleSampler:
    """State + one-iteration update of bayesNMF_sampler (R/bayesNMF_sampler.R:8-747),
    restricted to what the hot path touches."""

    def __init__(self, data, N, likelihood="poisson", prior="truncnormal", MH=None,
                 learning_rank=False, rank_method="SBFI", seed=0,
                 hyperprior_params=None, init_prior_params=None, init_params=None,
                 temperature_schedule=None, g0=0, G_total=None, mean_data=None, bits=32, reduce_fn=None,
                 state="f64"):
        self.data = np.asarray(data, dtype=np.float64)
        self.K, self.G = self.data.shape
        self.N = int(N)
        self.likelihood, self.prior = likelihood, prior
        self.MH = (likelihood == "poisson" and prior in ("truncnormal", "exponential")) if MH is None else bool(MH)
        check_model(likelihood, prior, self.MH)
        self.learning_rank, self.rank_method = bool(learning_rank), rank_method
        self.seed, self.bits = int(seed), bits
        self.f32 = state == "f32"          # BNMF_F32: state stored as float, arithmetic in double
        self._r = (lambda x: np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)) if self.f32 else (lambda x: x)
        self.g0 = int(g0)
        self.G_total = int(G_total) if G_total is not None else self.G
        # genome-sharded runs (Poisson latent-count models): `reduce_fn(array) -> array` sums an
        # array over the shards, at exactly the points where the CUDA path calls NCCL
        self.reduce = reduce_fn if reduce_fn is not None else (lambda a: a)
        self.iter = 1                      # state$iter  (R/bayesNMF_sampler.R:39-43)
        self.converged = False
        self.temperature_schedule = (np.ones(100000) if temperature_schedule is None
                                     else np.asarray(temperature_schedule, dtype=np.float64))
        mean_data = float(self.data.mean()) if mean_data is None else float(mean_data)
        # fill_hyperprior_params_  (R/setup.R:15-88): scalars broadcast to matrices
        hp = default_hyperprior_params(prior, mean_data, self.N)
        hp.update(hyperprior_params or {})
        self.hyper = _Store(self.f32)
        names = {"truncnormal": "MSAB", "exponential": "AB", "gamma": "ABCD"}[prior]
        for nm in names:
            for end, shp in (("_p", (self.K, self.N)), ("_e", (self.N, self.G))):
                full = nm + end
                if full in hp:
                    self.hyper[full] = np.asarray(hp[full], dtype=np.float64).reshape(shp)
                else:
                    self.hyper[full] = np.full(shp, float(hp[full.lower()]))
        self.prior_params = _Store(self.f32, {k: np.array(v, dtype=np.float64) for k, v in (init_prior_params or {}).items()})
        self.params = _Store(self.f32, {k: np.array(v, dtype=np.float64) for k, v in (init_params or {}).items()})
        self.acc = {"P": np.full((self.K, self.N), np.nan), "E": np.full((self.N, self.G), np.nan)}
        self.SP = np.zeros((self.K, self.N), dtype=np.int64)
        self.SE = np.zeros((self.N, self.G), dtype=np.int64)
        self.metrics = []
        self._initialize()

    # -- cell ids ---------------------------------------------------------------------
    def _cells_p(self, n=None):
        k = np.arange(self.K, dtype=np.uint64)
        if n is None:
            return k[:, None] + np.uint64(self.K) * np.arange(self.N, dtype=np.uint64)[None, :]
        return k + np.uint64(self.K) * np.uint64(n)

    def _cells_e(self, n=None):
        g = np.arange(self.G, dtype=np.uint64) + np.uint64(self.g0)
        if n is None:
            return np.arange(self.N, dtype=np.uint64)[:, None] + np.uint64(self.N) * g[None, :]
        return np.uint64(n) + np.uint64(self.N) * g

    # -- initialisation ---------------------------------------------------------------
    def _initialize(self):
        """initialize(): init_prior_params -> init_params -> sample_params(from_prior)
        -> record -> metrics  (R/bayesNMF_sampler.R:232-257)."""
        if self.likelihood == "normal":
            self.prior_params.setdefault("alpha", np.float64(3.0))
            self.prior_params.setdefault("beta", np.float64(3.0))
        self.init_prior_params_()
        skip = set(self.params.keys())
        self.init_params_()
        self.sample_params_(skip=skip, from_prior=True)
        self.update_sample_metrics_()

    def init_prior_params_(self):
        """R/sample_priors.R:15-141: draw every prior-parameter column that is missing or
        contains NA from its hyperprior (iteration 0 streams)."""
        pp, hy = self.prior_params, self.hyper
        K, N, G = self.K, self.N, self.G
        s = self.seed

        def need(name, shp, axis):
            if name not in pp:
                pp[name] = np.full(shp, np.nan)
            m = np.isnan(pp[name]).any(axis=axis, keepdims=True)   # per signature n
            return np.broadcast_to(m, shp)

        cp, ce = self._cells_p(), self._cells_e()
        if self.prior == "truncnormal":
            m = need("Mu_p", (K, N), 0)
            pp["Mu_p"] = np.where(m, dr.normal_draw(s, 0, px.PUR_HYP_P1, cp, hy["M_p"], np.sqrt(hy["S_p"])), pp["Mu_p"])
            m = need("Sigmasq_p", (K, N), 0)
            pp["Sigmasq_p"] = np.where(m, 1.0 / dr.gamma_draw(s, 0, px.PUR_HYP_P2, cp, hy["A_p"], hy["B_p"]), pp["Sigmasq_p"])
            m = need("Mu_e", (N, G), 1)
            pp["Mu_e"] = np.where(m, dr.normal_draw(s, 0, px.PUR_HYP_E1, ce, hy["M_e"], np.sqrt(hy["S_e"])), pp["Mu_e"])
            m = need("Sigmasq_e", (N, G), 1)
            pp["Sigmasq_e"] = np.where(m, 1.0 / dr.gamma_draw(s, 0, px.PUR_HYP_E2, ce, hy["A_e"], hy["B_e"]), pp["Sigmasq_e"])
        elif self.prior == "exponential":
            m = need("Lambda_p", (K, N), 0)
            pp["Lambda_p"] = np.where(m, dr.gamma_draw(s, 0, px.PUR_HYP_P1, cp, hy["A_p"], hy["B_p"]), pp["Lambda_p"])
            m = need("Lambda_e", (N, G), 1)
            pp["Lambda_e"] = np.where(m, dr.gamma_draw(s, 0, px.PUR_HYP_E1, ce, hy["A_e"], hy["B_e"]), pp["Lambda_e"])
        else:
            m = need("Beta_p", (K, N), 0)
            pp["Beta_p"] = np.where(m, dr.gamma_draw(s, 0, px.PUR_HYP_P1, cp, hy["A_p"], hy["B_p"]), pp["Beta_p"])
            m = need("Alpha_p", (K, N), 0)
            pp["Alpha_p"] = np.where(m, dr.gamma_draw(s, 0, px.PUR_HYP_P2, cp, hy["C_p"], hy["D_p"]), pp["Alpha_p"])
            m = need("Beta_e", (N, G), 1)
            pp["Beta_e"] = np.where(m, dr.gamma_draw(s, 0, px.PUR_HYP_E1, ce, hy["A_e"], hy["B_e"]), pp["Beta_e"])
            m = need("Alpha_e", (N, G), 1)
            pp["Alpha_e"] = np.where(m, dr.gamma_draw(s, 0, px.PUR_HYP_E2, ce, hy["C_e"], hy["D_e"]), pp["Alpha_e"])
        if self.likelihood == "normal":
            # R/sample_priors.R:133-140 (the %in% tests values, so these are always rebuilt)
            pp["Alpha"] = np.full(G, float(pp["alpha"]))
            pp["Beta"] = np.full(G, float(pp["beta"]))

    def init_params_(self):
        """R/sample_params.R:16-41."""
        p = self.params
        p.setdefault("P", np.full((self.K, self.N), np.nan))
        p.setdefault("E", np.full((self.N, self.G), np.nan))
        if self.likelihood == "normal":
            p.setdefault("sigmasq", np.zeros(self.G))
        p.setdefault("A", np.full(self.N, np.nan))
        p["A"] = np.asarray(p["A"], dtype=np.float64).reshape(-1)
        p.setdefault("R", float(self.N))

    # -- reconstruction / likelihood -----------------------------------------------------
    def get_Mhat(self, P=None, A=None, E=None):
        """get_Mhat_  (R/utils.R:29-49): P %*% diag(A) %*% E."""
        P = self.params["P"] if P is None else P
        A = self.params["A"] if A is None else A
        E = self.params["E"] if E is None else E
        return (P * np.asarray(A).reshape(1, -1)) @ E

    def loglik_matrix(self, P=None, A=None, E=None, sigmasq=None, likelihood=None):
        """get_loglik_(return_matrix = TRUE)  (R/utils.R:62-112)."""
        likelihood = likelihood or self.likelihood
        Mhat = self.get_Mhat(P, A, E)
        if likelihood == "normal":
            sg = self.params["sigmasq"] if sigmasq is None else sigmasq
            sg = np.asarray(sg, dtype=np.float64)
            if sg.ndim == 1:
                sg = np.broadcast_to(sg[None, :], Mhat.shape)
            return dnorm_log(self.data, Mhat, sg)
        return dpois_log(self.data, np.maximum(Mhat, 1e-6))

    def get_loglik(self, **kw):
        return float(self.loglik_matrix(**kw).sum())

    def log_prior_parts(self):
        """log prior part of get_logpost_  (R/utils.R:131-175), P side and E side; all N
        signatures count."""
        P, E, pp = self.params["P"], self.params["E"], self.prior_params
        if self.prior == "truncnormal":
            return (float(dtruncnorm0_log(P, pp["Mu_p"], np.sqrt(pp["Sigmasq_p"])).sum()),
                    float(dtruncnorm0_log(E, pp["Mu_e"], np.sqrt(pp["Sigmasq_e"])).sum()))
        if self.prior == "exponential":
            return float(dexp_log(P, pp["Lambda_p"]).sum()), float(dexp_log(E, pp["Lambda_e"]).sum())
        return (float(dgamma_log(P, pp["Alpha_p"], pp["Beta_p"]).sum()),
                float(dgamma_log(E, pp["Alpha_e"], pp["Beta_e"]).sum()))

    def log_prior(self):
        a, b = self.log_prior_parts()
        return a + b

    def compute_metrics_(self):
        """compute_metrics_ + update_sample_metrics_  (R/utils.R:412-455, :339-348)."""
        A = self.params["A"]
        Mhat = self.get_Mhat()
        n_params = A.sum() * (self.G_total + self.K)
        Mh = np.maximum(Mhat, 1e-6)
        Mp = np.maximum(self.data, 1e-6)
        lp_P, lp_E = self.log_prior_parts()
        # sums over this shard's genomes, then over shards (identity when not sharded)
        part = self.reduce(np.array([self.get_loglik(), np.sum((Mhat - self.data) ** 2), np.sum(Mp * np.log(Mp / Mh)), lp_E]))
        loglik = float(part[0])
        logpost = loglik + lp_P + float(part[3])
        m = dict(iter=self.iter,
                 RMSE=float(np.sqrt(part[1] / (self.K * self.G_total))),
                 KL=float(part[2]),        # padded_KL_, R/utils.R:467-471
                 loglikelihood=loglik, logposterior=logpost, n_params=float(n_params),
                 BIC=float(-2.0 * loglik + n_params * np.log(self.G_total)),
                 rank=float(A.sum()), temp=float(self.temperature_schedule[self.iter - 1]))
        if self.MH:
            act = A == 1
            m["P_mean_acceptance_rate"] = float(np.mean(self.acc["P"][:, act])) if act.any() else np.nan
            m["E_mean_acceptance_rate"] = float(np.mean(self.acc["E"][act, :])) if act.any() else np.nan
        return m

    def update_sample_metrics_(self):
        self.metrics.append(self.compute_metrics_())

    # -- prior parameters ----------------------------------------------------------------
    def sample_prior_params_(self):
        """sample_prior_params_ and its leaves  (R/sample_priors.R:150-397).  Every update
        is element-wise in (k,n) / (n,g) and reads only the previous P / E, so the
        reference's loop over n is evaluated for all n at once."""
        pp, hy, s, it = self.prior_params, self.hyper, self.seed, self.iter
        P, E = self.params["P"], self.params["E"]
        cp, ce = self._cells_p(), self._cells_e()
        if self.prior == "truncnormal":
            # sample_Mu_Pn / sample_Mu_En (:214-236): rnorm(., num/denom, 1/denom)  [quirk 1]
            num = hy["M_p"] / hy["S_p"] + P / pp["Sigmasq_p"]
            den = 1.0 / hy["S_p"] + 1.0 / pp["Sigmasq_p"]
            pp["Mu_p"] = dr.normal_draw(s, it, px.PUR_HYP_P1, cp, num / den, 1.0 / den)
            num = hy["M_e"] / hy["S_e"] + E / pp["Sigmasq_e"]
            den = 1.0 / hy["S_e"] + 1.0 / pp["Sigmasq_e"]
            pp["Mu_e"] = dr.normal_draw(s, it, px.PUR_HYP_E1, ce, num / den, 1.0 / den)
            # sample_Sigmasq_Pn / _En (:246-270): rate base is A_e on the E side  [quirk 2]
            pp["Sigmasq_p"] = 1.0 / dr.gamma_draw(s, it, px.PUR_HYP_P2, cp, hy["A_p"] + 0.5,
                                                  hy["B_p"] + (P - pp["Mu_p"]) ** 2 / 2.0)
            pp["Sigmasq_e"] = 1.0 / dr.gamma_draw(s, it, px.PUR_HYP_E2, ce, hy["A_e"] + 0.5,
                                                  hy["A_e"] + (E - pp["Mu_e"]) ** 2 / 2.0)
        elif self.prior == "exponential":
            # sample_Lambda_Pn / _En (:284-308)
            pp["Lambda_p"] = dr.gamma_draw(s, it, px.PUR_HYP_P1, cp, hy["A_p"] + 1.0, hy["B_p"] + P)
            pp["Lambda_e"] = dr.gamma_draw(s, it, px.PUR_HYP_E1, ce, hy["A_e"] + 1.0, hy["B_e"] + E)
        else:
            # sample_Beta_Pn (:323-329) then sample_Alpha_Pkn (:356-371) with the new Beta
            pp["Beta_p"] = dr.gamma_draw(s, it, px.PUR_HYP_P1, cp, hy["A_p"] + pp["Alpha_p"], hy["B_p"] + P)
            pp["Alpha_p"] = dr.alpha_draw(s, it, px.PUR_HYP_P2, cp, hy["C_p"], hy["D_p"], pp["Beta_p"], P, x0=pp["Alpha_p"])
            pp["Beta_e"] = dr.gamma_draw(s, it, px.PUR_HYP_E1, ce, hy["A_e"] + pp["Alpha_e"], hy["B_e"] + E)
            pp["Alpha_e"] = dr.alpha_draw(s, it, px.PUR_HYP_E2, ce, hy["C_e"], hy["D_e"], pp["Beta_e"], E, x0=pp["Alpha_e"])

    # -- P and E -------------------------------------------------------------------------
    def _prior_draw(self, side, n):
        """Draw column n of P / row n of E from its prior  (R/sample_Pn.R:12-30,56-74;
        R/sample_En.R:12-30,56-73)."""
        pp, s, it = self.prior_params, self.seed, self.iter
        if side == "P":
            c, pur, sl = self._cells_p(n), px.PUR_P, (slice(None), n)
            sfx = "_p"
        else:
            c, pur, sl = self._cells_e(n), px.PUR_E, (n, slice(None))
            sfx = "_e"
        if self.prior == "truncnormal":
            return dr.truncnorm0_draw(s, it, pur, c, pp["Mu" + sfx][sl], np.sqrt(pp["Sigmasq" + sfx][sl]), self.bits)
        if self.prior == "exponential":
            # rexp(rate) == Gamma(1, rate); drawn through the gamma generator like the kernels
            return dr.gamma_draw(s, it, pur, c, 1.0, pp["Lambda" + sfx][sl], self.bits)
        return dr.gamma_draw(s, it, pur, c, pp["Alpha" + sfx][sl], pp["Beta" + sfx][sl], self.bits)

    def sample_Pn(self, n, from_prior=False):
        """sample_Pn  (R/sample_Pn.R:11-42)."""
        A = self.params["A"]
        if from_prior or A[n] == 0:
            return self._prior_draw("P", n)
        if self.likelihood == "normal":
            return self.sample_Pn_normal(n, as_proposal=False)
        if self.MH:
            prop = self.sample_Pn_normal(n, as_proposal=True)
            return self.MH_Pn_poisson(prop, n)
        return self.sample_Pn_poisson(n)

    def sample_En(self, n, from_prior=False):
        """sample_En  (R/sample_En.R:11-42)."""
        A = self.params["A"]
        if from_prior or A[n] == 0:
            return self._prior_draw("E", n)
        if self.likelihood == "normal":
            return self.sample_En_normal(n, as_proposal=False)
        if self.MH:
            prop = self.sample_En_normal(n, as_proposal=True)
            return self.MH_En_poisson(prop, n)
        return self.sample_En_poisson(n)

    def sample_Pn_poisson(self, n):
        """R/sample_Pn.R:98-120: Gamma(shape + sum_g Z[k,n,g], rate + A_n sum_g E[n,g])."""
        pp, A, E = self.prior_params, self.params["A"], self.params["E"]
        if self.prior == "gamma":
            shape = pp["Alpha_p"][:, n] + self.SP[:, n]
            rate = pp["Beta_p"][:, n] + A[n] * self.rowsumE[n]
        else:
            shape = 1.0 + self.SP[:, n]
            rate = pp["Lambda_p"][:, n] + A[n] * self.rowsumE[n]
        return dr.gamma_draw(self.seed, self.iter, px.PUR_P, self._cells_p(n), shape, rate, self.bits)

    def sample_En_poisson(self, n):
        """R/sample_En.R:97-119: Gamma(shape + sum_k Z[k,n,g], rate + A_n sum_k P[k,n])."""
        pp, A, P = self.prior_params, self.params["A"], self.params["P"]
        csP = float(self._r(P[:, n].sum()))        # colSums(P) is stored in the state precision
        if self.prior == "gamma":
            shape = pp["Alpha_e"][n, :] + self.SE[n, :]
            rate = pp["Beta_e"][n, :] + A[n] * csP
        else:
            shape = 1.0 + self.SE[n, :]
            rate = pp["Lambda_e"][n, :] + A[n] * csP
        return dr.gamma_draw(self.seed, self.iter, px.PUR_E, self._cells_e(n), shape, rate, self.bits)

    def _sigmasq_matrix(self, Mhat, as_proposal):
        if as_proposal:
            return Mhat                                            # R/sample_Pn.R:137-139  [quirk 4]
        return np.broadcast_to(self.params["sigmasq"][None, :], Mhat.shape)   # :142-146

    def get_mu_sigmasq_Pn_normal(self, n, as_proposal=False):
        """R/sample_Pn.R:132-187."""
        P, E, A, pp = self.params["P"], self.params["E"], self.params["A"], self.prior_params
        Mhat = self.get_Mhat()
        sig = self._sigmasq_matrix(Mhat, as_proposal)
        A0 = A.copy(); A0[n] = 0
        Mhat_no_n = self.get_Mhat(A=A0)
        num1 = (E[n, :][None, :] * ((self.data - Mhat_no_n) / sig)).sum(axis=1)
        den = ((A[n] * E[n, :] ** 2)[None, :] * (1.0 / sig)).sum(axis=1)
        if self.prior == "exponential":
            mu = (num1 - pp["Lambda_p"][:, n]) / den
            return mu, 1.0 / den
        den = den + 1.0 / pp["Sigmasq_p"][:, n]
        mu = (num1 + pp["Mu_p"][:, n] / pp["Sigmasq_p"][:, n]) / den
        return mu, 1.0 / den

    def get_mu_sigmasq_En_normal(self, n, as_proposal=False):
        """R/sample_En.R:131-184."""
        P, E, A, pp = self.params["P"], self.params["E"], self.params["A"], self.prior_params
        Mhat = self.get_Mhat()
        sig = self._sigmasq_matrix(Mhat, as_proposal)
        A0 = A.copy(); A0[n] = 0
        Mhat_no_n = self.get_Mhat(A=A0)
        num1 = (P[:, n][:, None] * ((self.data - Mhat_no_n) / sig)).sum(axis=0)
        den = ((A[n] * P[:, n] ** 2)[:, None] * (1.0 / sig)).sum(axis=0)
        if self.prior == "exponential":
            mu = (num1 - pp["Lambda_e"][n, :]) / den
            return mu, 1.0 / den
        den = den + 1.0 / pp["Sigmasq_e"][n, :]
        mu = (num1 + pp["Mu_e"][n, :] / pp["Sigmasq_e"][n, :]) / den
        return mu, 1.0 / den

    def sample_Pn_normal(self, n, as_proposal=False):
        """R/sample_Pn.R:54-87."""
        if self.params["A"][n] == 0 or np.all(self.params["E"][n, :] == 0):
            return self._prior_draw("P", n)
        mu, v = self.get_mu_sigmasq_Pn_normal(n, as_proposal)
        self.last_cond = ("P", n, mu, v)
        return dr.truncnorm0_draw(self.seed, self.iter, px.PUR_P, self._cells_p(n), mu, np.sqrt(v), self.bits)

    def sample_En_normal(self, n, as_proposal=False):
        """R/sample_En.R:54-86."""
        if self.params["A"][n] == 0 or np.all(self.params["P"][:, n] == 0):
            return self._prior_draw("E", n)
        mu, v = self.get_mu_sigmasq_En_normal(n, as_proposal)
        self.last_cond = ("E", n, mu, v)
        return dr.truncnorm0_draw(self.seed, self.iter, px.PUR_E, self._cells_e(n), mu, np.sqrt(v), self.bits)

    def MH_Pn_poisson(self, proposal, n):
        """R/sample_Pn.R:199-248."""
        if not self.converged:
            self.acc["P"][:, n] = 1.0
            return proposal
        P = self.params["P"]
        Pprop = P.copy(); Pprop[:, n] = proposal
        Mhat = self.get_Mhat()
        Mhat_prop = self.get_Mhat(P=Pprop)
        lp_old = self.loglik_matrix(likelihood="poisson").sum(axis=1)
        lp_new = self.loglik_matrix(P=Pprop, likelihood="poisson").sum(axis=1)
        ln_old = self.loglik_matrix(sigmasq=np.maximum(Mhat_prop, 1.0), likelihood="normal").sum(axis=1)
        ln_new = self.loglik_matrix(P=Pprop, sigmasq=np.maximum(Mhat, 1.0), likelihood="normal").sum(axis=1)
        with np.errstate(over="ignore"):
            ratio = np.minimum(np.exp(lp_new + ln_old - (lp_old + ln_new)), 1.0)
        self.acc["P"][:, n] = ratio
        u = px.u01(px.words(self.seed, self.iter, px.PUR_MH_P, self._cells_p(n), 0)[0], self.bits)
        return np.where(u < ratio, proposal, P[:, n])

    def MH_En_poisson(self, proposal, n):
        """R/sample_En.R:196-241."""
        if not self.converged:
            self.acc["E"][n, :] = 1.0
            return proposal
        E = self.params["E"]
        Eprop = E.copy(); Eprop[n, :] = proposal
        Mhat = self.get_Mhat()
        Mhat_prop = self.get_Mhat(E=Eprop)
        lp_old = self.loglik_matrix(likelihood="poisson").sum(axis=0)
        lp_new = self.loglik_matrix(E=Eprop, likelihood="poisson").sum(axis=0)
        ln_old = self.loglik_matrix(sigmasq=np.maximum(Mhat_prop, 1.0), likelihood="normal").sum(axis=0)
        ln_new = self.loglik_matrix(E=Eprop, sigmasq=np.maximum(Mhat, 1.0), likelihood="normal").sum(axis=0)
        with np.errstate(over="ignore"):
            ratio = np.minimum(np.exp(lp_new + ln_old - (lp_old + ln_new)), 1.0)
        self.acc["E"][n, :] = ratio
        u = px.u01(px.words(self.seed, self.iter, px.PUR_MH_E, self._cells_e(n), 0)[0], self.bits)
        return np.where(u < ratio, proposal, E[n, :])

    # -- rank ----------------------------------------------------------------------------
    @staticmethod
    def compute_prior_prob_1(R, N, clip_val=0.4):
        """R/sample_params.R:178-187."""
        q = R / N
        q = max(q, clip_val / N)
        return min(q, 1.0 - clip_val / N)

    @staticmethod
    def sumLog(vec):
        """R/sample_params.R:199-206."""
        o = sorted(vec, reverse=True)
        s = o[0]
        for v in o[1:]:
            s = s + np.log(1.0 + np.exp(v - s))
        return s

    def sample_R(self, from_prior=False):
        """R/sample_params.R:217-241; `sample(range_R, 1, prob)` by inverse CDF in
        natural order with one uniform."""
        N = self.N
        u = px.u01(px.words(self.seed, self.iter, px.PUR_R, 0, 0)[0])
        if from_prior:
            return float(min(int(u * (N + 1)), N))
        T = self.temperature_schedule[self.iter - 1]
        sA = self.params["A"].sum()
        probs = np.empty(N + 1)
        for r in range(N + 1):
            q = self.compute_prior_prob_1(r, N)
            probs[r] = (1.0 / (N + 1)) * (q ** sA * (1.0 - q) ** (N - sA)) ** T
        probs = probs / probs.sum()
        cdf = np.cumsum(probs)
        r = int((cdf <= u).sum())
        return float(min(r, N))

    def sample_An(self, n, from_prior=False):
        """R/sample_params.R:101-166."""
        N, K, G = self.N, self.K, self.G_total
        q = self.compute_prior_prob_1(self.params["R"], N)
        u = px.u01(px.words(self.seed, self.iter, px.PUR_A, n, 0)[0])
        if from_prior:
            return 1.0 if u < q else 0.0
        T = self.temperature_schedule[self.iter - 1]
        A0 = self.params["A"].copy(); A0[n] = 0
        A1 = self.params["A"].copy(); A1[n] = 1
        l0, l1 = (float(v) for v in self.reduce(np.array([self.get_loglik(A=A0), self.get_loglik(A=A1)])))
        if self.rank_method == "SBFI":
            b0 = l0 - A0.sum() * (G + K) * np.log(G) / 2.0
            b1 = l1 - A1.sum() * (G + K) * np.log(G) / 2.0
            lp0 = np.log(1.0 - q) + T * b0
            lp1 = np.log(q) + T * b1
        else:
            lp0 = np.log(1.0 - q) + T * l0
            lp1 = np.log(q) + T * l1
        with np.errstate(all="ignore"):
            p = np.exp(lp1 - self.sumLog([lp0, lp1]))
        if np.isnan(p):
            if np.isnan(lp1) and np.isnan(lp0):
                p = 0.5
            elif np.isnan(lp1):
                p = 0.0
            elif np.isnan(lp0):
                p = 1.0
            elif lp1 > lp0:
                p = 1.0
            elif lp1 < lp0:
                p = 0.0
            else:
                p = 0.5
        self.last_pA = p
        return 1.0 if u < p else 0.0

    def sample_sigmasq(self):
        """R/sample_params.R:275-286: InvGamma(Alpha_g + K/2, rate = Beta_g + sum_k resid^2 / 2)."""
        Mhat = self.get_Mhat()
        ss = ((self.data - Mhat) ** 2).sum(axis=0)
        g = np.arange(self.G, dtype=np.uint64) + np.uint64(self.g0)
        shape = self.prior_params["Alpha"] + self.K / 2.0
        rate = self.prior_params["Beta"] + 0.5 * ss
        return 1.0 / dr.gamma_draw(self.seed, self.iter, px.PUR_SIGMASQ, g, shape, rate, self.bits)

    # -- one sweep -----------------------------------------------------------------------
    @property
    def rowsumE(self):
        """rowSums(E) as the kernels accumulate it: fixed point 2^-24, exact integer sum."""
        return self.reduce(np.rint(self.params["E"] * 16777216.0).sum(axis=1)) / 16777216.0

    def sample_params_(self, skip=(), from_prior=False):
        """sample_params_  (R/sample_params.R:51-89): P (n = 1..N) -> E (n = 1..N) -> R, A
        -> Z -> sigmasq."""
        p = self.params
        if "P" not in skip:
            self._rsE_cache = None
            for n in range(self.N):
                p["P"][:, n] = self._r(self.sample_Pn(n, from_prior))
        if "E" not in skip:
            for n in range(self.N):
                p["E"][n, :] = self._r(self.sample_En(n, from_prior))
        if "A" not in skip and self.learning_rank:
            p["R"] = self.sample_R(from_prior)
            for n in range(self.N):
                p["A"][n] = self.sample_An(n, from_prior)
        elif "A" not in skip and not self.learning_rank and from_prior:
            p["A"] = np.ones(self.N)
        if self.likelihood == "poisson" and not self.MH and "Z" not in skip:
            self.SP, self.SE = sample_Z_stats(self.data, p["P"], p["A"], p["E"], self.seed, self.iter,
                                              g0=self.g0, bits=self.bits, f32=self.f32)
            self.SP = self.reduce(self.SP)
        if self.likelihood == "normal" and "sigmasq" not in skip:
            p["sigmasq"] = self.sample_sigmasq()

    def step(self):
        """Loop body of run_gibbs_sampler  (R/bayesNMF_sampler.R:273-285)."""
        self.iter += 1
        self.sample_prior_params_()
        self.sample_params_()
        self.update_sample_metrics_()
        return self.metrics[-1]
