"""Host-side mirror of the reference's sampler API over the C ABI.

    bayesNMF(data, rank, likelihood, prior, rank_method, MH, convergence_control, ...)
        -> bayesNMF_sampler                                   (R/bayesNMF.R:24-138)

The reference is an R package; R is not available in this image, so the host layer that a
maintainer would patch in R is restated here in Python with the same names,
argument meaning, defaults and error messages, for the part of the object that the Gibbs
path touches: construction (R/bayesNMF_sampler.R:92-260), run_gibbs_sampler (:265-408),
get_MAP (R/utils.R:194-288), check_convergence_ (R/convergence.R:60-154) and
update_MAP_metrics_ (R/utils.R:356-397).  Every iteration itself runs on the GPU
(bnmf_step); this file only does the bookkeeping that happens every MAP_every iterations.
Logging, output directories, saveRDS and the plots stay R-only (out of scope).
"""
import time

import numpy as np

from ._lib import METRIC_NAMES, BnmfError, Handle
from .hyperpriors import fill_hyperprior_params


def new_convergence_control(MAP_over=1000, MAP_every=100, tol=0.001, Ninarow_nochange=5, Ninarow_nobest=10,
                            miniters=1000, maxiters=5000, minA=0, metric="logposterior"):
    """new_convergence_control (R/convergence.R:16-45)."""
    if miniters >= maxiters:
        import warnings
        warnings.warn("miniters >= maxiters, setting miniters to 0.")
        miniters = 0
    return dict(MAP_over=MAP_over, MAP_every=MAP_every, tol=tol, Ninarow_nochange=Ninarow_nochange,
                Ninarow_nobest=Ninarow_nobest, miniters=miniters, maxiters=maxiters, minA=minA, metric=metric)


def get_temp_sched(length, n_temp, rng):
    """get_temp_sched_ (R/utils.R:307-332); `sort(sample(sched, n_temp))` draws from the host
    RNG (numpy's Generator stands in for R's)."""
    nX = max(int(round(n_temp / 374.0)), 1)
    sched = [0.0] * nX
    for x in range(9, 4, -1):
        sched += [10.0 ** (-x)] * nX
    sched += [10.0 ** (-4)] * int(round(8 * nX))
    for y in range(4, 0, -1):
        for xi in range(90):
            sched += [(1 + xi * 0.1) * 10.0 ** (-y)] * nX
    sched = np.asarray(sched)
    if len(sched) > n_temp:
        sched = np.sort(rng.choice(sched, size=n_temp, replace=False))
    return np.concatenate([sched, np.ones(max(length - len(sched), 0))])


class bayesNMF_sampler:
    """bayesNMF_sampler (R/bayesNMF_sampler.R:8-747): fields `data`, `dims`, `specs`,
    `temperature_schedule`, `params`, `prior_params`, `hyperprior_params`, `acceptance_rates`,
    `samples`, `state`, `MAP`, `credible_intervals`, `time` keep the reference's names."""

    def __init__(self, data, rank, likelihood="poisson", prior="truncnormal", rank_method="SBFI", MH=None,
                 convergence_control=None, prop_temp=0.2, post_warmup=None, hyperprior_params=None,
                 init_prior_params=None, init_params=None, save_all_samples=False, seed=0, precision="f64", device=0):
        cc = dict(convergence_control or new_convergence_control())
        if MH is None:
            MH = likelihood == "poisson" and prior in ("truncnormal", "exponential")
        if post_warmup is None:
            post_warmup = 2 * cc["MAP_over"]
        rank = np.atleast_1d(np.asarray(rank, dtype=int))
        learning_rank = rank.size > 1                                    # :124-125
        if learning_rank and rank.min() != 0:
            rank = np.arange(0, rank.max() + 1)
        self.data = np.asarray(data, dtype=np.float64)
        self.dims = dict(K=self.data.shape[0], N=int(rank.max()), G=self.data.shape[1])
        self.specs = dict(rank=rank, likelihood=likelihood, prior=prior, MH=bool(MH), learning_rank=learning_rank,
                          convergence_control=cc, save_all_samples=bool(save_all_samples), seed=int(seed))
        if learning_rank:
            if rank_method not in ("SBFI", "BFI", "BIC"):
                raise ValueError("Rank method must be SBFI, BFI, or BIC")
            self.specs.update(prop_temp=prop_temp, rank_method=rank_method)
        if MH:
            self.specs["post_warmup"] = int(post_warmup)
        n_iters = cc["maxiters"] + (int(post_warmup) if MH else 0)       # :129-137
        host_rng = np.random.default_rng(seed)
        if learning_rank:
            self.temperature_schedule = get_temp_sched(n_iters, int(round(prop_temp * cc["maxiters"])), host_rng)
        else:
            self.temperature_schedule = np.ones(n_iters)
        self.state = dict(iter=1, converged=False, MAP_idx=np.arange(1, cc["MAP_over"] + 1),
                          sample_metrics={k: [] for k in METRIC_NAMES}, MAP_metrics=[])
        self.time = {}
        self.MAP = None
        self.credible_intervals = None

        # the device sampler: check_model happens inside bnmf_create (R/bayesNMF_sampler.R:217)
        K, N, G = self.dims["K"], self.dims["N"], self.dims["G"]
        self._h = Handle(self.data, N, likelihood=likelihood, prior=prior, MH=MH, learning_rank=learning_rank,
                         rank_method=rank_method if rank_method in ("SBFI", "BFI") else "SBFI", seed=seed,
                         precision=precision, device=device, ring_cap=cc["MAP_over"])
        self.hyperprior_params = fill_hyperprior_params(hyperprior_params, prior, float(self.data.mean()), N)
        for name, value in self.hyperprior_params.items():
            self._h.set_hyper(name, value)
        init_prior_params = dict(init_prior_params or {})
        if likelihood == "normal":                                        # :222-230
            self._h.set_hyper("alpha", init_prior_params.pop("alpha", 3.0))
            self._h.set_hyper("beta", init_prior_params.pop("beta", 3.0))
        self._h.set_temperature_schedule(self.temperature_schedule)
        init_params = dict(init_params or {})
        for name, value in init_prior_params.items():
            self._h.set_state(name, value)
        for name, value in init_params.items():
            self._h.set_state(name, value)
        row = self._h.init_from_prior(have=tuple(init_params), have_prior=tuple(init_prior_params))   # :241-257
        self.samples = {"P": [], "A": []}
        self._pull_state()
        self.samples["P"].append(self.params["P"].copy())
        self.samples["A"].append(self.params["A"].copy())
        for k in METRIC_NAMES:
            self.state["sample_metrics"][k].append(row[k])

    # -- state exchange ------------------------------------------------------------------
    def _names(self):
        lik, prior = self.specs["likelihood"], self.specs["prior"]
        pn = {"truncnormal": ["Mu", "Sigmasq"], "exponential": ["Lambda"], "gamma": ["Alpha", "Beta"]}[prior]
        params = ["P", "E", "A", "R"] + (["sigmasq"] if lik == "normal" else [])
        return params, [f"{p}_{s}" for p in pn for s in ("p", "e")]

    def _pull_state(self):
        """Fill self$params / prior_params / acceptance_rates from the device (what the R patch
        does before save_object or any post-hoc method reads them)."""
        params, priors = self._names()
        self.params = {n: self._h.get_state(n) for n in params}
        self.prior_params = {n: self._h.get_state(n) for n in priors}
        if self.specs["MH"]:
            self.acceptance_rates = {n: self._h.get_state(n) for n in ("P_acceptance_rate", "E_acceptance_rate")}

    def get_Mhat(self, P=None, A=None, E=None):
        """get_Mhat_ (R/utils.R:29-49)."""
        P = self.params["P"] if P is None else P
        A = self.params["A"] if A is None else A
        E = self.params["E"] if E is None else E
        return (P * np.asarray(A).reshape(1, -1)) @ E

    def _advance(self, n, converged):
        out = self._h.step(n, converged=converged, want_P=True, want_A=True)
        for j, k in enumerate(METRIC_NAMES):
            self.state["sample_metrics"][k].extend(out["metrics"][:, j].tolist())
        keep = None if self.specs["save_all_samples"] else self.specs["convergence_control"]["MAP_over"]
        for name in ("P", "A"):
            self.samples[name].extend(list(out[name]))
            if keep is not None and len(self.samples[name]) > keep:      # update_list, R/helpers.R:111-119
                del self.samples[name][:len(self.samples[name]) - keep]
        self.state["iter"] += n

    def sample_E(self, ago=0):
        """samples$E stays in the device ring (N x G per sample); `ago` = 0 is the newest."""
        return self._h.get_sample("E", ago)

    # -- MAP and convergence -------------------------------------------------------------
    def get_MAP(self, final=False, credible_interval=0.95):
        """get_MAP_ (R/utils.R:194-288) on the device ring."""
        cc = self.specs["convergence_control"]
        n_s = min(cc["MAP_over"], self._h.ring_count())
        P_map, E_map, A_map, n_match = self._h.get_map(n_s)
        keys = ["".join("1" if a else "0" for a in A) for A in self.samples["A"][-n_s:]]
        counts = {}
        for k in keys:
            counts[k] = counts.get(k, 0) + 1
        top = sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))
        mode_key = "".join("1" if a else "0" for a in A_map)
        assert top[0][0] == mode_key and top[0][1] == n_match
        first = self.state["iter"] - n_s + 1
        idx = np.array([first + i for i, k in enumerate(keys) if k == mode_key])
        keep_sigs = np.nonzero(A_map == 1)[0] if final else np.arange(self.dims["N"])
        self.MAP = dict(P=P_map[:, keep_sigs], A=A_map[keep_sigs].reshape(1, -1), E=E_map[keep_sigs, :], idx=idx,
                        A_counts=top[:5], keep_sigs=keep_sigs)
        if final or self.credible_intervals is not None:                 # :264-287, quantile type 7
            # element-wise quantiles over the matching samples, on the device ring: the E samples
            # (N x G each) never cross PCIe
            Pl, Ph, El, Eh, nm = self._h.get_credible_intervals(n_s, 0.5 - credible_interval / 2, 0.5 + credible_interval / 2)
            assert nm == n_match
            self.credible_intervals = dict(P=dict(lower=Pl[:, keep_sigs], upper=Ph[:, keep_sigs]),
                                           E=dict(lower=El[keep_sigs, :], upper=Eh[keep_sigs, :]))

    def assign_signatures_ensemble(self, reference_P, reference_names=None, credible_interval=0.95):
        """assign_signatures_ensemble_ (R/postprocessing.R:175-341) over the samples behind the current MAP:
        dict(assignments = rows of (sig_est, sig_ref, MAP_cosine, lower_cosine, upper_cosine),
             votes = rows of (sig_est, sig_ref, prop_votes)), stored in self.reference_comparison."""
        cc = self.specs["convergence_control"]
        n_s = min(cc["MAP_over"], self._h.ring_count())
        r = self._h.assign_signatures(n_s, reference_P, credible_interval)
        nm = (lambda j: reference_names[j]) if reference_names is not None else (lambda j: f"Ref{j + 1}")
        assignments = [dict(sig_est=int(k) + 1, sig_ref=nm(int(a)), MAP_cosine=float(c), lower_cosine=float(lo), upper_cosine=float(hi))
                       for k, a, c, lo, hi in zip(r["keep_sigs"], r["assignment"], r["MAP_cosine"], r["lower_cosine"], r["upper_cosine"])]
        votes = []
        for i, k in enumerate(r["keep_sigs"]):
            order = np.argsort(-r["votes"][i], kind="stable")
            votes += [dict(sig_est=int(k) + 1, sig_ref=nm(int(j)), prop_votes=float(r["votes"][i, j])) for j in order if r["votes"][i, j] > 0]
        self.reference_comparison = dict(reference_P=np.asarray(reference_P), assignments=assignments, votes=votes)
        return dict(assignments=assignments, votes=votes)

    def _update_MAP_metrics(self, final=False):
        """update_MAP_metrics_ + compute_metrics_(MAP = TRUE) (R/utils.R:356-397, :412-455)."""
        cc = self.specs["convergence_control"]
        A = np.ones(self.MAP["P"].shape[1]) if final else self.MAP["A"].reshape(-1)
        Mhat = self.get_Mhat(self.MAP["P"], A, self.MAP["E"])
        n_params = float(A.sum() * (self.dims["G"] + self.dims["K"]))
        it = self.state["iter"]
        sm = self.state["sample_metrics"]
        iters = np.asarray(sm["iter"])
        win = (iters > it - cc["MAP_over"]) & (iters <= it)
        loglik = float(np.mean(np.asarray(sm["loglikelihood"])[win]))
        logpost = float(np.mean(np.asarray(sm["logposterior"])[win]))
        Mp, Mh = np.maximum(self.data, 1e-6), np.maximum(Mhat, 1e-6)
        lo = max(it - cc["MAP_over"], 0)
        row = dict(iter=it, RMSE=float(np.sqrt(np.mean((Mhat - self.data) ** 2))), KL=float(np.sum(Mp * np.log(Mp / Mh))),
                   loglikelihood=loglik, logposterior=logpost, n_params=n_params,
                   BIC=-2.0 * loglik + n_params * np.log(self.dims["G"]), rank=float(np.sum(self.MAP["A"])),
                   MAP_A_counts=self.MAP["A_counts"][0][1], mean_temp=float(np.mean(self.temperature_schedule[lo:it])))
        if self.specs["MH"]:
            row["P_mean_acceptance_rate"] = sm["P_mean_acceptance_rate"][-1]
            row["E_mean_acceptance_rate"] = sm["E_mean_acceptance_rate"][-1]
        self.state["MAP_metrics"].append(row)

    def _check_convergence(self, final=False):
        """check_convergence_ (R/convergence.R:60-154)."""
        cc, st = self.specs["convergence_control"], self.state
        self._update_MAP_metrics(final=final)
        metric = st["MAP_metrics"][-1][cc["metric"]]
        if cc["metric"] in ("loglikelihood", "logposterior"):
            metric = -metric
        if "prev_MAP_metric" not in st:
            st.update(prev_MAP_metric=metric + 1, best_MAP_metric=metric + 1, inarow_na=0, inarow_no_change=0, inarow_no_best=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            pc = np.float64(metric - st["prev_MAP_metric"]) / np.float64(st["prev_MAP_metric"])
        st["prev_percent_change"], st["prev_MAP_metric"] = pc, metric
        if np.isnan(pc):
            st["inarow_no_change"] = 0; st["inarow_no_best"] = 0; st["inarow_na"] += 1
        elif abs(pc) < cc["tol"]:
            st["inarow_no_change"] += 1; st["inarow_na"] = 0
        else:
            st["inarow_no_change"] = 0; st["inarow_na"] = 0
        it = st["iter"]
        lo = max(it - cc["MAP_over"], 1)                                  # R's 1-based window [iter - MAP_over, iter]
        if np.all(self.temperature_schedule[lo - 1:it] == 1) and it >= cc["miniters"]:
            if metric < st["best_MAP_metric"]:
                st.update(best_MAP_metric=metric, best_iter=it, inarow_no_best=0)
            else:
                st["inarow_no_best"] += 1
            if st["inarow_no_change"] >= cc["Ninarow_nochange"]:
                st.update(converged=True, why="no change")
            elif st["inarow_no_best"] >= cc["Ninarow_nobest"]:
                st.update(converged=True, why="no best")
            elif it >= cc["maxiters"]:
                st.update(converged=True, why="max iters")

    # -- the driver loop -------------------------------------------------------------------
    def run_gibbs_sampler(self):
        """run_gibbs_sampler (R/bayesNMF_sampler.R:265-408): the loop body runs on the GPU in
        blocks that end at the next MAP check."""
        cc, st = self.specs["convergence_control"], self.state
        t0 = time.time()
        while not st["converged"] and st["iter"] < cc["maxiters"]:
            n = min(cc["MAP_every"] - st["iter"] % cc["MAP_every"], cc["maxiters"] - st["iter"])
            self._advance(n, converged=False)
            it = st["iter"]
            if (it % cc["MAP_every"] == 0 and it >= max(cc["MAP_over"], cc["MAP_every"])) or it >= cc["maxiters"]:
                if self.specs["save_all_samples"]:
                    st["MAP_idx"] = np.arange(it - cc["MAP_over"] + 1, it + 1)
                self.get_MAP()
                self._check_convergence()
                if st["converged"]:
                    st["converged_iter"] = it
        if self.specs["MH"]:
            t1 = time.time()
            self.time["warmup"] = (t1 - t0) / 60.0
            pw, done = self.specs["post_warmup"], 0
            while done < pw:
                n = min(cc["MAP_every"] - st["iter"] % cc["MAP_every"], pw - done)
                self._advance(n, converged=True)
                done += n
                final = done == pw
                if self.specs["save_all_samples"]:
                    st["MAP_idx"] = np.arange(st["iter"] - cc["MAP_over"] + 1, st["iter"] + 1)
                self.get_MAP(final=final)
                self._check_convergence(final=final)
            self.time["MH"] = (time.time() - t1) / 60.0
        else:
            self.get_MAP(final=True)
        self._pull_state()
        self.time["total"] = (time.time() - t0) / 60.0
        self.time["per_iter"] = self.time["total"] / st["iter"]
        return self

    def close(self):
        self._h.close()


def bayesNMF(data, rank, likelihood="poisson", prior="truncnormal", rank_method="SBFI", MH=None,
             convergence_control=None, prop_temp=0.2, post_warmup=None, hyperprior_params=None,
             init_prior_params=None, init_params=None, save_all_samples=True, seed=0, precision="f64", device=0,
             devices=None):
    """bayesNMF (R/bayesNMF.R:24-138).  Returns the sampler after run_gibbs_sampler(); with
    rank_method = "BIC" and several ranks, one fixed-rank sampler per rank is run (:66-127) and a
    dict(results, best_rank, sampler) returned.  `devices` = list of CUDA ordinals: the ranks of the
    BIC range are placed one per GPU and run side by side (bayesnmf_b200/replicas.py); every rank's
    chain is the one the serial loop would have run."""
    kw = dict(likelihood=likelihood, prior=prior, rank_method=rank_method, MH=MH, convergence_control=convergence_control,
              prop_temp=prop_temp, post_warmup=post_warmup, hyperprior_params=hyperprior_params,
              init_prior_params=init_prior_params, init_params=init_params, save_all_samples=save_all_samples,
              seed=seed, precision=precision, device=device)
    ranks = np.atleast_1d(np.asarray(rank, dtype=int))
    if ranks.size > 1 and rank_method == "BIC":
        results, best = [], None
        if devices is not None and len(devices) > 1:
            from .replicas import run_replicas
            kw.pop("device")
            samplers = run_replicas([(lambda device, k=int(k): bayesNMF_sampler(data, k, device=device, **kw).run_gibbs_sampler())
                                     for k in ranks], devices)
        else:
            samplers = (bayesNMF_sampler(data, int(k), **kw).run_gibbs_sampler() for k in ranks)
        for k, s in zip(ranks, samplers):
            bic = s.state["MAP_metrics"][-1]["BIC"]
            results.append(dict(rank=int(k), BIC=bic, time=s.time["total"]))
            if best is None or bic < best[0]:
                if best is not None:
                    best[1].close()
                best = (bic, s)
            else:
                s.close()
        results.sort(key=lambda r: r["BIC"])
        return dict(results=results, best_rank=results[0]["rank"], sampler=best[1])
    return bayesNMF_sampler(data, rank, **kw).run_gibbs_sampler()
