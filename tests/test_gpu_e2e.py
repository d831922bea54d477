"""End-to-end runs through the host mirror of the reference API (bayesnmf_b200.bayesNMF),
north-star level 3: posterior summaries within Monte-Carlo error of the reference's own
published outcome on its bundled example (vignettes/bayesNMF_tutorial.pdf pp.8-13,
SURVEY.md section 4): learned rank 4 of 1:10, planted COSMIC signatures SBS58, SBS40, SBS26,
SBS2 recovered with MAP cosine 0.9993 / 0.9642 / 0.9993 / 0.9996."""
import numpy as np
import pytest

from tests.util import example_data, synth_counts

pytestmark = pytest.mark.gpu


def _cos(A, B):
    return (A.T @ B) / np.outer(np.linalg.norm(A, axis=0), np.linalg.norm(B, axis=0))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_tutorial_example_rank_and_signatures(built_lib, seed):
    """bayesNMF(data$M, 1:10) with the defaults (Poisson-TruncNormal + MH, SBFI), shortened
    schedule (maxiters 2000 instead of 5000, MAP_over 500) so three seeds run in seconds."""
    from bayesnmf_b200 import bayesNMF, new_convergence_control
    M, Ptrue = example_data()
    cc = new_convergence_control(MAP_over=500, MAP_every=100, miniters=1000, maxiters=2000)
    s = bayesNMF(M, np.arange(1, 11), convergence_control=cc, post_warmup=500, seed=seed)
    assert s.dims == dict(K=96, N=10, G=64)
    assert s.state["converged"]
    assert s.MAP["P"].shape == (96, 4), s.MAP["A_counts"]            # learned rank 4 (tutorial p.10)
    c = _cos(s.MAP["P"], Ptrue)
    assert sorted(c.argmax(axis=0)) == [0, 1, 2, 3]
    best = c.max(axis=0)                                               # columns: SBS58, SBS40, SBS26, SBS2
    assert (best[[0, 2, 3]] > 0.99).all() and best[1] > 0.93, best
    np.testing.assert_allclose(s.MAP["P"].sum(axis=0), 1.0, rtol=1e-9)
    # reconstruction and the schema the unchanged R code reads
    Mhat = s.MAP["P"] @ s.MAP["E"]
    assert np.corrcoef(Mhat.ravel(), M.ravel())[0, 1] > 0.99
    sm = s.state["sample_metrics"]
    assert len(sm["iter"]) == s.state["iter"] and sm["iter"][0] == 1 and sm["iter"][-1] == s.state["iter"]
    assert s.state["iter"] == s.state["converged_iter"] + 500          # post_warmup MH samples
    acc = np.asarray(sm["P_mean_acceptance_rate"])
    warm = acc[1:s.state["converged_iter"]]       # NaN while a newly included signature has no rate yet (NA-filled matrix)
    assert np.all(warm[~np.isnan(warm)] == 1.0) and 0.0 < acc[-1] < 1.0               # R/sample_Pn.R:201-204
    lo, up = s.credible_intervals["P"]["lower"], s.credible_intervals["P"]["upper"]
    assert lo.shape == (96, 4) and np.all(lo <= s.MAP["P"] + 1e-12) and np.all(s.MAP["P"] <= up + 1e-12)
    assert s.credible_intervals["E"]["lower"].shape == (4, 64)
    assert s.state["MAP_metrics"][-1]["rank"] == 4
    # assign_signatures_ensemble_ (tutorial p.12-13): the four planted COSMIC signatures are recovered
    from tests.util import cosmic
    C, names, _ = cosmic()
    rc = s.assign_signatures_ensemble(C, names)
    planted = {names[int(np.argmax(_cos(Ptrue[:, [j]], C)))] for j in range(4)}
    assert {a["sig_ref"] for a in rc["assignments"]} == planted
    assert all(a["lower_cosine"] <= a["upper_cosine"] for a in rc["assignments"])
    s.close()


def test_fixed_rank_poisson_gamma(built_lib):
    """BASELINE config 1 (Poisson-Gamma, fixed rank 5, 96 x 100, 1000 iterations)."""
    from bayesnmf_b200 import bayesNMF, new_convergence_control
    M, P, E = synth_counts(96, 100, 5, 4000.0, seed=3)
    cc = new_convergence_control(MAP_over=200, MAP_every=100, miniters=300, maxiters=1000)
    s = bayesNMF(M, 5, likelihood="poisson", prior="gamma", convergence_control=cc, seed=5)
    assert not s.specs["MH"] and s.state["iter"] <= 1000
    c = _cos(s.MAP["P"], P)
    assert sorted(c.argmax(axis=0)) == [0, 1, 2, 3, 4] and (c.max(axis=0) > 0.9).all(), c.max(axis=0)
    assert s.state["MAP_metrics"][-1]["RMSE"] < 2.0 * np.sqrt(M.mean())
    s.close()


def test_bic_fan_out(built_lib):
    """rank_method = "BIC": one fixed-rank sampler per rank (R/bayesNMF.R:66-127)."""
    from bayesnmf_b200 import bayesNMF, new_convergence_control
    M, _, _ = synth_counts(96, 60, 3, 3000.0, seed=4)
    cc = new_convergence_control(MAP_over=100, MAP_every=50, miniters=200, maxiters=400)
    out = bayesNMF(M, [2, 3, 4], likelihood="normal", prior="exponential", rank_method="BIC", convergence_control=cc, seed=2)
    assert [r["BIC"] for r in out["results"]] == sorted(r["BIC"] for r in out["results"])
    assert out["best_rank"] in (3, 4) and out["sampler"].dims["N"] == out["best_rank"]
    out["sampler"].close()


@pytest.mark.parametrize("model", ["default", "fixed"])
def test_run_behind_the_abi_equals_host_loop(built_lib, model):
    """bnmf_run (run_gibbs_sampler + check_convergence_ + the MAP-metric windows behind the ABI,
    SURVEY.md section 8 f-4) makes the same decisions as the host mirror of the R loop: same stop
    iteration and reason, same sample_metrics rows, same MAP_metrics at every check."""
    from bayesnmf_b200 import bayesNMF_sampler, new_convergence_control
    M, _ = example_data()
    cc = new_convergence_control(MAP_over=300, MAP_every=100, miniters=600, maxiters=1500)
    if model == "default":
        kw = dict(rank=np.arange(1, 9), post_warmup=300)                          # Poisson-TruncNormal + MH, SBFI
    else:
        kw = dict(rank=4, likelihood="poisson", prior="gamma")                     # fixed rank, latent counts
    a = bayesNMF_sampler(M, convergence_control=cc, seed=5, **kw).run_gibbs_sampler()
    b = bayesNMF_sampler(M, convergence_control=cc, seed=5, **kw)
    r = b._h.run(cc, post_warmup=b.specs.get("post_warmup", 0))
    assert r["converged"] == int(a.state["converged"]) and r["why"] == a.state.get("why")
    assert r["converged_iter"] == a.state.get("converged_iter", 0) and r["iter"] == a.state["iter"]
    sm = a.state["sample_metrics"]
    np.testing.assert_array_equal(r["metrics"][:, 0], np.asarray(sm["iter"])[1:])
    np.testing.assert_allclose(r["metrics"][:, 4], np.asarray(sm["logposterior"])[1:], rtol=0, atol=0)
    assert len(r["MAP_metrics"]) == len(a.state["MAP_metrics"])
    for x, y in zip(r["MAP_metrics"], a.state["MAP_metrics"]):
        for key in ("iter", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "MAP_A_counts", "mean_temp"):
            np.testing.assert_allclose(x[key], y[key], rtol=1e-12, err_msg=key)
    assert r["best_iter"] == a.state.get("best_iter", 0)
    a.close(); b.close()
