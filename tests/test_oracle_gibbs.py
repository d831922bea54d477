"""CPU tests of the oracle (oracle/gibbs.py): invariants of the reference's updates,
independent re-derivations of the conditional parameters, the reference's end-to-end
known answer on its bundled example data, and shard-invariance of the statistics."""
import numpy as np
import pytest
from scipy import stats

from oracle import gibbs as og
from oracle.gibbs import OracleSampler, sample_Z_stats
from tests.util import cosmic, example_data, synth_counts


def test_default_hyperpriors_follow_setup_R():
    # R/setup.R:123-181
    h = og.default_hyperprior_params("truncnormal", 41.67, 10)
    assert h["m_p"] == 0 and h["a_e"] == 11 and np.isclose(h["s_p"], np.sqrt(4.167)) and np.isclose(h["b_p"], np.sqrt(10))
    h = og.default_hyperprior_params("exponential", 16.0, 4)
    assert h["a_p"] == 20 and h["b_e"] == 40
    h = og.default_hyperprior_params("gamma", 16.0, 4)
    assert (h["a_p"], h["b_p"], h["c_p"], h["d_p"]) == (20, 10, 40, 10)


def test_temperature_schedule_shape():
    # R/utils.R:307-332: ramp over the first n_temp entries, then 1
    t = og.get_temp_sched(5000, 1000)
    assert len(t) == 5000 and t[0] == 0 and (np.diff(t[:1000]) >= 0).all() and (t[1000:] == 1).all()
    assert t[:1000].max() < 1.0
    t = og.get_temp_sched(300, 100)          # schedule longer than n_temp: sorted sub-sample
    assert len(t) == 300 and (np.diff(t[:100]) >= 0).all() and (t[100:] == 1).all()


def test_Z_invariants_and_distribution():
    """sample_Zkg (R/sample_params.R:253-265): rows sum to M where Mhat > 0, zero for
    excluded signatures, and the picks follow probs/sum(probs)."""
    rng = np.random.default_rng(0)
    K, N, G = 6, 4, 5
    P = rng.gamma(1.0, 1.0, (K, N)); E = rng.gamma(1.0, 1.0, (N, G))
    A = np.array([1, 0, 1, 1.0])
    M = np.full((K, G), 4000.0)
    E[:, 3] = 0
    SP, SE, Z = sample_Z_stats(M, P, A, E, seed=1, it=2, return_Z=True)
    Mhat = (P * A) @ E
    assert (Z.sum(axis=1)[Mhat > 0] == 4000).all() and (Z[:, :, 3] == 0).all() and (Z[:, 1, :] == 0).all()
    assert np.array_equal(SP, Z.sum(axis=2)) and np.array_equal(SE, Z.sum(axis=0))
    prob = (P * A)[:, :, None] * E[None]
    for k, g in [(0, 0), (5, 4), (2, 1)]:
        p = prob[k, :, g] / prob[k, :, g].sum()
        keep = p > 0
        assert stats.chisquare(Z[k, keep, g], 4000 * p[keep]).pvalue > 1e-4


def test_Z_shard_invariance():
    """Statistics of a G-sharded run sum to the unsharded ones bit-for-bit (draws are
    addressed by the global genome index)."""
    M, P, E = synth_counts(96, 40, 5, 300.0, seed=3)
    A = np.ones(5)
    SP, SE = sample_Z_stats(M, P, A, E, seed=5, it=3)
    parts = [sample_Z_stats(M[:, a:b], P, A, E[:, a:b], seed=5, it=3, g0=a) for a, b in [(0, 13), (13, 14), (14, 40)]]
    assert np.array_equal(sum(p[0] for p in parts), SP)
    assert np.array_equal(np.concatenate([p[1] for p in parts], axis=1), SE)


@pytest.mark.parametrize("prior", ["truncnormal", "exponential"])
def test_proposal_moments_match_closed_form(prior):
    """get_mu_sigmasq_{Pn,En}_normal (R/sample_Pn.R:132-187, R/sample_En.R:131-184) with
    sigmasq = Mhat, against the algebraically reduced form
    num1 = sum E M/Mhat - sum E + P_kn den  (SURVEY.md section 8a row 6)."""
    M, _, _ = synth_counts(96, 30, 4, 800.0, seed=2)
    o = OracleSampler(M, 4, "poisson", prior, seed=4)
    o.step()
    P, E, pp = o.params["P"], o.params["E"], o.prior_params
    Mhat = P @ E
    for n in range(4):
        mu, v = o.get_mu_sigmasq_Pn_normal(n, as_proposal=True)
        den = (E[n] ** 2 / Mhat).sum(axis=1)
        num1 = (E[n] * M / Mhat).sum(axis=1) - E[n].sum() + P[:, n] * den
        if prior == "exponential":
            np.testing.assert_allclose(mu, (num1 - pp["Lambda_p"][:, n]) / den, rtol=1e-9)
            np.testing.assert_allclose(v, 1 / den, rtol=1e-12)
        else:
            d2 = den + 1 / pp["Sigmasq_p"][:, n]
            np.testing.assert_allclose(mu, (num1 + pp["Mu_p"][:, n] / pp["Sigmasq_p"][:, n]) / d2, rtol=1e-9)
        mu, v = o.get_mu_sigmasq_En_normal(n, as_proposal=True)
        den = (P[:, n][:, None] ** 2 / Mhat).sum(axis=0)
        num1 = (P[:, n][:, None] * M / Mhat).sum(axis=0) - P[:, n].sum() + E[n] * den
        if prior == "exponential":
            np.testing.assert_allclose(mu, (num1 - pp["Lambda_e"][n]) / den, rtol=1e-9)


def test_poisson_gamma_conditionals():
    """sample_Pn_poisson / sample_En_poisson (R/sample_Pn.R:98-120, R/sample_En.R:97-119):
    the draws are Gamma(Alpha + S, Beta + margins) of the previous iteration's Z."""
    from oracle import draws as dr, philox as px
    M, _, _ = synth_counts(20, 15, 3, 200.0, seed=1)
    o = OracleSampler(M, 3, "poisson", "gamma", seed=6)
    SP, SE, E0 = o.SP.copy(), o.SE.copy(), o.params["E"].copy()
    o.step()
    pp = o.prior_params
    rs = np.rint(E0 * 2 ** 24).sum(axis=1) / 2 ** 24
    for n in range(3):
        want = dr.gamma_draw(6, 2, px.PUR_P, o._cells_p(n), pp["Alpha_p"][:, n] + SP[:, n], pp["Beta_p"][:, n] + rs[n])
        assert np.array_equal(o.params["P"][:, n], want)
        want = dr.gamma_draw(6, 2, px.PUR_E, o._cells_e(n), pp["Alpha_e"][n] + SE[n], pp["Beta_e"][n] + o.params["P"][:, n].sum())
        assert np.array_equal(o.params["E"][n], want)


def test_MH_warmup_accepts_everything_then_rejects():
    # R/sample_Pn.R:201-204 / R/sample_En.R:198-201
    M, _, _ = synth_counts(30, 20, 3, 500.0, seed=1)
    o = OracleSampler(M, 3, "poisson", "truncnormal", seed=2)
    m = o.step()
    assert m["P_mean_acceptance_rate"] == 1.0 and m["E_mean_acceptance_rate"] == 1.0
    o.converged = True
    for _ in range(5):
        m = o.step()
    assert 0.0 < m["P_mean_acceptance_rate"] < 1.0 and 0.0 < m["E_mean_acceptance_rate"] < 1.0


def test_metrics_row():
    M, _, _ = synth_counts(30, 20, 3, 500.0, seed=1)
    o = OracleSampler(M, 3, "poisson", "gamma", seed=2)
    m = o.step()
    Mhat = o.params["P"] @ o.params["E"]
    assert m["iter"] == 2 and m["rank"] == 3 and m["n_params"] == 3 * (20 + 30)
    np.testing.assert_allclose(m["BIC"], -2 * m["loglikelihood"] + m["n_params"] * np.log(20))
    np.testing.assert_allclose(m["loglikelihood"], stats.poisson.logpmf(M, np.maximum(Mhat, 1e-6)).sum(), rtol=1e-10)
    np.testing.assert_allclose(m["RMSE"], np.sqrt(((M - Mhat) ** 2).mean()))


def test_rank_learning_prunes_extra_signatures():
    """SBFI (R/sample_params.R:101-166) on data with 2 planted signatures and N = 5."""
    C = cosmic()[0]
    rng = np.random.default_rng(0)
    P = C[:, [1, 4]]
    E = rng.gamma(2.0, 800.0, (2, 40))
    M = rng.poisson(P @ E).astype(float)
    o = OracleSampler(M, 5, "poisson", "truncnormal", learning_rank=True, seed=3,
                      temperature_schedule=og.get_temp_sched(400, 150))
    for _ in range(300):
        m = o.step()
    assert m["rank"] == 2


def test_reference_example_known_answer():
    """The reference's one end-to-end known answer (vignettes/bayesNMF_tutorial.pdf pp.10-13,
    SURVEY.md section 4): on inst/extdata/example_data.rds the planted signatures
    SBS58/SBS40/SBS26/SBS2 are recovered with MAP cosine 0.9993/0.9642/0.9993/0.9996."""
    M, Ptrue = example_data()
    o = OracleSampler(M, 4, "poisson", "truncnormal", seed=1)
    acc = 0
    for i in range(300):
        o.step()
        if i >= 150:
            acc = acc + o.params["P"]
    c = (acc.T @ Ptrue) / np.outer(np.linalg.norm(acc, axis=0), np.linalg.norm(Ptrue, axis=0))
    best = c.max(axis=0)
    assert sorted(c.argmax(axis=0)) == [0, 1, 2, 3]
    assert (best[[0, 2, 3]] > 0.995).all() and best[1] > 0.93     # SBS40 is the flat, hard one


def test_get_mode_and_MAP():
    """get_mode ties go to the alphabetically first pattern (table() order + stable sort,
    R/helpers.R:63-79); get_MAP averages the renormalised matching samples (R/utils.R:236-250)."""
    rng = np.random.default_rng(0)
    A = [np.array([1, 0, 1]), np.array([1, 1, 1]), np.array([1, 1, 1]), np.array([1, 0, 1])]
    mode, idx = og.get_mode(A)
    assert list(mode) == [1, 0, 1] and idx == [0, 3]
    P = [rng.gamma(1.0, 1.0, (6, 3)) for _ in A]
    E = [rng.gamma(1.0, 5.0, (3, 4)) for _ in A]
    Pm, Em, Am, idx = og.get_MAP(P, E, A)
    want_P = (P[0] / P[0].sum(0) + P[3] / P[3].sum(0)) / 2
    want_E = (E[0] * P[0].sum(0)[:, None] + E[3] * P[3].sum(0)[:, None]) / 2
    np.testing.assert_allclose(Pm, want_P, rtol=1e-14)
    np.testing.assert_allclose(Em, want_E, rtol=1e-14)
    np.testing.assert_allclose(Pm.sum(0), 1.0)
    for p, e in zip(P, E):                      # renormalize keeps the product
        rp, re = og.renormalize(p, e)
        np.testing.assert_allclose(rp @ re, p @ e, rtol=1e-12)


def test_z_thresholds_contract_edges():
    """Integer pick thresholds of the latent-count draw (csrc/bnmf_poisson.cuh, oracle z_thresholds):
    monotone, 0 for leading zero-probability categories, saturated at 2^32 - 1 once the CDF reaches
    the total (so a trailing zero-probability category is never picked), NaN -> 0 when the total is 0."""
    from oracle.gibbs import z_cdf, z_thresholds, sample_Z_stats
    P = np.array([[0.0, 0.2, 0.0, 0.8, 0.0]])                 # K = 1, N = 5
    E = np.ones((5, 3)); E[:, 2] = 0.0                        # third genome: all probabilities zero
    A = np.ones(5)
    cdf = z_cdf(P, A, E)
    thr = z_thresholds(cdf, 5)
    assert thr.shape == (1, 4, 3)
    assert (np.diff(thr.astype(np.int64), axis=1) >= 0).all()
    assert thr[0, 0, 0] == 0                                  # category 0 has no mass: always skipped
    assert thr[0, 1, 0] == thr[0, 2, 0] == np.uint64(int(0.2 * 2 ** 32))
    assert thr[0, 3, 0] == 0xFFFFFFFF                         # CDF complete: category 4 unreachable
    assert (thr[0, :, 2] == 0).all()                          # 0/0 -> NaN -> 0 (cell has no picks anyway)
    M = np.array([[1000, 7, 5]])
    SP, SE, Z = sample_Z_stats(M, P, A, E, seed=1, it=2, return_Z=True)
    assert Z[0, [0, 2, 4], :].sum() == 0 and Z[0, :, 2].sum() == 0
    assert Z[0, :, 0].sum() == 1000 and abs(Z[0, 1, 0] - 200) < 60
    # float32 state: same construction in float32 arithmetic
    thr32 = z_thresholds(z_cdf(P, A, E, f32=True), 5, f32=True)
    assert thr32[0, 3, 0] == 0xFFFFFFFF and thr32[0, 0, 0] == 0
