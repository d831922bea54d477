"""Sample ring (record_sample, R/bayesNMF_sampler.R:651-672) and get_MAP_ on the device
(R/utils.R:194-288) against the oracle's restatement, through the C ABI."""
import numpy as np
import pytest

from tests.util import synth_counts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lik,prior,MH,learn", [("poisson", "truncnormal", True, True), ("poisson", "gamma", False, False),
                                                ("normal", "exponential", False, True)])
def test_ring_and_map(built_lib, lik, prior, MH, learn):
    from bayesnmf_b200 import BnmfError, Handle
    from oracle.gibbs import get_MAP, get_temp_sched
    K, G, N, cap = 96, 40, 5, 24
    M, _, _ = synth_counts(K, G, 3, 1500.0, seed=2)
    h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, learning_rank=learn, seed=4, ring_cap=cap)
    h.set_temperature_schedule(get_temp_sched(80, 30))     # hyperparameters: the library's defaults (R/setup.R:123-181)
    h.init_from_prior()
    Ps, Es, As = [h.get_state("P")], [h.get_state("E")], [h.get_state("A")]
    for _ in range(37):
        h.step(1)
        Ps.append(h.get_state("P")); Es.append(h.get_state("E")); As.append(h.get_state("A"))
    assert h.ring_count() == cap
    for ago in (0, 1, 7, cap - 1):       # update_list keeps the newest MAP_over samples (R/helpers.R:111-119)
        np.testing.assert_array_equal(h.get_sample("P", ago), Ps[-1 - ago])
        np.testing.assert_array_equal(h.get_sample("E", ago), Es[-1 - ago])
        np.testing.assert_array_equal(h.get_sample("A", ago), As[-1 - ago])
    with pytest.raises(BnmfError):
        h.get_sample("P", cap)
    for n_s in (cap, 10, 1):
        P_map, E_map, A_map, n_match = h.get_map(n_s)
        oP, oE, oA, idx = get_MAP(Ps[-n_s:], Es[-n_s:], As[-n_s:])
        assert n_match == len(idx)
        np.testing.assert_array_equal(A_map, oA)
        np.testing.assert_allclose(P_map, oP, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(E_map, oE, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(P_map.sum(axis=0), 1.0, rtol=1e-12)     # renormalised signatures
    with pytest.raises(BnmfError):
        h.get_map(cap + 1)
    # credible intervals (R/utils.R:264-287): quantile type 7 of the renormalised matching samples
    for n_s, (lo, hi) in ((cap, (0.025, 0.975)), (10, (0.1, 0.5)), (1, (0.025, 0.975))):
        _, _, oA, idx = get_MAP(Ps[-n_s:], Es[-n_s:], As[-n_s:])
        sel = [i for i in range(n_s) if np.array_equal(As[-n_s:][i], oA)]
        Pm = np.stack([Ps[-n_s:][i] for i in sel]); cs = Pm.sum(axis=1)
        Pr = Pm / cs[:, None, :]
        Er = np.stack([Es[-n_s:][i] for i in sel]) * cs[:, :, None]
        Pl, Ph, El, Eh, nm = h.get_credible_intervals(n_s, lo, hi)
        assert nm == len(sel)
        np.testing.assert_allclose(Pl, np.quantile(Pr, lo, axis=0), rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(Ph, np.quantile(Pr, hi, axis=0), rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(El, np.quantile(Er, lo, axis=0), rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(Eh, np.quantile(Er, hi, axis=0), rtol=1e-12, atol=1e-300)


def test_map_without_ring_fails(built_lib):
    from bayesnmf_b200 import BnmfError, Handle
    M, _, _ = synth_counts(96, 10, 3, 500.0, seed=0)
    h = Handle(M, 3, likelihood="poisson", prior="gamma", MH=False, seed=1)
    h.init_from_prior()
    with pytest.raises(BnmfError):
        h.get_map(1)


def test_assign_signatures_ensemble(built_lib):
    """assign_signatures_ensemble_ (R/postprocessing.R:175-341) on the retained samples against a host
    restatement with scipy's Hungarian solver: votes, winners, MAP cosine, credible interval of the
    samples' cosines; and the tutorial's known answer -- the planted COSMIC signatures win."""
    from scipy.optimize import linear_sum_assignment
    from bayesnmf_b200 import Handle
    from oracle.gibbs import get_MAP
    from tests.util import cosmic, example_data
    M, Ptrue = example_data()
    C, names, _ = cosmic()
    N, cap = 4, 40
    h = Handle(M, N, likelihood="poisson", prior="gamma", MH=False, seed=3, ring_cap=cap)
    h.init_from_prior()
    h.step(400)
    Ps = [h.get_sample("P", cap - 1 - i) for i in range(cap)]           # oldest first
    Es = [h.get_sample("E", cap - 1 - i) for i in range(cap)]
    As = [h.get_sample("A", cap - 1 - i) for i in range(cap)]
    r = h.assign_signatures(cap, C, credible_interval=0.9)
    P_map, _, A_map, idx = get_MAP(Ps, Es, As)
    keep = np.nonzero(A_map == 1)[0]
    assert r["n_match"] == len(idx) and np.array_equal(r["keep_sigs"], keep)
    cosf = lambda A, B: (A.T @ B) / np.outer(np.linalg.norm(A, axis=0), np.linalg.norm(B, axis=0))
    V = np.zeros((len(keep), C.shape[1]))
    sims = []
    for i in idx:
        S = cosf(Ps[i][:, keep], C)
        rows, cols = linear_sum_assignment(-S)
        V[rows, cols] += S[rows, cols]
        sims.append(S)
    prop = V / V.sum(axis=1, keepdims=True)
    np.testing.assert_allclose(r["votes"], prop, rtol=1e-10, atol=1e-15)
    win = prop.argmax(axis=1)
    assert np.array_equal(r["assignment"], win)
    np.testing.assert_allclose(r["MAP_cosine"], cosf(P_map[:, keep], C)[np.arange(len(keep)), win], rtol=1e-10)
    sc = np.stack([S[np.arange(len(keep)), win] for S in sims])
    np.testing.assert_allclose(r["lower_cosine"], np.quantile(sc, 0.05, axis=0), rtol=1e-10)
    np.testing.assert_allclose(r["upper_cosine"], np.quantile(sc, 0.95, axis=0), rtol=1e-10)
    # the planted signatures of the example (columns of Ptrue are COSMIC SBS58, SBS40, SBS26, SBS2)
    planted = {names[int(np.argmax(cosf(Ptrue[:, [j]], C)))] for j in range(Ptrue.shape[1])}
    assert {names[j] for j in r["assignment"]} == planted, (r["assignment"], planted)
    assert (r["MAP_cosine"] > 0.93).all() and (r["lower_cosine"] <= r["MAP_cosine"] + 0.02).all()
    h.close()
