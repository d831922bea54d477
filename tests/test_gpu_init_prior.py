"""init_prior_params (R/sample_priors.R:15-141) through the C ABI: a prior-parameter matrix the
user supplies is kept verbatim, except for the signatures whose column (P side) / row (E side)
holds an NA, which are drawn from the hyperprior; matrices that are not supplied are drawn whole.
`have_prior` is the bit mask of include/bnmf.h (BNMF_HAVE_PRIOR_*), the one r/bayesNMF_sampler_b200.R
builds from names(init_prior_params).  Checked against the oracle: 1e-6 relative."""
import numpy as np
import pytest

from tests.util import synth_counts

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _supplied(prior, K, N, G, rng):
    """Two supplied matrices per prior: one with an NA signature, one whole."""
    pos = lambda shp: rng.gamma(2.0, 1.5, size=shp) + 0.05           # noqa: E731
    if prior == "truncnormal":
        a, b = pos((K, N)), rng.normal(1.0, 0.5, size=(N, G))
        a[3, 1] = np.nan                                             # signature 1 of Sigmasq_p is redrawn
        return {"Sigmasq_p": a, "Mu_e": b}
    if prior == "exponential":
        a, b = pos((K, N)), pos((N, G))
        a[:, 0] = np.nan
        return {"Lambda_p": a, "Lambda_e": b}
    a, b = pos((K, N)), pos((N, G))
    b[2, 5] = np.nan                                                 # row 2 of Beta_e is redrawn
    return {"Alpha_p": a, "Beta_e": b}


@pytest.mark.parametrize("lik,prior,MH", [("poisson", "truncnormal", True), ("poisson", "exponential", True),
                                          ("poisson", "gamma", False), ("normal", "truncnormal", False)])
def test_init_prior_params_kept_and_na_signatures_drawn(built_lib, lik, prior, MH):
    from bayesnmf_b200 import Handle
    from oracle.gibbs import OracleSampler
    K, G, N = 96, 40, 4
    M, _, _ = synth_counts(K, G, N, 1500.0, seed=2)
    ipp = _supplied(prior, K, N, G, np.random.default_rng(11))
    o = OracleSampler(M, N, lik, prior, MH=MH, seed=9, init_prior_params={k: v.copy() for k, v in ipp.items()})
    h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, seed=9)
    for k, v in o.hyper.items():
        h.set_hyper(k, v[0, 0])
    for k, v in ipp.items():
        h.set_state(k, v)
    row = h.init_from_prior(have_prior=tuple(ipp))
    every = {"truncnormal": ["Mu_p", "Sigmasq_p", "Mu_e", "Sigmasq_e"], "exponential": ["Lambda_p", "Lambda_e"],
             "gamma": ["Alpha_p", "Beta_p", "Alpha_e", "Beta_e"]}[prior]
    for nm in every:
        got, ref = h.get_state(nm), o.prior_params[nm]
        assert np.isfinite(got).all(), nm
        np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-300, err_msg=nm)
        if nm in ipp:                        # supplied entries of signatures without NA: verbatim
            axis = 0 if nm.endswith("_p") else 1
            keep = ~np.isnan(ipp[nm]).any(axis=axis, keepdims=True) & np.ones(ipp[nm].shape, bool)
            assert keep.any() and not keep.all() or not np.isnan(ipp[nm]).any()
            np.testing.assert_array_equal(got[keep], ipp[nm][keep], err_msg=f"{nm} kept")
            drawn = ~keep
            if drawn.any():
                assert not np.array_equal(got[drawn], ipp[nm][drawn])
    np.testing.assert_allclose(h.get_state("P"), o.params["P"], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(h.get_state("E"), o.params["E"], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(row["logposterior"], o.metrics[0]["logposterior"], rtol=RTOL)
    om = o.step()
    met = h.step(1)["metrics"][0]
    np.testing.assert_allclose(h.get_state("P"), o.params["P"], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(met[4], om["logposterior"], rtol=RTOL)
