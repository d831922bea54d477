"""oracle/cpu_port.cpp (the C++ / OpenMP port bench.py times as the CPU arm) against the numpy oracle:
latent-count margins bit for bit, every draw and metric to 1e-10 -- two independent executors of the
same restatement of R/sample_params.R:253-265, R/sample_Pn.R:98-120, R/sample_En.R:97-119,
R/sample_priors.R:284-397, R/utils.R:412-471."""
import numpy as np
import pytest

from tests.util import synth_counts


@pytest.mark.parametrize("prior", ["gamma", "exponential"])
@pytest.mark.parametrize("K,G,N", [(96, 40, 5), (30, 17, 3)])
def test_cpu_port_equals_numpy_oracle(prior, K, G, N):
    from oracle.cpu_port import CpuPort
    from oracle.gibbs import OracleSampler
    M, _, _ = synth_counts(K, G, N, 900.0, seed=3)
    M[:, 1] = 0                                   # an empty genome
    o = OracleSampler(M, N, "poisson", prior, MH=False, seed=21)
    c = CpuPort(M, N, prior, seed=21)
    row = c.init_from_prior()
    names = ["P", "E"] + (["Alpha_p", "Beta_p", "Alpha_e", "Beta_e"] if prior == "gamma" else ["Lambda_p", "Lambda_e"])

    def check(tag, row, om):
        assert np.array_equal(c.get("SP"), o.SP), tag
        assert np.array_equal(c.get("SE"), o.SE), tag
        for nm in names:
            ref = o.params[nm] if nm in o.params else o.prior_params[nm]
            np.testing.assert_allclose(c.get(nm), ref, rtol=1e-10, atol=1e-300, err_msg=f"{tag} {nm}")
        for key in ("iter", "RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "temp"):
            np.testing.assert_allclose(row[key], om[key], rtol=1e-9, atol=1e-9, err_msg=f"{tag} {key}")

    check("init", row, o.metrics[0])
    for it in range(3):
        om = o.step()
        r = c.step(1)[0]
        check(f"iter {o.iter}", dict(zip(["iter", "RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "temp"], r)), om)
    c.close()


def test_cpu_port_shard_offsets():
    """g0 / G_total: a shard draws what the whole matrix draws for its genomes (Philox keyed by global cell)."""
    from oracle.cpu_port import CpuPort
    M, _, _ = synth_counts(96, 24, 4, 700.0, seed=5)
    hyper = {"a_p": 20.0, "b_p": 10.0, "c_p": 30.0, "d_p": 10.0, "a_e": 20.0, "b_e": 10.0, "c_e": 30.0, "d_e": 10.0}
    whole = CpuPort(M, 4, "gamma", seed=2, hyper=hyper)
    part = CpuPort(M[:, 8:], 4, "gamma", seed=2, hyper=hyper, g0=8, G_total=24)
    whole.init_from_prior(); part.init_from_prior()
    np.testing.assert_array_equal(part.get("E"), whole.get("E")[:, 8:])
    np.testing.assert_array_equal(part.get("SE"), whole.get("SE")[:, 8:])
    np.testing.assert_array_equal(part.get("P"), whole.get("P"))
