"""GPU parity tests of the sweep-based models (Normal likelihood, Poisson + MH proposals,
rank learning) against the oracle, through the C ABI.  Tolerance: 1e-6 relative on every
conditional draw, prior parameter, acceptance rate and metric (north star level 2)."""
import numpy as np
import pytest

from tests.util import synth_counts

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _pair(M, N, lik, prior, MH, seed=5, learning_rank=False, rank_method="SBFI", temps=None, **okw):
    from bayesnmf_b200 import Handle
    from oracle.gibbs import OracleSampler
    o = OracleSampler(M, N, lik, prior, MH=MH, seed=seed, learning_rank=learning_rank, rank_method=rank_method,
                      temperature_schedule=temps, **okw)
    h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, seed=seed, learning_rank=learning_rank, rank_method=rank_method)
    for k, v in o.hyper.items():
        h.set_hyper(k, v[0, 0])
    if temps is not None:
        h.set_temperature_schedule(temps)
    return o, h


def _names(lik, prior, MH):
    nm = ["P", "E"]
    nm += {"truncnormal": ["Mu_p", "Sigmasq_p", "Mu_e", "Sigmasq_e"], "exponential": ["Lambda_p", "Lambda_e"],
           "gamma": ["Alpha_p", "Beta_p", "Alpha_e", "Beta_e"]}[prior]
    if lik == "normal":
        nm.append("sigmasq")
    return nm


def _check_state(o, h, names, tag, MH):
    for nm in names:
        ref = o.params[nm] if nm in o.params else o.prior_params[nm]
        np.testing.assert_allclose(h.get_state(nm), ref, rtol=RTOL, atol=1e-300, err_msg=f"{tag} {nm}")
    np.testing.assert_array_equal(h.get_state("A"), o.params["A"], err_msg=f"{tag} A")
    if MH:
        np.testing.assert_allclose(h.get_state("P_acceptance_rate"), o.acc["P"], rtol=RTOL, atol=1e-12, err_msg=f"{tag} P acc")
        np.testing.assert_allclose(h.get_state("E_acceptance_rate"), o.acc["E"], rtol=RTOL, atol=1e-12, err_msg=f"{tag} E acc")


def _check_row(row, om, tag, MH):
    from bayesnmf_b200._lib import METRIC_NAMES
    got = dict(zip(METRIC_NAMES, row))
    keys = ["iter", "RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "temp"]
    if MH:
        keys += ["P_mean_acceptance_rate", "E_mean_acceptance_rate"]
    for key in keys:
        np.testing.assert_allclose(got[key], om[key], rtol=RTOL, atol=1e-9, err_msg=f"{tag} {key}", equal_nan=True)


def _mhat_atol(o, lik):
    """Normal likelihood: Mhat comes from the tensor cores (csrc/bnmf_tc.cuh) with the inputs rounded to fixed
    point (40 bits) per row of P / column of E: |dMhat[k,g]| <= N 2^-40 max_n P[k,n] A_n x 2 max_n E[n,g] x 2 (the scales are the
    next powers of two).  The Poisson models keep the fp64 kernel."""
    if lik != "normal":
        return 1e-9
    P, E, A = o.params["P"], o.params["E"], o.params["A"]
    return float(o.N * 2.0 ** -40 * 4.0 * (P * A[None, :]).max() * E.max()) + 1e-9


CASES = [("poisson", "truncnormal", True), ("poisson", "exponential", True),
         ("normal", "truncnormal", False), ("normal", "exponential", False)]


@pytest.mark.parametrize("lik,prior,MH", CASES)
@pytest.mark.parametrize("K,G,N", [(96, 64, 5), (50, 37, 3), (200, 300, 7)])
def test_sweep_iteration_parity(built_lib, lik, prior, MH, K, G, N):
    M, _, _ = synth_counts(K, G, N, 2000.0, seed=4)
    if lik == "normal":
        M = M + np.random.default_rng(0).normal(0, 2.0, M.shape)      # real-valued data
    o, h = _pair(M, N, lik, prior, MH)
    row = h.init_from_prior()
    names = _names(lik, prior, MH)
    _check_state(o, h, names, "init", MH)
    _check_row([row[k] for k in row], o.metrics[0], "init", MH)
    for it in range(3):
        om = o.step()
        met = h.step(1)["metrics"][0]
        _check_state(o, h, names, f"iter {o.iter}", MH)
        _check_row(met, om, f"iter {o.iter}", MH)
    if MH:                                      # the real accept step (R/sample_Pn.R:206-247)
        o.converged = True
        for it in range(3):
            om = o.step()
            met = h.step(1, converged=True)["metrics"][0]
            _check_state(o, h, names, f"MH iter {o.iter}", MH)
            _check_row(met, om, f"MH iter {o.iter}", MH)
        acc = h.get_state("P_acceptance_rate")
        assert 0.0 < acc.mean() < 1.0
    np.testing.assert_allclose(h.get_state("Mhat"), o.get_Mhat(), rtol=1e-9, atol=_mhat_atol(o, lik))


@pytest.mark.parametrize("lik,prior,MH,method", [("poisson", "truncnormal", True, "SBFI"), ("poisson", "exponential", True, "BFI"),
                                                 ("normal", "truncnormal", False, "SBFI"), ("poisson", "gamma", False, "SBFI")])
def test_rank_learning_parity(built_lib, lik, prior, MH, method):
    """R, A sweep (R/sample_params.R:101-241) with a temperature ramp: identical inclusion
    indicators, expected rank and metrics, iteration by iteration."""
    from oracle.gibbs import get_temp_sched
    K, G, N = 96, 48, 6
    M, _, _ = synth_counts(K, G, 3, 1500.0, seed=8)
    temps = get_temp_sched(60, 25)
    o, h = _pair(M, N, lik, prior, MH, learning_rank=True, rank_method=method, temps=temps, seed=12)
    h.init_from_prior()
    names = _names(lik, prior, MH)
    _check_state(o, h, names, "init", MH)
    assert h.get_state("R")[0] == o.params["R"]
    for it in range(30):
        om = o.step()
        met = h.step(1)["metrics"][0]
        np.testing.assert_array_equal(h.get_state("A"), o.params["A"], err_msg=f"iter {o.iter} A")
        assert h.get_state("R")[0] == o.params["R"], f"iter {o.iter} R"
        _check_row(met, om, f"iter {o.iter}", MH)
    _check_state(o, h, names, "final", MH)
    if not MH and lik == "poisson":
        assert np.array_equal(h.get_state("SP"), o.SP) and np.array_equal(h.get_state("SE"), o.SE)


def test_user_supplied_initial_values(built_lib):
    """samples[[name]][[1]] is exactly the user's init (vignettes/advanced.qmd:181-185)."""
    from bayesnmf_b200 import Handle
    from oracle.gibbs import OracleSampler
    M, P0, E0 = synth_counts(96, 30, 4, 1000.0, seed=1)
    o = OracleSampler(M, 4, "poisson", "truncnormal", seed=3, init_params={"P": P0.copy(), "E": E0.copy()})
    h = Handle(M, 4, likelihood="poisson", prior="truncnormal", MH=True, seed=3)
    for k, v in o.hyper.items():
        h.set_hyper(k, v[0, 0])
    h.set_state("P", P0); h.set_state("E", E0)
    row = h.init_from_prior(have=("P", "E"))
    np.testing.assert_array_equal(h.get_state("P"), P0)
    np.testing.assert_array_equal(h.get_state("E"), E0)
    np.testing.assert_allclose(row["loglikelihood"], o.metrics[0]["loglikelihood"], rtol=RTOL)
    om = o.step()
    met = h.step(1)["metrics"][0]
    np.testing.assert_allclose(h.get_state("P"), o.params["P"], rtol=RTOL)
    np.testing.assert_allclose(met[3], om["loglikelihood"], rtol=RTOL)


@pytest.mark.parametrize("lik,prior,MH", [("poisson", "truncnormal", True), ("normal", "exponential", False)])
def test_sweep_parity_f32_state(built_lib, lik, prior, MH):
    """BNMF_F32 state for the sweep models: conditional draws within 1e-4 relative of the oracle's
    float32-state emulation (north star: 1e-4 for fp32)."""
    from bayesnmf_b200 import Handle
    from oracle.gibbs import OracleSampler
    K, G, N = 96, 64, 4
    M, _, _ = synth_counts(K, G, N, 2000.0, seed=4)
    if lik == "normal":
        M = M + np.random.default_rng(0).normal(0, 2.0, M.shape)
        M = M.astype(np.float32).astype(np.float64)          # the data are stored in the state precision too
    o = OracleSampler(M, N, lik, prior, MH=MH, seed=6, state="f32")
    h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, seed=6, precision="f32")
    h.init_from_prior()
    for it in range(3):
        om = o.step()
        met = h.step(1)["metrics"][0]
        np.testing.assert_allclose(h.get_state("P"), o.params["P"], rtol=1e-4, atol=1e-7, err_msg=f"iter {o.iter} P")
        np.testing.assert_allclose(h.get_state("E"), o.params["E"], rtol=1e-4, atol=1e-5, err_msg=f"iter {o.iter} E")
        np.testing.assert_allclose(met[3], om["loglikelihood"], rtol=1e-4)


@pytest.mark.parametrize("K,G,N,prior,what", [
    (1536, 64, 40, "exponential", "C5-like: SBS1536 context, 18 KB columns in k_e_sweep, 40 sequential conditionals"),
    (6, 40000, 4, "truncnormal", "row-resident P sweep over a multi-block cluster (40,000 genomes per mutation type)"),
    (96, 4099, 15, "exponential", "ragged genome count, 15 signatures"),
])
def test_sweep_parity_large_shapes(built_lib, K, G, N, prior, what):
    """Oracle parity (not GPU-vs-GPU) at the shapes where the sweep kernels change their decomposition:
    Poisson + MH, warm-up iterations and the real accept step (converged = TRUE)."""
    M, _, _ = synth_counts(K, G, min(N, 8), 3000.0, seed=21)
    o, h = _pair(M, N, "poisson", prior, True, seed=31)
    row = h.init_from_prior()
    names = _names("poisson", prior, True)
    _check_state(o, h, names, "init", True)
    _check_row([row[k] for k in row], o.metrics[0], "init", True)
    for it in range(2):
        om = o.step()
        met = h.step(1)["metrics"][0]
        _check_state(o, h, names, f"{what}: iter {o.iter}", True)
        _check_row(met, om, f"iter {o.iter}", True)
    o.converged = True
    for it in range(2):
        om = o.step()
        met = h.step(1, converged=True)["metrics"][0]
        _check_state(o, h, names, f"{what}: MH iter {o.iter}", True)
        _check_row(met, om, f"MH iter {o.iter}", True)
    np.testing.assert_allclose(h.get_state("Mhat"), o.get_Mhat(), rtol=1e-9, atol=1e-9)
    h.close()


def test_normal_parity_c4_like(built_lib):
    """Normal likelihood through the Gram-matrix P sweep at a C4-like shape (N = 15, thousands of genomes)."""
    K, G, N = 96, 3001, 15
    M, _, _ = synth_counts(K, G, 8, 3000.0, seed=22)
    M = M + np.random.default_rng(1).normal(0, 2.0, M.shape)
    o, h = _pair(M, N, "normal", "truncnormal", False, seed=32)
    h.init_from_prior()
    names = _names("normal", "truncnormal", False)
    _check_state(o, h, names, "init", False)
    for it in range(3):
        om = o.step()
        met = h.step(1)["metrics"][0]
        _check_state(o, h, names, f"iter {o.iter}", False)
        _check_row(met, om, f"iter {o.iter}", False)
    h.close()
