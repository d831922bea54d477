"""GPU parity tests of the Poisson (non-MH) path against the oracle, through the C ABI."""
import os

import numpy as np
import pytest

from tests.util import synth_counts

pytestmark = pytest.mark.gpu


def _handle(M, N, prior="gamma", seed=3, **kw):
    from bayesnmf_b200 import Handle
    return Handle(M, N, likelihood="poisson", prior=prior, MH=False, seed=seed, **kw)


@pytest.mark.parametrize("K,G,N,mu", [(96, 64, 5, 4000.0), (96, 100, 20, 4000.0), (130, 77, 7, 300.0),
                                      (33, 1, 3, 50.0), (96, 300, 40, 100.0), (200, 45, 64, 2000.0)])
def test_zstat_bit_exact(built_lib, K, G, N, mu):
    """Latent-count margins are bit-exact under shared Philox draws (north star level 1)."""
    from oracle.gibbs import sample_Z_stats
    rng = np.random.default_rng(K + G + N)
    M, _, _ = synth_counts(K, G, N, mu, seed=1)
    P = rng.gamma(1.0, 0.02, size=(K, N))
    E = rng.gamma(1.0, mu / N, size=(N, G))
    A = np.ones(N)
    if N > 2:
        A[1] = 0                      # an excluded signature: its Z must be 0
    if G > 3:
        E[:, 2] = 0.0                 # all-zero probabilities: the whole column of Z is 0
    M[0, 0] = 0
    h = _handle(M, N)
    h.set_state("P", P); h.set_state("E", E); h.set_state("A", A)
    ms = h.sample_z(7)
    SP, SE = h.get_state("SP"), h.get_state("SE")
    oSP, oSE, Z = sample_Z_stats(M, P, A, E, seed=3, it=7, return_Z=True)
    assert np.array_equal(SP, oSP)
    assert np.array_equal(SE, oSE)
    # invariants of R/sample_params.R:253-265
    Mhat = (P * A) @ E
    assert np.array_equal(Z.sum(axis=1)[Mhat > 0], M[Mhat > 0])
    assert SP.sum() == SE.sum() == M[Mhat > 0].sum()
    if N > 2:
        assert SP[:, 1].sum() == 0 and SE[1].sum() == 0
    assert ms > 0


def test_zstat_changes_with_iteration_and_seed(built_lib):
    M, P, E = synth_counts(96, 40, 5, 500.0, seed=2)
    h = _handle(M, 5, seed=11)
    h.set_state("P", P); h.set_state("E", E); h.set_state("A", np.ones(5))
    h.sample_z(2); a = h.get_state("SP")
    h.sample_z(3); b = h.get_state("SP")
    h.sample_z(2); c = h.get_state("SP")
    assert np.array_equal(a, c) and not np.array_equal(a, b)


@pytest.mark.parametrize("prior", ["gamma", "exponential"])
@pytest.mark.parametrize("K,G,N", [(96, 64, 5), (50, 37, 3)])
def test_iteration_parity(built_lib, prior, K, G, N):
    """Conditional draws, prior parameters and metrics within 1e-6 relative (fp64),
    latent-count margins bit-exact, over several full Gibbs iterations."""
    from oracle.gibbs import OracleSampler
    M, _, _ = synth_counts(K, G, N, 2000.0, seed=5)
    o = OracleSampler(M, N, "poisson", prior, MH=False, seed=9)
    h = _handle(M, N, prior=prior, seed=9)
    for k, v in o.hyper.items():
        h.set_hyper(k, v[0, 0])
    row = h.init_from_prior()
    names = ["P", "E"] + (["Alpha_p", "Beta_p", "Alpha_e", "Beta_e"] if prior == "gamma" else ["Lambda_p", "Lambda_e"])

    def check(tag):
        for nm in names:
            ref = o.params[nm] if nm in o.params else o.prior_params[nm]
            np.testing.assert_allclose(h.get_state(nm), ref, rtol=1e-6, atol=1e-300, err_msg=f"{tag} {nm}")
        assert np.array_equal(h.get_state("SP"), o.SP), tag
        assert np.array_equal(h.get_state("SE"), o.SE), tag

    check("init")
    om = o.metrics[0]
    for key in ("RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank"):
        np.testing.assert_allclose(row[key], om[key], rtol=1e-6, err_msg=f"init {key}")
    for it in range(4):
        om = o.step()
        met = h.step(1)["metrics"][0]
        check(f"iter {o.iter}")
        from bayesnmf_b200._lib import METRIC_NAMES
        got = dict(zip(METRIC_NAMES, met))
        assert got["iter"] == o.iter
        for key in ("RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "temp"):
            np.testing.assert_allclose(got[key], om[key], rtol=1e-6, err_msg=f"iter {o.iter} {key}")


def test_step_chunks_equal_single_steps(built_lib):
    """bnmf_step(n) == n x bnmf_step(1): the draw streams depend on the iteration only."""
    M, _, _ = synth_counts(96, 50, 4, 1000.0, seed=7)
    a = _handle(M, 4, seed=21); a.init_from_prior()
    b = _handle(M, 4, seed=21); b.init_from_prior()
    ra = a.step(6, want_P=True)
    rows = [b.step(1, want_P=True) for _ in range(6)]
    np.testing.assert_array_equal(ra["metrics"], np.concatenate([r["metrics"] for r in rows]))
    np.testing.assert_array_equal(ra["P"], np.concatenate([r["P"] for r in rows]))
    np.testing.assert_array_equal(a.get_state("E"), b.get_state("E"))


@pytest.mark.parametrize("lik,prior,MH", [("poisson", "gamma", False), ("poisson", "truncnormal", True), ("normal", "exponential", False)])
def test_default_hyperparameters(built_lib, lik, prior, MH):
    """Without bnmf_set_hyper the library uses the reference's data-dependent defaults
    (get_default_*_hyperprior_params_, R/setup.R:123-181), like the oracle."""
    from bayesnmf_b200 import Handle
    from oracle.gibbs import OracleSampler
    M, _, _ = synth_counts(96, 30, 4, 800.0, seed=6)
    o = OracleSampler(M, 4, lik, prior, MH=MH, seed=2)
    h = Handle(M, 4, likelihood=lik, prior=prior, MH=MH, seed=2)
    h.init_from_prior()
    o.step(); h.step(1)
    np.testing.assert_allclose(h.get_state("P"), o.params["P"], rtol=1e-6, atol=1e-300)
    np.testing.assert_allclose(h.get_state("E"), o.params["E"], rtol=1e-6, atol=1e-300)


@pytest.mark.parametrize("K,G,N,mu", [(96, 100, 20, 4000.0), (130, 77, 7, 300.0), (96, 300, 40, 100.0)])
def test_zstat_bit_exact_f32_state(built_lib, K, G, N, mu):
    """BNMF_F32: state and CDF in float32, 24 random bits per pick -- still bit-exact against
    the oracle's float32 restatement of the same arithmetic."""
    from oracle.gibbs import sample_Z_stats
    rng = np.random.default_rng(K + G + N)
    M, _, _ = synth_counts(K, G, N, mu, seed=1)
    P = rng.gamma(1.0, 0.02, size=(K, N)).astype(np.float32).astype(np.float64)
    E = rng.gamma(1.0, mu / N, size=(N, G)).astype(np.float32).astype(np.float64)
    A = np.ones(N); A[1] = 0
    h = _handle(M, N, precision="f32")
    h.set_state("P", P); h.set_state("E", E); h.set_state("A", A)
    h.sample_z(5)
    oSP, oSE = sample_Z_stats(M, P, A, E, seed=3, it=5, f32=True)
    assert np.array_equal(h.get_state("SP"), oSP) and np.array_equal(h.get_state("SE"), oSE)


@pytest.mark.parametrize("prior", ["gamma", "exponential"])
def test_iteration_parity_f32_state(built_lib, prior):
    """Full iterations with float32 state: every stored quantity equals the oracle's float32
    emulation to 1e-6 (identical floats up to draws that land on a rounding boundary), margins
    bit-exact, metrics to 1e-5 (they are formed from the float32 CDF total)."""
    from oracle.gibbs import OracleSampler
    K, G, N = 96, 64, 5
    M, _, _ = synth_counts(K, G, N, 2000.0, seed=5)
    o = OracleSampler(M, N, "poisson", prior, MH=False, seed=9, state="f32")
    h = _handle(M, N, prior=prior, seed=9, precision="f32")
    h.init_from_prior()
    names = ["P", "E"] + (["Alpha_p", "Beta_p", "Alpha_e", "Beta_e"] if prior == "gamma" else ["Lambda_p", "Lambda_e"])
    for it in range(4):
        om = o.step()
        met = h.step(1)["metrics"][0]
        for nm in names:
            ref = o.params[nm] if nm in o.params else o.prior_params[nm]
            np.testing.assert_allclose(h.get_state(nm), ref, rtol=1e-6, atol=1e-300, err_msg=f"iter {o.iter} {nm}")
        assert np.array_equal(h.get_state("SP"), o.SP) and np.array_equal(h.get_state("SE"), o.SE)
        for j, key in ((1, "RMSE"), (2, "KL"), (3, "loglikelihood"), (4, "logposterior")):
            np.testing.assert_allclose(met[j], om[key], rtol=1e-5, err_msg=f"iter {o.iter} {key}")


def test_count_limits_rejected(built_lib):
    """Inputs the kernels cannot represent exactly fail loudly at creation."""
    from bayesnmf_b200 import BnmfError
    M = np.ones((8, 4))
    for bad in (-1.0, 0.5, 2.0 ** 24 + 1, np.nan):
        Mb = M.copy(); Mb[3, 2] = bad
        with pytest.raises(BnmfError):
            _handle(Mb, 2)
    h = _handle(np.full((8, 4), 2.0 ** 24), 2)      # the largest supported cell
    h.close()


@pytest.mark.parametrize("lik,prior,MH,rank", [("poisson", "gamma", False, False), ("poisson", "truncnormal", True, True), ("normal", "exponential", False, False)])
def test_recycled_device_blocks_give_the_same_chain(built_lib, lik, prior, MH, rank):
    """bnmf_destroy keeps a handle's device blocks for the next handle of the process (include/bnmf.h:
    bnmf_release_cached_memory); a sampler built on blocks another chain left dirty runs the same
    chain as one built on fresh memory."""
    from bayesnmf_b200 import Handle, release_cached_memory
    M, _, _ = synth_counts(96, 300, 5, 800.0, seed=3)

    def chain(seed):
        h = Handle(M.astype(np.float64), 5, likelihood=lik, prior=prior, MH=MH, learning_rank=rank, seed=seed, ring_cap=4)
        h.init_from_prior()
        out = h.step(5, want_P=True, want_A=True)
        E = h.get_state("E")
        Pm, Em, Am, nm = h.get_map(4)
        h.close()
        return out["metrics"], out["P"], E, Pm, Em

    release_cached_memory()
    fresh = chain(11)
    chain(12)                       # leaves its state in the cached blocks
    recycled = chain(11)
    release_cached_memory()
    again = chain(11)
    for a, b, c in zip(fresh, recycled, again):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)


def test_block_cache_eviction(built_lib):
    """BNMF_CACHE_MB bounds the cache of device blocks; when it is full the oldest blocks are handed
    back to the driver.  (The limit is read once per process: run in a child.)"""
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import numpy as np
        from tests.util import synth_counts
        from bayesnmf_b200 import Handle
        M, _, _ = synth_counts(96, 400, 5, 800.0, seed=3)
        def make(seed):
            h = Handle(M.astype(np.float64), 5, seed=seed); h.init_from_prior(); return h
        a = make(1); ra = a.step(4)["metrics"]; a.close()          # one 64 MiB slab goes to the cache (= the limit)
        b, c = make(2), make(3); b.step(2); c.step(2); b.close(); c.close()   # the second slab evicts the first
        d = make(1); rd = d.step(4)["metrics"]; d.close()
        assert np.array_equal(ra, rd)
        print("ok")
    """)
    env = dict(os.environ, BNMF_CACHE_MB="64", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=env, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
