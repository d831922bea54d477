"""Properties of the CUDA path at the full sizes of BASELINE.json (where the numpy oracle would
take minutes), through the C ABI: conservation of the latent counts, independence of the draws
from the shard decomposition, and a bit-exact oracle check of a slice at an arbitrary genome
offset.  Plus: the two implementations of the P sweep (cluster-resident rows, pass per signature)
against each other on a row that spans several thread blocks."""
import os

import numpy as np
import pytest

from tests.util import synth_counts

pytestmark = pytest.mark.gpu


def _handle(M, N, **kw):
    from bayesnmf_b200 import Handle
    return Handle(M, N, likelihood="poisson", prior="gamma", MH=False, seed=9, **kw)


def test_c3_full_size_latent_count_margins(built_lib):
    """C3 shape (96 x 100,000, N = 20): sum_n SE[n,g] = colSums(M), sum_n SP[k,n] = rowSums(M)
    wherever Mhat > 0 (R/sample_params.R:253-265), reproducible, and identical for a shard."""
    from oracle.gibbs import sample_Z_stats
    K, G, N = 96, 100_000, 20
    M, _, _ = synth_counts(K, G, N, 4000.0, seed=3)
    rng = np.random.default_rng(1)
    P = rng.gamma(1.0, 0.02, size=(K, N))
    E = rng.gamma(1.0, 4000.0 / N, size=(N, G))
    A = np.ones(N); A[3] = 0
    h = _handle(M, N)
    h.set_state("P", P); h.set_state("E", E); h.set_state("A", A)
    h.sample_z(4)
    SP, SE = h.get_state("SP"), h.get_state("SE")
    assert np.array_equal(SE.sum(axis=0), M.sum(axis=0))
    assert np.array_equal(SP.sum(axis=1), M.sum(axis=1))
    assert SP[:, 3].sum() == 0 and SE[3].sum() == 0          # an excluded signature gets no counts
    h.sample_z(4)
    assert np.array_equal(SP, h.get_state("SP")) and np.array_equal(SE, h.get_state("SE"))
    h.close()
    # a shard [g0, g1) of the same problem draws the same counts (Philox is keyed by the global cell)
    g0, g1 = 41_237, 47_301
    hs = _handle(M[:, g0:g1], N, g0=g0, G_total=G)
    hs.set_state("P", P); hs.set_state("E", E[:, g0:g1]); hs.set_state("A", A)
    hs.sample_z(4)
    assert np.array_equal(hs.get_state("SE"), SE[:, g0:g1])
    hs.close()
    # and the oracle agrees bit for bit on a slice of it
    s0, s1 = 41_300, 41_364
    oSP, oSE = sample_Z_stats(M[:, s0:s1], P, A, E[:, s0:s1], seed=9, it=4, g0=s0)
    assert np.array_equal(oSE, SE[:, s0:s1])


def test_c3_full_size_iterations_conserve_counts(built_lib):
    """Three full Gibbs iterations at the C3 shape: every iteration's margins add up to the data,
    metrics are finite and the log-likelihood improves on the prior draw."""
    from bayesnmf_b200.hyperpriors import fill_hyperprior_params
    K, G, N = 96, 100_000, 20
    M, _, _ = synth_counts(K, G, N, 4000.0, seed=4)
    h = _handle(M, N)
    for k, v in fill_hyperprior_params(None, "gamma", float(M.mean()), N).items():
        h.set_hyper(k, v)
    row1 = h.init_from_prior()
    out = h.step(3)
    assert np.isfinite(out["metrics"]).all()
    assert out["metrics"][-1][3] > row1["loglikelihood"]
    SP, SE = h.get_state("SP"), h.get_state("SE")
    assert SP.sum() == SE.sum() == M.sum()
    assert np.array_equal(SE.sum(axis=0), M.sum(axis=0))
    h.close()


@pytest.mark.parametrize("lik,prior,MH", [("poisson", "exponential", True), ("normal", "truncnormal", False)])
def test_p_sweep_implementations_agree(built_lib, lik, prior, MH):
    """The P sweep exists three times: k_p_rows (a cluster of blocks per mutation type, rows resident
    in shared memory), the pass-per-signature kernels, and -- Normal likelihood -- the Gram-matrix
    form.  Same conditionals, different summation orders (Gram: a different but algebraically equal
    expression): states agree to 1e-9 (1e-7 for Gram) relative over several iterations, before and
    after `converged`.  G is large enough that a row spans several blocks of a cluster."""
    from bayesnmf_b200 import Handle
    K, G, N = 6, 40_000, 4
    M, _, _ = synth_counts(K, G, N, 600.0, seed=2)
    if lik == "normal":
        M = M + np.random.default_rng(5).normal(0.0, 1.0, M.shape)
    variants = [("rows", {"BNMF_GRAM": "0"}), ("passes", {"BNMF_GRAM": "0", "BNMF_P_ROWS": "0"})]
    if lik == "normal":
        variants.append(("gram", {}))
    res = {}
    for name, env in variants:
        os.environ.update(env)
        try:
            h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, seed=4)
        finally:
            for k in env:
                os.environ.pop(k, None)
        h.init_from_prior()
        h.step(3)
        out = h.step(2, converged=True) if MH else h.step(2)
        res[name] = (h.get_state("P"), h.get_state("E"), h.get_state("Mhat"), out["metrics"][-1], h.timing()["launches"])
        h.close()
    assert res["rows"][4] < res["passes"][4]                   # the cluster kernel replaced 2N (4N) launches
    # Normal likelihood: Mhat is rebuilt on the tensor cores from inputs rounded to 32-bit fixed point
    # (csrc/bnmf_tc.cuh): a last-bit difference between two variants can move a rounding, i.e. Mhat by
    # N 2^-32 x (row scale of P) x (column scale of E) -- the variants agree to the north star's 1e-6, not to 1e-9
    tc = lik == "normal"
    for name, rtol in (("rows", 1e-6 if tc else 1e-9), ("gram", 1e-6 if tc else 1e-7)):
        if name not in res:
            continue
        P1, E1, H1, m1, _ = res[name]
        P0, E0, H0, m0, _ = res["passes"]
        np.testing.assert_allclose(P1, P0, rtol=rtol, atol=1e-300, err_msg=name)
        np.testing.assert_allclose(E1, E0, rtol=rtol, atol=1e-300, err_msg=name)
        np.testing.assert_allclose(H1, H0, rtol=rtol, atol=1e-5 if tc else 1e-9, err_msg=name)
        np.testing.assert_allclose(m1, m0, rtol=rtol, atol=1e-7, equal_nan=True, err_msg=name)


def test_overlapped_iterations_equal_sequential_ones(built_lib):
    """Count matrices above 2 M cells run the hyper-draws of iteration t+1 on a side stream under the
    latent-count kernel of iteration t (k_eside_hyper: attempt 0 of every Alpha draw in place, rejected
    cells parked and finished by k_alpha_retry).  One call of 5 iterations (overlapped) equals 5 calls
    of one iteration (a call's last iteration never speculates: the fused sequential kernels) bit for
    bit; so does a run whose parking list is too short (BNMF_ALPHA_CAP: the overflow path), and one whose
    iterations start with k_begin_iter instead of having k_pside's last block do its work (BNMF_FOLD_BEGIN=0), and
    one whose steady-state iterations launch k_pside and k_eside instead of the merged k_sides (BNMF_SIDES=0)."""
    from bayesnmf_b200 import Handle
    K, G, N = 96, 24_000, 8
    M, _, _ = synth_counts(K, G, N, 2000.0, seed=6)
    names = ("P", "E", "Alpha_e", "Beta_e", "Alpha_p", "Beta_p", "SP", "SE")

    def chain(calls, env=None):
        os.environ.update(env or {})
        try:
            h = Handle(M.astype(np.float64), N, likelihood="poisson", prior="gamma", MH=False, seed=13)
            h.init_from_prior()
            rows = np.concatenate([h.step(n)["metrics"] for n in calls])
        finally:
            for k in (env or {}):
                os.environ.pop(k, None)
        st = {n: h.get_state(n) for n in names}
        launches = h.timing()["launches"]
        h.close()
        return rows, st, launches

    rows_a, st_a, l_a = chain([5])
    rows_b, st_b, l_b = chain([1] * 5)
    rows_c, st_c, _ = chain([5], {"BNMF_ALPHA_CAP": "16"})
    rows_d, st_d, l_d = chain([5], {"BNMF_FOLD_BEGIN": "0"})     # k_begin_iter, k_pside, k_eside: three launches
    rows_e, st_e, l_e = chain([5], {"BNMF_SIDES": "0"})          # k_pside (+ k_begin_iter's work), k_eside: two launches
    assert l_a > l_b * 5            # the overlapped call launched the side-stream kernels, the single steps did not
    assert l_e == l_a + 4           # iterations 2..5 of the call ran k_sides: one launch for the two sides
    assert l_d == l_e + 5
    for rows, st in ((rows_b, st_b), (rows_c, st_c), (rows_d, st_d), (rows_e, st_e)):
        np.testing.assert_array_equal(rows_a, rows)
        for n in names:
            np.testing.assert_array_equal(st_a[n], st[n], err_msg=n)
