"""The draw code the kernels run (bayesnmf_b200/csrc/bnmf_rng.cuh, compiled for the host
by g++) against the numpy oracle, value for value.  CPU only."""
import ctypes

import numpy as np
import pytest

from oracle import draws as dr
from oracle import philox as px


@pytest.fixture(scope="module")
def hc():
    from tests.hostcheck.build_hostcheck import build
    return ctypes.CDLL(build())


def _p(a, ty=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(ty))


N = 4000
CELLS = (np.arange(N, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(2 ** 40)


def test_philox(hc):
    out = (ctypes.c_uint32 * 4)()
    for ctr, key in [((0, 0, 0, 0), (0, 0)), ((1, 2, 3, 4), (5, 6)), ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2)]:
        hc.hc_philox(*[ctypes.c_uint32(c) for c in ctr], *[ctypes.c_uint32(k) for k in key], out)
        assert list(out) == [int(x) for x in px.philox4x32_10(*ctr, *key)]


def _run(fn, seed, it, pur, *cols):
    out = np.empty(N)
    fn(ctypes.c_uint64(seed), ctypes.c_uint32(it), ctypes.c_uint32(pur), _p(CELLS, ctypes.c_uint64),
       *[_p(c) for c in cols], _p(out), ctypes.c_long(N))
    return out


def test_gamma(hc):
    rng = np.random.default_rng(0)
    shape = np.exp(rng.uniform(np.log(0.02), np.log(5000.0), N))
    rate = np.exp(rng.uniform(-5, 5, N))
    got = _run(hc.hc_gamma, 42, 3, px.PUR_P, shape, rate)
    np.testing.assert_allclose(got, dr.gamma_draw(42, 3, px.PUR_P, CELLS, shape, rate), rtol=1e-12)


def test_truncnorm(hc):
    rng = np.random.default_rng(1)
    mean = rng.normal(0, 5, N)
    sd = np.exp(rng.uniform(-3, 2, N))
    got = _run(hc.hc_truncnorm, 9, 5, px.PUR_E, mean, sd)
    np.testing.assert_allclose(got, dr.truncnorm0_draw(9, 5, px.PUR_E, CELLS, mean, sd), rtol=1e-11, atol=1e-300)


def test_normal_exponential(hc):
    rng = np.random.default_rng(2)
    mean, sd = rng.normal(0, 5, N), np.exp(rng.uniform(-3, 2, N))
    np.testing.assert_allclose(_run(hc.hc_normal, 1, 2, px.PUR_HYP_P1, mean, sd),
                               dr.normal_draw(1, 2, px.PUR_HYP_P1, CELLS, mean, sd), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(_run(hc.hc_exp, 1, 2, px.PUR_P, sd), dr.exponential_draw(1, 2, px.PUR_P, CELLS, sd), rtol=1e-13)


def test_alpha(hc):
    rng = np.random.default_rng(3)
    C = np.exp(rng.uniform(np.log(0.3), np.log(300.0), N))
    D = np.exp(rng.uniform(-2, 3, N))
    beta = np.exp(rng.uniform(-4, 6, N))
    X = np.exp(rng.uniform(-12, 8, N))
    x0 = np.exp(rng.uniform(np.log(1e-4), np.log(1e5), N))     # incl. starts outside [1e-3, 1e4]
    got = _run(hc.hc_alpha, 77, 4, px.PUR_HYP_E2, C, D, beta, X, x0)
    want = dr.alpha_draw(77, 4, px.PUR_HYP_E2, CELLS, C, D, beta, X, x0=x0)
    # lgamma/digamma come from libm vs scipy: equal to rounding, and the accept test is discrete
    close = np.isclose(got, want, rtol=1e-8)
    assert close.mean() > 0.999, (~close).sum()


def test_digamma(hc):
    x = np.exp(np.random.default_rng(4).uniform(np.log(1e-3), np.log(1e4), N))
    d, t = np.empty(N), np.empty(N)
    hc.hc_digamma(_p(x), _p(d), _p(t), ctypes.c_long(N))
    np.testing.assert_allclose(d, dr.digamma(x), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(t, dr.trigamma(x), rtol=1e-12)
