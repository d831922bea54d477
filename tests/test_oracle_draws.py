"""The oracle's random-variate generators against scipy CDFs (KS) -- the distributions
are the reference's (stats::rgamma / rexp / rnorm, truncnorm::rtruncnorm,
armspp::arms of R/sample_priors.R:356-397); the algorithms are the GPU build's."""
import numpy as np
import pytest
from scipy import integrate, stats

from oracle import draws as dr
from oracle import philox as px

N = 20000
CELLS = np.arange(N, dtype=np.uint64)
P_MIN = 1e-4


@pytest.mark.parametrize("shape,rate", [(0.05, 2.0), (0.7, 0.3), (1.0, 5.0), (3.3, 0.01), (250.0, 7.0)])
def test_gamma(shape, rate):
    x = dr.gamma_draw(11, 3, px.PUR_P, CELLS, shape, rate)
    assert stats.kstest(x, stats.gamma(shape, scale=1.0 / rate).cdf).pvalue > P_MIN


@pytest.mark.parametrize("mean,sd", [(2.0, 1.0), (0.0, 3.0), (-0.4, 1.0), (-3.0, 1.5), (-40.0, 2.0), (1e-3, 1e-2)])
def test_truncnorm(mean, sd):
    x = dr.truncnorm0_draw(5, 2, px.PUR_E, CELLS, mean, sd)
    assert (x >= 0).all()
    a = -mean / sd
    if a < 30:
        ref = stats.truncnorm(a, np.inf, loc=mean, scale=sd)
        assert stats.kstest(x, ref.cdf).pvalue > P_MIN
    else:   # far tail: z - alpha is ~ Exp(alpha) to first order; check the mean
        assert abs(x.mean() / (sd / a) - 1.0) < 0.05


def test_exponential_and_normal():
    x = dr.exponential_draw(1, 1, px.PUR_P, CELLS, 2.5)
    assert stats.kstest(x, stats.expon(scale=0.4).cdf).pvalue > P_MIN
    z = dr.normal_draw(1, 1, px.PUR_HYP_P1, CELLS, -1.0, 3.0)
    assert stats.kstest(z, stats.norm(-1.0, 3.0).cdf).pvalue > P_MIN


@pytest.mark.parametrize("C,D,beta,X", [(64.5, 10.0, 3.0, 0.02), (22.0, 10.0, 9.0, 45.0), (1.5, 0.5, 1.0, 1.0),
                                        (0.5, 2.0, 0.2, 1e-4), (10.0, 10.0, 1e3, 1e3)])
@pytest.mark.parametrize("start", ["default", "random"])
def test_alpha_conditional(C, D, beta, X, start):
    """Exact draw from log f(x) = (C-1) log x - D x + x log(beta) + (x-1) log X - lgamma(x)
    on [1e-3, 1e4] (R/sample_priors.R:357-365), wherever the mode search starts (the tangent
    points of the envelope only change the acceptance rate)."""
    x0 = None if start == "default" else np.exp(np.random.default_rng(5).uniform(np.log(2e-3), np.log(5e3), CELLS.shape))
    x = dr.alpha_draw(3, 9, px.PUR_HYP_P2, CELLS, C, D, beta, X, x0=x0)
    assert (x >= 1e-3).all() and (x <= 1e4).all()
    lo, hi = max(1e-3, x.min() * 0.2), min(1e4, x.max() * 3.0)
    grid = np.unique(np.concatenate([np.geomspace(1e-3, 1e4, 20001), np.linspace(lo, hi, 20001)]))
    lf = dr.alpha_logpdf(grid, C, D, beta, X)
    f = np.exp(lf - lf.max())
    cdf = integrate.cumulative_trapezoid(f, grid, initial=0.0)
    cdf /= cdf[-1]
    assert stats.kstest(x, lambda q: np.interp(q, grid, cdf)).pvalue > P_MIN


def test_digamma_trigamma():
    from scipy.special import digamma, polygamma
    x = np.concatenate([np.geomspace(1e-3, 1e4, 500), [1.0, 2.0, 6.0]])
    np.testing.assert_allclose(dr.digamma(x), digamma(x), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(dr.trigamma(x), polygamma(1, x), rtol=1e-8)


def test_draws_are_pure_functions_of_their_address():
    a = dr.gamma_draw(7, 4, px.PUR_E, CELLS[:100], 2.0, 1.0)
    b = dr.gamma_draw(7, 4, px.PUR_E, CELLS[:100][::-1].copy(), 2.0, 1.0)[::-1]
    assert np.array_equal(a, b)
    assert not np.array_equal(a, dr.gamma_draw(7, 5, px.PUR_E, CELLS[:100], 2.0, 1.0))
    assert not np.array_equal(a, dr.gamma_draw(8, 4, px.PUR_E, CELLS[:100], 2.0, 1.0))
