"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Vectorised fp64 restatements of the random-variate generators, addressed by the same
Philox streams as the CUDA kernels (bayesnmf_b200/csrc/bnmf_rng.cuh), so that every
draw can be compared value-for-value.

The reference takes its variates from third-party code that is NOT in
/root/reference (R `stats`/nmath: rgamma = Ahrens-Dieter GD/GS, rnorm = inversion,
rexp, runif, rbinom/rmultinom; `truncnorm::rtruncnorm`; `invgamma::rinvgamma` =
1/rgamma; `armspp::arms`; no versions pinned in DESCRIPTION:11-23).  What is restated
here are the *published algorithms the GPU build uses instead* (north star:
Marsaglia-Tsang gamma, truncated normal by rejection, exact adaptive-rejection draw of
the log-concave Alpha conditional); the distributions are the reference's
(call sites cited per function) and are checked against scipy CDFs in
tests/test_oracle_draws.py.  Parity with R itself is therefore distributional only:
"parity unpinned" at the draw level.
"""
import numpy as np
from scipy.special import gammaln

from . import philox as px

MAX_ATTEMPTS = 4096


def normal_from(w0, w1, bits=32):
    u1 = px.u01(w0, bits)
    u2 = px.u01(w1, bits)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586 * u2)


def exponential_draw(seed, it, purpose, cell, rate, bits=32):
    """stats::rexp(n, rate)  (R/sample_Pn.R:21, R/sample_En.R:21)."""
    w = px.words(seed, it, purpose, cell, 0)
    return -np.log(px.u01(w[0], bits)) / rate


def normal_draw(seed, it, purpose, cell, mean, sd, bits=32):
    """stats::rnorm(n, mean, sd)  (R/sample_priors.R:34-38, :219, :235)."""
    w = px.words(seed, it, purpose, cell, 0)
    return mean + sd * normal_from(w[0], w[1], bits)


def gamma_draw(seed, it, purpose, cell, shape, rate, bits=32, sub0=0):
    """stats::rgamma(n, shape, rate) (R/sample_Pn.R:23-27,116-118; R/sample_En.R:23-27,
    115-117; R/sample_priors.R:72-129,285-344), by Marsaglia & Tsang (2000) with the
    U^(1/shape) boost for shape < 1; attempt t uses Philox block sub0 + t."""
    cell = np.asarray(cell, dtype=np.uint64)
    shape = np.broadcast_to(np.asarray(shape, dtype=np.float64), cell.shape).copy()
    rate = np.broadcast_to(np.asarray(rate, dtype=np.float64), cell.shape)
    boost = shape < 1.0
    a = np.where(boost, shape + 1.0, shape)
    d = a - (1.0 / 3.0)
    c = 1.0 / np.sqrt(9.0 * d)
    out = d.copy()
    todo = np.ones(cell.shape, dtype=bool)
    flat = lambda x: x  # noqa: E731
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in range(MAX_ATTEMPTS):
            if not todo.any():
                break
            idx = np.nonzero(todo)
            w = px.words(seed, it, purpose, cell[idx], sub0 + t)
            x = normal_from(w[0], w[1], bits)
            v = 1.0 + c[idx] * x
            ok = v > 0.0
            v3 = v * v * v
            u = px.u01(w[2], bits)
            x2 = x * x
            acc = u < 1.0 - 0.0331 * (x2 * x2)
            acc2 = np.log(u) < 0.5 * x2 + d[idx] * (1.0 - v3 + np.log(np.where(ok, v3, 1.0)))
            acc = ok & (acc | acc2)
            g = d[idx] * v3
            bi = boost[idx]
            g = np.where(bi, g * np.power(px.u01(w[3], bits), 1.0 / shape[idx]), g)
            sel = tuple(i[acc] for i in idx)
            out[sel] = g[acc]
            todo[sel] = False
    out = out / rate
    return np.maximum(out, np.finfo(np.float64).tiny)


def truncnorm0_draw(seed, it, purpose, cell, mean, sd, bits=32):
    """truncnorm::rtruncnorm(n, a = 0, b = Inf, mean, sd) -- every call site in the
    reference has a = 0, b = Inf (R/sample_Pn.R:14-19,59-64,79-85; R/sample_En.R:14-19,
    59-64,78-84).  alpha = -mean/sd <= 0.45: normal rejection; else Robert (1995)."""
    cell = np.asarray(cell, dtype=np.uint64)
    mean = np.broadcast_to(np.asarray(mean, dtype=np.float64), cell.shape)
    sd = np.broadcast_to(np.asarray(sd, dtype=np.float64), cell.shape)
    alpha = -mean / sd
    plain = alpha <= 0.45
    lam = 0.5 * (alpha + np.sqrt(alpha * alpha + 4.0))
    z = alpha.copy()           # plain branch result (z), fallback alpha
    e = np.zeros(cell.shape)   # Robert branch result (z - alpha)
    todo = np.ones(cell.shape, dtype=bool)
    for t in range(MAX_ATTEMPTS):
        if not todo.any():
            break
        idx = np.nonzero(todo)
        w = px.words(seed, it, purpose, cell[idx], t)
        pl = plain[idx]
        zz = normal_from(w[0], w[1], bits)
        acc_plain = zz >= alpha[idx]
        ee = -np.log(px.u01(w[0], bits)) / lam[idx]
        dz = (alpha[idx] + ee) - lam[idx]
        acc_rob = np.log(px.u01(w[1], bits)) <= -0.5 * (dz * dz)
        acc = np.where(pl, acc_plain, acc_rob)
        sel = tuple(i[acc] for i in idx)
        z[sel] = zz[acc]
        e[sel] = ee[acc]
        todo[sel] = False
    x_plain = np.maximum(mean + sd * z, 0.0)
    x_rob = sd * e
    return np.where(plain, x_plain, x_rob)


def digamma_trigamma(x):
    """bnmf_rng.cuh::digamma_trigamma, operation for operation (shared reciprocals)."""
    x = np.asarray(x, dtype=np.float64).copy()
    r1 = np.zeros_like(x)
    r2 = np.zeros_like(x)
    for _ in range(8):
        m = x < 6.0
        if not m.any():
            break
        inv = 1.0 / x
        r1 = np.where(m, r1 - inv, r1)
        r2 = np.where(m, r2 + inv * inv, r2)
        x = np.where(m, x + 1.0, x)
    inv = 1.0 / x
    f = inv * inv
    t1 = f * ((-1.0 / 12.0) + f * ((1.0 / 120.0) + f * ((-1.0 / 252.0) + f * ((1.0 / 240.0) + f * (-1.0 / 132.0)))))
    psi = r1 + np.log(x) - 0.5 * inv + t1
    t2 = inv + 0.5 * f + (f * inv) * ((1.0 / 6.0) + f * ((-1.0 / 30.0) + f * ((1.0 / 42.0) + f * (-1.0 / 30.0))))
    return psi, r2 + t2


def digamma(x):
    return digamma_trigamma(x)[0]


def trigamma(x):
    return digamma_trigamma(x)[1]


def _seg_unit(s, w):
    """integral of exp(-|s| y) over [0, w]: the mass of a linear-exponent segment of width w
    relative to its higher end (never overflows)."""
    sw = np.abs(s) * w
    small = sw < 1e-8
    with np.errstate(divide="ignore", invalid="ignore"):
        big = -np.expm1(-sw) / np.where(small, 1.0, np.abs(s))
    return np.where(small, w * (1.0 - 0.5 * sw), big)


def _seg_inv(s, a, b, q):
    """x in [a, b] at partial mass fraction q of the segment above (same anchoring)."""
    w = b - a
    sw = s * w
    small = np.abs(sw) < 1e-8
    pos = sw > 0.0
    em = np.expm1(-np.abs(sw))
    with np.errstate(divide="ignore", invalid="ignore"):
        sd = np.where(small, 1.0, s)
        big = np.where(pos, b + np.log1p((1.0 - q) * em) / sd, a + np.log1p(q * em) / sd)
    return np.where(small, a + q * w, big)


def alpha_logpdf(x, C, D, beta, X):
    """log f(x) of R/sample_priors.R:357-363 / :383-389 up to the (dropped) -log X."""
    cm1 = C - 1.0
    b = D - np.log(beta) - np.log(X)
    return cm1 * np.log(x) - b * x - gammaln(x)


LAST_ALPHA_STATS = {}
PSI_LO = -1000.5755719318103     # digamma(1e-3), literal shared with bnmf_rng.cuh
PSI_HI = 9.21029037114285        # digamma(1e4)
ALPHA_NEWTON = 16         # at most; stops at ALPHA_TOL relative (about 2 steps from the previous Alpha)
ALPHA_TOL = 1e-2


def alpha_draw(seed, it, purpose, cell, C, D, beta, X, x0=None):
    """armspp::arms(n_samples = 1, log_pdf, lower = 1e-3, upper = 1e4) for the shape
    parameter of the Gamma prior (R/sample_priors.R:356-397).  The target is
    log-concave (f'' = -(C-1)/x^2 - trigamma(x) < -C/x^2), so ARMS is exact ARS; drawn
    here exactly with a fixed three-tangent envelope (same construction as
    bnmf_rng.cuh::alpha_draw).  x0 = start of the (approximate) mode search = the current
    Alpha; the tangent points only affect the acceptance rate, not the distribution."""
    LO, HI = 1e-3, 1e4
    cell = np.asarray(cell, dtype=np.uint64)
    shp = cell.shape
    C = np.broadcast_to(np.asarray(C, dtype=np.float64), shp)
    D = np.broadcast_to(np.asarray(D, dtype=np.float64), shp)
    beta = np.broadcast_to(np.asarray(beta, dtype=np.float64), shp)
    X = np.broadcast_to(np.asarray(X, dtype=np.float64), shp)
    cm1 = C - 1.0
    b = D - np.log(beta) - np.log(X)
    h = lambda x: cm1 * np.log(x) - b * x - gammaln(x)  # noqa: E731
    hp = lambda x: cm1 / x - b - digamma(x)  # noqa: E731
    hpp = lambda x: -cm1 / (x * x) - trigamma(x)  # noqa: E731

    at_lo = cm1 / LO - b - PSI_LO <= 0.0
    at_hi = (~at_lo) & (cm1 / HI - b - PSI_HI >= 0.0)
    a = np.full(shp, LO)
    bb = np.full(shp, HI)
    xdef = np.where(C > 1.0, np.where(C < HI, C, 0.5 * HI), 1.0)
    if x0 is None:
        x = xdef
    else:
        x0 = np.broadcast_to(np.asarray(x0, dtype=np.float64), shp)
        x = np.where((x0 > LO) & (x0 < HI), x0, xdef)
    done = np.zeros(shp, dtype=bool)
    steps = np.zeros(shp, dtype=np.int64)
    for _ in range(ALPHA_NEWTON):
        steps += ~done
        f = hp(x)
        a = np.where(f > 0.0, x, a)
        bb = np.where(f > 0.0, bb, x)
        with np.errstate(over="ignore", invalid="ignore"):
            xn = x * np.exp(-f / (x * hpp(x)))   # Newton step in log x
        bad = ~((xn > a) & (xn < bb))
        xn = np.where(bad, np.sqrt(a * bb), xn)
        conv = np.abs(xn - x) <= ALPHA_TOL * x
        x = np.where(done, x, xn)            # the converging step is still taken
        done = done | conv
    m = np.where(at_lo, LO, np.where(at_hi, HI, x))
    hp_m = hp(m)
    s = 1.0 / np.sqrt(-hpp(m))
    x0 = np.where(m - s > 0.5 * m, m - s, 0.5 * m)
    x1 = m.copy()
    x2 = m + s
    lo_case = m <= LO
    x0 = np.where(lo_case, LO, x0)
    x1 = np.where(lo_case, LO + s, x1)
    x2 = np.where(lo_case, LO + 2.0 * s, x2)
    hi_case = m >= HI
    hx0 = HI - 2.0 * s
    hx1 = HI - s
    shrink = hx0 < 0.5 * HI
    hx0 = np.where(shrink, 0.5 * HI, hx0)
    hx1 = np.where(shrink, 0.75 * HI, hx1)
    x0 = np.where(hi_case, hx0, x0)
    x1 = np.where(hi_case, hx1, x1)
    x2 = np.where(hi_case, HI, x2)
    xs = [x0, x1, x2]
    hv = [h(v) for v in xs]
    sl = [np.where(v == m, hp_m, hp(v)) for v in xs]
    z = [np.full(shp, LO), None, None, np.full(shp, HI)]
    for j in range(2):
        den = sl[j] - sl[j + 1]
        zz = (hv[j + 1] - hv[j] + sl[j] * xs[j] - sl[j + 1] * xs[j + 1]) / den
        zz = np.where(zz >= xs[j], zz, xs[j])
        zz = np.where(zz <= xs[j + 1], zz, xs[j + 1])
        zz = np.where(zz < LO, LO, zz)
        zz = np.where(zz > HI, HI, zz)
        z[j + 1] = zz
    # masses of the three envelope segments, each relative to the envelope's overall maximum
    # (a piecewise-linear exponent peaks at a segment end), so nothing can overflow even when
    # the tangent points do not bracket the mode
    top = [np.maximum(hv[j] + sl[j] * (z[j] - xs[j]), hv[j] + sl[j] * (z[j + 1] - xs[j])) for j in range(3)]
    hmax = np.maximum(np.maximum(top[0], top[1]), top[2])
    mass = []
    for j in range(3):
        mj = np.where(z[j + 1] > z[j], np.exp(top[j] - hmax) * _seg_unit(sl[j], z[j + 1] - z[j]), 0.0)
        mass.append(mj)
    tot = (mass[0] + mass[1]) + mass[2]
    out = m.copy()
    todo = np.ones(shp, dtype=bool)
    attempts = np.zeros(shp, dtype=np.int64)
    XS = np.stack(xs); HV = np.stack(hv); SL = np.stack(sl); Z = np.stack(z); MS = np.stack(mass)
    for t in range(MAX_ATTEMPTS):
        if not todo.any():
            break
        idx = np.nonzero(todo)
        attempts[idx] += 1
        w = px.words(seed, it, purpose, cell[idx], t)
        r = px.u01(w[0]) * tot[idx]
        m0 = MS[0][idx]; m1 = MS[1][idx]
        j = np.zeros(r.shape, dtype=np.int64)
        ge0 = r >= m0
        r = np.where(ge0, r - m0, r)
        j = np.where(ge0, 1, j)
        ge1 = ge0 & (r >= m1)
        r = np.where(ge1, r - m1, r)
        j = np.where(ge1, 2, j)
        pick = lambda A: np.choose(j, [A[0][idx], A[1][idx], A[2][idx]])  # noqa: E731
        mj = pick(MS)
        good = mj > 0.0
        with np.errstate(divide="ignore", invalid="ignore"):
            q = r / np.where(good, mj, 1.0)
        q = np.where(q >= 1.0, 1.0 - 1e-16, q)
        zl = np.choose(j, [Z[0][idx], Z[1][idx], Z[2][idx]])
        zr = np.choose(j, [Z[1][idx], Z[2][idx], Z[3][idx]])
        slj = pick(SL); hvj = pick(HV); xsj = pick(XS)
        xc = _seg_inv(slj, zl, zr, q)
        xc = np.minimum(np.maximum(xc, zl), zr)
        env = hvj + slj * (xc - xsj)
        hx = cm1[idx] * np.log(xc) - b[idx] * xc - gammaln(xc)
        acc = good & (np.log(px.u01(w[1])) <= hx - env)
        sel = tuple(i[acc] for i in idx)
        out[sel] = xc[acc]
        todo[sel] = False
    LAST_ALPHA_STATS.update(newton_steps=steps, attempts=attempts)   # diagnostics for tests / tuning
    return out
