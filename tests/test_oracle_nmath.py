"""The oracle's restatement of R's nmath against independent evaluations (scipy / mpmath)."""
import numpy as np
import pytest

from oracle.oracle import lib


def _v(f, *cols):
    return np.array([f(*[float(c) for c in row]) for row in zip(*cols)])


def test_pgamma_against_scipy_and_mpmath():
    from scipy.special import gammainc
    import mpmath as mp
    L = lib()
    rng = np.random.default_rng(0)
    sod = np.concatenate([rng.uniform(0.02, 1.3, 300), [0.85, 0.1, 0.999]])
    mod = rng.uniform(5, 40, sod.size)
    for s, m in zip(sod, mod):
        a, sc = 1 / s ** 2, m * s ** 2
        x = np.arange(0, 301, 7, dtype=float)
        got = _v(lambda xx: L.cco_pgamma(xx, a, sc, 1, 0), x)
        want = gammainc(a, x / sc)
        np.testing.assert_allclose(got, want, rtol=2e-12, atol=1e-300)
    mp.mp.dps = 40
    for a, x in ((1.384, 0.3), (1.384, 20.0), (100.0, 80.0), (100.0, 130.0), (2500.0, 2400.0), (2500.0, 2600.0), (0.35, 2.0)):
        want = float(mp.gammainc(a, 0, x, regularized=True))
        assert L.cco_pgamma(x, a, 1.0, 1, 0) == pytest.approx(want, rel=5e-14)


def test_pnorm_cody_tables_and_log_tail():
    """R's pnorm_both (Cody 1969) as restated in the oracle: the coefficient tables are pinned against erfc over the central
    range and against mpmath in both tails, on the plain and on the log scale (the log tail is what pgamma_raw's log-space
    redo of the asymptotic branch uses, nmath_port.cpp)."""
    import math
    import mpmath as mp
    L = lib()
    xs = np.linspace(-8, 8, 4001)
    ref = np.array([0.5 * math.erfc(-x / math.sqrt(2)) for x in xs])
    got = _v(lambda x: L.cco_pnorm(x, 1, 0), xs)
    np.testing.assert_allclose(got, ref, rtol=2e-14)
    mp.mp.dps = 80
    for x in (-0.3, 0.5, -0.6744, -0.675, -3.0, 2.0, -5.65, -5.66, -9.0, -20.0, -36.0, -37.4, -37.6, -38.5, -60.0, -500.0, 12.0):
        for lower in (1, 0):
            t = x if lower else -x                                     # the requested probability is Phi(t)
            want = mp.log(mp.ncdf(t)) if t < 0 else mp.log1p(-mp.ncdf(-t))
            got = L.cco_pnorm(x, lower, 1)
            assert abs(got - want) <= 4e-15 * max(abs(want), mp.mpf("1e-300")) + 1e-300, (x, lower, got, float(want))
    # plain scale: R returns exact 0 / 1 beyond its cut-offs (-37.5193, 8.2924)
    assert L.cco_pnorm(-37.6, 1, 0) == 0.0 and L.cco_pnorm(8.3, 1, 0) == 1.0 and L.cco_pnorm(-37.0, 1, 0) > 0
    # pgamma far in the lower tail of a very peaked shape: the log-space redo keeps full relative accuracy below 1e-292
    mp.mp.dps = 60
    for a, target in ((60000.0, -295), (60000.0, -306), (120000.0, -315), (250000.0, -321), (1e6, -300)):
        lo_, hi_ = a - 40 * math.sqrt(a), a - 30 * math.sqrt(a)
        for _ in range(60):                                             # x with P(a, x) = 10^target
            mid = 0.5 * (lo_ + hi_)
            if mp.log10(mp.gammainc(a, 0, mid, regularized=True)) < target:
                lo_ = mid
            else:
                hi_ = mid
        x = float(hi_)
        want = mp.gammainc(a, 0, x, regularized=True)
        got = L.cco_pgamma(x, a, 1.0, 1, 0)
        assert 0 < float(want) < 1.1e-292
        assert abs(got - want) <= 2e-7 * want + 5e-324, (a, x, got, float(want))   # R's Temme expansion itself is ~1e-8 here


def test_dpois_dbinom_against_scipy():
    from scipy import stats
    L = lib()
    rng = np.random.default_rng(1)
    x = np.floor(rng.uniform(0, 400, 500))
    lam = rng.uniform(1e-3, 500, 500)
    np.testing.assert_allclose(_v(lambda a, b: L.cco_dpois(a, b, 1), x, lam), stats.poisson.logpmf(x, lam), rtol=1e-11, atol=1e-11)
    n = np.floor(rng.uniform(1, 3e6, 500))
    p = rng.uniform(1e-4, 0.9, 500)
    k = np.floor(n * np.clip(p * rng.uniform(0.5, 1.5, 500), 0, 1))
    np.testing.assert_allclose(_v(lambda a, b, c: L.cco_dbinom(a, b, c, 1), k, n, p), stats.binom.logpmf(k, n, p), rtol=1e-9, atol=1e-9)
    # edge semantics of R
    assert L.cco_dbinom(2.5, 10, 0.5, 1) == -np.inf
    assert np.isnan(L.cco_dbinom(2, 10, 1.5, 1))
    assert np.isnan(L.cco_dpois(3, -1.0, 1))
    assert L.cco_dpois(3, 0.0, 1) == -np.inf and L.cco_dpois(0, 0.0, 1) == 0.0


def test_priors_against_scipy():
    from scipy import stats
    L = lib()
    x = np.linspace(0.7, 0.99, 50)
    for a, b in ((850.5, 150.5), (950.5, 50.5), (2550.0, 450.0), (8500.0, 1500.0), (1.5, 3.0)):
        np.testing.assert_allclose(_v(lambda v: L.cco_dbeta(v, a, b, 1), x), stats.beta.logpdf(x, a, b), rtol=1e-10, atol=1e-9)
    assert L.cco_dnorm(18.0, 18.3, 0.1, 1) == pytest.approx(stats.norm.logpdf(18.0, 18.3, 0.1), rel=1e-14)
    assert L.cco_dunif(0.2, 0.0, 0.4, 1) == pytest.approx(-np.log(0.4)) and L.cco_dunif(0.5, 0.0, 0.4, 1) == -np.inf
    assert L.cco_pweibull(50.0, 3.0, 140.0, 1, 0) == pytest.approx(stats.weibull_min.cdf(50.0, 3.0, scale=140.0), rel=1e-14)
    assert L.cco_pweibull(0.0, 3.0, 140.0, 1, 0) == 0.0


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    import ctypes as C
    L = lib()

    def kat(ctr, key):
        c, k, o = (C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), (C.c_uint * 4)()
        L.cco_philox_kat(c, k, o)
        return list(o)
    assert kat([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
