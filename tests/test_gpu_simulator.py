"""Device forward simulator (R/simulators.R:114-263): distributional checks - the reference's stream is
R's Mersenne-Twister, so parity is in distribution, not draw by draw."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_poisson_sampler_matches_the_pmf(built_lib):
    from scipy import stats
    from covidcurve_b200 import _lib
    n = 200_000
    for lam in (0.3, 4.0, 9.99, 10.0, 37.5, 1.0e4):
        x = _lib.nmath("rpois", np.full(n, lam), np.arange(n) // 1000, 12345.0)
        assert np.all(x == np.floor(x)) and x.min() >= 0
        assert abs(x.mean() - lam) < 5 * np.sqrt(lam / n) and abs(x.var() / lam - 1) < 0.02
        # chi-square against the pmf on bins with expected count >= 20
        lo, hi = int(stats.poisson.ppf(1e-4, lam)), int(stats.poisson.ppf(1 - 1e-4, lam))
        ks = np.arange(lo, hi + 1)
        exp = n * stats.poisson.pmf(ks, lam)
        obs = np.array([(x == k).sum() for k in ks]) if len(ks) < 400 else np.histogram(x, bins=np.arange(lo, hi + 2) - 0.5)[0]
        if len(ks) >= 400:                                   # pool to ~60 bins
            g = np.array_split(np.arange(len(ks)), 60)
            exp, obs = np.array([exp[i].sum() for i in g]), np.array([obs[i].sum() for i in g])
        keep = exp >= 20
        chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-6, keep.sum() - 1), (lam, chi2, keep.sum())
    # different replicate ids / seeds give different streams, equal ones the same draws
    a = _lib.nmath("rpois", np.full(1000, 50.0), 1.0, 7.0)
    assert np.array_equal(a, _lib.nmath("rpois", np.full(1000, 50.0), 1.0, 7.0))
    assert not np.array_equal(a, _lib.nmath("rpois", np.full(1000, 50.0), 2.0, 7.0))
    assert not np.array_equal(a, _lib.nmath("rpois", np.full(1000, 50.0), 1.0, 8.0))
    assert np.isnan(_lib.nmath("rpois", np.array([-1.0]), 0.0, 1.0)[0]) and _lib.nmath("rpois", np.array([0.0]), 0.0, 1.0)[0] == 0


def test_simulated_tables_have_the_line_list_models_moments(built_lib):
    from scipy import stats
    from covidcurve_b200 import _lib
    T, A = 150, 4
    t = np.arange(1, T + 1)
    infections = 4000.0 / (1 + np.exp(-(t - 60) / 9.0)) * np.exp(-np.maximum(t - 100, 0) / 25.0)
    ifr, rho, pop = np.array([0.001, 0.01, 0.05, 0.2]), np.array([1.0, 1.2, 0.9, 0.7]), np.array([4e5, 3e5, 2e5, 1e5])
    m_od, s_od, rate, frac = 14.26, 0.79, 18.3, 0.25
    R = 400
    sim = _lib.simulate_infxn_2_death(infections, ifr, rho, pop, m_od, s_od, rate, frac, seed=11, n_rep=R)
    # expected cell means from an independent host restatement of the line-list model
    G = stats.gamma.cdf(np.arange(T + 1), a=1 / s_od ** 2, scale=m_od * s_od ** 2)
    H = 1 - np.exp(-np.arange(T + 1) / rate)
    P = pop * rho / (pop * rho).sum()
    mu_d = np.convolve(infections, np.diff(G))[:T][:, None] * (P * ifr)[None, :]
    mu_s = np.convolve(infections, np.diff(H))[:T][:, None] * (P * frac)[None, :]
    np.testing.assert_allclose(sim["mean_deaths"], mu_d, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(sim["mean_serocon"], mu_s, rtol=1e-11)
    for got, mu in ((sim["deaths"], mu_d), (sim["serocon"], mu_s)):
        assert got.shape == (R, T, A) and np.all(got == np.floor(got)) and got.min() >= 0
        z = (got.mean(axis=0) - mu) / np.sqrt(np.maximum(mu, 1e-12) / R)
        big = mu > 0.05
        assert np.abs(z[big]).max() < 5.5 and abs(z[big].mean()) < 0.2          # cell means
        ratio = got.var(axis=0, ddof=1)[mu > 5] / mu[mu > 5]
        assert abs(ratio.mean() - 1) < 0.02                                       # Poisson dispersion
        tot = got.sum(axis=(1, 2))
        assert abs(tot.mean() / mu.sum() - 1) < 4 / np.sqrt(mu.sum() * R) + 1e-4
    # cells are independent: neighbouring days of one stratum are uncorrelated
    d = sim["deaths"][:, 90:130, 3]
    resid = d - mu_d[None, 90:130, 3]
    assert abs(np.mean(resid[:, :-1] * resid[:, 1:]) / np.mean(resid ** 2)) < 0.03
    # same seed -> same tables; another seed -> different tables
    again = _lib.simulate_infxn_2_death(infections, ifr, rho, pop, m_od, s_od, rate, frac, seed=11, n_rep=3)
    assert np.array_equal(again["deaths"], sim["deaths"][:3])
    other = _lib.simulate_infxn_2_death(infections, ifr, rho, pop, m_od, s_od, rate, frac, seed=12, n_rep=3)
    assert not np.array_equal(other["deaths"], sim["deaths"][:3])


def test_Agesim_infxn_2_death_mirror(built_lib):
    import pandas as pd
    from covidcurve_b200 import simulators
    T = 200
    infxns = np.floor(5e3 / (1 + np.exp(-np.linspace(-5, 7, T))))          # the curve of tests/testthat/test-check_natcub_cpp.R:6-11
    fat = pd.DataFrame({"Strata": ["ma1", "ma2", "ma3"], "IFR": [1e-3, 0.05, 0.1], "Rho": [1.0, 1.0, 1.0]})
    demog = pd.DataFrame({"Strata": ["ma1", "ma2", "ma3"], "popN": [1.5e6, 2.25e6, 1.25e6]})
    out = simulators.Agesim_infxn_2_death(fat, infxns, m_od=19.8, s_od=0.85, curr_day=T, spec=0.95, sens=0.85, demog=demog,
                                          sero_delay_rate=18.3, simulate_seroreversion=False, smplfrac=1e-3, seed=5)
    d, agg, sero = out["StrataAgg_TimeSeries_Death"], out["Agg_TimeSeries_Death"], out["StrataAgg_Seroprev"]
    assert list(d.columns) == ["ObsDay", "Strata", "Deaths"] and len(d) == 3 * T and len(agg) == T
    assert list(sero.columns) == ["Strata", "ObsDay", "testedN", "TrueSeroCount", "TruePrev", "ObsPrev"] and len(sero) == 3 * T
    assert np.array_equal(d.groupby("ObsDay")["Deaths"].sum().to_numpy(), agg["Deaths"].to_numpy())
    s1 = sero[sero["Strata"] == "ma1"]
    assert s1["TrueSeroCount"].is_monotonic_increasing and np.allclose(s1["testedN"], 1500.0)
    np.testing.assert_allclose(sero["ObsPrev"], 0.85 * sero["TruePrev"] + 0.05 * (1 - sero["TruePrev"]))
    # total deaths against the expectation of the line-list model
    from scipy import stats
    G = stats.gamma.cdf(np.arange(T + 1), a=1 / 0.85 ** 2, scale=19.8 * 0.85 ** 2)
    P = demog["popN"].to_numpy() / demog["popN"].sum()
    want = np.convolve(infxns, np.diff(G))[:T].sum() * (P * fat["IFR"].to_numpy()).sum()
    assert abs(agg["Deaths"].sum() - want) < 5 * np.sqrt(want)
    rev = simulators.Agesim_infxn_2_death(fat, infxns, m_od=19.8, s_od=0.85, curr_day=T, spec=0.95, sens=0.85, demog=demog,
                                          sero_delay_rate=18.3, simulate_seroreversion=True, sero_rev_shape=3, sero_rev_scale=140,
                                          smplfrac=1e-3, seed=5)["StrataAgg_Seroprev"]
    assert np.all(rev["ObsPrev"] <= 0.85 * rev["TruePrev"] + 0.05 * (1 - rev["TruePrev"]) + 1e-15)   # reversions only lower it
    assert np.any(rev["ObsPrev"] < 0.85 * rev["TruePrev"] + 0.05 * (1 - rev["TruePrev"]) - 1e-12)    # and somebody did revert
    with pytest.raises(ValueError):
        simulators.Agesim_infxn_2_death(fat, infxns[:-1], curr_day=T, spec=0.95, sens=0.85, demog=demog, sero_delay_rate=18.3)


def test_seroreversion_cells(built_lib):
    """Joint (conversion day, reversion day) cells: nobody reverts before converting, and the expected cumulative
    reversions equal the line-list model's  sum_t inf[t] w_a P(Exp + Weibull <= j - t + 1)."""
    from scipy import integrate, stats
    from covidcurve_b200 import _lib
    T, A = 120, 2
    t = np.arange(1, T + 1)
    infections = 3000.0 / (1 + np.exp(-(t - 40) / 6.0))
    ifr, rho, pop = np.array([0.01, 0.1]), np.array([1.0, 1.0]), np.array([3e5, 1e5])
    rate, shape, scale, frac = 12.0, 2.5, 60.0, 0.5
    R = 300
    sim = _lib.simulate_infxn_2_death(infections, ifr, rho, pop, 14.26, 0.79, rate, frac, seed=3, n_rep=R, sero_rev_shape=shape,
                                      sero_rev_scale=scale)
    conv, rev = np.cumsum(sim["serocon"], axis=1), np.cumsum(sim["serorev"], axis=1)
    assert np.all(rev <= conv) and np.all(sim["serorev"] >= 0) and rev[:, -1].sum() > 0
    # conversions keep their marginal law
    mu_c = sim["mean_serocon"]
    z = (sim["serocon"].mean(axis=0) - mu_c) / np.sqrt(np.maximum(mu_c, 1e-12) / R)
    assert np.abs(z[mu_c > 0.05]).max() < 5.5
    ratio = sim["serocon"].var(axis=0, ddof=1)[mu_c > 5] / mu_c[mu_c > 5]
    assert abs(ratio.mean() - 1) < 0.03
    # expected cumulative reversions by day j
    cdf = lambda x: integrate.quad(lambda s: np.exp(-s / rate) / rate * stats.weibull_min.cdf(x - s, shape, scale=scale), 0, x)[0]
    F = np.array([cdf(float(k)) for k in range(T + 1)])
    w = pop * rho / (pop * rho).sum() * frac
    for j in (60, 90, 120):
        want = sum(infections[tt - 1] * F[j - tt + 1] for tt in range(1, j + 1)) * w
        got = rev[:, j - 1, :].mean(axis=0)
        se = np.sqrt(want / R)
        assert np.all(np.abs(got - want) < 5 * se + 1e-9), (j, got, want)
    # reproducible
    again = _lib.simulate_infxn_2_death(infections, ifr, rho, pop, 14.26, 0.79, rate, frac, seed=3, n_rep=2, sero_rev_shape=shape,
                                        sero_rev_scale=scale)
    assert np.array_equal(again["serorev"], sim["serorev"][:2]) and np.array_equal(again["serocon"], sim["serocon"][:2])
