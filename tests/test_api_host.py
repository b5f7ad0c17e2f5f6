"""Host-side mirror of the R interface (covidcurve_b200/api.py, summaries.py): no GPU needed."""
import numpy as np
import pandas as pd
import pytest

from covidcurve_b200 import api, summaries
from covidcurve_b200.spec import PRIOR_DBETA, PRIOR_DNORM, PRIOR_DUNIF, PRIOR_NONE


def _model_from(spec, with_rev=False):
    A, S = spec.A, spec.S
    m = api.make_IFRmodel_age.new()
    m.set_MeanTODparam("mod"); m.set_CoefVarOnsetTODparam("sod")
    m.set_IFRparams(spec.IFRparams); m.set_maxMa(spec.IFRparams[-1])
    m.set_Knotparams(spec.Knotparams); m.set_relKnot(spec.Knotparams[-1])
    m.set_Infxnparams(spec.Infxnparams); m.set_relInfxn(spec.Infxnparams[-1])
    m.set_Noiseparams(spec.Noiseparams)
    m.set_Serotestparams(["sens", "spec", "sero_con_rate"])
    sero = pd.DataFrame({"SeroStartSurvey": np.repeat(spec.sero_survey_start, A), "SeroEndSurvey": np.repeat(spec.sero_survey_end, A),
                         "Strata": spec.IFRparams * S, "SeroPos": spec.sero_a, "SeroN": spec.sero_b,
                         "SeroPrev": spec.sero_a / spec.sero_b, "SeroLCI": spec.sero_a / spec.sero_b * 0.9,
                         "SeroUCI": spec.sero_a / spec.sero_b * 1.1})
    m.set_data({"obs_deaths": pd.DataFrame({"ObsDay": np.arange(1, spec.T + 1), "Deaths": spec.obs_deaths}),
                "prop_deaths": pd.DataFrame({"Strata": spec.IFRparams, "PropDeaths": spec.prop_strata_obs_deaths}),
                "obs_serology": sero})
    m.set_demog(pd.DataFrame({"Strata": spec.IFRparams, "popN": spec.demog}))
    m.set_paramdf(pd.DataFrame({"name": spec.names, "min": spec.pmin, "init": spec.pinit, "max": spec.pmax,
                                "dsc1": spec.dsc1, "dsc2": spec.dsc2}))
    return m


def test_build_spec_reproduces_the_stored_fit_inputs(modfit):
    spec, g = modfit
    m = _model_from(spec)
    s2 = api.build_spec(m, binomial_likelihood=True, reparamIFR=False, reparamInfxn=False, reparamKnots=False)
    assert s2.names == spec.names and s2.T == 200 and s2.K == 5 and s2.M == 165 and s2.rcensor_day == 2147483647
    for k in ("obs_deaths", "prop_strata_obs_deaths", "sero_a", "sero_b", "demog", "pmin", "pmax", "dsc1", "dsc2"):
        assert np.array_equal(getattr(s2, k), getattr(spec, k)), k
    assert s2.role_map() == spec.role_map()
    fam, jr, jc = s2.prior_table()
    assert fam[s2.index("sod")] == PRIOR_NONE                       # FFF quirk: the sod Beta prior is dropped
    assert fam[s2.index("ma1")] == PRIOR_DUNIF and fam[s2.index("mod")] == PRIOR_DNORM and fam[s2.index("sens")] == PRIOR_DBETA
    assert list(jr) == [-1, -1, -1]


def test_reparameterised_role_map_follows_paramdf_order(modfit):
    spec, _ = modfit
    m = _model_from(spec)
    s3 = api.build_spec(m, reparamIFR=True, reparamInfxn=True, reparamKnots=True)
    rm = s3.role_map()
    assert rm["idx_max_ma"] == s3.index("ma3") and rm["idx_rel_knot"] == s3.index("x4") and rm["idx_rel_infxn"] == s3.index("y5")
    fam, jr, jc = s3.prior_table()
    assert fam[s3.index("sod")] == PRIOR_DBETA and list(jr) == [s3.index("ma3"), s3.index("x4"), s3.index("y5")]
    assert list(jc) == [2.0, 3.0, 4.0]                               # (A-1, K-2, K-1) scalars multiplied by each reference


def test_logit_data_list_matches_main_R(modfit):
    spec, _ = modfit
    m = _model_from(spec)
    s = api.build_spec(m, binomial_likelihood=False, reparamIFR=False, reparamInfxn=False, reparamKnots=False)
    p = spec.sero_a / spec.sero_b
    logit = lambda v: np.log(v / (1 - v))
    np.testing.assert_allclose(s.sero_a, logit(p), rtol=1e-14)                                  # main.R:146 SeroMu
    np.testing.assert_allclose(s.sero_b, (logit(p * 1.1) - logit(p * 0.9)) / (2 * 1.96), rtol=1e-12)   # main.R:145 SeroSE


def test_model_validation_messages(modfit):
    spec, _ = modfit
    m = api.make_IFRmodel_age.new()
    with pytest.raises(ValueError, match="Must specify IFR parameters before specifying data"):
        m.set_data({})
    m.set_IFRparams(["ma1", "ma2"])
    with pytest.raises(ValueError):
        m.set_maxMa("ma9")
    with pytest.raises(ValueError, match="Serology test parameters"):
        m.set_Serotestparams(["sens", "oops"])
    full = _model_from(spec)
    full.set_Serotestparams(["sens", "spec", "sero_con_rate", "sero_rev_shape"])
    with pytest.raises(ValueError, match="both the Weibull"):
        api.build_spec(full)
    one = _model_from(spec)
    one.IFRparams = ["ma1"]
    with pytest.raises(ValueError):
        api._checks(one, True, False, False, False)


def test_gelman_rubin_and_ess_on_known_series():
    rng = np.random.default_rng(0)
    iid = rng.normal(size=(4, 4000, 2))
    r = summaries.rhat_per_parameter(iid)
    assert np.all(np.abs(r - 1) < 0.01)
    est, upper = summaries.gelman_diag(iid)
    assert np.all(np.abs(est - 1) < 0.01) and np.all(upper >= est)
    shifted = iid.copy(); shifted[0, :, 0] += 3.0
    assert summaries.rhat_per_parameter(shifted)[0] > 1.5 and summaries.gelman_diag(shifted)[0][0] > 1.5
    assert 3000 < summaries.effective_size(iid[0, :, 0]) < 5000
    ar = np.zeros(20000); e = rng.normal(size=20000)
    for t in range(1, 20000):
        ar[t] = 0.9 * ar[t - 1] + e[t]
    ess = summaries.effective_size(ar)                                  # AR(1), rho = 0.9: n (1 - rho) / (1 + rho) ~ 1050
    assert 700 < ess < 1500
    assert abs(summaries.geweke_z(iid[1, :, 1])) < 4


def test_geweke_windows_are_codas(monkeypatch):
    """coda::geweke.diag: first window 1..ceiling(1 + 0.1 (n-1)), second floor(n - 0.5 (n-1))..n - 1 001 and 5 001 points for
    n = 10 000 (what get_cred_intervals' GewekeZ is computed on, R/summarize_plot.R:150)."""
    seen = []
    monkeypatch.setattr(summaries, "_spectrum0_ar", lambda w: (seen.append((w[0], w[-1], w.size)), 1.0)[1])
    summaries.geweke_z(np.arange(1.0, 10001.0))
    assert seen == [(1.0, 1001.0, 1001), (5000.0, 10000.0, 5001)]
    seen.clear()
    summaries.geweke_z(np.arange(1.0, 41.0))                     # n = 40: 1..ceiling(4.9) = 5 and floor(20.5) = 20..40
    assert seen == [(1.0, 5.0, 5), (20.0, 40.0, 21)]


def test_spectrum0_ar_against_a_direct_yule_walker_solve():
    """coda::spectrum0.ar = R's ar.yw (order <= 10 log10 n by AIC, var.pred scaled by n / (n - (p + 1))) evaluated at frequency zero.
    The Levinson-Durbin recursion of summaries._spectrum0_ar (and of the device kernel that restates it) against an independent
    route: the Yule-Walker equations of every order solved as Toeplitz systems, the same AIC and the same variance scaling."""
    from scipy.linalg import solve_toeplitz
    rng = np.random.default_rng(4)
    for rho, n in ((0.0, 700), (0.6, 1500), (0.95, 4000), (-0.5, 333)):
        x = np.zeros(n); e = rng.normal(size=n)
        for t in range(1, n):
            x[t] = rho * x[t - 1] + e[t] + (0.3 * e[t - 1] if rho == 0.6 else 0.0)
        xc = x - x.mean()
        omax = int(min(n - 1, np.floor(10 * np.log10(n))))
        r = np.array([np.dot(xc[: n - k], xc[k:]) / n for k in range(omax + 1)])
        best = (n * np.log(r[0]), r[0], np.zeros(0))
        for p in range(1, omax + 1):
            phi = solve_toeplitz(r[:p], r[1:p + 1])
            v = r[0] - np.dot(phi, r[1:p + 1])
            aic = n * np.log(v) + 2 * p
            if aic < best[0]:
                best = (aic, v, phi)
        _, v, phi = best
        want = v * n / (n - (len(phi) + 1)) / (1 - phi.sum()) ** 2
        got = summaries._spectrum0_ar(x)
        assert abs(got - want) <= 1e-8 * want, (rho, n, got, want)
        assert abs(summaries.effective_size(x) - n * x.var(ddof=1) / want) <= 1e-8 * n


def test_rhat_from_moments_equals_rhat_from_draws():
    x = np.random.default_rng(1).normal(size=(5, 300, 3)) + np.arange(5)[:, None, None] * 0.1
    np.testing.assert_allclose(summaries.rhat_from_moments(*summaries.chain_moments(x)), summaries.rhat_per_parameter(x))


def test_posterior_downsample_is_weighted_without_replacement():
    """`sample(1:n, dwnsmpl, prob = convert_post_probs(logpost))` (R/summarize_plot.R:296-303): no repeats, increasing
    row order, rows whose probability underflows in linear space are still reachable, high-posterior rows dominate."""
    from covidcurve_b200 import posterior
    n = 400
    rng = np.random.default_rng(0)
    ll = rng.normal(-500.0, 3.0, n)
    ll[:50] -= 5000.0                                       # exp() underflows to exactly zero for these rows
    out = pd.DataFrame({"chain": "chain1", "rung": "rung1", "iteration": np.arange(1, n + 1), "stage": "sampling",
                        "logprior": np.zeros(n), "loglikelihood": ll})
    fit = {"mcmcout": {"output": out, "diagnostics": {"rung_details": pd.DataFrame({"rung": [1], "thermodynamic_power": [1.0]})}}}
    got = posterior._downsample(fit, None, 100, seed=3)
    assert len(got) == 100 and got["iteration"].is_monotonic_increasing and got["iteration"].is_unique
    assert (got["iteration"] > 50).all()
    allrows = posterior._downsample(fit, None, n, seed=3)  # every row, including the underflowed ones
    assert len(allrows) == n
    top = np.argsort(-ll)[:40] + 1
    hits = np.mean([np.isin(top, posterior._downsample(fit, None, 100, seed=s)["iteration"]).mean() for s in range(20)])
    assert hits > 0.6                                       # uniform sampling would give 0.25
    with pytest.raises(ValueError):
        posterior._downsample(fit, None, n + 1, seed=3)
    np.testing.assert_allclose(posterior.convert_post_probs(ll).sum(), 1.0, rtol=1e-12)


def test_simulator_mirror_validates_like_the_reference():
    """R/simulators.R:124-147: the argument checks run before any device work."""
    from covidcurve_b200 import simulators
    fat = pd.DataFrame({"Strata": ["ma1", "ma2"], "IFR": [0.01, 0.1], "Rho": [1.0, 1.0]})
    demog = pd.DataFrame({"Strata": ["ma1", "ma2"], "popN": [1e5, 2e5]})
    inf = np.full(50, 10.0)
    kw = dict(curr_day=50, spec=0.95, sens=0.85, demog=demog, sero_delay_rate=18.3)
    with pytest.raises(ValueError, match="Strata, IFR, Rho"):
        simulators.Agesim_infxn_2_death(fat.rename(columns={"Rho": "rho"}), inf, **kw)
    with pytest.raises(ValueError, match="same order"):
        simulators.Agesim_infxn_2_death(fat, inf, **dict(kw, demog=demog.iloc[::-1].reset_index(drop=True)))
    with pytest.raises(ValueError, match="one entry per day"):
        simulators.Agesim_infxn_2_death(fat, inf[:-1], **kw)
    with pytest.raises(ValueError, match="spec"):
        simulators.Agesim_infxn_2_death(fat, inf, **dict(kw, spec=1.5))
    with pytest.raises(ValueError, match="smplfrac"):
        simulators.Agesim_infxn_2_death(fat, inf, smplfrac=0.0, **kw)
    with pytest.raises(ValueError, match="sero_rev_shape"):
        simulators.Agesim_infxn_2_death(fat, inf, simulate_seroreversion=True, sero_rev_shape=None, sero_rev_scale=140.0, **kw)
    with pytest.raises(NotImplementedError):
        simulators.Agesim_infxn_2_death(fat, inf, return_linelist=True, **kw)
