"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the reference's stored values.

Tolerance: north_star asks for 1e-10 relative in double precision; in-band failure values must agree exactly.
"""
import dataclasses

import numpy as np
import pytest

from tests.conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _model(spec):
    from covidcurve_b200 import _lib
    return _lib.Model(spec, device=0)


def _oracle(spec):
    from oracle.oracle import Oracle
    return Oracle(spec)


def test_golden_rows_of_the_stored_reference_fit(built_lib, modfit):
    spec, g = modfit
    m = _model(spec)
    ll = m.loglike(g["theta"])
    assert relerr(ll, g["loglike"]) <= TOL          # values produced by the reference C++ under R 4.0.3
    assert relerr(ll, _oracle(spec).loglike(g["theta"], nthreads=8)) <= TOL
    lp = m.logprior(g["theta"])
    assert relerr(lp, g["logprior"]) <= TOL
    assert ll[0] == pytest.approx(-231967.69462831024, rel=1e-12) or g["rows"][0] != 0


def test_reference_unit_test_vectors(built_lib, simdata_binomial, simdata_logit):
    """tests/testthat/test-check_natcub_cpp.R: morelikely > lesslikely for both likelihoods (+ survey-time anchors)."""
    for (spec, g), anchors in ((simdata_binomial, (-4441.569153555149, -544840.7303090655)),
                               (simdata_logit, (-3317.7333646918946, -118054.91121335802))):
        m = _model(spec)
        th = np.stack([g["morelikely"], g["lesslikely"]])
        ll = m.loglike(th)
        assert ll[0] > ll[1]
        assert relerr(ll, _oracle(spec).loglike(th)) <= TOL
        assert relerr(ll, np.array(anchors)) <= 1e-9
        assert m.loglike_scalar(th[0], param_i=3) == ll[0]   # param_i is ignored, like binomial.cpp:7


def test_reference_call_shape_params_param_i_data_misc(built_lib, simdata_binomial, simdata_logit):
    """The static twins' own 4-argument shape (src/RcppExports.cpp:10-35): a NAMED parameter vector plus the data and misc lists,
    exactly as tests/testthat/test-check_natcub_cpp.R:90-150 calls COVIDCurve:::natcubspline_loglike_binomial / _logit."""
    from covidcurve_b200 import _lib
    for (spec, g), binomial in ((simdata_binomial, True), (simdata_logit, False)):
        misc = dict(rcensor_day=int(g["rcensor_day"]), days_obsd=int(g["days_obsd"]), n_knots=int(g["n_knots"]), n_sero_obs=2,
                    sero_survey_start=g["sero_survey_start"], sero_survey_end=g["sero_survey_end"],
                    max_seroday_obsd=int(g["max_seroday_obsd"]), demog=g["demog"], account_serorev=False)
        data = dict(obs_deaths=g["obs_deaths"], prop_strata_obs_deaths=g["prop_strata_obs_deaths"])
        data.update(dict(obs_serologypos=g["obs_serologypos"], obs_serologyn=g["obs_serologyn"]) if binomial else
                    dict(obs_serologymu=g["obs_serologymu"], obs_serologyse=g["obs_serologyse"]))
        names = [str(n) for n in g["param_names"]]
        # the test's vectors carry two NA entries "just to make visible to cpp" and are in another order than df_params
        order = ["mod", "sod", "ma1", "ma2", "ma3", "x1", "x2", "x3", "x4", "y1", "y2", "y3", "y4", "y5", "ne1", "ne2", "ne3", "sens", "spec",
                 "sero_con_rate"]
        mk = lambda v: {**{n: float(v[names.index(n)]) for n in order}, "sero_rev_shape": np.nan, "sero_rev_scale": np.nan}
        more = _lib.natcubspline_loglike(mk(g["morelikely"]), 1, data, misc, binomial=binomial)["LogLik"]
        less = _lib.natcubspline_loglike(mk(g["lesslikely"]), 1, data, misc, binomial=binomial)["LogLik"]
        assert more > less                                                   # the reference test's own assertion (:135, :152)
        want = _oracle(spec).loglike(np.stack([g["morelikely"], g["lesslikely"]]))
        assert relerr(np.array([more, less]), want) <= TOL
        assert _lib.natcubspline_loglike(mk(g["morelikely"]), 7, data, misc, binomial=binomial)["LogLik"] == more   # param_i ignored
        bad = mk(g["morelikely"])
        del bad["x3"]
        with pytest.raises(_lib.CCError, match="no element named 'x3'"):
            _lib.natcubspline_loglike(bad, 1, data, misc, binomial=binomial)
        # other data -> another cached model, not a stale answer
        data2 = dict(data, obs_deaths=np.where(g["obs_deaths"] > 0, g["obs_deaths"] + 1, g["obs_deaths"]))
        assert _lib.natcubspline_loglike(mk(g["morelikely"]), 1, data2, misc, binomial=binomial)["LogLik"] != more
        assert _lib.natcubspline_loglike(mk(g["morelikely"]), 1, data, misc, binomial=binomial)["LogLik"] == more


@pytest.mark.parametrize("cfg,kw", [
    ("cfg2", {}), ("cfg2", dict(reparam=True)), ("cfg2", dict(binomial=True)),
    ("cfg3", {}), ("cfg3", dict(serorev=True)), ("cfg3", dict(reparam=True)), ("cfg4", {}), ("cfg4", dict(serorev=True, binomial=False)),
])
def test_synthetic_configs_random_vectors(built_lib, cfg, kw):
    from covidcurve_b200 import synth
    spec, truth = synth.make_spec(cfg, **kw)
    n = 1500 if cfg != "cfg4" else 600
    th = synth.random_theta(spec, n, seed=11, bad_frac=0.02)
    if not kw.get("reparam"):
        near = synth.random_theta(spec, n, seed=12, bad_frac=0.0, near_truth=synth.true_theta(spec, truth), jitter=0.02)
        th = np.concatenate([th, near])
    m, o = _model(spec), _oracle(spec)
    want = o.loglike(th, nthreads=8)
    got = m.loglike(th)
    assert relerr(got, want) <= TOL
    clamp = -(np.finfo(float).max / 100)
    assert 0 < np.sum(want == clamp) < len(want)
    assert relerr(m.logprior(th), o.logprior(th)) <= TOL


def test_edge_cases_censoring_missing_skips(built_lib, modfit):
    spec, g = modfit
    th = g["theta"][:300]
    variants = []
    variants.append(dataclasses.replace(spec, rcensor_day=150))                      # right censoring (binomial.cpp:205,216)
    od = spec.obs_deaths.copy(); od[::7] = -1; od[3] = -1
    variants.append(dataclasses.replace(spec, obs_deaths=od))                         # missing deaths sentinel
    sa, sb = spec.sero_a.copy(), spec.sero_b.copy(); sa[1] = sb[1] = -1; sa[4] = -1
    variants.append(dataclasses.replace(spec, sero_a=sa, sero_b=sb))                 # skip only if BOTH are -1
    variants.append(dataclasses.replace(spec, rcensor_day=1))                         # everything censored
    variants.append(dataclasses.replace(spec, sero_survey_start=np.array([1, 150]), sero_survey_end=np.array([1, 165])))
    for v in variants:
        got, want = _model(v).loglike(th), _oracle(v).loglike(th)
        assert relerr(got, want) <= TOL


def test_early_exits_and_non_finite_inputs(built_lib, modfit):
    spec, g = modfit
    base = g["theta"][100].copy()
    ix = {n: i for i, n in enumerate(spec.names)}
    th = np.tile(base, (8, 1))
    th[0, ix["x2"]] = th[0, ix["x1"]]             # equal knots -> not strictly increasing
    th[1, ix["x3"]] = 1.0                          # knot at/below day 1
    th[2, ix["y5"]] = 30.0                         # more infections than people
    th[3, ix["sod"]] = np.nan
    th[4, ix["ma1"]] = -0.1                        # negative Poisson mean -> NaN -> clamp
    th[5, ix["spec"]] = -5.0                      # p outside [0,1] -> NaN -> clamp
    th[6, ix["x1"]] = 2.0000001; th[6, ix["x2"]] = 2.5   # two knots inside one day (interval-lag rule)
    th[7, ix["sod"]] = 1.7                         # shape < 1 (continued-fraction branch of pgamma)
    got, want = _model(spec).loglike(th), _oracle(spec).loglike(th)
    clamp = -(np.finfo(float).max / 100)
    assert list(want[:6] == clamp) == [True] * 6
    assert relerr(got, want) <= TOL


def test_gamma_table_regimes(built_lib, modfit):
    """sod sweeps the incomplete-gamma regimes: shape 1/sod^2 from ~1 to 1e4 (series, CF, Temme asymptotics)."""
    spec, g = modfit
    spec = dataclasses.replace(spec, pmin=spec.pmin.copy(), pmax=spec.pmax.copy())
    base = g["theta"][g["loglike"].argmax()].copy()
    i_sod, i_mod = spec.index("sod"), spec.index("mod")
    sods = np.concatenate([np.geomspace(0.01, 0.999, 160), [0.05, 0.1, 0.2, 0.5, 1.0, 1.3]])
    th = np.tile(base, (len(sods) * 3, 1))
    th[:, i_sod] = np.tile(sods, 3)
    th[:, i_mod] = np.repeat([5.0, 19.8, 60.0], len(sods))
    got, want = _model(spec).loglike(th), _oracle(spec).loglike(th, nthreads=8)
    assert relerr(got, want) <= TOL


def test_gamma_table_far_lower_tail_log_space(built_lib, modfit):
    """sod in (0.001, 0.0055): shapes above 33 500, where gamma-cdf values below 1.002e-292 come out of the Temme-type
    asymptotic branch and R redoes them in log space through pnorm's log tail (Cody 1969): device = oracle element by element
    and through the whole likelihood."""
    from covidcurve_b200 import _lib
    from oracle.oracle import lib
    Lo = lib()
    rng = np.random.default_rng(5)
    n = 6000
    a = np.exp(rng.uniform(np.log(33500.0), np.log(1e6), n))
    z = rng.uniform(-37.5, -34.0, n)                                   # P between 0 (below 1e-324) and ~1e-260
    x = a + z * np.sqrt(a)
    want = np.array([Lo.cco_pgamma(xi, ai, 1.0, 1, 0) for xi, ai in zip(x, a)])
    got = _lib.nmath("pgamma", x, a, 1.0)
    tiny = want < 1.002e-292
    assert tiny.sum() > 1500 and (want[tiny] > 0).sum() > 1000
    assert np.all(np.abs(got - want) <= np.maximum(2 * 4.9406564584124654e-324, 1e-11 * want))
    spec, g = modfit
    base = g["theta"][g["loglike"].argmax()].copy()
    i_sod, i_mod = spec.index("sod"), spec.index("mod")
    sods = np.geomspace(0.001, 0.0056, 80)
    mods = np.array([5.0, 12.3, 18.0, 19.8, 33.0, 60.0])
    th = np.tile(base, (len(sods) * len(mods), 1))
    th[:, i_sod] = np.tile(sods, len(mods))
    th[:, i_mod] = np.repeat(mods, len(sods))
    got, want = _model(spec).loglike(th), _oracle(spec).loglike(th, nthreads=8)
    assert relerr(got, want) <= TOL


def test_device_pointer_path_and_soa_leading_dimension(built_lib, modfit):
    import torch
    from covidcurve_b200 import _lib
    spec, g = modfit
    m = _model(spec)
    n, ld = 777, 1024
    th = torch.zeros((spec.d, ld), dtype=torch.float64)
    th[:, :n] = torch.from_numpy(np.ascontiguousarray(g["theta"][:n].T))
    d_th = th.cuda()
    d_out = torch.empty(n, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    m.loglike_ptr(d_th.data_ptr(), n, ld, d_out.data_ptr(), _lib.CC_MEM_DEVICE, st)
    torch.cuda.synchronize()
    assert relerr(d_out.cpu().numpy(), g["loglike"][:n]) <= TOL
    ms, launches = m.last_kernel_ms()
    assert ms > 0 and launches == 5   # phase-1 pass, bucket offsets, scatter, phase-2 pass, phase-3 pass
    m.logprior_ptr(d_th.data_ptr(), n, ld, d_out.data_ptr(), _lib.CC_MEM_DEVICE, st)
    torch.cuda.synchronize()
    assert relerr(d_out.cpu().numpy(), g["logprior"][:n]) <= TOL
    # empty batch is a no-op, ragged leading dimension is rejected
    m.loglike_ptr(d_th.data_ptr(), 0, ld, d_out.data_ptr(), _lib.CC_MEM_DEVICE, st)
    with pytest.raises(_lib.CCError):
        m.loglike_ptr(d_th.data_ptr(), n, n - 1, d_out.data_ptr(), _lib.CC_MEM_DEVICE, st)


def test_large_host_batch_is_chunked_consistently(built_lib, modfit):
    """> 1 staging chunk (65 536 vectors) through the host path: same values as small batches, order preserved."""
    spec, g = modfit
    m = _model(spec)
    reps = 70000 // len(g["theta"]) + 1
    th = np.tile(g["theta"], (reps, 1))[:70000]
    ll = m.loglike(th)
    assert relerr(ll, np.tile(g["loglike"], reps)[:70000]) <= TOL


def test_full_size_linearity_property(built_lib):
    """Size-independent property at cfg3 shape: scaling every IFR by c scales expected deaths by c, so with
    ma -> c*ma the Poisson term changes but sero term does not; checked against the oracle on a subsample and by
    the invariance of the value under a permutation of the batch (no cross-vector leakage) at 200k vectors."""
    from covidcurve_b200 import synth
    spec, truth = synth.make_spec("cfg3")
    n = 200_000
    th = synth.random_theta(spec, n, seed=5, bad_frac=0.01)
    m = _model(spec)
    ll = m.loglike(th)
    perm = np.random.default_rng(0).permutation(n)
    ll_p = m.loglike(th[perm])
    assert np.array_equal(ll[perm], ll_p)          # bit-identical, independent of position in the batch
    sub = perm[:3000]
    assert relerr(ll[sub], _oracle(spec).loglike(th[sub], nthreads=8)) <= TOL


def test_baseline_size_batch_properties(built_lib):
    """BASELINE.json cfg5 at its full size (2^20 box-uniform vectors): the value of a vector does not depend on how the
    batch is cut - one device launch, two half launches, the host-buffer path with its geometric chunks on two streams and
    the reversed batch all agree bit for bit - and a random subset agrees with the oracle."""
    import torch
    from covidcurve_b200 import _lib, synth
    spec, truth = synth.make_spec("cfg3")
    n = 1 << 20
    th = synth.random_theta(spec, n, seed=synth.SEED_THETA, bad_frac=0.01)
    m = _model(spec)
    soa = np.ascontiguousarray(th.T)
    host = m.loglike(th)                                              # host buffers: chunks of 64k .. 512k vectors
    d_th = torch.from_numpy(soa).cuda()
    d_out = torch.empty(n, dtype=torch.float64, device="cuda")
    m.loglike_ptr(d_th.data_ptr(), n, n, d_out.data_ptr(), _lib.CC_MEM_DEVICE, 0)
    torch.cuda.synchronize()
    one = d_out.cpu().numpy()
    assert np.array_equal(one, host)
    h = n // 2 + 12345                                                # two ragged launches into the same output
    d_out.zero_()
    m.loglike_ptr(d_th.data_ptr(), h, n, d_out.data_ptr(), _lib.CC_MEM_DEVICE, 0)
    m.loglike_ptr(d_th.data_ptr() + 8 * h, n - h, n, d_out.data_ptr() + 8 * h, _lib.CC_MEM_DEVICE, 0)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), one)
    assert np.array_equal(m.loglike(th[::-1])[::-1], one)
    fail = one == -np.finfo(float).max / 100
    assert 0.02 < fail.mean() < 0.06 and np.isfinite(one).all()       # the 1 % bad slice + population-check failures
    sub = np.random.default_rng(1).choice(n, 2000, replace=False)
    # force the hardest band of the box into the sample: for sod in [0.0214, 0.0232] expected counts become subnormal and the
    # reference's gradual-underflow semantics decide the value
    sod_all = th[:, spec.index("sod")]
    hard = np.nonzero((sod_all > 0.021) & (sod_all < 0.0235))[0][:400]
    sub = np.unique(np.concatenate([sub, hard]))
    assert len(hard) >= 300 and relerr(one[sub], _oracle(spec).loglike(th[sub], nthreads=8)) <= TOL


def test_subnormal_gamma_cdf_follows_the_reference_rounding(built_lib):
    """R's pgamma_raw redoes results below DBL_MIN / DBL_EPSILON in log space (one exp, one rounding).  Where the value is
    subnormal with fewer than ~1e9 quanta that single rounding decides every bit, so device and oracle must agree EXACTLY;
    above that the two differ by their ordinary 1e-13 relative rounding."""
    from scipy.special import gammaln
    from covidcurve_b200 import _lib
    from oracle.oracle import lib
    Lo = lib()
    rng = np.random.default_rng(0)
    n = 12000
    a = rng.uniform(1800, 2300, n)
    u = rng.uniform(-744.4, -700.0, n)                       # log P: the subnormal range and a little above
    f = lambda xx: a * np.log(xx) - xx - gammaln(a + 1) - np.log1p(-xx / (a + 1)) - u
    lo_, hi_ = np.full(n, 1.0), a - 1.0
    for _ in range(70):
        mid = 0.5 * (lo_ + hi_)
        neg = f(mid) < 0
        lo_, hi_ = np.where(neg, mid, lo_), np.where(neg, hi_, mid)
    x = 0.5 * (lo_ + hi_)
    want = np.array([Lo.cco_pgamma(xi, ai, 1.0, 1, 0) for xi, ai in zip(x, a)])
    got = _lib.nmath("pgamma", x, a, 1.0)
    quanta = want / 4.9406564584124654e-324
    few = (want > 0) & (quanta < 1e9)
    assert few.sum() > 3000 and np.array_equal(got[few], want[few])
    rest = (want > 0) & ~few
    assert np.all(np.abs(got[rest] - want[rest]) <= np.maximum(2 * 4.9406564584124654e-324, 5e-12 * want[rest]))
    assert np.array_equal(got == 0, want == 0)


def test_curves_match_oracle_intermediates(built_lib, modfit):
    spec, g = modfit
    m, o = _model(spec), _oracle(spec)
    th = g["theta"][:16]
    cv = m.curves(th)
    for r in range(len(th)):
        tr = o.trace(th[r])
        np.testing.assert_allclose(cv["infxn"][r], tr["infxn"], rtol=1e-12)
        np.testing.assert_allclose(cv["ne"][r], tr["ne"], rtol=1e-13)
        np.testing.assert_allclose(cv["auc"][r], tr["auc"], rtol=1e-10, atol=1e-300)
        # sero_full over the first M days times ne[a], window-averaged, equals the oracle's survey numerators
        for s in range(spec.S):
            w = cv["sero_full"][r][spec.sero_survey_start[s] - 1: spec.sero_survey_end[s]].mean()
            np.testing.assert_allclose(w * cv["ne"][r], tr["sero_num"][s * spec.A:(s + 1) * spec.A], rtol=1e-10)
    # more infections than people: the likelihood exits early (no auc), the sero-curve extraction drops that check
    # (R/summarize_plot.R:729) and an invalid knot order leaves no curve at all
    big, bad = th[0].copy(), th[0].copy()
    big[[spec.index(n) for n in spec.Infxnparams]] = 12.0
    bad[spec.index("x1")] = 199.0
    cv2 = m.curves(np.stack([big, bad, th[1]]))
    assert np.isnan(cv2["auc"][0]).all() and np.isfinite(cv2["sero_full"][0]).all() and np.isfinite(cv2["infxn"][0]).all()
    assert np.isnan(cv2["infxn"][1]).all() and np.isnan(cv2["sero_full"][1]).all() and np.isfinite(cv2["ne"][1]).all()
    np.testing.assert_array_equal(cv2["sero_full"][2], cv["sero_full"][1])
    hz = 1 - np.exp(-(np.arange(spec.T) + 1) / big[spec.index("sero_con_rate")])
    np.testing.assert_allclose(cv2["sero_full"][0], np.convolve(cv2["infxn"][0], hz)[:spec.T], rtol=1e-12)


def test_device_nmath_against_oracle_nmath(built_lib):
    """The device restatements of R's nmath (and the table-driven log of the Poisson loop) element by element."""
    from covidcurve_b200 import _lib
    from oracle.oracle import lib
    L = lib()
    rng = np.random.default_rng(3)
    x = np.floor(rng.uniform(0, 500, 4000)); lam = rng.uniform(1e-3, 600, 4000)
    want = np.array([L.cco_dpois(a, b, 1) for a, b in zip(x, lam)])
    np.testing.assert_allclose(_lib.nmath("dpois", x, lam), want, rtol=1e-12, atol=1e-12)
    n = np.floor(rng.uniform(1, 3e6, 4000)); p = rng.uniform(1e-4, 0.9, 4000); k = np.floor(n * p * rng.uniform(0.5, 1.5, 4000).clip(0, 1 / p))
    k = np.minimum(k, n)
    want = np.array([L.cco_dbinom(a, b, c, 1) for a, b, c in zip(k, n, p)])
    np.testing.assert_allclose(_lib.nmath("dbinom", k, n, p), want, rtol=1e-11, atol=1e-10)
    sod = rng.uniform(0.03, 1.2, 3000); mod = rng.uniform(5, 40, 3000); day = np.floor(rng.uniform(0, 300, 3000))
    want = np.array([L.cco_pgamma(d, 1 / s ** 2, mm * s ** 2, 1, 0) for d, s, mm in zip(day, sod, mod)])
    np.testing.assert_allclose(_lib.nmath("pgamma", day, 1 / sod ** 2, mod * sod ** 2), want, rtol=1e-10, atol=1e-300)
    xs = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2, 20000), [1.0, 5e-324, 1e-310, np.inf]])
    got = _lib.nmath("fastlog", xs)
    assert np.max(np.abs(got[:-1] - np.log(xs[:-1])) / np.maximum(1.0, np.abs(np.log(xs[:-1])))) < 5e-16
    assert got[-1] == np.inf and np.isnan(_lib.nmath("fastlog", np.array([-1.0])))[0] and _lib.nmath("fastlog", np.array([0.0]))[0] == -np.inf

