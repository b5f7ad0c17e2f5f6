"""Pin the CPU oracle (oracle/) to the reference: the stored fit's (parameter vector -> loglikelihood, logprior) rows
written by the reference C++ under R 4.0.3 + drjacoby (data/covidcurve_modfit.rda), and the inputs of the
reference's own unit test (tests/testthat/test-check_natcub_cpp.R) rebuilt from data/covidcurve_simdata.rda."""
import numpy as np

from oracle.oracle import Oracle


def test_oracle_reproduces_the_stored_reference_fit(modfit):
    spec, g = modfit
    o = Oracle(spec)
    ll = o.loglike(g["theta"], nthreads=8)
    lp = o.logprior(g["theta"], nthreads=8)
    assert np.max(np.abs(ll - g["loglike"]) / np.abs(g["loglike"])) <= 1e-12
    assert np.max(np.abs(lp - g["logprior"]) / np.abs(g["logprior"])) <= 1e-12
    # row 0 of the fit is the init vector (SURVEY.md 8c smoke value)
    i0 = int(np.nonzero(g["rows"] == 0)[0][0])
    assert ll[i0] == -231967.69462831024 or abs(ll[i0] / -231967.69462831024 - 1) < 1e-15
    assert abs(lp[i0] / -44.18760904148496 - 1) < 1e-14


def test_fff_quirk_sod_prior_is_dropped(modfit):
    """With all reparam flags FALSE the generated prior ends before the sod Beta term (Agecpp2strings.R:202)."""
    spec, g = modfit
    assert str(g["logprior_string"]).rstrip().count("R::dbeta(sod") == 1
    o = Oracle(spec)
    th = g["theta"][:5].copy()
    lp0 = o.logprior(th)
    th[:, spec.index("sod")] = 0.5   # far outside where Beta(2550, 450) has mass
    assert np.array_equal(o.logprior(th), lp0)


def test_reference_unit_test_ordering(simdata_binomial, simdata_logit):
    """test-check_natcub_cpp.R:135,152: morelikely > lesslikely for both serology likelihoods."""
    for (spec, g), anchors in ((simdata_binomial, (-4441.569153555149, -544840.7303090655)),
                               (simdata_logit, (-3317.7333646918946, -118054.91121335802))):
        ll = Oracle(spec).loglike(np.stack([g["morelikely"], g["lesslikely"]]))
        assert ll[0] > ll[1]
        np.testing.assert_allclose(ll, anchors, rtol=1e-9)


def test_param_i_is_ignored_and_failures_are_in_band(modfit):
    spec, g = modfit
    o = Oracle(spec)
    th = g["theta"][7].copy()
    import ctypes as C
    f = lambda i: o.L.cco_loglike(C.byref(o.desc), th.ctypes.data_as(C.POINTER(C.c_double)), i, None)
    assert f(1) == f(13)
    th[spec.index("x2")] = th[spec.index("x1")]
    assert f(1) == -(np.finfo(float).max / 100)


def test_oracle_is_immune_to_flush_to_zero_hosts(modfit):
    """Denormal gamma-cdf values must not flush to zero even if the host process runs FTZ/DAZ (import torch can)."""
    import torch  # noqa: F401
    torch.set_flush_denormal(True)
    try:
        spec, g = modfit
        th = np.tile(g["theta"][200], (40, 1))
        th[:, spec.index("sod")] = np.geomspace(0.004, 0.03, 40)
        a = Oracle(spec).loglike(th, nthreads=4)
    finally:
        torch.set_flush_denormal(False)
    b = Oracle(spec).loglike(th, nthreads=1)
    assert np.array_equal(a, b)
