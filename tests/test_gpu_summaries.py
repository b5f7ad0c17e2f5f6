"""Device-side posterior summaries (csrc/cc_summ.cu) and the expected-deaths check (cc_post_deaths_batch) through the C ABI,
against the host restatements of the R-side consumers: covidcurve_b200/summaries.py (R/summarize_plot.R:46-264 + coda) and the
oracle's get_post_deaths loop (R/summarize_plot.R:605-641).  Tolerances: these are floating-point reductions in a different
order than numpy's, so 1e-9 relative (1e-7 for ESS / Geweke, whose AR fit amplifies the autocovariance rounding)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ar1_draws(C=6, n=1500, P=5, seed=3):
    rng = np.random.default_rng(seed)
    rho = np.linspace(0.0, 0.95, P)
    x = np.zeros((C, n, P))
    e = rng.standard_normal((C, n, P))
    for t in range(1, n):
        x[:, t, :] = rho * x[:, t - 1, :] + e[:, t, :]
    return x * np.array([1.0, 1e-3, 50.0, 1.0, 2.0]) + np.array([0.0, 0.01, -300.0, 5.0, 1e4]) + rng.standard_normal((C, 1, P)) * 0.05


def _host_table(x):
    from scipy import stats
    from covidcurve_b200.summaries import effective_size, geweke_z
    q = np.quantile(x, [0.025, 0.5, 0.975])
    z = geweke_z(x)
    return np.array([x.min(), q[0], q[1], x.mean(), q[2], x.max(), effective_size(x), z, stats.norm.pdf(z)])


def _close(got, want, rtol):
    return np.allclose(got, want, rtol=rtol, atol=rtol * 1e-3)


def test_cred_interval_table_by_chain_and_pooled(built_lib):
    from covidcurve_b200 import _lib
    x = _ar1_draws()
    C, n, P = x.shape
    tab = _lib.summarize_draws(x, by_chain=True)
    assert tab.shape == (C, P, len(_lib.SUMM_FIELDS))
    for c in range(C):
        for p in range(P):
            want = _host_table(x[c, :, p])
            assert _close(tab[c, p, :6], want[:6], 1e-12), (c, p)
            assert _close(tab[c, p, 6:9], want[6:9], 1e-7), (c, p, tab[c, p, 6:9], want[6:9])
            assert tab[c, p, 11] == n and _close(tab[c, p, 9], x[c, :, p].var(ddof=1), 1e-10)
    pooled = _lib.summarize_draws(x, by_chain=False)
    assert pooled.shape == (1, P, len(_lib.SUMM_FIELDS))
    for p in range(P):
        want = _host_table(x[:, :, p].reshape(-1))          # chains concatenated in chain order
        assert _close(pooled[0, p, :6], want[:6], 1e-12) and _close(pooled[0, p, 6:9], want[6:9], 1e-7)


def test_quantiles_are_exact_order_statistics(built_lib):
    """the radix select returns the same order statistics as a sort: ties, negative values, zeros, tiny and huge magnitudes"""
    from covidcurve_b200 import _lib
    rng = np.random.default_rng(5)
    x = np.stack([np.concatenate([rng.standard_normal(500), np.zeros(40), -np.zeros(3), [1e300, -1e300, 5e-324, -5e-324]]),
                  np.repeat(rng.integers(-3, 4, 547 // 7 + 1), 7)[:547].astype(float)], axis=1)[None]
    tab = _lib.summarize_draws(x, by_chain=True)
    for p in range(2):
        q = np.quantile(x[0, :, p], [0.025, 0.5, 0.975])
        assert np.array_equal(tab[0, p, [1, 2, 4]], q)
        assert tab[0, p, 0] == x[0, :, p].min() and tab[0, p, 5] == x[0, :, p].max()


def test_short_and_constant_series(built_lib):
    from covidcurve_b200 import _lib
    from covidcurve_b200.summaries import effective_size
    x = np.zeros((2, 7, 2))
    x[:, :, 0] = 3.25                                        # constant: spectrum 0 -> ESS 0 (summaries.effective_size)
    x[:, :, 1] = np.arange(14).reshape(2, 7) ** 2
    tab = _lib.summarize_draws(x, by_chain=True)
    assert np.all(tab[:, 0, 6] == 0.0) and np.all(tab[:, 0, [0, 1, 2, 3, 4, 5]] == 3.25)
    assert _close(tab[1, 1, 6], effective_size(x[1, :, 1]), 1e-9)
    one = _lib.summarize_draws(np.array([[[2.0]], [[4.0]]]), by_chain=False)     # two chains x one draw, pooled
    assert one[0, 0, 0] == 2.0 and one[0, 0, 5] == 4.0 and one[0, 0, 2] == 3.0 and one[0, 0, 6] == 0.0


def test_gelman_diag_and_rhat(built_lib):
    from covidcurve_b200 import summaries
    x = _ar1_draws(C=8, n=1001)
    for burn in (True, False):
        est, upper, rhat = summaries.device_gelman_diag(x, autoburnin=burn)
        e0, u0 = summaries.gelman_diag(x, autoburnin=burn)
        assert _close(est, e0, 1e-10) and _close(upper, u0, 1e-9)
        xs = x[:, x.shape[1] // 2:, :] if burn else x
        assert _close(rhat, summaries.rhat_per_parameter(xs), 1e-10)
    with pytest.raises(Exception):
        summaries.device_gelman_diag(x[:1])                  # Gelman-Rubin requires at least two chains


def test_overall_ifr_both_standards(built_lib):
    from covidcurve_b200 import _lib, summaries
    rng = np.random.default_rng(11)
    C, n, A = 3, 400, 4
    names = ["ma%d" % a for a in range(A)] + ["x"] + ["ne%d" % a for a in range(A)]
    x = rng.uniform(0.01, 0.4, (C, n, len(names)))
    pop = np.array([1e5, 3e5, 2.5e5, 5e4])
    ma, ne = x[:, :, :A], x[:, :, A + 1:]
    want_pop = (ma * (pop / pop.sum())).sum(-1)
    w = ne * pop
    want_ar = (ma * (w / w.sum(-1, keepdims=True))).sum(-1)
    assert _close(_lib.overall_ifr_draws(x, range(A), pop), want_pop, 1e-13)
    assert _close(_lib.overall_ifr_draws(x, range(A), pop, range(A + 1, 2 * A + 1)), want_ar, 1e-13)
    df = summaries.device_overall_IFR_cred_intervals(x, names, names[:A], pop, Noiseparams=names[A + 1:], by_chain=True)
    assert list(df["chain"]) == [1, 2, 3]
    for c in range(C):
        assert _close(df.loc[c, ["min", "LCI", "median", "mean", "UCI", "max"]].to_numpy(dtype=float), _host_table(want_ar[c])[:6], 1e-12)


def test_summaries_of_a_device_resident_run(built_lib):
    """the recorded draws of a GPU run summarised in place (strided view of the run state, no copy to the host) equal the
    summaries of the fetched draws; the sharded helpers reduce to the same tables in a single process"""
    import torch
    from covidcurve_b200 import _lib, dist, summaries
    from tests.golden import specs
    spec, g = specs.modfit()
    model = _lib.Model(spec, device=0)
    chains, burn, samp = 4, 150, 300
    mc = _lib.Mcmc(model, burnin=burn, samples=samp, rungs=3, chains=chains, seed=21, record_rungs=False, init=np.tile(g["theta"][0], (chains, 1)))
    mc.advance(burn + samp - 1)
    rows = mc.rows
    dev = torch.device("cuda", 0)
    draws = dist.draws_tensor(mc, dev)[:, 0, rows - samp:rows, :]          # [chains, samples, d], strided, in HBM
    assert not draws.is_contiguous()
    host = mc.fetch()["theta"][:, 0, rows - samp:rows, :]
    assert np.array_equal(draws.cpu().numpy(), host)
    t_dev = _lib.summarize_draws(draws, by_chain=True)
    t_host = _lib.summarize_draws(host, by_chain=True)
    assert np.array_equal(t_dev, t_host)
    moving = [p for p in range(spec.d) if host[:, :, p].std() > 0]
    for p in moving[:6]:
        assert _close(t_dev[1, p, :6], _host_table(host[1, :, p])[:6], 1e-12)
        assert _close(t_dev[1, p, 6:9], _host_table(host[1, :, p])[6:9], 1e-6)
    est, upper, rhat = summaries.device_gelman_diag(draws)
    e0, u0 = summaries.gelman_diag(host[:, :, moving])
    assert _close(est[moving], e0, 1e-9) and _close(upper[moving], u0, 1e-8)
    e1, u1, r1 = dist.gelman_allgather(draws, chains)                        # no process group: the gather is the identity
    assert np.array_equal(e1, est) and np.array_equal(u1, upper) and np.array_equal(r1, rhat)
    df = dist.cred_intervals_allgather(draws, chains, spec.names)
    assert len(df) == chains * spec.d and np.array_equal(df["median"].to_numpy().reshape(chains, spec.d), t_dev[:, :, 2])
    mc.close()
    model.close()


def test_expected_deaths_check_against_the_reference_loop(built_lib):
    """cc_post_deaths_batch against the oracle's loop-for-loop get_post_deaths (R/summarize_plot.R:605-641) on draws around the
    truth of two shapes; a draw with unordered knots has no curve (NaN block)"""
    from covidcurve_b200 import _lib, synth
    from oracle.oracle import Oracle
    for cfg in ("cfg2", "cfg3"):
        spec, truth = synth.make_spec(cfg)
        th = synth.random_theta(spec, 6, seed=77, bad_frac=0.0, near_truth=synth.true_theta(spec, truth), jitter=0.05)
        model = _lib.Model(spec, device=0)
        got = model.post_deaths(th)
        o = Oracle(spec)
        for r in range(len(th)):
            want = o.post_deaths(th[r])
            rel = np.abs(got[r] - want) / np.maximum(np.abs(want), 1e-300)
            assert rel.max() <= 1e-10, (cfg, r, rel.max())
        bad = th[:1].copy()
        k = [spec.index(p) for p in spec.Knotparams]
        bad[0, k[0]], bad[0, k[1]] = bad[0, k[1]], bad[0, k[0]]
        assert np.isnan(model.post_deaths(bad)).all()
        model.close()
