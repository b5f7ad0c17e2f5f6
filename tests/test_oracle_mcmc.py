"""The CPU restatement of drjacoby's sampler (oracle/cc_oracle_mcmc.cpp): bookkeeping and statistical sanity.
Parity with drjacoby itself is unpinned (its source is not available); what is checked here are the facts the stored
reference fit exposes: iteration 1 is the init vector, per-parameter move rate ~0.23 after adaptation,
and that the chain climbs to the likelihood level the reference fit reaches."""
import numpy as np

from oracle.oracle import Oracle


def test_single_rung_chain_bookkeeping_and_adaptation(modfit):
    spec, g = modfit
    o = Oracle(spec)
    r = o.mcmc(burnin=600, samples=400, rungs=1, chains=2, seed=11, nthreads=2)
    th, ll, lp = r["theta"], r["loglike"], r["logprior"]
    assert th.shape == (2, 1, 1000, spec.d)
    assert np.array_equal(th[:, 0, 0, :], np.tile(spec.pinit, (2, 1)))          # iteration 1 = init
    assert ll[0, 0, 0] == o.loglike(spec.pinit)[0] and lp[0, 0, 0] == o.logprior(spec.pinit)[0]
    # recorded (ll, lp) are those of the recorded state
    k = 777
    assert ll[1, 0, k] == o.loglike(th[1, 0, k])[0] and lp[1, 0, k] == o.logprior(th[1, 0, k])[0]
    # single-site moves: consecutive rows differ in at most d coordinates, every state inside the box
    assert np.all(th >= spec.pmin) and np.all(th <= spec.pmax)
    moved = (np.diff(th[:, 0, 500:], axis=1) != 0).mean(axis=(0, 1))
    assert np.all(moved > 0.08) and np.all(moved < 0.5), moved                    # Robbins-Monro target 0.234
    assert ll[:, 0, -1].min() > -2000 > ll[0, 0, 0]                               # the stored fit's mode is ~ -722
    assert np.all(r["beta"] == [1.0])


def test_tempered_ladder_and_swaps(modfit):
    spec, g = modfit
    r = Oracle(spec).mcmc(burnin=150, samples=150, rungs=4, chains=1, seed=5, GTI_pow=2.0)
    np.testing.assert_allclose(r["beta"], (np.arange(4) / 3.0) ** 2.0)
    acc = r["mc_accept"][0]
    assert acc.shape == (3,) and np.all((acc >= 0) & (acc <= 1))
    # the cold rung (last) reaches higher likelihood than the prior-only rung (first, beta = 0)
    assert r["loglike"][0, 3, -50:].mean() > r["loglike"][0, 0, -50:].mean()


def test_chain_streams_are_keyed_by_global_chain_id(modfit):
    spec, _ = modfit
    o = Oracle(spec)
    both = o.mcmc(burnin=30, samples=10, rungs=2, chains=2, seed=3)
    second = o.mcmc(burnin=30, samples=10, rungs=2, chains=1, seed=3, chain_offset=1)
    assert np.array_equal(both["theta"][1], second["theta"][0])
