"""Deterministic synthetic workloads of the shapes BASELINE.json names - no R, no network.

Data generation only: a small numpy/scipy forward model (natural cubic spline -> gamma death delay -> exponential /
Weibull sero delays) produces expected deaths and prevalences from a "true" parameter vector, which are then sampled.
This is NOT the product path (that is the CUDA library) and NOT the oracle (oracle/): it only fabricates inputs, the
way the reference's tests fabricate theirs with R/simulators.R (tests/testthat/test-check_natcub_cpp.R:6-50).

Shapes (A strata, K knots, T days, S surveys, M last survey day):
  cfg2: A=5,  T=200, S=2, logit-normal serology, seroreversion, 4 chains x 20 rungs
  cfg3: A=10, T=300, S=3, binomial serology, 64 chains x 50 rungs           (also cfg5: 1M random vectors)
  cfg4: A=17, T=400, S=3, binomial serology, 4096 chains x 64 rungs
  long: A=4,  T=500, K=6, S=3, binomial serology, seroreversion (test shape for the T > 448 kernels; not a BASELINE config)
"""
from __future__ import annotations

import numpy as np

from .spec import INT_MAX, HotPathSpec

SEED_DATA, SEED_THETA, SEED_MCMC = 20201, 20202, 20203

CONFIGS = {
    "cfg2": dict(A=5, T=200, K=5, surveys=((130, 140), (165, 175)), binomial=False, serorev=True, chains=4, rungs=20),
    "cfg3": dict(A=10, T=300, K=5, surveys=((120, 130), (190, 200), (270, 280)), binomial=True, serorev=False, chains=64, rungs=50),
    "cfg4": dict(A=17, T=400, K=5, surveys=((150, 160), (250, 260), (370, 380)), binomial=True, serorev=False, chains=4096, rungs=64),
}
CONFIGS["cfg5"] = dict(CONFIGS["cfg3"], chains=0, rungs=0)
# not a BASELINE config: a long series (T > 448 -> the 16-tile kernels) with six knots and seroreversion, for the parity tests
CONFIGS["long"] = dict(A=4, T=500, K=6, surveys=((200, 210), (330, 340), (470, 480)), binomial=True, serorev=True, chains=2, rungs=3)


def _natural_spline(x, y):
    """Natural cubic spline coefficients (b, c, d) through (x, y)."""
    n = len(x) - 1
    h = np.diff(x)
    al = np.zeros(n + 1)
    for i in range(1, n):
        al[i] = 3 / h[i] * (y[i + 1] - y[i]) - 3 / h[i - 1] * (y[i] - y[i - 1])
    l, mu, z = np.ones(n + 1), np.zeros(n + 1), np.zeros(n + 1)
    for i in range(1, n):
        l[i] = 2 * (x[i + 1] - x[i - 1]) - h[i - 1] * mu[i - 1]
        mu[i] = h[i] / l[i]
        z[i] = (al[i] - h[i - 1] * z[i - 1]) / l[i]
    b, c, d = np.zeros(n), np.zeros(n + 1), np.zeros(n)
    for j in range(n - 1, -1, -1):
        c[j] = z[j] - mu[j] * c[j + 1]
        b[j] = (y[j + 1] - y[j]) / h[j] - h[j] * (c[j + 1] + 2 * c[j]) / 3
        d[j] = (c[j + 1] - c[j]) / (3 * h[j])
    return b, c[:-1], d


def _forward(T, M, node_x, node_y, ne_w, ma, mod, sod, rate, serorev, shape, scale):
    """Expected daily deaths by stratum [T, A] and seroconverted numbers [M, A] for a true parameter set."""
    from scipy.special import gammainc
    b, c, d = _natural_spline(node_x, node_y)
    days = np.arange(1, T + 1, dtype=float)
    j = np.clip(np.searchsorted(node_x, days, side="right") - 1, 0, len(node_x) - 2)
    dx = days - node_x[j]
    infxn = np.exp(node_y[j] + b[j] * dx + c[j] * dx ** 2 + d[j] * dx ** 3)
    pg = gammainc(1 / sod ** 2, np.arange(T + 1) / (mod * sod ** 2))
    auc = np.convolve(infxn, np.diff(pg))[:T]
    deaths = auc[:, None] * (ne_w * ma)[None, :]
    dd = np.arange(M, dtype=float)
    if serorev:
        con = (1 / rate) * np.exp(-(dd + 1) / rate)
        surv = np.exp(-np.power(dd / scale, shape))
        haz = np.convolve(con, surv)[:M]
    else:
        haz = 1 - np.exp(-(dd + 1) / rate)
    sero = np.convolve(infxn[:M], haz)[:M]
    return infxn, deaths, sero[:, None] * ne_w[None, :]


def make_spec(cfg: str = "cfg3", serorev=None, binomial=None, reparam: bool = False, seed: int = SEED_DATA):
    """HotPathSpec + truth dict for one of the BASELINE configs."""
    c = dict(CONFIGS[cfg])
    if serorev is not None:
        c["serorev"] = bool(serorev)
    if binomial is not None:
        c["binomial"] = bool(binomial)
    A, T, K = c["A"], c["T"], c["K"]
    starts = np.array([s for s, _ in c["surveys"]], dtype=np.int32)
    ends = np.array([e for _, e in c["surveys"]], dtype=np.int32)
    S, M = len(starts), int(ends.max())
    rng = np.random.default_rng(seed)

    # demography: decreasing age pyramid of 5e6 people; IFR log-linear 1e-4 .. 0.1
    w = np.linspace(1.6, 0.4, A)
    demog = np.round(5e6 * w / w.sum())
    ma = np.exp(np.linspace(np.log(1e-4), np.log(0.1), A))
    ne = np.ones(A)
    ne_w = ne * demog / np.sum(ne * demog)
    # true curve: natural cubic spline through K knots fitted to log(5e3 * sigmoid(linspace(-5, 7, T)))
    win = np.linspace(2, T, K)  # K-1 disjoint increasing knot windows covering (2, T)
    node_x = np.concatenate([[1.0], 0.5 * (win[:-2] + win[1:-1]), [T - 5.5]])  # window centres, last knot near the end
    grid = np.linspace(-5, 7, T)
    target = np.log(5e3 / (1 + np.exp(-grid)))
    node_y = np.interp(node_x, np.arange(1, T + 1), target)
    mod, sod, sens, spec_, rate, shape, scale = 19.8, 0.85, 0.85, 0.95, 18.3, 3.0, 140.0
    infxn, deaths, sero = _forward(T, M, node_x, node_y, ne_w, ma, mod, sod, rate, c["serorev"], shape, scale)

    obs_deaths = rng.poisson(deaths.sum(axis=1)).astype(float)
    prop = deaths.sum(axis=0)
    prop = prop / prop.sum()
    sero_n = np.round(demog * 1e-3)
    a_vec, b_vec = [], []
    for s0, e0 in zip(starts, ends):
        true_prev = sero[s0 - 1:e0].mean(axis=0) / demog
        p_obs = sens * true_prev + (1 - spec_) * (1 - true_prev)
        pos = rng.binomial(sero_n.astype(np.int64), p_obs).astype(float)
        if c["binomial"]:
            a_vec.append(pos)
            b_vec.append(sero_n)
        else:
            p = np.clip(pos / sero_n, 1e-6, 1 - 1e-6)
            z = 1.96
            den = 1 + z * z / sero_n
            ctr = (p + z * z / (2 * sero_n)) / den
            half = z * np.sqrt(p * (1 - p) / sero_n + z * z / (4 * sero_n ** 2)) / den
            lo, hi = np.clip(ctr - half, 1e-9, 1 - 1e-9), np.clip(ctr + half, 1e-9, 1 - 1e-9)
            logit = lambda v: np.log(v / (1 - v))
            a_vec.append(logit(p))
            b_vec.append((logit(hi) - logit(lo)) / (2 * 1.96))
    sero_a, sero_b = np.concatenate(a_vec), np.concatenate(b_vec)

    # df_params: (name, min, init, max, dsc1, dsc2) in the row order of the reference's deploy scripts
    rows = []
    ifr = ["ma%d" % (i + 1) for i in range(A)]
    infx = ["y%d" % (i + 1) for i in range(K)]
    knots = ["x%d" % (i + 1) for i in range(K - 1)]
    noise = ["ne%d" % (i + 1) for i in range(A)]
    if not reparam:
        rows += [(n, 0.0, min(0.2, 2 * ma[i] + 0.01), 0.4, 0.0, 0.4) for i, n in enumerate(ifr)]
        rows += [(n, 0.0, float(np.clip(node_y[i], 0.5, 9.5)), 10.0, 0.0, 10.0) for i, n in enumerate(infx)]
        rows += [(n, float(win[i]) + (0.5 if i else 0.0), float(node_x[i + 1]), float(win[i + 1]) - 0.5,
                  float(win[i]) + (0.5 if i else 0.0), float(win[i + 1]) - 0.5) for i, n in enumerate(knots)]
    else:
        # scalars in (0, 1) times the reference parameter (last of each group), R/Agecpp2strings.R:277-355
        rows += [(n, 0.0, 0.2, 1.0, 0.0, 1.0) for n in ifr[:-1]] + [(ifr[-1], 0.0, 0.2, 0.4, 0.0, 0.4)]
        rows += [(n, 0.0, float(np.clip(node_y[i] / node_y[-1], 0.05, 0.95)), 1.0, 0.0, 1.0) for i, n in enumerate(infx[:-1])]
        rows += [(infx[-1], 0.0, float(np.clip(node_y[-1], 0.5, 9.5)), 10.0, 0.0, 10.0)]
        rows += [(n, 0.0, float(node_x[i + 1] / node_x[-1]), 1.0, 0.0, 1.0) for i, n in enumerate(knots[:-1])]
        rows += [(knots[-1], float(win[-2]) + 0.5, float(node_x[-1]), float(T), float(win[-2]) + 0.5, float(T))]
    rows += [("sens", 0.5, 0.85, 1.0, 850.5, 150.5), ("spec", 0.5, 0.95, 1.0, 950.5, 50.5)]
    rows += [(n, 0.5, 1.0, 1.5, 1.0, 0.05) for n in noise]
    rows += [("mod", 18.0, 19.0, 20.0, 19.8, 0.1), ("sod", 0.0, 0.9, 1.0, 2550.0, 450.0), ("sero_con_rate", 16.0, 18.0, 21.0, 18.3, 0.1)]
    if c["serorev"]:
        rows += [("sero_rev_shape", 1.0, 3.0, 5.0, 3.0, 1.0), ("sero_rev_scale", 135.0, 140.0, 145.0, 140.0, 1.0)]
    names = [r[0] for r in rows]
    arr = np.array([r[1:] for r in rows], dtype=np.float64)
    spec = HotPathSpec(
        names=names, pmin=arr[:, 0], pinit=arr[:, 1], pmax=arr[:, 2], dsc1=arr[:, 3], dsc2=arr[:, 4],
        obs_deaths=obs_deaths, prop_strata_obs_deaths=prop, sero_a=sero_a, sero_b=sero_b, binomial_likelihood=c["binomial"],
        demog=demog, n_knots=K, rcensor_day=INT_MAX, days_obsd=T, sero_survey_start=starts, sero_survey_end=ends,
        account_serorev=c["serorev"], IFRparams=ifr, Knotparams=knots, Infxnparams=infx, Noiseparams=noise,
        reparamIFR=reparam, reparamKnots=reparam, reparamInfxn=reparam,
        maxMa=ifr[-1] if reparam else None, relKnot=knots[-1] if reparam else None, relInfxn=infx[-1] if reparam else None)
    truth = dict(node_x=node_x, node_y=node_y, ma=ma, ne=ne, mod=mod, sod=sod, sens=sens, spec=spec_, sero_con_rate=rate,
                 sero_rev_shape=shape, sero_rev_scale=scale, infxn=infxn, chains=c["chains"], rungs=c["rungs"], cfg=cfg)
    return spec, truth


def true_theta(spec: HotPathSpec, truth: dict) -> np.ndarray:
    """The generating parameter vector in df_params order (non-reparameterised specs only)."""
    assert not (spec.reparamIFR or spec.reparamKnots or spec.reparamInfxn)
    th = np.array(spec.pinit, dtype=np.float64)
    for i, n in enumerate(spec.IFRparams):
        th[spec.index(n)] = truth["ma"][i]
    for i, n in enumerate(spec.Infxnparams):
        th[spec.index(n)] = truth["node_y"][i]
    for i, n in enumerate(spec.Knotparams):
        th[spec.index(n)] = truth["node_x"][i + 1]
    for i, n in enumerate(spec.Noiseparams):
        th[spec.index(n)] = truth["ne"][i]
    for n in ("mod", "sod", "sens", "spec", "sero_con_rate"):
        th[spec.index(n)] = truth[n]
    if spec.account_serorev:
        th[spec.index("sero_rev_shape")] = truth["sero_rev_shape"]
        th[spec.index("sero_rev_scale")] = truth["sero_rev_scale"]
    return th


def _popn_ok(spec: HotPathSpec, th: np.ndarray) -> np.ndarray:
    """True where the knots are ordered and the implied infections fit in every stratum's population (vectorised
    restatement of the two validity checks, used only to condition the synthetic draws)."""
    rm = spec.role_map()
    n, K, T = th.shape[0], spec.K, spec.T

    def role(rows, reparam, rel):
        v = th[:, rows].copy()
        if reparam:
            for q, r in enumerate(rows):
                if r != rel:
                    v[:, q] *= th[:, rel]
        return v
    x = np.concatenate([np.ones((n, 1)), role(rm["idx_knots"], spec.reparamKnots, rm["idx_rel_knot"])], axis=1)
    y = role(rm["idx_infxn"], spec.reparamInfxn, rm["idx_rel_infxn"])
    ordered = np.all(np.diff(x, axis=1) > 0, axis=1)
    x = np.where(ordered[:, None], x, np.arange(1, K + 1)[None, :].astype(float))
    h = np.diff(x, axis=1)
    al = np.zeros((n, K))
    for i in range(1, K - 1):
        al[:, i] = 3 / h[:, i] * (y[:, i + 1] - y[:, i]) - 3 / h[:, i - 1] * (y[:, i] - y[:, i - 1])
    l, mu, z = np.ones((n, K)), np.zeros((n, K)), np.zeros((n, K))
    for i in range(1, K - 1):
        l[:, i] = 2 * (x[:, i + 1] - x[:, i - 1]) - h[:, i - 1] * mu[:, i - 1]
        mu[:, i] = h[:, i] / l[:, i]
        z[:, i] = (al[:, i] - h[:, i - 1] * z[:, i - 1]) / l[:, i]
    b, c, dd = np.zeros((n, K - 1)), np.zeros((n, K)), np.zeros((n, K - 1))
    for j in range(K - 2, -1, -1):
        c[:, j] = z[:, j] - mu[:, j] * c[:, j + 1]
        b[:, j] = (y[:, j + 1] - y[:, j]) / h[:, j] - h[:, j] * (c[:, j + 1] + 2 * c[:, j]) / 3
        dd[:, j] = (c[:, j + 1] - c[:, j]) / (3 * h[:, j])
    days = np.arange(1, T + 1, dtype=float)[None, :]
    j = np.zeros((n, T), dtype=np.int64)
    for m in range(1, K - 1):
        j += (days >= x[:, m:m + 1])
    take = lambda a: np.take_along_axis(a, j, axis=1)
    dx = days - take(x[:, :-1])
    with np.errstate(over="ignore"):
        tot = np.exp(take(y[:, :-1]) + take(b) * dx + take(c[:, :-1]) * dx ** 2 + take(dd) * dx ** 3).sum(axis=1)
    if spec.A > 1:
        ne = th[:, rm["idx_noise"]] * spec.demog[None, :]
        ne = ne / ne.sum(axis=1, keepdims=True)
    else:
        ne = np.ones((n, 1))
    fits = np.all(ne * tot[:, None] <= 0.98 * np.floor(spec.demog)[None, :], axis=1)
    return ordered & fits


def random_theta(spec: HotPathSpec, n: int, seed: int = SEED_THETA, bad_frac: float = 0.01, near_truth=None, jitter: float = 0.0,
                 condition: bool = True):
    """[n, d] parameter vectors.

    Uniform in the (min, max) boxes, conditioned (by rejection, `condition`) on the two validity checks of the
    likelihood - ordered knots and infections that fit in the population - i.e. uniform over the region where the
    reference does a full evaluation; then a `bad_frac` slice gets deliberately disordered knots or a huge last
    infection knot so that both early-exit branches (binomial.cpp:88-96, 170-184) are exercised.  With `near_truth`
    (a vector) the draws are that vector times (1 + jitter * N(0,1)), clipped into the box - the regime an MCMC
    actually visits.
    """
    rng = np.random.default_rng(seed)
    lo, hi = spec.pmin, spec.pmax
    if near_truth is None:
        th = lo + (hi - lo) * rng.random((n, spec.d))
        pending = np.arange(n) if condition else np.arange(0)
        for _ in range(500):
            if not len(pending):
                break
            bad = np.concatenate([~_popn_ok(spec, th[pending[c0:c0 + 65536]]) for c0 in range(0, len(pending), 65536)])
            pending = pending[bad]
            th[pending] = lo + (hi - lo) * rng.random((len(pending), spec.d))
    else:
        th = np.asarray(near_truth)[None, :] * (1 + jitter * rng.standard_normal((n, spec.d)))
        eps = 1e-9 * (hi - lo)
        th = np.clip(th, lo + eps, hi - eps)
    nb = int(round(bad_frac * n))
    if nb:
        rows = rng.choice(n, size=nb, replace=False)
        kx = [spec.index(k) for k in spec.Knotparams]
        ky = spec.index(spec.Infxnparams[-1])
        for q, r in enumerate(rows):
            if q % 2 == 0 and len(kx) >= 2:
                th[r, kx[0]], th[r, kx[1]] = th[r, kx[1]], th[r, kx[0]]  # knot disorder
            else:
                th[r, ky] = 25.0  # exp(25) infections per day -> exceeds every stratum's population
    return th
