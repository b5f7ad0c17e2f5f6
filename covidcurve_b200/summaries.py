"""Posterior summaries over the gathered draws (the consumers of the fit object, R/summarize_plot.R:46-264).

  rhat_per_parameter            drjacoby's `diagnostics$rhat` (Gelman-Rubin over chains, sampling draws of the cold rung)
  get_gelman_rubin_diagnostic   R/summarize_plot.R:46-58  -> coda::gelman.diag (autoburnin = TRUE: second half of each series)
  get_cred_intervals            R/summarize_plot.R:68-153 -> min / 2.5% / median / mean / 97.5% / max / ESS / Geweke
  get_overall_IFR_cred_intervals R/summarize_plot.R:162-264 -> demography- or attack-rate-weighted overall IFR

All functions take plain arrays `draws[chain, iteration, param]`, so they work on draws fetched from one GPU or
all-gathered from many (covidcurve_b200/dist.py).  Moment-based statistics (R-hat) are also available in a
"sufficient statistics" form so that a multi-GPU run only has to all-reduce (count, mean, M2) per chain.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


# ----------------------------------------------------------------------------------------------------------- R-hat
def chain_moments(draws: np.ndarray):
    """Per-chain (n, mean, unbiased variance) of draws[chain, iteration, param] -> three arrays [chain, param]."""
    n = draws.shape[1]
    mean = draws.mean(axis=1)
    var = draws.var(axis=1, ddof=1) if n > 1 else np.zeros_like(mean)
    return np.full(mean.shape, float(n)), mean, var


def rhat_from_moments(n, mean, var):
    """Gelman-Rubin potential scale reduction factor from per-chain moments (drjacoby's `gelman_rubin`):
    W = mean_c var_c;  B = n/(C-1) sum_c (mean_c - mean)^2;  V = (1 - 1/n) W + B/n;  rhat = sqrt(V / W)."""
    C = mean.shape[0]
    n = n[0]
    W = var.mean(axis=0)
    B = n / (C - 1) * ((mean - mean.mean(axis=0, keepdims=True)) ** 2).sum(axis=0)
    V = (1 - 1 / n) * W + B / n
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(V / W)


def rhat_per_parameter(draws: np.ndarray) -> np.ndarray:
    return rhat_from_moments(*chain_moments(np.asarray(draws, dtype=np.float64)))


def gelman_diag(draws: np.ndarray, confidence: float = 0.95, autoburnin: bool = True):
    """coda::gelman.diag(transform = FALSE, multivariate = FALSE): point estimate and upper confidence limit of the
    PSRF per parameter, with coda's degrees-of-freedom correction (Brooks & Gelman 1998)."""
    from scipy import stats
    x = np.asarray(draws, dtype=np.float64)
    if autoburnin:
        n0 = x.shape[1]
        x = x[:, n0 // 2:, :]  # coda: window(x, start = end/2 + 1)
    C, n, P = x.shape
    xbar = x.mean(axis=1)
    s2 = x.var(axis=1, ddof=1)
    W = s2.mean(axis=0)
    B = n * xbar.var(axis=0, ddof=1)
    muhat = xbar.mean(axis=0)
    var_w = s2.var(axis=0, ddof=1) / C
    var_b = 2 * B ** 2 / (C - 1)
    cov_wb = np.array([(n / C) * (np.cov(s2[:, p], xbar[:, p] ** 2)[0, 1] - 2 * muhat[p] * np.cov(s2[:, p], xbar[:, p])[0, 1])
                       for p in range(P)])
    V = (n - 1) * W / n + (1 + 1 / C) * B / n
    var_V = ((n - 1) ** 2 * var_w + (1 + 1 / C) ** 2 * var_b + 2 * (n - 1) * (1 + 1 / C) * cov_wb) / n ** 2
    df_V = 2 * V ** 2 / var_V
    df_adj = (df_V + 3) / (df_V + 1)
    B_df, W_df = C - 1, 2 * W ** 2 / var_w
    R2_fixed = (n - 1) / n
    R2_random = (1 + 1 / C) * (1 / n) * (B / W)
    R2_est = R2_fixed + R2_random
    R2_upper = R2_fixed + stats.f.ppf((1 + confidence) / 2, B_df, W_df) * R2_random
    return np.sqrt(df_adj * R2_est), np.sqrt(df_adj * R2_upper)


def get_gelman_rubin_diagnostic(modout, autoburnin: bool = True) -> pd.DataFrame:
    """R/summarize_plot.R:46-58: drops rung/iteration/stage/logprior/loglikelihood, one series per chain (no stage or rung
    filter, exactly like the reference), then coda::gelman.diag."""
    out = modout["mcmcout"]["output"]
    params = [c for c in out.columns if c not in ("chain", "rung", "iteration", "stage", "logprior", "loglikelihood")]
    chains = list(pd.unique(out["chain"]))
    _assert_multi(chains)
    series = np.stack([out.loc[out["chain"] == c, params].to_numpy(dtype=float) for c in chains])
    est, upper = gelman_diag(series, autoburnin=autoburnin)
    return pd.DataFrame({"param": params, "Point est.": est, "Upper C.I.": upper})


def _assert_multi(chains):
    if len(chains) < 2:
        raise ValueError("Gelman-Rubin requires at least two chains")


# ----------------------------------------------------------------------------------------- ESS / Geweke (coda)
def _spectrum0_ar(x: np.ndarray) -> float:
    """coda::spectrum0.ar: spectral density at frequency zero from an AR(p) fit (Yule-Walker, order by AIC)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 3 or np.var(x) == 0:
        return 0.0
    xc = x - x.mean()
    order_max = int(min(n - 1, np.floor(10 * np.log10(n))))
    acov = np.array([np.dot(xc[: n - k], xc[k:]) / n for k in range(order_max + 1)])
    # Levinson-Durbin recursions, keeping the innovations variance of every order
    best_aic, best = None, (acov[0], np.zeros(0))
    phi = np.zeros(0)
    v = acov[0]
    aic0 = n * np.log(v)
    best_aic, best = aic0, (v, phi)
    for k in range(1, order_max + 1):
        if v <= 0:
            break
        kappa = (acov[k] - np.dot(phi, acov[k - 1:0:-1][: k - 1] if k > 1 else np.zeros(0))) / v
        phi = np.concatenate([phi - kappa * phi[::-1], [kappa]])
        v = v * (1 - kappa ** 2)
        if v <= 0:
            break
        aic = n * np.log(v) + 2 * k
        if aic < best_aic:
            best_aic, best = aic, (v, phi.copy())
    v, phi = best
    v = v * n / (n - (len(phi) + 1))  # R's ar.yw variance scaling
    return float(v / (1 - phi.sum()) ** 2)


def effective_size(x: np.ndarray) -> float:
    """coda::effectiveSize for one series."""
    x = np.asarray(x, dtype=np.float64)
    s0 = _spectrum0_ar(x)
    return 0.0 if s0 == 0 else float(x.size * np.var(x, ddof=1) / s0)


def geweke_z(x: np.ndarray, frac1: float = 0.1, frac2: float = 0.5) -> float:
    """coda::geweke.diag z-score for one series.  Windows as coda takes them (1-based, inclusive):
    first 1 .. ceiling(1 + frac1 (n - 1)), second floor(n - frac2 (n - 1)) .. n  (n = 10 000: 1 001 and 5 001 points)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    end1 = int(np.ceil(1 + frac1 * (n - 1)))
    start2 = int(np.floor(n - frac2 * (n - 1)))
    a, b = x[:max(end1, 1)], x[max(start2, 1) - 1:]
    va, vb = _spectrum0_ar(a) / a.size, _spectrum0_ar(b) / b.size
    with np.errstate(divide="ignore", invalid="ignore"):
        return float((a.mean() - b.mean()) / np.sqrt(va + vb))


# -------------------------------------------------------------------------------------------- credible intervals
def _sampling_rows(modout, whichrung):
    out = modout["mcmcout"]["output"]
    return out[(out["stage"] == "sampling") & (out["rung"] == whichrung)]


def default_cold_rung(modout) -> str:
    """Label of the cold chain: the rung whose thermodynamic power is 1 (the reference defaults to "rung1", which is only
    unambiguous for rungs = 1, SURVEY.md 8c open question)."""
    rd = modout["mcmcout"]["diagnostics"]["rung_details"]
    return "rung%d" % int(rd.loc[rd["thermodynamic_power"].idxmax(), "rung"])


def get_cred_intervals(IFRmodel_inf, what: str, whichrung: str | None = None, by_chain: bool = True) -> pd.DataFrame:
    """R/summarize_plot.R:68-153.  what in {"IFRparams", "Knotparams", "Infxnparams", "Serotestparams", "Noiseparams",
    "DeathDelayparams"}: per (chain?, param) min, LCI (2.5 %), median, mean, UCI (97.5 %), max, ESS, Geweke z and its density."""
    from scipy import stats
    m = IFRmodel_inf["inputs"]["IFRmodel"]
    groups = dict(IFRparams=m.IFRparams, Knotparams=m.Knotparams, Infxnparams=m.Infxnparams, Serotestparams=m.Serotestparams,
                  Noiseparams=m.Noiseparams or [], DeathDelayparams=[m.modparam, m.sodparam])   # summarize_plot.R:72,124
    groups["TODparams"] = groups["DeathDelayparams"]                                            # alias kept from round 1
    if what not in groups:
        raise ValueError("what must be one of %s" % list(groups))
    rung = whichrung or default_cold_rung(IFRmodel_inf)
    df = _sampling_rows(IFRmodel_inf, rung)
    rows = []
    keys = list(pd.unique(df["chain"])) if by_chain else [None]
    for ch in keys:
        sub = df if ch is None else df[df["chain"] == ch]
        for p in groups[what]:
            x = sub[p].to_numpy(dtype=float)
            z = geweke_z(x)
            q = np.quantile(x, [0.025, 0.5, 0.975])  # numpy's default = R's type 7
            row = dict(param=p, min=x.min(), LCI=q[0], median=q[1], mean=x.mean(), UCI=q[2], max=x.max(), ESS=effective_size(x),
                       GewekeZ=z, GewekeP=float(stats.norm.pdf(z)))
            if ch is not None:
                row = dict(chain=ch, **row)
            rows.append(row)
    return pd.DataFrame(rows)


def overall_ifr_weights(m, whichstandard: str):
    """(popN, use_noise) of R/summarize_plot.R:162-264: "pop" weights by demography alone, "arpop" by attack rate x demography."""
    if whichstandard not in ("pop", "arpop"):
        raise ValueError("Standardization must either be by populatin (pop) or by population and attack rate (arpop)")
    if whichstandard == "arpop" and not m.Noiseparams:
        raise ValueError("arpop needs the attack-rate (noise) parameters")
    return m.demog["popN"].to_numpy(dtype=float), whichstandard == "arpop"


def get_overall_IFR_cred_intervals(IFRmodel_inf, whichstandard: str = "pop", whichrung: str | None = None, by_chain: bool = True) -> pd.DataFrame:
    """R/summarize_plot.R:162-264: overall IFR per draw = sum_a ma_a * w_a with w = popN / sum(popN) (whichstandard = "pop", the
    reference's default) or the attack-rate re-weighted demography ("arpop"), then the summary columns of get_cred_intervals."""
    from scipy import stats
    m = IFRmodel_inf["inputs"]["IFRmodel"]
    pop, use_noise = overall_ifr_weights(m, whichstandard)
    rung = whichrung or default_cold_rung(IFRmodel_inf)
    df = _sampling_rows(IFRmodel_inf, rung)
    ma = df[m.IFRparams].to_numpy(dtype=float)
    ne = df[m.Noiseparams].to_numpy(dtype=float) if use_noise else np.ones_like(ma)
    w = ne * pop[None, :]
    w = w / w.sum(axis=1, keepdims=True)
    overall = (ma * w).sum(axis=1)
    rows = []
    keys = list(pd.unique(df["chain"])) if by_chain else [None]
    for ch in keys:
        x = overall if ch is None else overall[(df["chain"] == ch).to_numpy()]
        q = np.quantile(x, [0.025, 0.5, 0.975])
        z = geweke_z(x)
        row = dict(min=x.min(), LCI=q[0], median=q[1], mean=x.mean(), UCI=q[2], max=x.max(), ESS=effective_size(x),
                   GewekeZ=z, GewekeP=float(stats.norm.pdf(z)))
        if ch is not None:
            row = dict(chain=ch, **row)
        rows.append(row)
    return pd.DataFrame(rows)


# ------------------------------------------------------------------------------------- the same summaries on the device
# (covidcurve_b200/csrc/cc_summ.cu through the C ABI).  `draws` is [chains, iterations, params]: a numpy array (uploaded) or a
# CUDA torch tensor (used where it lives - the recorded draws of a run, or the NCCL all-gathered tensor of dist.allgather_draws).
_COLS = ("min", "LCI", "median", "mean", "UCI", "max", "ESS", "GewekeZ", "GewekeP")


def device_cred_intervals(draws, names, by_chain: bool = True, chain_ids=None, device: int = 0) -> pd.DataFrame:
    """get_cred_intervals' table (R/summarize_plot.R:136-152) for every parameter in `names` (the last axis of `draws`)."""
    from . import _lib
    tab = _lib.summarize_draws(draws, by_chain=by_chain, device=device)                  # [groups, params, fields]
    G, P = tab.shape[:2]
    if len(names) != P:
        raise ValueError("names must label the last axis of draws")
    cols = {}
    if by_chain:
        ids = np.arange(1, G + 1) if chain_ids is None else np.asarray(chain_ids)
        cols["chain"] = np.repeat(ids, P)
    cols["param"] = np.tile(np.asarray(names, dtype=object), G)
    for f, c in enumerate(_COLS):
        cols[c] = tab[:, :, f].reshape(-1)
    return pd.DataFrame(cols)


def device_gelman_diag(draws, confidence: float = 0.95, autoburnin: bool = True, device: int = 0):
    """coda::gelman.diag(transform = FALSE, multivariate = FALSE) from device-side moments: (point estimate, upper limit, drjacoby
    rhat) per parameter.  Only the F quantile of the upper limit (one scalar per parameter) is taken on the host."""
    from . import _lib
    n = draws.shape[1]
    first = n // 2 if autoburnin else 0                      # coda: window(x, start = end/2 + 1)
    mean, var = _lib.chain_moments(draws, first_iter=first, device=device)
    return gelman_from_chain_moments(mean, var, n - first, confidence, device)


def gelman_from_chain_moments(mean, var, n_used: int, confidence: float = 0.95, device: int = 0):
    from scipy import stats
    from . import _lib
    g = _lib.gelman_from_moments(mean, var, n_used, device=device)
    C = mean.shape[0]
    f = dict(zip(_lib.GELMAN_FIELDS, g.T))
    upper = np.sqrt(f["df_adj"] * (f["R2_fixed"] + stats.f.ppf((1 + confidence) / 2, C - 1, f["W_df"]) * f["R2_random"]))
    return f["psrf"], upper, f["rhat"]


def device_overall_IFR_cred_intervals(draws, names, IFRparams, popN, Noiseparams=None, by_chain: bool = True, device: int = 0) -> pd.DataFrame:
    """get_overall_IFR_cred_intervals on the device: the per-draw overall IFR (cc_overall_ifr_draws) summarised like a parameter.
    Noiseparams = None is the reference's default standard ("pop"), a list of names "arpop"."""
    from . import _lib
    names = list(names)
    idx_ifr = [names.index(p) for p in IFRparams]
    idx_noise = None if Noiseparams is None else [names.index(p) for p in Noiseparams]
    est = _lib.overall_ifr_draws(draws, idx_ifr, popN, idx_noise, device=device)
    est3 = est[:, :, None]
    out = device_cred_intervals(est3, ["overall_IFR"], by_chain=by_chain, device=device)
    return out.drop(columns=["param"])
