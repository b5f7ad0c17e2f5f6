"""Posterior curve re-evaluation (R/summarize_plot.R:276-884), batched over the down-sampled draws.

The reference re-compiles a surgically edited copy of the generated likelihood per call and maps it over the draws one
by one (`Rcpp::cppFunction` + `purrr::map`); here the same intermediates come out of ONE batched call of
`cc_curves_batch` (covidcurve_b200/csrc: curves_kernel).  Plot objects are not produced - the functions return the
`curvedata` data frames the reference's plots are drawn from.

  draw_posterior_infxn_cubic_splines   :276-585   ne[a] * infxn[t] per draw (per stratum, or summed)
  draw_posterior_sero_curves           :687-884   seroconverted numbers, crude and test-adjusted prevalence over ALL days
  posterior_check_infxns_to_death      :591-680   expected deaths per stratum with the gamma DENSITY kernel (cc_post_deaths_batch)
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .api import build_spec
from .summaries import default_cold_rung


def convert_post_probs(logpost):
    """exp(logpost - logsumexp(logpost)) (R/summarize_plot.R:9-11)."""
    lp = np.asarray(logpost, dtype=np.float64)
    m = lp.max()
    return np.exp(lp - (m + np.log(np.sum(np.exp(lp - m)))))


def _downsample(IFRmodel_inf, whichrung, dwnsmpl, seed):
    """Rows of the sampling stage of one rung, drawn WITHOUT replacement with probability proportional to the posterior
    (`sample(1:n, size = dwnsmpl, prob = probs)`), in increasing row order."""
    out = IFRmodel_inf["mcmcout"]["output"]
    rung = whichrung or default_cold_rung(IFRmodel_inf)
    nodes = out[(out["stage"] == "sampling") & (out["rung"] == rung)].reset_index(drop=True)
    if dwnsmpl < 1 or dwnsmpl > len(nodes):
        raise ValueError("dwnsmpl must be between 1 and the number of sampling rows of the rung")
    # sequential sampling without replacement, probability proportional to the posterior, done in log space (Gumbel
    # top-k == R's ProbSampleNoReplace in distribution) so that rows whose probability underflows stay drawable
    logpost = nodes["loglikelihood"].to_numpy(dtype=np.float64) + nodes["logprior"].to_numpy(dtype=np.float64)
    keys = logpost + np.random.default_rng(seed).gumbel(size=len(nodes))
    rows = np.sort(np.argsort(-keys, kind="stable")[: int(dwnsmpl)])
    return nodes.iloc[rows].reset_index(drop=True)


def _curves(IFRmodel_inf, nodes, device, **want):
    from . import _lib
    inp = IFRmodel_inf["inputs"]
    # reparam flags FALSE: the stored posterior was already lifted back (R/main.R:208-242; summarize_plot.R:307-312)
    spec = build_spec(inp["IFRmodel"], inp["binomial_likelihood"], False, False, False)
    model = _lib.Model(spec, device=device)
    try:
        return spec, model.curves(nodes[spec.names].to_numpy(dtype=np.float64), **want)
    finally:
        model.close()


def _sim_index(nodes, by_chain):
    if by_chain:
        return nodes.groupby("chain").cumcount().to_numpy() + 1
    return np.arange(1, len(nodes) + 1)


def draw_posterior_infxn_cubic_splines(IFRmodel_inf, whichrung=None, dwnsmpl=100, by_chain=True, by_strata=False, seed=1, device=0):
    """curvedata of R/summarize_plot.R:276: one row per (draw, day) with `infxns_<stratum>` columns (= ne[a] * infxn[t]),
    `totinfxns` when not by_strata, the IFR / mod / sod columns of the draw, `sim` (per chain when by_chain) and `time`."""
    nodes = _downsample(IFRmodel_inf, whichrung, dwnsmpl, seed)
    spec, cv = _curves(IFRmodel_inf, nodes, device, infxn=True, ne=True, auc=False, sero_full=False)
    m = IFRmodel_inf["inputs"]["IFRmodel"]
    n, T, A = len(nodes), spec.T, spec.A
    strata = cv["ne"][:, None, :] * cv["infxn"][:, :, None]            # [draw, day, stratum]
    keep = m.IFRparams + [m.modparam, m.sodparam]
    base = {"sim": np.repeat(_sim_index(nodes, by_chain), T), "time": np.tile(np.arange(1, T + 1), n)}
    if by_chain:
        base = {"chain": np.repeat(nodes["chain"].to_numpy(), T), **base}
    for k in keep:
        base[k] = np.repeat(nodes[k].to_numpy(), T)
    flat = strata.reshape(n * T, A)
    for a, nm in enumerate(m.IFRparams):
        base["infxns_" + nm] = flat[:, a]
    df = pd.DataFrame(base)
    if not by_strata:
        df["totinfxns"] = flat.sum(axis=1)
    return {"curvedata": df}


def draw_posterior_sero_curves(IFRmodel_inf, whichrung=None, dwnsmpl=100, by_chain=True, seed=1, device=0):
    """R/summarize_plot.R:687: per (draw, day, stratum) the seroconverted number, the crude prevalence (number / popN) and
    the prevalence a test with the draw's sensitivity / specificity would observe, over ALL observed days (the reference
    drops the population check in this extraction, :729)."""
    nodes = _downsample(IFRmodel_inf, whichrung, dwnsmpl, seed)
    spec, cv = _curves(IFRmodel_inf, nodes, device, infxn=False, ne=True, auc=False, sero_full=True)
    m = IFRmodel_inf["inputs"]["IFRmodel"]
    n, T, A = len(nodes), spec.T, spec.A
    num = cv["ne"][:, None, :] * cv["sero_full"][:, :, None]
    crude = num / np.floor(spec.demog)[None, None, :]
    sens, spc = nodes["sens"].to_numpy()[:, None, None], nodes["spec"].to_numpy()[:, None, None]
    adj = sens * crude + (1 - spc) * (1 - crude)
    base = {"sim": np.repeat(_sim_index(nodes, by_chain), T), "ObsDay": np.tile(np.arange(1, T + 1), n)}
    if by_chain:
        base = {"chain": np.repeat(nodes["chain"].to_numpy(), T), **base}
    for a, nm in enumerate(m.IFRparams):
        base["sero_con_num_" + nm] = num[:, :, a].reshape(-1)
        base["crude_pd_seroprev_" + nm] = crude[:, :, a].reshape(-1)
        base["RG_pd_seroprev_" + nm] = adj[:, :, a].reshape(-1)
    return pd.DataFrame(base)


def posterior_check_infxns_to_death(IFRmodel_inf, whichrung=None, dwnsmpl=100, by_chain=False, seed=1, device=0):
    """R/summarize_plot.R:591: expected deaths per stratum and day for each down-sampled draw,
    deaths[j][a] = sum_{i <= j} infxns[i][a] * ifr[a] * dgamma(j - i; 1/sod^2, mod sod^2)  (the DENSITY, :605-641),
    evaluated on the device for all draws at once (cc_post_deaths_batch).  One row per (draw, day) with the draw's IFR
    columns and `deaths_<stratum>`."""
    from . import _lib
    nodes = _downsample(IFRmodel_inf, whichrung, dwnsmpl, seed)
    inp = IFRmodel_inf["inputs"]
    m = inp["IFRmodel"]
    spec = build_spec(m, inp["binomial_likelihood"], False, False, False)
    model = _lib.Model(spec, device=device)
    try:
        deaths = model.post_deaths(nodes[spec.names].to_numpy(dtype=np.float64))          # [draw][day][stratum]
    finally:
        model.close()
    n, T = len(nodes), spec.T
    cols = {"sim": np.repeat(_sim_index(nodes, by_chain), T), "time": np.tile(np.arange(1, T + 1), n)}
    if by_chain:
        cols = {"chain": np.repeat(nodes["chain"].to_numpy(), T), **cols}
    for a, nm in enumerate(m.IFRparams):
        cols[nm] = np.repeat(nodes[nm].to_numpy(), T)
    for a, nm in enumerate(m.IFRparams):
        cols["deaths_" + nm] = deaths[:, :, a].reshape(-1)
    return pd.DataFrame(cols)
