"""Host-side mirror of the reference's R interface for the hot path (R is absent from this image, so the host side
above the C ABI is Python; the R shim that does the same from R is in covidcurve_b200/r/).

  make_IFRmodel_age          R/model.R:11-355     the R6 model object (same public fields and setters)
  run_IFRmodel_age           R/main.R:12-274      same arguments, same checks, same returned structure
                                                  (`IFRmodel_inf` = {inputs, mcmcout}; mcmcout = `drjacoby_output` =
                                                  {output, diagnostics, parameters})

Only lines 179-193 of R/main.R (the call to drjacoby::run_mcmc with two generated C++ strings) are re-pointed: the
generated strings become a flat model descriptor (covidcurve_b200/spec.py) and the run happens on the GPU.
data.frames are pandas DataFrames; R's NULL is None; R's NA is NaN.
"""
from __future__ import annotations

import warnings
from typing import Optional, Sequence

import numpy as np
import pandas as pd

from .spec import INT_MAX, HotPathSpec

_SERO_NAMES = ("sens", "spec", "sero_con_rate", "sero_rev_shape", "sero_rev_scale")


def _assert(cond, msg):
    if not cond:
        raise ValueError(msg)


def _assert_strings(val, what):
    _assert(isinstance(val, (list, tuple)) and all(isinstance(v, str) for v in val), "%s must be character strings" % what)
    _assert(len(set(val)) == len(val), "%s must be unique" % what)


def _assert_in(x, y, what):
    _assert(all(v in y for v in x), "%s: all values must be in %s" % (what, list(y)))


def _assert_data(val, IFRparams):
    _assert(isinstance(val, dict), "data must be a list (dict)")
    _assert_in(val.keys(), ("obs_deaths", "prop_deaths", "obs_serology"), "names(data)")
    od, pdth, sero = val["obs_deaths"], val["prop_deaths"], val["obs_serology"]
    for df, cols, nm in ((od, ("ObsDay", "Deaths"), "obs_deaths"), (pdth, ("Strata", "PropDeaths"), "prop_deaths")):
        _assert(isinstance(df, pd.DataFrame), "%s must be a data frame" % nm)
        _assert_in(df.columns, cols, "colnames(%s)" % nm)
    _assert(np.all(np.diff(od["ObsDay"].to_numpy(dtype=float)) > 0), "obs_deaths$ObsDay must be increasing")
    p = pdth["PropDeaths"].to_numpy(dtype=float)
    _assert(np.all((p >= 0) & (p <= 1)), "PropDeaths must be in [0, 1]")
    _assert(abs(p.sum() - 1) < 1.5e-8, "Proportion of deaths by strata must sum to 1")
    _assert_in(pdth["Strata"], IFRparams, "prop_deaths$Strata")
    _assert(isinstance(sero, pd.DataFrame), "obs_serology must be a data frame")
    _assert_in(sero.columns, ("SeroStartSurvey", "SeroEndSurvey", "Strata", "SeroPos", "SeroN", "SeroPrev", "SeroLCI", "SeroUCI"),
               "colnames(obs_serology)")
    _assert_in(sero["Strata"], IFRparams, "obs_serology$Strata")
    for c in ("SeroStartSurvey", "SeroEndSurvey"):
        v = sero[c].to_numpy(dtype=float)
        _assert(np.all(v == np.floor(v)) and np.all(v > 0), "%s must be positive integers" % c)
    _assert(len(od) > 1, "Time-Series data has only one observed day specified")
    _assert(od["ObsDay"].min() == 1, "Time-Series Data must start on day 1")


class make_IFRmodel_age:
    """The `IFRmodel` R6 class (R/model.R:11-355): a validated holder of data, demography, parameter roles and the
    parameter table `paramdf` (columns name, min, init, max, dsc1, dsc2; row order = parameter order)."""

    classname = "IFRmodel"

    def __init__(self, data=None, maxObsDay=None, modparam=None, sodparam=None, IFRparams=None, maxMa=None, Infxnparams=None,
                 relInfxn=None, Knotparams=None, relKnot=None, Serotestparams=None, Noiseparams=None, paramdf=None, demog=None,
                 rcensor_day=None, IFRdictkey=None):
        self.data = self.maxObsDay = self.modparam = self.sodparam = self.IFRparams = self.maxMa = None
        self.Knotparams = self.relKnot = self.Infxnparams = self.relInfxn = self.Noiseparams = self.paramdf = None
        self.Serotestparams = self.demog = self.IFRdictkey = None
        self.rcensor_day = INT_MAX  # .Machine$integer.max = no censoring (model.R:180-184)
        if IFRparams is not None:
            self.set_IFRparams(list(IFRparams))
        if maxMa is not None:
            # model.R:154 assigns to a non-existent field (self$maxMA), so in the reference the constructor argument is
            # silently lost and callers use set_maxMa(); here the argument simply works.
            self.set_maxMa(maxMa)
        if Noiseparams is not None and self.IFRparams is not None and len(self.IFRparams) > 1:
            self.set_Noiseparams(list(Noiseparams))
        if Knotparams is not None:
            self.set_Knotparams(list(Knotparams))
        if relKnot is not None:
            self.set_relKnot(relKnot)
        if Infxnparams is not None:
            self.set_Infxnparams(list(Infxnparams))
        if relInfxn is not None:
            self.set_relInfxn(relInfxn)
        if modparam is not None:
            self.set_MeanTODparam(modparam)
        if sodparam is not None:
            self.set_CoefVarOnsetTODparam(sodparam)
        if Serotestparams is not None:
            self.set_Serotestparams(list(Serotestparams))
        if data is not None:
            self.set_data(data)
        if demog is not None:
            self.set_demog(demog)
        if paramdf is not None:
            self.set_paramdf(paramdf)
        if rcensor_day is not None:
            self.set_rcensor_day(rcensor_day)
        if IFRdictkey is not None:
            self.set_IFRdictkey(IFRdictkey)

    @classmethod
    def new(cls, *a, **k):  # R6: make_IFRmodel_age$new()
        return cls(*a, **k)

    def set_IFRparams(self, val):
        _assert_strings(val, "IFRparams")
        self.IFRparams = list(val)

    def set_maxMa(self, val):
        _assert(self.IFRparams, "Must specify IFR parameters before specifying the maximum mortality strata (maxMa)")
        _assert(isinstance(val, str) and val in self.IFRparams, "maxMa must be one of IFRparams")
        self.maxMa = val

    def set_Noiseparams(self, val):
        _assert_strings(val, "Noiseparams")
        _assert(self.IFRparams is not None, "Must specificy IFR parmaeters before Noise Effect parameters")
        _assert(len(self.IFRparams) > 1, "Cannot specify Noise Effect parameters when only one strata is being considered")
        _assert(len(val) == len(self.IFRparams), "Noiseparams and IFRparams must have the same length")
        self.Noiseparams = list(val)

    def set_Knotparams(self, val):
        _assert_strings(val, "Knotparams")
        self.Knotparams = list(val)

    def set_relKnot(self, val):
        _assert(self.Knotparams, "Must specify Knotparams before specifying the relative knot point (relKnot)")
        _assert(isinstance(val, str) and val in self.Knotparams, "relKnot must be one of Knotparams")
        self.relKnot = val

    def set_Infxnparams(self, val):
        _assert_strings(val, "Infxnparams")
        self.Infxnparams = list(val)

    def set_relInfxn(self, val):
        _assert(self.Infxnparams, "Must specify Infxnparams before specifying the relative infection point (relInfxn)")
        _assert(isinstance(val, str) and val in self.Infxnparams, "relInfxn must be one of Infxnparams")
        self.relInfxn = val

    def set_MeanTODparam(self, val):
        _assert(isinstance(val, str), "modparam must be a string")
        self.modparam = val

    def set_CoefVarOnsetTODparam(self, val):
        _assert(isinstance(val, str), "sodparam must be a string")
        self.sodparam = val

    def set_Serotestparams(self, val):
        _assert_strings(val, "Serotestparams")
        _assert_in(val, _SERO_NAMES, "Serology test parameters")
        self.Serotestparams = list(val)

    def set_data(self, val):
        _assert(self.IFRparams is not None, "Must specify IFR parameters before specifying data")
        _assert_data(val, self.IFRparams)
        self.data = val
        self.maxObsDay = int(val["obs_deaths"]["ObsDay"].max())

    def set_paramdf(self, val):
        _assert(self.IFRparams and self.Knotparams and self.Infxnparams and self.Serotestparams and self.modparam and self.sodparam,
                "Must specify modparam, sodparam, IFRparams, Knotparams, Infxnparams, Serotestparams, and Noiseparams before "
                "specifying the param dataframe")
        _assert(isinstance(val, pd.DataFrame), "paramdf must be a data frame")
        _assert_in(val.columns, ("name", "init", "min", "max", "dsc1", "dsc2"), "colnames(paramdf)")
        _assert(val["name"].is_unique, "paramdf$name must be unique")
        _assert_in(val["name"], [self.modparam, self.sodparam] + self.IFRparams + self.Knotparams + self.Infxnparams +
                   self.Serotestparams + (self.Noiseparams or []), "paramdf$name")
        self.paramdf = val.reset_index(drop=True)

    def set_rcensor_day(self, val):
        _assert(float(val) == int(val) and int(val) > 0, "rcensor_day must be a positive integer")
        self.rcensor_day = int(val)

    def set_demog(self, val):
        _assert(self.IFRparams is not None, "Must specificy IFR parameters prior to specifying demography data")
        _assert(isinstance(val, pd.DataFrame), "demog must be a data frame")
        _assert_in(val.columns, ("Strata", "popN"), "colnames(demog)")
        _assert_in(val["Strata"], self.IFRparams, "demog$Strata")
        v = val["popN"].to_numpy(dtype=float)
        _assert(np.all(v == np.floor(v)) and np.all(v > 0), "popN must be positive integers")
        self.demog = val.reset_index(drop=True)

    def set_IFRdictkey(self, val):
        _assert(isinstance(val, pd.DataFrame), "IFRdictkey must be a data frame")
        self.IFRdictkey = val


def _logit(p):
    return np.log(p / (1 - p))


def build_spec(IFRmodel: make_IFRmodel_age, binomial_likelihood=True, reparamIFR=True, reparamInfxn=True, reparamKnots=True):
    """data_list + misc_list + df_params + role/prior tables (R/main.R:99-172, R/Agecpp2strings.R) as a HotPathSpec."""
    m = IFRmodel
    A = len(m.IFRparams)
    sero = m.data["obs_serology"]
    have = [n in m.Serotestparams for n in ("sero_rev_shape", "sero_rev_scale")]
    if sum(have) == 1:
        raise ValueError("Must specify both the Weibull Shape and Scale paramter for seroreversion to be considered. "
                         "Specifying only one parameter is not an option")
    account_serorev = all(have)
    deaths = m.data["obs_deaths"]["Deaths"].to_numpy(dtype=float).copy()
    deaths[np.isnan(deaths)] = -1  # main.R:132
    if binomial_likelihood:
        sa, sb = sero["SeroPos"].to_numpy(dtype=float), sero["SeroN"].to_numpy(dtype=float)
    else:
        xmin = np.finfo(np.float64).tiny  # .Machine$double.xmin
        sb = (_logit(sero["SeroUCI"].to_numpy(dtype=float) + xmin) - _logit(sero["SeroLCI"].to_numpy(dtype=float) + xmin)) / (2 * 1.96)
        sa = _logit(sero["SeroPrev"].to_numpy(dtype=float) + xmin)
    starts = pd.unique(sero["SeroStartSurvey"]).astype(np.int32)
    ends = pd.unique(sero["SeroEndSurvey"]).astype(np.int32)
    pdf = m.paramdf
    return HotPathSpec(
        names=list(pdf["name"]), pmin=pdf["min"].to_numpy(dtype=float), pinit=pdf["init"].to_numpy(dtype=float),
        pmax=pdf["max"].to_numpy(dtype=float), dsc1=pdf["dsc1"].to_numpy(dtype=float), dsc2=pdf["dsc2"].to_numpy(dtype=float),
        obs_deaths=deaths, prop_strata_obs_deaths=m.data["prop_deaths"]["PropDeaths"].to_numpy(dtype=float), sero_a=sa, sero_b=sb,
        binomial_likelihood=bool(binomial_likelihood), demog=m.demog["popN"].to_numpy(dtype=float), n_knots=len(m.Knotparams) + 1,
        rcensor_day=m.rcensor_day, days_obsd=int(m.maxObsDay), sero_survey_start=starts, sero_survey_end=ends,
        account_serorev=account_serorev, IFRparams=m.IFRparams, Knotparams=m.Knotparams, Infxnparams=m.Infxnparams,
        Noiseparams=m.Noiseparams if A > 1 else [], modparam=m.modparam, sodparam=m.sodparam,
        reparamIFR=bool(reparamIFR), reparamKnots=bool(reparamKnots), reparamInfxn=bool(reparamInfxn),
        maxMa=m.maxMa if reparamIFR else None, relKnot=m.relKnot if reparamKnots else None, relInfxn=m.relInfxn if reparamInfxn else None,
        max_seroday_obsd=int(sero["SeroEndSurvey"].max()))


def _checks(IFRmodel, binomial_likelihood, reparamIFR, reparamInfxn, reparamKnots):
    """The assertions of R/main.R:29-112."""
    m = IFRmodel
    _assert(isinstance(m, make_IFRmodel_age), "IFRmodel must inherit from class 'IFRmodel'")
    for f in ("data", "IFRparams", "Infxnparams", "Knotparams", "paramdf", "Serotestparams", "modparam", "sodparam", "maxObsDay", "demog"):
        _assert(getattr(m, f) is not None, "IFRmodel$%s cannot be null" % f)
    A = len(m.IFRparams)
    ds, ps, ss = (list(map(str, m.demog["Strata"])), list(map(str, m.data["prop_deaths"]["Strata"][:A])),
                  list(map(str, m.data["obs_serology"]["Strata"][:A])))
    _assert(ds == ps, "Strata within the demography data-frame must be in the same order as the strata in the observed deaths data frame")
    _assert(ss == ps, "Strata within the observed serology data-frame must be in the same order as the strata in the observed deaths data frame")
    sero = m.data["obs_serology"]
    if binomial_likelihood:
        if sero[["SeroPos", "SeroN"]].isna().any().any():
            raise ValueError("Cannot have missing values of SeroPos or SeroN when considering the binomial likelihood")
    elif sero[["SeroPrev", "SeroUCI", "SeroLCI"]].isna().any().any():
        raise ValueError("Cannot have missing values of SeroPrev, SeroUCI, or SeroLCI when considering the logit likelihood")
    if m.data["prop_deaths"]["PropDeaths"].isna().any():
        raise ValueError("Cannot have missing proportions of cumulative deaths")
    if m.data["obs_deaths"]["Deaths"].isna().any():
        warnings.warn("Missing daily deaths -- will skip over in likelihood")
    if A == 1:
        # the reference means to stop here too (main.R:79-81) but tests a non-existent field, so the guard never fires
        raise ValueError("One strata currently not supported. Must specify more than strate to be estimated.")
    if reparamIFR:
        _assert(m.maxMa is not None, "If performing reparameterization, must set a maximum Ma in the R6 class object")
    if reparamInfxn:
        _assert(m.relInfxn is not None, "If performing reparameterization, must set a relative infection point in the R6 class object")
    if reparamKnots:
        _assert(m.relKnot is not None, "If performing reparameterization, must set a relative knot point in the R6 class object")
    if A != 1 and (m.Noiseparams is None or len(m.Noiseparams) != A):
        raise ValueError("You must specificy the same number of IFR params and Noise params to estimate. Having a mix of no noise "
                         "params and noise params among IFR params is not currently supported")


def run_IFRmodel_age(IFRmodel, binomial_likelihood=True, reparamIFR=True, reparamInfxn=True, reparamKnots=True, thinning=0,
                     burnin=1000, samples=10000, rungs=1, chains=5, coupling_on=True, GTI_pow=1.0, beta_manual=None, cluster=None,
                     pb_markdown=False, silent=False, *, seed: int = 1, device: int = 0, record_all_rungs: bool = True,
                     cold_rung_last: bool = False, chain_offset: int = 0):
    """Run the age-stratified IFR model (R/main.R:12).  Same arguments as the reference; `cluster`, `pb_markdown` and
    `silent` are accepted and ignored (chains run concurrently on the GPU, there is no progress bar).  Keyword-only
    extras: RNG `seed`, CUDA `device`, `record_all_rungs` (drjacoby records every rung), `cold_rung_last` (default False:
    "rung1" is the cold chain, the label every consumer in R/summarize_plot.R:68,164,276,591,687 reads; `rung_details` states
    the thermodynamic power of every label), `chain_offset` (global id of the first chain when a run is sharded over several
    processes / GPUs).

    Returns an `IFRmodel_inf`-shaped dict: {"inputs": {...}, "mcmcout": {"output", "diagnostics", "parameters"}}.
    """
    from . import _lib
    from .summaries import rhat_per_parameter

    _checks(IFRmodel, binomial_likelihood, reparamIFR, reparamInfxn, reparamKnots)
    burnin, samples, rungs, chains = int(burnin), int(samples), int(rungs), int(chains)
    spec = build_spec(IFRmodel, binomial_likelihood, reparamIFR, reparamInfxn, reparamKnots)
    model = _lib.Model(spec, device=device)
    mc = _lib.Mcmc(model, burnin=burnin, samples=samples, rungs=rungs, chains=chains, chain_offset=chain_offset,
                   coupling_on=bool(coupling_on), GTI_pow=float(GTI_pow), beta_manual=beta_manual, seed=seed,
                   cold_rung_last=cold_rung_last, record_rungs=record_all_rungs)
    mc.advance(burnin + samples - 1)
    draws = mc.fetch()
    diag = mc.diagnostics()
    mc.close()
    model.close()
    mcmcout = assemble_drjacoby_output(spec, draws, diag, burnin, samples, rungs, chains, coupling_on, GTI_pow, beta_manual,
                                       record_all_rungs=record_all_rungs, cold_rung_last=cold_rung_last, chain_offset=chain_offset,
                                       rhat_fn=rhat_per_parameter)
    out = mcmcout["output"]
    if thinning and thinning > 0:  # main.R:196-202
        keep = np.arange(thinning, burnin + samples + 1, thinning)
        out = out[out["iteration"].isin(keep)].reset_index(drop=True)
    for flag, role, rel in ((reparamIFR, IFRmodel.IFRparams, IFRmodel.maxMa), (reparamKnots, IFRmodel.Knotparams, IFRmodel.relKnot),
                            (reparamInfxn, IFRmodel.Infxnparams, IFRmodel.relInfxn)):
        if flag:  # main.R:208-242: scalar columns times their reference column
            scalars = [n for n in role if n != rel]
            out[scalars] = out[scalars].mul(out[rel], axis=0)
    mcmcout["output"] = out
    inputs = dict(IFRmodel=IFRmodel, reparamIFR=reparamIFR, reparamInfxn=reparamInfxn, reparamKnots=reparamKnots,
                  account_seroreversion=spec.account_serorev, binomial_likelihood=binomial_likelihood, burnin=burnin, samples=samples,
                  chains=chains)
    if rungs > 1:
        inputs.update(rungs=rungs, GTI_pow=GTI_pow, beta_manual=beta_manual, coupling_on=coupling_on)
    return {"class": "IFRmodel_inf", "inputs": inputs, "mcmcout": mcmcout}


def save_fit_rds(IFRmodel_inf, path, compress=True):
    """saveRDS(IFRmodel_inf, path): an R-readable fit object (covidcurve_b200/rdata.py; R side: r/ccb200.R ccb200_read_fit)."""
    from . import rdata
    rdata.write_fit_rds(path, IFRmodel_inf, compress=compress)


def assemble_drjacoby_output(spec, draws, diag, burnin, samples, rungs, chains, coupling_on, GTI_pow, beta_manual, record_all_rungs=True,
                             cold_rung_last=True, chain_offset=0, rhat_fn=None):
    """`drjacoby_output` = list(output, diagnostics, parameters) (field list: SURVEY.md 8c).

    output: one row per (chain, rung, iteration) with columns chain ("chain1"..), rung ("rung1"..), iteration (1-based,
    restarting per chain), stage ("burnin"/"sampling"), logprior, loglikelihood, then the parameters in df_params order.
    """
    theta, ll, lp, it = draws["theta"], draws["loglike"], draws["logprior"], draws["iteration"]
    C_, Rr, rows, d = theta.shape
    # the cold chain is wherever the ladder says it is (beta_manual overrides the orientation flag on the C side)
    cold_pos = int(np.argmax(diag["beta"])) if len(np.atleast_1d(diag["beta"])) == rungs else ((rungs - 1) if cold_rung_last else 0)
    rung_ids = np.arange(1, rungs + 1) if record_all_rungs else np.array([cold_pos + 1])
    chain_col = np.repeat(np.array(["chain%d" % (chain_offset + c + 1) for c in range(C_)]), Rr * rows)
    rung_col = np.tile(np.repeat(np.array(["rung%d" % r for r in rung_ids]), rows), C_)
    iter_col = np.tile(it.astype(np.int64), C_ * Rr)
    stage_col = np.where(iter_col <= burnin, "burnin", "sampling")
    cols = {"chain": chain_col, "rung": rung_col, "iteration": iter_col, "stage": stage_col, "logprior": lp.reshape(-1),
            "loglikelihood": ll.reshape(-1)}
    flat = theta.reshape(-1, d)
    for j, n in enumerate(spec.names):
        cols[n] = flat[:, j]
    output = pd.DataFrame(cols)
    rhat = None
    if rhat_fn is not None and chains > 1 and samples > 1:
        q = cold_pos if record_all_rungs else 0
        samp = it > burnin
        rhat = rhat_fn(theta[:, q][:, samp, :])
        rhat = pd.Series(rhat, index=spec.names)
    rung_details = pd.DataFrame({"rung": np.arange(1, rungs + 1), "thermodynamic_power": diag["beta"]})
    mc_accept = np.nan if rungs == 1 else pd.DataFrame(
        {"link": np.tile(np.arange(1, rungs), C_), "chain": np.repeat(np.arange(1, C_ + 1) + chain_offset, rungs - 1),
         "value": diag["mc_accept"].reshape(-1)})
    tt = np.where(np.isfinite(spec.pmin), np.where(np.isfinite(spec.pmax), 3, 2), np.where(np.isfinite(spec.pmax), 1, 0))
    df_params = pd.DataFrame({"name": spec.names, "min": spec.pmin, "init": spec.pinit, "max": spec.pmax, "trans_type": tt})
    data = {"obs_deaths": spec.obs_deaths, "prop_strata_obs_deaths": spec.prop_strata_obs_deaths}
    data.update({"obs_serologypos": spec.sero_a, "obs_serologyn": spec.sero_b} if spec.binomial_likelihood
                else {"obs_serologymu": spec.sero_a, "obs_serologyse": spec.sero_b})
    parameters = dict(data=data, df_params=df_params, loglike="covidcurve_b200::cc_loglike_batch", logprior="covidcurve_b200::cc_logprior_batch",
                      burnin=burnin, samples=samples, rungs=rungs, chains=chains, coupling_on=coupling_on, GTI_pow=GTI_pow,
                      beta_manual=beta_manual)
    diagnostics = dict(rhat=rhat, rung_details=rung_details, mc_accept=mc_accept, move_accept=diag["move_accept"], bw=diag["bw"])
    return {"class": "drjacoby_output", "output": output, "diagnostics": diagnostics, "parameters": parameters}
