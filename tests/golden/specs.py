"""Build HotPathSpec objects from the committed golden fixtures (no /root/reference access)."""
import os

import numpy as np

from covidcurve_b200.spec import HotPathSpec

HERE = os.path.dirname(os.path.abspath(__file__))


def modfit():
    """(spec, fixture) of the stored reference fit (cfg1: A=3, K=5, T=200, S=2, M=165, binomial, FFF)."""
    g = np.load(os.path.join(HERE, "modfit_golden.npz"))
    spec = HotPathSpec(
        names=list(g["param_names"]), pmin=g["pmin"], pinit=g["pinit"], pmax=g["pmax"], dsc1=g["dsc1"], dsc2=g["dsc2"],
        obs_deaths=g["obs_deaths"], prop_strata_obs_deaths=g["prop_strata_obs_deaths"],
        sero_a=g["obs_serologypos"], sero_b=g["obs_serologyn"], binomial_likelihood=True,
        demog=g["demog"], n_knots=int(g["n_knots"]), rcensor_day=int(g["rcensor_day"]), days_obsd=int(g["days_obsd"]),
        sero_survey_start=g["sero_survey_start"], sero_survey_end=g["sero_survey_end"], account_serorev=False,
        IFRparams=list(g["IFRparams"]), Knotparams=list(g["Knotparams"]), Infxnparams=list(g["Infxnparams"]),
        Noiseparams=list(g["Noiseparams"]))
    return spec, g


def simdata(binomial=True):
    """(spec, fixture) of tests/testthat/test-check_natcub_cpp.R rebuilt from covidcurve_simdata.rda."""
    g = np.load(os.path.join(HERE, "simdata_golden.npz"))
    m, _ = modfit()
    spec = HotPathSpec(
        names=list(g["param_names"]), pmin=m.pmin, pinit=m.pinit, pmax=m.pmax, dsc1=m.dsc1, dsc2=m.dsc2,
        obs_deaths=g["obs_deaths"], prop_strata_obs_deaths=g["prop_strata_obs_deaths"],
        sero_a=g["obs_serologypos"] if binomial else g["obs_serologymu"],
        sero_b=g["obs_serologyn"] if binomial else g["obs_serologyse"], binomial_likelihood=binomial,
        demog=g["demog"], n_knots=int(g["n_knots"]), rcensor_day=int(g["rcensor_day"]), days_obsd=int(g["days_obsd"]),
        sero_survey_start=g["sero_survey_start"], sero_survey_end=g["sero_survey_end"], account_serorev=False,
        IFRparams=["ma1", "ma2", "ma3"], Knotparams=["x1", "x2", "x3", "x4"], Infxnparams=["y1", "y2", "y3", "y4", "y5"],
        Noiseparams=["ne1", "ne2", "ne3"], max_seroday_obsd=int(g["max_seroday_obsd"]))
    return spec, g
