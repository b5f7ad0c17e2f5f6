"""Generate the committed golden fixtures from the reference's packaged .rda files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Outputs (committed, small):
  tests/golden/modfit_golden.npz   - the stored reference fit data/covidcurve_modfit.rda:
        the exact data_list / misc / df_params / prior table it was run with and a
        3 420-row sample of its 60 000 (parameter vector -> loglikelihood, logprior) rows,
        which were produced by the reference C++ under R 4.0.3 + drjacoby.
        Also posterior summaries of the sampling stage (for statistical MCMC checks).
  tests/golden/simdata_golden.npz  - inputs of tests/testthat/test-check_natcub_cpp.R rebuilt
        from data/covidcurve_simdata.rda, both parameter vectors.
  tests/golden/r403_stream.rda     - data/covidcurve_simdata.rda itself (13 kB, a byte stream written by R 4.0.3): the
        known-answer vector of the R serialisation writer (parse -> re-serialise must reproduce it byte for byte).
  tests/golden/modfit_schema.json  - field names / SEXP types / classes of the stored fit's `mcmcout` and `inputs`
        (no values): what the fit-object writer must reproduce field for field.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
from rdx import env_items, get, load, names  # noqa: E402

REF = "/root/reference/data"


def col(df, k):
    return df["v"][names(df).index(k)]


def factor_or_str(c):
    if c["t"] == 16:
        return list(c["v"])
    lev = c["attr"]["levels"]["v"]
    return [lev[i - 1] for i in c["v"]]


def main():
    fit = load(os.path.join(REF, "covidcurve_modfit.rda"))["covidcurve_modfit"]
    mc = get(fit, "mcmcout")
    out = get(mc, "output")
    par = get(mc, "parameters")
    model = env_items(get(get(fit, "inputs"), "IFRmodel"))

    pnames = list(col(get(par, "df_params"), "name")["v"])
    theta = np.array([col(out, p)["v"] for p in pnames], dtype=np.float64).T  # [60000, d]
    ll = np.array(col(out, "loglikelihood")["v"])
    lp = np.array(col(out, "logprior")["v"])
    chain = np.array([int(s.replace("chain", "")) for s in factor_or_str(col(out, "chain"))])
    iteration = np.array(col(out, "iteration")["v"])
    stage = np.array([0 if s == "burnin" else 1 for s in factor_or_str(col(out, "stage"))])
    n = theta.shape[0]
    assert n == 60000 and theta.shape[1] == 20

    rng = np.random.default_rng(8)
    pick = set()
    for c in (1, 2, 3):
        idx = np.nonzero(chain == c)[0]
        pick.update(idx[:100].tolist())
    order = np.argsort(ll)
    pick.update(order[:60].tolist())
    pick.update(order[-60:].tolist())
    pick.update(rng.choice(n, size=3000, replace=False).tolist())
    pick = np.array(sorted(pick))

    dfp = get(par, "df_params")
    pdf = model["paramdf"]
    assert list(col(pdf, "name")["v"]) == pnames
    data = get(par, "data")
    sero = get(model["data"], "obs_serology")
    start = sorted(set(int(v) for v in col(sero, "SeroStartSurvey")["v"]))
    end = sorted(set(int(v) for v in col(sero, "SeroEndSurvey")["v"]))

    samp = stage == 1
    q = np.quantile(theta[samp], [0.025, 0.5, 0.975], axis=0)
    # per-parameter move rate of the single-site sampler (fraction of iterations a column changed)
    moved = []
    for c in (1, 2, 3):
        th = theta[chain == c]
        moved.append((np.diff(th, axis=0) != 0).mean(axis=0))
    np.savez_compressed(
        os.path.join(HERE, "modfit_golden.npz"),
        param_names=np.array(pnames),
        pmin=np.array(col(dfp, "min")["v"]), pinit=np.array(col(dfp, "init")["v"]),
        pmax=np.array(col(dfp, "max")["v"]),
        dsc1=np.array(col(pdf, "dsc1")["v"]), dsc2=np.array(col(pdf, "dsc2")["v"]),
        obs_deaths=np.array(col(data, "obs_deaths")["v"]),
        prop_strata_obs_deaths=np.array(col(data, "prop_strata_obs_deaths")["v"]),
        obs_serologypos=np.array(col(data, "obs_serologypos")["v"]),
        obs_serologyn=np.array(col(data, "obs_serologyn")["v"]),
        demog=np.array(col(model["demog"], "popN")["v"], dtype=np.float64),
        days_obsd=np.int64(model["maxObsDay"]["v"][0]),
        rcensor_day=np.int64(model["rcensor_day"]["v"][0]),
        n_knots=np.int64(len(model["Knotparams"]["v"]) + 1),
        sero_survey_start=np.array(start), sero_survey_end=np.array(end),
        max_seroday_obsd=np.int64(max(end)),
        IFRparams=np.array(model["IFRparams"]["v"]), Knotparams=np.array(model["Knotparams"]["v"]),
        Infxnparams=np.array(model["Infxnparams"]["v"]), Noiseparams=np.array(model["Noiseparams"]["v"]),
        Serotestparams=np.array(model["Serotestparams"]["v"]),
        rows=pick, theta=theta[pick], loglike=ll[pick], logprior=lp[pick],
        row_chain=chain[pick], row_iteration=iteration[pick], row_stage=stage[pick],
        post_q025_q50_q975=q, post_mean=theta[samp].mean(axis=0),
        move_rate=np.array(moved),
        stored_rhat=np.array(get(get(mc, "diagnostics"), "rhat")["v"]),
        logprior_string=np.array(get(par, "logprior")["v"][0]),
        burnin=np.int64(get(par, "burnin")["v"][0]), samples=np.int64(get(par, "samples")["v"][0]),
    )

    sim = load(os.path.join(REF, "covidcurve_simdata.rda"))["covidcurve_simdata"]
    agg = get(sim, "Agg_TimeSeries_Death")
    strat = get(sim, "StrataAgg_TimeSeries_Death")
    sp = get(sim, "StrataAgg_Seroprev")
    deaths = np.array(col(agg, "Deaths")["v"], dtype=np.float64)
    s_strata = factor_or_str(col(strat, "Strata"))
    s_deaths = np.array(col(strat, "Deaths")["v"], dtype=np.float64)
    strata = ["ma1", "ma2", "ma3"]
    prop = np.array([s_deaths[[s == a for s in s_strata]].sum() for a in strata]) / deaths.sum()
    sp_strata = factor_or_str(col(sp, "Strata"))
    sp_day = np.array(col(sp, "ObsDay")["v"])
    sp_n = np.array(col(sp, "testedN")["v"], dtype=np.float64)
    sp_prev = np.array(col(sp, "ObsPrev")["v"], dtype=np.float64)
    pos, nn, prev = [], [], []
    for day in (135, 160):  # survey-major, then strata (test-check_natcub_cpp.R:53-65)
        for a in strata:
            k = [i for i in range(len(sp_day)) if sp_day[i] == day and sp_strata[i] == a]
            assert len(k) == 1
            k = k[0]
            # R's round() is half-to-even on the representable value; products here are not ties
            pos.append(float(np.round(sp_prev[k] * sp_n[k])))
            nn.append(sp_n[k])
            prev.append(sp_prev[k])
    prev = np.array(prev)
    # test-check_natcub_cpp.R:102-108 and :121-127, in the order used by the stored fit's df_params
    order20 = ["ma1", "ma2", "ma3", "y1", "y2", "y3", "y4", "y5", "ne1", "ne2", "ne3", "x1", "x2", "x3", "x4",
               "sens", "spec", "mod", "sod", "sero_con_rate"]
    more = dict(mod=19.8, sod=0.85, ma1=0.01, ma2=0.05, ma3=0.1, x1=30, x2=60, x3=90, x4=120,
                y1=2.8, y2=5.7, y3=7.7, y4=8.4, y5=8.5, ne1=0.33, ne2=0.33, ne3=0.33,
                sens=0.85, spec=0.95, sero_con_rate=18)
    less = dict(mod=19.8, sod=0.85, ma1=0.1, ma2=0.39, ma3=0.98, x1=22.25, x2=54.47, x3=109.9, x4=145.58,
                y1=2.98, y2=4.52, y3=6.74, y4=7.82, y5=7.88, ne1=0.1, ne2=0.4, ne3=0.5,
                sens=0.85, spec=0.99, sero_con_rate=10)
    np.savez_compressed(
        os.path.join(HERE, "simdata_golden.npz"),
        param_names=np.array(order20),
        morelikely=np.array([more[k] for k in order20], dtype=np.float64),
        lesslikely=np.array([less[k] for k in order20], dtype=np.float64),
        obs_deaths=deaths, prop_strata_obs_deaths=prop,
        obs_serologypos=np.array(pos), obs_serologyn=np.array(nn),
        obs_serologymu=np.log(prev / (1 - prev)), obs_serologyse=np.full(6, 0.05),
        demog=np.array([1500000.0, 2250000.0, 1250000.0]),
        days_obsd=np.int64(200), rcensor_day=np.int64(2147483647), n_knots=np.int64(5),
        sero_survey_start=np.array([105, 130]), sero_survey_end=np.array([115, 140]),
        max_seroday_obsd=np.int64(140),
    )
    print("rows kept:", len(pick), "row0 ll/lp:", ll[0], lp[0])
    print("simdata: total deaths", deaths.sum(), "prop", prop, "pos", pos)


def schema(o, depth=0):
    """Names, SEXP types and classes of an object tree (values dropped; data-frame columns kept, long vectors cut)."""
    if o is None:
        return {"t": "NULL"}
    if "env" in o:
        return {"t": "env"}
    a = o.get("attr", {})
    d = {"t": o["t"], "len": len(o["v"]) if o["t"] != 19 else None}
    if "class" in a:
        d["class"] = list(a["class"]["v"])
    if "row.names" in a:
        d["compact_row_names"] = a["row.names"]["v"][0] == -2147483648
    if o["t"] == 19:
        d["names"] = list(a["names"]["v"])
        d["items"] = [schema(v, depth + 1) for v in o["v"]]
        del d["len"]
    elif d["len"] > 64:
        d["len"] = "n"
    return d


def extras():
    import json
    import shutil
    shutil.copyfile(os.path.join(REF, "covidcurve_simdata.rda"), os.path.join(HERE, "r403_stream.rda"))
    fit = load(os.path.join(REF, "covidcurve_modfit.rda"))["covidcurve_modfit"]
    sch = {"class": list(fit["attr"]["class"]["v"]), "names": names(fit), "inputs": schema(get(fit, "inputs")),
           "mcmcout": schema(get(fit, "mcmcout"))}
    for node in (sch["mcmcout"]["items"][2]["items"][2], sch["mcmcout"]["items"][2]["items"][3]):
        node["len"] = 1  # the generated C++ strings: one element each
    with open(os.path.join(HERE, "modfit_schema.json"), "w") as f:
        json.dump(sch, f, indent=1)


if __name__ == "__main__":
    main()
    extras()
