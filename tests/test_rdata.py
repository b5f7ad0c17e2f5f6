"""R serialisation writer / reader (covidcurve_b200/rdata.py): pinned to a stream written by R 4.0.3,
and the fit object checked field for field against the schema of the reference's stored fit."""
import bz2
import json
import os

import numpy as np
import pandas as pd

from covidcurve_b200 import api, rdata

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def test_writer_reproduces_an_R_written_stream_byte_for_byte():
    raw = bz2.decompress(open(os.path.join(GOLD, "r403_stream.rda"), "rb").read())
    parsed = rdata.RDX(raw=raw)
    assert parsed.kind == "rda" and parsed.hdr == (2, 0x00040003, 0x00020300)
    objs = dict(parsed.top["pairlist"])
    assert list(objs) == ["covidcurve_simdata"]
    assert rdata.dumps_rda(objs) == raw                       # data frames, factor, chr / int / double columns, symbol refs
    sd = rdata.r_to_python(objs["covidcurve_simdata"])
    assert list(sd) == ["StrataAgg_TimeSeries_Death", "Agg_TimeSeries_Death", "StrataAgg_Seroprev"]
    assert sd["Agg_TimeSeries_Death"].shape == (200, 2) and sd["StrataAgg_TimeSeries_Death"]["Strata"].iloc[1] == "ma2"


def _fake_fit(spec, chains=3, rungs=1, burnin=5, samples=7, seed=0):
    rng = np.random.default_rng(seed)
    rows = burnin + samples
    draws = {"theta": rng.uniform(spec.pmin, spec.pmax, (chains, rungs, rows, spec.d)), "loglike": -rng.uniform(700, 900, (chains, rungs, rows)),
             "logprior": -rng.uniform(20, 50, (chains, rungs, rows)), "iteration": np.arange(1, rows + 1)}
    diag = {"beta": np.linspace(0, 1, rungs) if rungs > 1 else np.ones(1), "mc_accept": rng.uniform(0, 1, (chains, max(rungs - 1, 0))),
            "move_accept": rng.uniform(0, 1, (chains, rungs, spec.d)), "bw": np.ones((chains, rungs, spec.d))}
    from covidcurve_b200.summaries import rhat_per_parameter
    mcmcout = api.assemble_drjacoby_output(spec, draws, diag, burnin, samples, rungs, chains, True, 1.0, None, rhat_fn=rhat_per_parameter)
    import test_api_host
    inputs = dict(IFRmodel=test_api_host._model_from(spec), reparamIFR=False, reparamInfxn=False, reparamKnots=False,
                  account_seroreversion=False, binomial_likelihood=True, burnin=burnin, samples=samples, chains=chains)
    if rungs > 1:
        inputs.update(rungs=rungs, GTI_pow=1.0, beta_manual=None, coupling_on=True)
    return {"class": "IFRmodel_inf", "inputs": inputs, "mcmcout": mcmcout}


def _same_schema(ours, ref, path=""):
    """every field of the reference object exists in ours with the same SEXP type / class, in the same order"""
    if ref["t"] == "env":
        return                                               # the R6 model: written as a list of fields (see rdata.py)
    assert ours is not None, path
    assert ours["t"] == ref["t"], (path, ours["t"], ref["t"])
    cls = ours.get("attr", {}).get("class", {}).get("v")
    assert (cls or None) == ref.get("class"), (path, cls, ref.get("class"))
    if ref.get("compact_row_names"):
        rn = ours["attr"]["row.names"]["v"]
        assert rn[0] == rdata.NA_INTEGER and rn[1] == -len(ours["v"][0]["v"]), path
    if ref["t"] == 19:
        nm = ours["attr"]["names"]["v"]
        assert nm[:len(ref["names"])] == ref["names"], (path, nm, ref["names"])
        for n, a, b in zip(ref["names"], ours["v"], ref["items"]):
            _same_schema(a, b, path + "$" + n)
    elif isinstance(ref["len"], int):
        assert len(ours["v"]) == ref["len"], (path, len(ours["v"]), ref["len"])


def test_fit_object_matches_the_stored_fit_field_for_field(modfit, tmp_path):
    spec, _ = modfit
    fit = _fake_fit(spec)
    robj = rdata.fit_to_r(fit)
    sch = json.load(open(os.path.join(GOLD, "modfit_schema.json")))
    assert robj["attr"]["class"]["v"] == sch["class"] and robj["attr"]["names"]["v"] == sch["names"]
    _same_schema(robj["v"][0], sch["inputs"], "inputs")
    _same_schema(robj["v"][1], sch["mcmcout"], "mcmcout")
    # through a file and back: values survive bit for bit, labels and integer columns keep their R types
    path = str(tmp_path / "fit.rds")
    rdata.write_fit_rds(path, fit)
    assert open(path, "rb").read(2) == b"\x1f\x8b"
    back = rdata.r_to_python(rdata.read_rds(path))
    out0, out1 = fit["mcmcout"]["output"], back["mcmcout"]["output"]
    assert list(out1.columns) == list(out0.columns) and len(out1) == len(out0)
    for c in out0.columns:
        assert np.array_equal(out1[c].to_numpy(), out0[c].to_numpy()), c
    assert out1["iteration"].dtype == np.int64 and out1["chain"].iloc[0] == "chain1"
    assert back["mcmcout"]["diagnostics"]["mc_accept"] is None            # logical NA at rungs = 1
    assert back["mcmcout"]["parameters"]["rungs"] == 1 and back["inputs"]["burnin"] == 5.0
    mf = back["inputs"]["IFRmodel"]
    assert mf["IFRparams"].tolist() == spec.IFRparams and mf["maxMa"] == "ma3" and mf["rcensor_day"] == 2147483647
    assert np.array_equal(mf["paramdf"]["init"].to_numpy(), spec.pinit) and mf["data"]["obs_deaths"].shape == (spec.T, 2)
    np.testing.assert_array_equal(back["mcmcout"]["parameters"]["data"]["obs_deaths"], spec.obs_deaths.astype(float))


def test_tempered_fit_and_special_values(modfit, tmp_path):
    spec, _ = modfit
    fit = _fake_fit(spec, chains=2, rungs=3)
    fit["mcmcout"]["output"].loc[0, "loglikelihood"] = -np.finfo(float).max / 100     # the in-band failure value
    fit["mcmcout"]["output"].loc[1, "loglikelihood"] = np.nan
    path = str(tmp_path / "fit.rds")
    rdata.write_fit_rds(path, fit, compress=False)
    assert open(path, "rb").read(2) == b"X\n"
    back = rdata.r_to_python(rdata.read_rds(path))
    ll = back["mcmcout"]["output"]["loglikelihood"].to_numpy()
    assert ll[0] == -np.finfo(float).max / 100 and np.isnan(ll[1])
    acc = back["mcmcout"]["diagnostics"]["mc_accept"]
    assert list(acc.columns) == ["link", "chain", "value"] and len(acc) == 2 * 2
    assert back["inputs"]["rungs"] == 3.0 and back["inputs"]["coupling_on"] is True
    assert back["mcmcout"]["diagnostics"]["move_accept"].shape == (2, 3, spec.d)    # dim attribute, column-major
    np.testing.assert_array_equal(back["mcmcout"]["diagnostics"]["move_accept"], fit["mcmcout"]["diagnostics"]["move_accept"])
    # strings outside ASCII and NA_character_ carry the right CHARSXP flags
    s = rdata.dumps_rds(rdata.r_str(["a", None, "é"]))
    v = rdata.RDX(raw=s).top["v"]
    assert v == ["a", None, "é"] and b"\x00\x04\x00\x09\x00\x00\x00\x01a" in s and b"\x00\x00\x00\x09\xff\xff\xff\xff" in s
    assert b"\x00\x00\x80\x09\x00\x00\x00\x02\xc3\xa9" in s
