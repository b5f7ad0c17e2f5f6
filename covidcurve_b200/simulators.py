"""Host mirror of the reference's forward simulator (R/simulators.R:114-263 `Agesim_infxn_2_death`, :1-92 `sim_seroprev`)
above `cc_simulate_infxn_2_death`: same arguments, same three aggregated tables, but the (day, stratum) counts are drawn
directly on the device as the independent Poisson variables they are, instead of through a line list with one row per
infection (csrc/cc_sim.cuh explains the equivalence).  A cfg4-size epidemic (millions of infections) costs the same as a
small one.  Differences, all deliberate:
  * `return_linelist = TRUE` is not available (there is no line list);
  * with `simulate_seroreversion = TRUE` the conversion and reversion day of a person are coupled, so the independent
    Poisson cells are the (conversion day, reversion day) pairs (csrc/cc_sim.cuh);
  * random numbers come from the device's counter-based generator (`seed`), not from R's stream.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def Agesim_infxn_2_death(fatalitydata, infections, m_od=14.26, s_od=0.79, curr_day=None, spec=None, sens=None, demog=None,
                         sero_delay_rate=None, simulate_seroreversion=False, sero_rev_shape=None, sero_rev_scale=None, smplfrac=1,
                         return_linelist=False, seed=1, device=0):
    """Returns {"StrataAgg_TimeSeries_Death": (ObsDay, Strata, Deaths), "Agg_TimeSeries_Death": (ObsDay, Deaths),
    "StrataAgg_Seroprev": (Strata, ObsDay, testedN, TrueSeroCount, TruePrev, ObsPrev)} like simulators.R:243-258."""
    from . import _lib
    # the reference's own assertions (simulators.R:124-147)
    if list(fatalitydata.columns) != ["Strata", "IFR", "Rho"]:
        raise ValueError("fatalitydata must have the columns Strata, IFR, Rho")
    if list(demog.columns) != ["Strata", "popN"]:
        raise ValueError("demog must have the columns Strata, popN")
    if list(demog["Strata"]) != list(fatalitydata["Strata"]):
        raise ValueError("demog$Strata must equal fatalitydata$Strata -- check that your strata are in the same order")
    infections = np.asarray(infections, dtype=np.float64)
    if curr_day is None or int(curr_day) != curr_day or len(infections) != int(curr_day):
        raise ValueError("infections must have one entry per day up to curr_day")
    for name, v in (("spec", spec), ("sens", sens)):
        if v is None or not (0 <= v <= 1):
            raise ValueError("%s must be in [0, 1]" % name)
    if not (0 < smplfrac <= 1):
        raise ValueError("smplfrac must be in (0, 1]")
    if simulate_seroreversion and (sero_rev_shape is None or sero_rev_scale is None or not (sero_rev_shape > 0 and sero_rev_scale > 0)):
        raise ValueError("sero_rev_shape and sero_rev_scale must be positive numbers when simulate_seroreversion is TRUE")
    if return_linelist:
        raise NotImplementedError("the aggregate sampler has no line list")
    T, strata = int(curr_day), list(fatalitydata["Strata"])
    A = len(strata)
    sim = _lib.simulate_infxn_2_death(infections, fatalitydata["IFR"].to_numpy(), fatalitydata["Rho"].to_numpy(),
                                      demog["popN"].to_numpy(), m_od, s_od, sero_delay_rate, smplfrac, seed=seed, n_rep=1, device=device,
                                      sero_rev_shape=sero_rev_shape if simulate_seroreversion else None,
                                      sero_rev_scale=sero_rev_scale if simulate_seroreversion else None)
    deaths, conv = sim["deaths"][0], sim["serocon"][0]                       # [T][A]
    days = np.arange(1, T + 1)
    death_strata = pd.DataFrame({"ObsDay": np.repeat(days, A).astype(float), "Strata": pd.Categorical(strata * T, categories=strata),
                                 "Deaths": deaths.reshape(-1).astype(np.int64)})
    death_agg = pd.DataFrame({"ObsDay": days.astype(float), "Deaths": deaths.sum(axis=1).astype(np.int64)})
    tested = demog["popN"].to_numpy(dtype=np.float64) * smplfrac             # simulators.R:57-58
    cum = np.cumsum(conv, axis=0)                                            # TrueSeroCount = cumulative seroconversions
    true_prev = cum / tested[None, :]
    still = cum - (np.cumsum(sim["serorev"][0], axis=0) if simulate_seroreversion else 0.0)   # RevertSeroCount, simulators.R:80
    obs_prev = still / tested[None, :]
    obs_prev = sens * obs_prev + (1 - spec) * (1 - obs_prev)                  # simulators.R:82-84
    # the reference groups by Strata, then day
    sero = pd.DataFrame({"Strata": np.repeat(np.array(strata, dtype=object), T), "ObsDay": np.tile(days, A).astype(float),
                         "testedN": np.repeat(tested, T), "TrueSeroCount": cum.T.reshape(-1).astype(np.int64),
                         "TruePrev": true_prev.T.reshape(-1), "ObsPrev": obs_prev.T.reshape(-1)})
    return {"StrataAgg_TimeSeries_Death": death_strata, "Agg_TimeSeries_Death": death_agg, "StrataAgg_Seroprev": sero}
