"""HotPathSpec: the flattened (data, misc, df_params, role map, prior table) bundle.

This is what `R/Agecpp2strings.R` + `R/main.R:132-172` produce in the reference - there as two generated C++
source strings plus `data_list`/`misc_list`/`df_params`; here as plain arrays that the C-ABI model descriptor
(`include/covidcurve_b200.h: cc_model_desc`) takes positionally.

Reference mapping (all under /root/reference):
  * parameter order            = rows of `df_params <- paramdf[, 1:4]`            R/main.R:172
  * data_list                  = obs_deaths / prop_strata_obs_deaths / sero pair   R/main.R:136-152
  * misc_list                  = rcensor_day ... account_serorev                   R/main.R:158-166
  * role map (name -> array)   = the generated prologue                            R/Agecpp2strings.R:255-368
  * prior table                = the generated logprior                            R/Agecpp2strings.R:66-219
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

PRIOR_NONE, PRIOR_DUNIF, PRIOR_DNORM, PRIOR_DBETA = 0, 1, 2, 3
INT_MAX = 2147483647


@dataclass
class HotPathSpec:
    # df_params
    names: List[str]
    pmin: np.ndarray
    pinit: np.ndarray
    pmax: np.ndarray
    dsc1: np.ndarray
    dsc2: np.ndarray
    # data_list
    obs_deaths: np.ndarray               # [T], -1 = missing
    prop_strata_obs_deaths: np.ndarray   # [A]
    sero_a: np.ndarray                   # [S*A] obs_serologypos | obs_serologymu  (survey-major)
    sero_b: np.ndarray                   # [S*A] obs_serologyn   | obs_serologyse
    binomial_likelihood: bool
    # misc_list
    demog: np.ndarray                    # [A]
    n_knots: int
    rcensor_day: int
    days_obsd: int
    sero_survey_start: np.ndarray        # [S]
    sero_survey_end: np.ndarray          # [S]
    account_serorev: bool
    # role lists (model order)
    IFRparams: List[str]
    Knotparams: List[str]
    Infxnparams: List[str]
    Noiseparams: List[str]
    modparam: str = "mod"
    sodparam: str = "sod"
    reparamIFR: bool = False
    reparamKnots: bool = False
    reparamInfxn: bool = False
    maxMa: Optional[str] = None
    relKnot: Optional[str] = None
    relInfxn: Optional[str] = None
    # derived
    max_seroday_obsd: int = field(default=0)

    def __post_init__(self):
        f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        self.names = [str(n) for n in self.names]
        for k in ("pmin", "pinit", "pmax", "dsc1", "dsc2", "obs_deaths", "prop_strata_obs_deaths", "sero_a", "sero_b", "demog"):
            setattr(self, k, f64(getattr(self, k)))
        self.sero_survey_start = np.ascontiguousarray(np.asarray(self.sero_survey_start, dtype=np.int32))
        self.sero_survey_end = np.ascontiguousarray(np.asarray(self.sero_survey_end, dtype=np.int32))
        for k in ("IFRparams", "Knotparams", "Infxnparams", "Noiseparams"):
            setattr(self, k, [str(s) for s in getattr(self, k)])
        if not self.max_seroday_obsd:
            self.max_seroday_obsd = int(self.sero_survey_end.max())
        self.validate()

    # ------------------------------------------------------------------ shapes
    @property
    def d(self) -> int:
        return len(self.names)

    @property
    def A(self) -> int:
        return len(self.demog)

    @property
    def K(self) -> int:
        return int(self.n_knots)

    @property
    def T(self) -> int:
        return int(self.days_obsd)

    @property
    def S(self) -> int:
        return len(self.sero_survey_start)

    @property
    def M(self) -> int:
        return int(self.max_seroday_obsd)

    def index(self, name: str) -> int:
        try:
            return self.names.index(name)
        except ValueError:
            raise KeyError("parameter %r is not a row of df_params" % name) from None

    def validate(self):
        d, A, K, T, S, M = self.d, self.A, self.K, self.T, self.S, self.M
        if len(set(self.names)) != d:
            raise ValueError("parameter names must be unique")
        for k in ("pmin", "pinit", "pmax", "dsc1", "dsc2"):
            if getattr(self, k).shape != (d,):
                raise ValueError("%s must have one entry per parameter" % k)
        if self.obs_deaths.shape != (T,):
            raise ValueError("obs_deaths must have days_obsd entries")
        if self.prop_strata_obs_deaths.shape != (A,):
            raise ValueError("prop_strata_obs_deaths must have one entry per stratum")
        if self.sero_a.shape != (S * A,) or self.sero_b.shape != (S * A,):
            raise ValueError("serology vectors must have n_sero_obs * strata entries (survey-major)")
        if len(self.IFRparams) != A or (A > 1 and len(self.Noiseparams) != A):
            raise ValueError("IFRparams / Noiseparams must match the number of strata")
        if len(self.Knotparams) != K - 1 or len(self.Infxnparams) != K:
            raise ValueError("need n_knots-1 Knotparams and n_knots Infxnparams")
        if K < 3:
            raise ValueError("natural cubic spline needs at least 3 knots")
        if len(self.sero_survey_end) != S or S < 1:
            raise ValueError("serosurvey start/end must have the same length >= 1")
        if M > T:
            raise ValueError("max_seroday_obsd must not exceed days_obsd (the reference reads infxn_spline[0..M-1])")
        if np.any(self.sero_survey_start < 1) or np.any(self.sero_survey_end > M) or np.any(self.sero_survey_end < self.sero_survey_start):
            raise ValueError("serosurvey windows must satisfy 1 <= start <= end <= max_seroday_obsd")
        needed = ["sens", "spec", "sero_con_rate", self.modparam, self.sodparam]
        if self.account_serorev:
            needed += ["sero_rev_shape", "sero_rev_scale"]
        needed += self.IFRparams + self.Knotparams + self.Infxnparams + (self.Noiseparams if A > 1 else [])
        for nm in needed:
            self.index(nm)
        if self.reparamIFR and self.maxMa not in self.IFRparams:
            raise ValueError("reparamIFR requires maxMa in IFRparams")
        if self.reparamKnots and self.relKnot not in self.Knotparams:
            raise ValueError("reparamKnots requires relKnot in Knotparams")
        if self.reparamInfxn and self.relInfxn not in self.Infxnparams:
            raise ValueError("reparamInfxn requires relInfxn in Infxnparams")

    # ------------------------------------------------------------- role map
    def _in_paramdf_order(self, role: Sequence[str]) -> List[str]:
        return [n for n in self.names if n in role]

    def _role_rows(self, role: Sequence[str], reparam: bool, rel: Optional[str]) -> List[int]:
        """Row supplying each position of a role vector (R/Agecpp2strings.R:277-355).

        With re-parameterisation the non-reference positions are filled, in model-list order, from the scalar
        names taken in *paramdf* order - the two orders may differ and the reference follows exactly this rule.
        """
        if not reparam:
            return [self.index(n) for n in role]
        scalars = [n for n in self._in_paramdf_order(role) if n != rel]
        rows, c = [], 0
        for n in role:
            if n == rel:
                rows.append(self.index(rel))
            else:
                rows.append(self.index(scalars[c]))
                c += 1
        return rows

    def role_map(self) -> dict:
        A = self.A
        noise = [self.index(n) for n in self._in_paramdf_order(self.Noiseparams)] if A > 1 else [-1] * A
        return dict(
            idx_sens=self.index("sens"), idx_spec=self.index("spec"), idx_sero_con_rate=self.index("sero_con_rate"),
            idx_sero_rev_shape=self.index("sero_rev_shape") if self.account_serorev else -1,
            idx_sero_rev_scale=self.index("sero_rev_scale") if self.account_serorev else -1,
            idx_mod=self.index(self.modparam), idx_sod=self.index(self.sodparam),
            idx_knots=self._role_rows(self.Knotparams, self.reparamKnots, self.relKnot),
            idx_infxn=self._role_rows(self.Infxnparams, self.reparamInfxn, self.relInfxn),
            idx_ifr=self._role_rows(self.IFRparams, self.reparamIFR, self.maxMa),
            idx_noise=noise,
            idx_rel_knot=self.index(self.relKnot) if self.reparamKnots else -1,
            idx_rel_infxn=self.index(self.relInfxn) if self.reparamInfxn else -1,
            idx_max_ma=self.index(self.maxMa) if self.reparamIFR else -1,
        )

    # ----------------------------------------------------------- prior table
    def prior_table(self):
        """(family[d], jacobian rows[3], jacobian counts[3]) of the generated logprior.

        Families: IFR/knot/infxn -> dunif(dsc1, dsc2); sero_con_rate, sero_rev_*, noise, mod -> dnorm;
        sens, spec, sod -> dbeta (R/Agecpp2strings.R:73-148).  With all reparam flags FALSE the sod Beta prior is
        NOT part of the returned sum (R/Agecpp2strings.R:202) -> family NONE.
        Parameters that appear in df_params but in no prior line (e.g. sero_rev_* rows present while
        seroreversion is off) get NONE.
        """
        fam = np.zeros(self.d, dtype=np.int32)
        for n in self.IFRparams + self.Knotparams + self.Infxnparams:
            fam[self.index(n)] = PRIOR_DUNIF
        fam[self.index("sero_con_rate")] = PRIOR_DNORM
        fam[self.index("sens")] = PRIOR_DBETA
        fam[self.index("spec")] = PRIOR_DBETA
        if self.account_serorev:
            fam[self.index("sero_rev_shape")] = PRIOR_DNORM
            fam[self.index("sero_rev_scale")] = PRIOR_DNORM
        if self.A > 1:
            for n in self.Noiseparams:
                fam[self.index(n)] = PRIOR_DNORM
        fam[self.index(self.modparam)] = PRIOR_DNORM
        fff = not (self.reparamIFR or self.reparamKnots or self.reparamInfxn)
        fam[self.index(self.sodparam)] = PRIOR_NONE if fff else PRIOR_DBETA
        jac_rows = np.array([
            self.index(self.maxMa) if self.reparamIFR else -1,
            self.index(self.relKnot) if self.reparamKnots else -1,
            self.index(self.relInfxn) if self.reparamInfxn else -1], dtype=np.int32)
        jac_counts = np.array([self.A - 1, self.K - 2, self.K - 1], dtype=np.float64)
        return fam, jac_rows, jac_counts
