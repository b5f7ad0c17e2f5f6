"""ctypes wrapper of the CPU oracle (oracle/libcc_oracle.so) - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcc_oracle.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_spp = C.POINTER(C.c_char_p)


class _Desc(C.Structure):
    _fields_ = [
        ("d", C.c_int), ("names", c_spp), ("pmin", c_dp), ("pmax", c_dp), ("dsc1", c_dp), ("dsc2", c_dp),
        ("A", C.c_int), ("demog", c_dp), ("n_knots", C.c_int), ("rcensor_day", C.c_int), ("days_obsd", C.c_int),
        ("n_sero_obs", C.c_int), ("sero_survey_start", c_ip), ("sero_survey_end", c_ip), ("max_seroday_obsd", C.c_int),
        ("account_serorev", C.c_int),
        ("obs_deaths", c_dp), ("prop_strata_obs_deaths", c_dp), ("sero_a", c_dp), ("sero_b", c_dp),
        ("binomial_likelihood", C.c_int),
        ("ifr_names", c_spp), ("knot_names", c_spp), ("infxn_names", c_spp), ("noise_names", c_spp),
        ("mod_name", C.c_char_p), ("sod_name", C.c_char_p),
        ("reparam_ifr", C.c_int), ("reparam_knots", C.c_int), ("reparam_infxn", C.c_int),
        ("max_ma", C.c_char_p), ("rel_knot", C.c_char_p), ("rel_infxn", C.c_char_p),
    ]


class _Trace(C.Structure):
    _fields_ = [("L1", C.c_double), ("L2", C.c_double), ("L3", C.c_double), ("infxn", c_dp), ("auc", c_dp),
                ("ne", c_dp), ("hazard", c_dp), ("sero_num", c_dp), ("status", C.c_int)]


class _McmcOpts(C.Structure):
    _fields_ = [("burnin", C.c_int), ("samples", C.c_int), ("rungs", C.c_int), ("chains", C.c_int),
                ("coupling_on", C.c_int), ("GTI_pow", C.c_double), ("beta_manual", c_dp),
                ("seed", C.c_ulonglong), ("nthreads", C.c_int), ("cold_rung_last", C.c_int),
                ("init_per_chain", C.c_int), ("chain_offset", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the oracle with its committed Makefile (g++ -O2); returns the .so path."""
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["make", "-C", HERE, "libcc_oracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dbl = C.c_double
        for name, nargs in (("cco_stirlerr", 1), ("cco_bd0", 2), ("cco_lgamma1p", 1), ("cco_log1pmx", 1)):
            getattr(L, name).restype = dbl
            getattr(L, name).argtypes = [dbl] * nargs
        for name, nd in (("cco_dpois_raw", 2), ("cco_dpois", 2), ("cco_dbinom", 3), ("cco_dnorm", 3), ("cco_dunif", 3),
                         ("cco_dbeta", 3), ("cco_dgamma", 3)):
            getattr(L, name).restype = dbl
            getattr(L, name).argtypes = [dbl] * nd + [C.c_int]
        L.cco_dbinom_raw.restype = dbl
        L.cco_dbinom_raw.argtypes = [dbl] * 4 + [C.c_int]
        for name in ("cco_pgamma", "cco_pweibull"):
            getattr(L, name).restype = dbl
            getattr(L, name).argtypes = [dbl] * 3 + [C.c_int, C.c_int]
        L.cco_pnorm.restype = dbl
        L.cco_pnorm.argtypes = [dbl, C.c_int, C.c_int]
        L.cco_loglike.restype = dbl
        L.cco_loglike.argtypes = [C.POINTER(_Desc), c_dp, C.c_int, C.POINTER(_Trace)]
        L.cco_logprior.restype = dbl
        L.cco_logprior.argtypes = [C.POINTER(_Desc), c_dp, C.c_int]
        for name in ("cco_loglike_many", "cco_logprior_many"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.POINTER(_Desc), c_dp, C.c_long, c_dp, C.c_int]
        L.cco_mcmc.restype = C.c_int
        L.cco_mcmc.argtypes = [C.POINTER(_Desc), c_dp, C.POINTER(_McmcOpts), c_dp, c_dp, c_dp, c_dp, c_dp]
        L.cco_post_deaths.restype = None
        L.cco_post_deaths.argtypes = [c_dp, c_dp, C.c_int, C.c_int, dbl, dbl, c_dp]
        L.cco_philox_kat.restype = C.c_int
        L.cco_philox_kat.argtypes = [C.POINTER(C.c_uint)] * 3
        L.cco_draw.restype = None
        L.cco_draw.argtypes = [C.c_ulonglong] + [C.c_uint] * 5 + [c_dp, c_dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _strs(lst):
    arr = (C.c_char_p * max(1, len(lst)))(*[s.encode() for s in lst])
    return arr


class Oracle:
    """The oracle bound to one HotPathSpec (covidcurve_b200.spec.HotPathSpec or any duck-typed equivalent)."""

    def __init__(self, spec):
        self.spec = spec
        self.L = lib()
        s = spec
        self._keep = keep = {}
        keep["names"] = _strs(s.names)
        keep["ifr"] = _strs(s.IFRparams)
        keep["knot"] = _strs(s.Knotparams)
        keep["infxn"] = _strs(s.Infxnparams)
        keep["noise"] = _strs(s.Noiseparams if s.Noiseparams else [])
        keep["start"] = np.ascontiguousarray(s.sero_survey_start, dtype=np.int32)
        keep["end"] = np.ascontiguousarray(s.sero_survey_end, dtype=np.int32)
        opt = lambda v: v.encode() if v else None
        self.desc = _Desc(
            d=s.d, names=C.cast(keep["names"], c_spp), pmin=_dp(s.pmin), pmax=_dp(s.pmax), dsc1=_dp(s.dsc1), dsc2=_dp(s.dsc2),
            A=s.A, demog=_dp(s.demog), n_knots=s.K, rcensor_day=int(s.rcensor_day), days_obsd=s.T, n_sero_obs=s.S,
            sero_survey_start=keep["start"].ctypes.data_as(c_ip), sero_survey_end=keep["end"].ctypes.data_as(c_ip),
            max_seroday_obsd=s.M, account_serorev=int(s.account_serorev),
            obs_deaths=_dp(s.obs_deaths), prop_strata_obs_deaths=_dp(s.prop_strata_obs_deaths), sero_a=_dp(s.sero_a),
            sero_b=_dp(s.sero_b), binomial_likelihood=int(s.binomial_likelihood),
            ifr_names=C.cast(keep["ifr"], c_spp), knot_names=C.cast(keep["knot"], c_spp),
            infxn_names=C.cast(keep["infxn"], c_spp), noise_names=C.cast(keep["noise"], c_spp),
            mod_name=s.modparam.encode(), sod_name=s.sodparam.encode(),
            reparam_ifr=int(s.reparamIFR), reparam_knots=int(s.reparamKnots), reparam_infxn=int(s.reparamInfxn),
            max_ma=opt(s.maxMa), rel_knot=opt(s.relKnot), rel_infxn=opt(s.relInfxn))

    def _rows(self, theta):
        theta = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
        assert theta.shape[1] == self.spec.d, "theta must be [n, d] in df_params order"
        return theta

    def loglike(self, theta, nthreads: int = 1):
        th = self._rows(theta)
        out = np.empty(th.shape[0])
        self.L.cco_loglike_many(C.byref(self.desc), _dp(th), th.shape[0], _dp(out), nthreads)
        return out

    def logprior(self, theta, nthreads: int = 1):
        th = self._rows(theta)
        out = np.empty(th.shape[0])
        self.L.cco_logprior_many(C.byref(self.desc), _dp(th), th.shape[0], _dp(out), nthreads)
        return out

    def trace(self, theta):
        """One evaluation with all intermediates (dict)."""
        s = self.spec
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1))
        bufs = dict(infxn=np.zeros(s.T), auc=np.zeros(s.T), ne=np.zeros(s.A), hazard=np.zeros(s.M), sero_num=np.zeros(s.S * s.A))
        tr = _Trace(infxn=_dp(bufs["infxn"]), auc=_dp(bufs["auc"]), ne=_dp(bufs["ne"]), hazard=_dp(bufs["hazard"]),
                    sero_num=_dp(bufs["sero_num"]))
        ll = self.L.cco_loglike(C.byref(self.desc), _dp(th), 1, C.byref(tr))
        bufs.update(loglike=ll, L1=tr.L1, L2=tr.L2, L3=tr.L3, status=tr.status)
        return bufs

    def post_deaths(self, theta):
        """Expected deaths [T][A] of one posterior draw (R/summarize_plot.R:591-680): per-stratum infection curves ne[a] * infxn[t]
        from the oracle's own evaluation, then the reference's get_post_deaths loop.  `theta` holds stored (lifted) values."""
        s = self.spec
        assert not (s.reparamIFR or s.reparamKnots or s.reparamInfxn), "posterior draws are stored un-reparameterised"
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1))
        tr = self.trace(th)
        strata = np.ascontiguousarray((tr["ne"][:, None] * tr["infxn"][None, :]))           # [A][T]: stratum-major like unlist()
        ifr = np.ascontiguousarray([th[s.index(nm)] for nm in s.IFRparams])
        out = np.empty((s.T, s.A))
        self.L.cco_post_deaths(_dp(strata), _dp(ifr), s.T, s.A, float(th[s.index(s.modparam)]), float(th[s.index(s.sodparam)]), _dp(out))
        return out

    def mcmc(self, burnin, samples, rungs=1, chains=1, coupling_on=True, GTI_pow=1.0, beta_manual=None, seed=1,
             nthreads=1, cold_rung_last=True, init=None, chain_offset=0):
        s = self.spec
        init = np.ascontiguousarray(s.pinit if init is None else init, dtype=np.float64)
        per_chain = init.ndim == 2
        assert init.shape == ((chains, s.d) if per_chain else (s.d,))
        n_it = burnin + samples
        theta = np.zeros((chains, rungs, n_it, s.d))
        ll = np.zeros((chains, rungs, n_it))
        lp = np.zeros((chains, rungs, n_it))
        acc = np.zeros((chains, max(rungs - 1, 1)))
        beta = np.zeros(rungs)
        bm = None if beta_manual is None else np.ascontiguousarray(beta_manual, dtype=np.float64)
        o = _McmcOpts(burnin=burnin, samples=samples, rungs=rungs, chains=chains, coupling_on=int(coupling_on),
                      GTI_pow=GTI_pow, beta_manual=_dp(bm) if bm is not None else None, seed=seed, nthreads=nthreads,
                      cold_rung_last=int(cold_rung_last), init_per_chain=int(per_chain), chain_offset=chain_offset)
        rc = self.L.cco_mcmc(C.byref(self.desc), _dp(init), C.byref(o), _dp(theta), _dp(ll), _dp(lp), _dp(acc), _dp(beta))
        if rc != 0:
            raise RuntimeError("oracle mcmc failed rc=%d" % rc)
        return dict(theta=theta, loglike=ll, logprior=lp, mc_accept=acc, beta=beta)
