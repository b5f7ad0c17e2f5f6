"""ctypes binding of the C ABI in include/covidcurve_b200.h (libcovidcurve_b200.so, built in-tree by
`__graft_entry__.build()` / `python -m covidcurve_b200.build`).

This is the only place Python touches the native library.  There is NO CPU fallback: if the shared object is
missing, or no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcovidcurve_b200.so")
LIB_PATH = os.environ.get("CCB200_LIB", LIB_PATH)  # developer override: an alternative build of the same ABI

CC_ABI_VERSION = 1
CC_MEM_HOST, CC_MEM_DEVICE = 0, 1
CC_OVERFLO_LOGLIKE = -(np.finfo(np.float64).max / 100.0)

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class CCError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_params", C.c_int32), ("n_strata", C.c_int32), ("n_knots", C.c_int32),
        ("days_obsd", C.c_int32), ("n_sero_obs", C.c_int32), ("max_seroday_obsd", C.c_int32), ("rcensor_day", C.c_int32),
        ("account_serorev", C.c_int32), ("binomial_likelihood", C.c_int32),
        ("reparam_ifr", C.c_int32), ("reparam_knots", C.c_int32), ("reparam_infxn", C.c_int32),
        ("demog", c_dp), ("obs_deaths", c_dp), ("prop_strata_obs_deaths", c_dp), ("sero_a", c_dp), ("sero_b", c_dp),
        ("sero_survey_start", c_ip), ("sero_survey_end", c_ip),
        ("idx_sens", C.c_int32), ("idx_spec", C.c_int32), ("idx_sero_con_rate", C.c_int32),
        ("idx_sero_rev_shape", C.c_int32), ("idx_sero_rev_scale", C.c_int32), ("idx_mod", C.c_int32), ("idx_sod", C.c_int32),
        ("idx_knots", c_ip), ("idx_infxn", c_ip), ("idx_ifr", c_ip), ("idx_noise", c_ip),
        ("idx_rel_knot", C.c_int32), ("idx_rel_infxn", C.c_int32), ("idx_max_ma", C.c_int32),
        ("pmin", c_dp), ("pmax", c_dp),
        ("prior_family", c_ip), ("prior_p1", c_dp), ("prior_p2", c_dp),
        ("jac_rows", C.c_int32 * 3), ("jac_counts", C.c_double * 3),
    ]


class CallArgs(C.Structure):
    """cc_call_args: the reference's (params, param_i, data, misc) call shape as plain fields."""
    _fields_ = [
        ("n_params", C.c_int32), ("names", C.POINTER(C.c_char_p)),
        ("obs_deaths", c_dp), ("prop_strata_obs_deaths", c_dp), ("sero_a", c_dp), ("sero_b", c_dp),
        ("demog", c_dp), ("n_strata", C.c_int32), ("n_knots", C.c_int32), ("rcensor_day", C.c_int32), ("days_obsd", C.c_int32),
        ("n_sero_obs", C.c_int32), ("sero_survey_start", c_ip), ("sero_survey_end", c_ip),
        ("max_seroday_obsd", C.c_int32), ("account_serorev", C.c_int32), ("binomial_likelihood", C.c_int32),
    ]


class McmcOpts(C.Structure):
    _fields_ = [
        ("burnin", C.c_int32), ("samples", C.c_int32), ("rungs", C.c_int32), ("n_chains", C.c_int32),
        ("chain_offset", C.c_int32), ("coupling_on", C.c_int32), ("GTI_pow", C.c_double), ("beta_manual", c_dp),
        ("seed", C.c_uint64), ("cold_rung_last", C.c_int32), ("record_rungs", C.c_int32), ("thin", C.c_int32),
        ("init", c_dp), ("init_default", c_dp), ("bw_per_particle", C.c_int32), ("no_adapt_in_sampling", C.c_int32),
    ]


# every symbol include/covidcurve_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "cc_last_error": (C.c_char_p, []),
    "cc_abi_version": (C.c_int, []),
    "cc_device_count": (C.c_int, []),
    "cc_model_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(_vp)]),
    "cc_model_destroy": (None, [_vp]),
    "cc_loglike_batch": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, _vp, C.c_int, _vp]),
    "cc_logprior_batch": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, _vp, C.c_int, _vp]),
    "cc_loglike": (C.c_int, [_vp, c_dp, C.c_int, c_dp]),
    "cc_logprior": (C.c_int, [_vp, c_dp, C.c_int, c_dp]),
    "cc_loglike_call": (C.c_int, [C.POINTER(CallArgs), c_dp, C.c_int, C.c_int, c_dp]),
    "cc_curves_batch": (C.c_int, [_vp, c_dp, C.c_int64, C.c_int64, c_dp, c_dp, c_dp, c_dp]),
    "cc_post_deaths_batch": (C.c_int, [_vp, c_dp, C.c_int64, C.c_int64, c_dp]),
    "cc_summarize_draws": (C.c_int, [C.c_int, _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int, _vp, c_dp]),
    "cc_chain_moments": (C.c_int, [C.c_int, _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp, _vp, _vp]),
    "cc_gelman_from_moments": (C.c_int, [C.c_int, _vp, _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int, _vp, c_dp]),
    "cc_overall_ifr_draws": (C.c_int, [C.c_int, _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, c_ip, c_ip, c_dp, C.c_int32,
                                       C.c_int, _vp, _vp]),
    "cc_mcmc_create": (C.c_int, [_vp, C.POINTER(McmcOpts), C.POINTER(_vp)]),
    "cc_mcmc_advance": (C.c_int, [_vp, C.c_int32]),
    "cc_mcmc_last_advance_ms": (C.c_int, [_vp, c_dp]),
    "cc_mcmc_iterations": (C.c_int64, [_vp]),
    "cc_mcmc_rows": (C.c_int64, [_vp]),
    "cc_mcmc_fetch": (C.c_int, [_vp, c_dp, c_dp, c_dp, c_ip]),
    "cc_mcmc_device_draws": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "cc_mcmc_diagnostics": (C.c_int, [_vp, c_dp, c_dp, c_dp, c_dp]),
    "cc_mcmc_loglike_evals": (C.c_int64, [_vp]),
    "cc_mcmc_destroy": (None, [_vp]),
    "cc_nmath_batch": (C.c_int, [C.c_int, C.c_int, c_dp, c_dp, c_dp, C.c_int64, c_dp]),
    "cc_simulate_infxn_2_death": (C.c_int, [C.c_int, c_dp, C.c_int32, c_dp, c_dp, c_dp, C.c_int32, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "cc_fp64_peak_probe": (C.c_int, [C.c_int, C.c_int, c_dp, c_dp]),
    "cc_dmma_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]),
    "cc_last_kernel_ms": (C.c_int, [_vp, c_dp, C.POINTER(C.c_int32)]),
}

_lib = None


def load():
    """dlopen the native library and bind every exported symbol.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CCError("%s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.cc_abi_version() != CC_ABI_VERSION:
            raise CCError("ABI version mismatch between _lib.py and libcovidcurve_b200.so")
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise CCError("covidcurve_b200 error %d: %s" % (rc, load().cc_last_error().decode()))


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def make_desc(spec):
    """cc_model_desc for a HotPathSpec; returns (desc, keepalive)."""
    rm = spec.role_map()
    fam, jac_rows, jac_counts = spec.prior_table()
    keep = dict(
        idx_knots=np.ascontiguousarray(rm["idx_knots"], dtype=np.int32), idx_infxn=np.ascontiguousarray(rm["idx_infxn"], dtype=np.int32),
        idx_ifr=np.ascontiguousarray(rm["idx_ifr"], dtype=np.int32), idx_noise=np.ascontiguousarray(rm["idx_noise"], dtype=np.int32),
        fam=np.ascontiguousarray(fam, dtype=np.int32), start=spec.sero_survey_start, end=spec.sero_survey_end, spec=spec)
    de = ModelDesc(
        abi_version=CC_ABI_VERSION, n_params=spec.d, n_strata=spec.A, n_knots=spec.K, days_obsd=spec.T, n_sero_obs=spec.S,
        max_seroday_obsd=spec.M, rcensor_day=int(spec.rcensor_day), account_serorev=int(spec.account_serorev),
        binomial_likelihood=int(spec.binomial_likelihood), reparam_ifr=int(spec.reparamIFR), reparam_knots=int(spec.reparamKnots),
        reparam_infxn=int(spec.reparamInfxn),
        demog=_dp(spec.demog), obs_deaths=_dp(spec.obs_deaths), prop_strata_obs_deaths=_dp(spec.prop_strata_obs_deaths),
        sero_a=_dp(spec.sero_a), sero_b=_dp(spec.sero_b), sero_survey_start=_ip(keep["start"]), sero_survey_end=_ip(keep["end"]),
        idx_sens=rm["idx_sens"], idx_spec=rm["idx_spec"], idx_sero_con_rate=rm["idx_sero_con_rate"],
        idx_sero_rev_shape=rm["idx_sero_rev_shape"], idx_sero_rev_scale=rm["idx_sero_rev_scale"], idx_mod=rm["idx_mod"],
        idx_sod=rm["idx_sod"], idx_knots=_ip(keep["idx_knots"]), idx_infxn=_ip(keep["idx_infxn"]), idx_ifr=_ip(keep["idx_ifr"]),
        idx_noise=_ip(keep["idx_noise"]), idx_rel_knot=rm["idx_rel_knot"], idx_rel_infxn=rm["idx_rel_infxn"],
        idx_max_ma=rm["idx_max_ma"], pmin=_dp(spec.pmin), pmax=_dp(spec.pmax), prior_family=_ip(keep["fam"]),
        prior_p1=_dp(spec.dsc1), prior_p2=_dp(spec.dsc2))
    for j in range(3):
        de.jac_rows[j] = int(jac_rows[j])
        de.jac_counts[j] = float(jac_counts[j])
    return de, keep


class Model:
    """A model handle bound to one CUDA device (cc_model)."""

    def __init__(self, spec, device: int = 0):
        self.L = load()
        self.spec = spec
        self.device = device
        de, self._keep = make_desc(spec)
        h = _vp()
        check(self.L.cc_model_create(C.byref(de), device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.cc_model_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- host-buffer calls: theta is [n, d] rows (one parameter vector per row, df_params column order)
    def _soa(self, theta):
        th = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        if th.shape[1] != self.spec.d:
            raise ValueError("theta must be [n, %d] in df_params order" % self.spec.d)
        return np.ascontiguousarray(th.T)  # [d][n]

    def loglike(self, theta):
        soa = self._soa(theta)
        n = soa.shape[1]
        out = np.empty(n)
        check(self.L.cc_loglike_batch(self.h, soa.ctypes.data, n, n, out.ctypes.data, CC_MEM_HOST, None))
        return out

    def logprior(self, theta):
        soa = self._soa(theta)
        n = soa.shape[1]
        out = np.empty(n)
        check(self.L.cc_logprior_batch(self.h, soa.ctypes.data, n, n, out.ctypes.data, CC_MEM_HOST, None))
        return out

    def loglike_scalar(self, params, param_i: int = 1):
        p = np.ascontiguousarray(params, dtype=np.float64)
        out = C.c_double()
        check(self.L.cc_loglike(self.h, _dp(p), param_i, C.byref(out)))
        return out.value

    def logprior_scalar(self, params, param_i: int = 1):
        p = np.ascontiguousarray(params, dtype=np.float64)
        out = C.c_double()
        check(self.L.cc_logprior(self.h, _dp(p), param_i, C.byref(out)))
        return out.value

    # ---- raw-pointer calls (SoA [d][ld]); pointers are ints (host or device addresses)
    def loglike_ptr(self, theta_ptr: int, n: int, ld: int, out_ptr: int, mem: int, stream: int = 0):
        check(self.L.cc_loglike_batch(self.h, theta_ptr, n, ld, out_ptr, mem, stream or None))

    def logprior_ptr(self, theta_ptr: int, n: int, ld: int, out_ptr: int, mem: int, stream: int = 0):
        check(self.L.cc_logprior_batch(self.h, theta_ptr, n, ld, out_ptr, mem, stream or None))

    def last_kernel_ms(self):
        ms, nl = C.c_double(), C.c_int32()
        check(self.L.cc_last_kernel_ms(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def curves(self, theta, infxn=True, ne=True, auc=True, sero_full=True):
        soa = self._soa(theta)
        n = soa.shape[1]
        s = self.spec
        o = dict(infxn=np.empty((n, s.T)) if infxn else None, ne=np.empty((n, s.A)) if ne else None,
                 auc=np.empty((n, s.T)) if auc else None, sero_full=np.empty((n, s.T)) if sero_full else None)
        ptr = lambda a: _dp(a) if a is not None else None
        check(self.L.cc_curves_batch(self.h, _dp(soa), n, n, ptr(o["infxn"]), ptr(o["ne"]), ptr(o["auc"]), ptr(o["sero_full"])))
        return {k: v for k, v in o.items() if v is not None}


    def post_deaths(self, theta):
        """cc_post_deaths_batch: expected deaths [n][T][A] per draw with the gamma density kernel (R/summarize_plot.R:605-641)."""
        soa = self._soa(theta)
        n = soa.shape[1]
        out = np.empty((n, self.spec.T, self.spec.A))
        check(self.L.cc_post_deaths_batch(self.h, _dp(soa), n, n, _dp(out)))
        return out


# ---- posterior summaries on the device (cc_summ.cu).  `draws` is either a numpy array [chains, iterations, params] (uploaded)
# or a CUDA torch tensor of that shape (any chain / iteration stride, unit parameter stride; used in place)
SUMM_FIELDS = ("min", "LCI", "median", "mean", "UCI", "max", "ESS", "GewekeZ", "GewekeP", "var", "spec0", "n")
GELMAN_FIELDS = ("rhat", "psrf", "W", "B", "W_df", "df_adj", "R2_fixed", "R2_random")


def _draws_arg(draws):
    """(pointer, C, n, P, stride_chain, stride_iter, mem, device or None, stream, keepalive)"""
    if isinstance(draws, np.ndarray) or not hasattr(draws, "data_ptr"):
        a = np.ascontiguousarray(draws, dtype=np.float64)
        if a.ndim != 3:
            raise ValueError("draws must be [chains, iterations, params]")
        Cn, n, P = a.shape
        return a.ctypes.data, Cn, n, P, n * P, P, CC_MEM_HOST, None, None, a
    import torch
    t = draws
    if t.ndim != 3 or t.dtype != torch.float64 or not t.is_cuda:
        raise ValueError("device draws must be a CUDA float64 tensor [chains, iterations, params]")
    if t.stride(2) != 1:
        t = t.contiguous()
    stream = torch.cuda.current_stream(t.device).cuda_stream
    return t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.stride(0), t.stride(1), CC_MEM_DEVICE, t.device.index, stream or None, t


def summarize_draws(draws, by_chain=True, device=0):
    """cc_summarize_draws -> array [groups, params, len(SUMM_FIELDS)] (groups = chains, or 1 when not by_chain)."""
    ptr, Cn, n, P, sc, si, mem, dev, stream, _keep = _draws_arg(draws)
    out = np.empty((Cn if by_chain else 1, P, len(SUMM_FIELDS)))
    check(load().cc_summarize_draws(device if dev is None else dev, ptr, Cn, n, P, sc, si, int(bool(by_chain)), mem, stream, _dp(out)))
    return out


def chain_moments(draws, first_iter=0, device=0):
    """cc_chain_moments -> (mean, var) [chains, params]; numpy arrays for numpy draws, CUDA tensors for CUDA draws."""
    ptr, Cn, n, P, sc, si, mem, dev, stream, _keep = _draws_arg(draws)
    if mem == CC_MEM_HOST:
        mean, var = np.empty((Cn, P)), np.empty((Cn, P))
        check(load().cc_chain_moments(device, ptr, Cn, n, P, sc, si, int(first_iter), mem, None, mean.ctypes.data, var.ctypes.data))
        return mean, var
    import torch
    mean = torch.empty((Cn, P), dtype=torch.float64, device=draws.device)
    var = torch.empty_like(mean)
    check(load().cc_chain_moments(dev, ptr, Cn, n, P, sc, si, int(first_iter), mem, stream, mean.data_ptr(), var.data_ptr()))
    return mean, var


def gelman_from_moments(mean, var, n_iter_used, device=0):
    """cc_gelman_from_moments -> array [params, len(GELMAN_FIELDS)] from per-chain moments (numpy or CUDA tensors [chains, params])."""
    if isinstance(mean, np.ndarray):
        m, v = np.ascontiguousarray(mean, dtype=np.float64), np.ascontiguousarray(var, dtype=np.float64)
        Cn, P = m.shape
        out = np.empty((P, len(GELMAN_FIELDS)))
        check(load().cc_gelman_from_moments(device, m.ctypes.data, v.ctypes.data, Cn, int(n_iter_used), P, CC_MEM_HOST, None, _dp(out)))
        return out
    import torch
    m, v = mean.contiguous(), var.contiguous()
    Cn, P = m.shape
    out = np.empty((P, len(GELMAN_FIELDS)))
    stream = torch.cuda.current_stream(m.device).cuda_stream
    check(load().cc_gelman_from_moments(m.device.index, m.data_ptr(), v.data_ptr(), Cn, int(n_iter_used), P, CC_MEM_DEVICE, stream or None, _dp(out)))
    return out


def overall_ifr_draws(draws, idx_ifr, popN, idx_noise=None, device=0):
    """cc_overall_ifr_draws -> est [chains, iterations] (numpy for numpy draws, a CUDA tensor for CUDA draws)."""
    ptr, Cn, n, P, sc, si, mem, dev, stream, _keep = _draws_arg(draws)
    ii = np.ascontiguousarray(idx_ifr, dtype=np.int32)
    nn = None if idx_noise is None else np.ascontiguousarray(idx_noise, dtype=np.int32)
    pop = np.ascontiguousarray(popN, dtype=np.float64)
    if pop.size != ii.size or (nn is not None and nn.size != ii.size):
        raise ValueError("idx_ifr, idx_noise and popN must have one entry per stratum")
    if mem == CC_MEM_HOST:
        est = np.empty((Cn, n))
        check(load().cc_overall_ifr_draws(device, ptr, Cn, n, P, sc, si, _ip(ii), None if nn is None else _ip(nn), _dp(pop), ii.size, mem, None,
                                          est.ctypes.data))
        return est
    import torch
    est = torch.empty((Cn, n), dtype=torch.float64, device=draws.device)
    check(load().cc_overall_ifr_draws(dev, ptr, Cn, n, P, sc, si, _ip(ii), None if nn is None else _ip(nn), _dp(pop), ii.size, mem, stream,
                                      est.data_ptr()))
    return est


class Mcmc:
    """GPU-resident Metropolis-coupled MCMC run state (cc_mcmc)."""

    def __init__(self, model: Model, burnin: int, samples: int, rungs: int = 1, chains: int = 1, chain_offset: int = 0,
                 coupling_on: bool = True, GTI_pow: float = 1.0, beta_manual=None, seed: int = 1, cold_rung_last: bool = True,
                 record_rungs: bool = True, thin: int = 1, init=None, bw_per_particle: bool = False,
                 adapt_in_sampling: bool = True):
        self.model, self.L = model, model.L
        s = model.spec
        self.rungs, self.chains, self.d = rungs, chains, s.d
        self.n_rec_rungs = rungs if record_rungs else 1
        keep = self._keep = {}
        keep["init_default"] = np.ascontiguousarray(s.pinit, dtype=np.float64)
        if init is not None:
            init = np.ascontiguousarray(init, dtype=np.float64)
            if init.shape != (chains, s.d):
                raise ValueError("init must be [chains, d]")
            keep["init"] = init
        if beta_manual is not None:
            bm = np.ascontiguousarray(beta_manual, dtype=np.float64)
            if bm.shape != (rungs,):
                raise ValueError("beta_manual must have one entry per rung")
            keep["beta"] = bm
        o = McmcOpts(burnin=burnin, samples=samples, rungs=rungs, n_chains=chains, chain_offset=chain_offset,
                     coupling_on=int(coupling_on), GTI_pow=GTI_pow, beta_manual=_dp(keep["beta"]) if "beta" in keep else None,
                     seed=seed, cold_rung_last=int(cold_rung_last), record_rungs=int(record_rungs), thin=thin,
                     init=_dp(keep["init"]) if "init" in keep else None, init_default=_dp(keep["init_default"]),
                     bw_per_particle=int(bw_per_particle), no_adapt_in_sampling=int(not adapt_in_sampling))
        h = _vp()
        check(self.L.cc_mcmc_create(model.h, C.byref(o), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.cc_mcmc_destroy(self.h)
            self.h = None

    __del__ = close

    def advance(self, n_iter: int):
        check(self.L.cc_mcmc_advance(self.h, n_iter))

    @property
    def last_advance_ms(self):
        """Device time (CUDA events on the library's stream) of the last advance()."""
        ms = C.c_double()
        check(self.L.cc_mcmc_last_advance_ms(self.h, C.byref(ms)))
        return ms.value

    @property
    def iterations(self):
        return int(self.L.cc_mcmc_iterations(self.h))

    @property
    def rows(self):
        return int(self.L.cc_mcmc_rows(self.h))

    @property
    def loglike_evals(self):
        return int(self.L.cc_mcmc_loglike_evals(self.h))

    def fetch(self):
        rows = self.rows
        theta = np.empty((self.chains, self.n_rec_rungs, rows, self.d))
        ll = np.empty((self.chains, self.n_rec_rungs, rows))
        lp = np.empty((self.chains, self.n_rec_rungs, rows))
        it = np.empty(rows, dtype=np.int32)
        check(self.L.cc_mcmc_fetch(self.h, _dp(theta), _dp(ll), _dp(lp), _ip(it)))
        return dict(theta=theta, loglike=ll, logprior=lp, iteration=it)

    def device_draws(self):
        """(theta_ptr, loglike_ptr, logprior_ptr, rows_cap): raw device addresses of the recorded draws."""
        a, b, c, cap = _vp(), _vp(), _vp(), C.c_int64()
        check(self.L.cc_mcmc_device_draws(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(cap)))
        return a.value, b.value, c.value, cap.value

    def diagnostics(self):
        P = self.chains * self.rungs
        beta = np.empty(self.rungs)
        mc = np.full((self.chains, max(self.rungs - 1, 1)), np.nan)
        mv = np.empty((self.chains, self.rungs, self.d))
        bw = np.empty((P, self.d))
        check(self.L.cc_mcmc_diagnostics(self.h, _dp(beta), _dp(mc), _dp(mv), _dp(bw)))
        return dict(beta=beta, mc_accept=mc, move_accept=mv, bw=bw.reshape(self.chains, self.rungs, self.d))


def natcubspline_loglike(params: dict, param_i: int, data: dict, misc: dict, binomial: bool = True, device: int = 0):
    """COVIDCurve:::natcubspline_loglike_binomial / _logit (R/RcppExports.R:4-10) with the reference's own arguments: `params` a
    name -> value mapping (an R named vector), `data` and `misc` the lists of R/main.R:136-166.  Returns {"LogLik": value}."""
    names = [str(k).encode() for k in params]
    vals = np.ascontiguousarray([float(v) for v in params.values()], dtype=np.float64)
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))
    keep = dict(names=(C.c_char_p * len(names))(*names), obs=f64(data["obs_deaths"]), prop=f64(data["prop_strata_obs_deaths"]),
                sa=f64(data["obs_serologypos" if binomial else "obs_serologymu"]), sb=f64(data["obs_serologyn" if binomial else "obs_serologyse"]),
                demog=f64(misc["demog"]), st=i32(misc["sero_survey_start"]), en=i32(misc["sero_survey_end"]))
    a = CallArgs(n_params=len(names), names=C.cast(keep["names"], C.POINTER(C.c_char_p)), obs_deaths=_dp(keep["obs"]),
                 prop_strata_obs_deaths=_dp(keep["prop"]), sero_a=_dp(keep["sa"]), sero_b=_dp(keep["sb"]), demog=_dp(keep["demog"]),
                 n_strata=keep["demog"].size, n_knots=int(misc["n_knots"]), rcensor_day=int(misc["rcensor_day"]),
                 days_obsd=int(misc["days_obsd"]), n_sero_obs=int(misc["n_sero_obs"]), sero_survey_start=_ip(keep["st"]),
                 sero_survey_end=_ip(keep["en"]), max_seroday_obsd=int(misc["max_seroday_obsd"]),
                 account_serorev=int(bool(misc["account_serorev"])), binomial_likelihood=int(binomial))
    out = C.c_double()
    check(load().cc_loglike_call(C.byref(a), _dp(vals), int(param_i), device, C.byref(out)))
    return {"LogLik": out.value}


FN = dict(pgamma=1, dpois=2, dbinom=3, dnorm=4, dbeta=5, dunif=6, pweibull=7, stirlerr=8, bd0=9, fastlog=10, rpois=11)


def nmath(fn: str, a, b=0.0, c=0.0, device: int = 0):
    """Element-wise device evaluation of one of the R nmath routines on the path (test hook)."""
    a, b, c = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), np.asarray(c, dtype=np.float64))
    a, b, c = (np.ascontiguousarray(v).ravel() for v in (a, b, c))
    out = np.empty(a.size)
    check(load().cc_nmath_batch(device, FN[fn], _dp(a), _dp(b), _dp(c), a.size, _dp(out)))
    return out


def fp64_peak_probe(device: int = 0, iters: int = 4096):
    """Measured FP64 FMA peak of the device (TFLOP/s, FMA = 2 flops) - the roofline denominator of this path."""
    t, ms = C.c_double(), C.c_double()
    check(load().cc_fp64_peak_probe(device, iters, C.byref(t), C.byref(ms)))
    return t.value, ms.value


def dmma_probe(device: int = 0, iters: int = 4096, fma_per_mma: int = 0):
    """(mma TFLOP/s, concurrent fma TFLOP/s, ms) of the FP64 tensor-core probe."""
    a, b, ms = C.c_double(), C.c_double(), C.c_double()
    check(load().cc_dmma_probe(device, iters, fma_per_mma, C.byref(a), C.byref(b), C.byref(ms)))
    return a.value, b.value, ms.value


def simulate_infxn_2_death(infections, ifr, rho, popN, m_od, s_od, sero_delay_rate, smplfrac=1.0, seed=1, n_rep=1, device=0,
                           sero_rev_shape=None, sero_rev_scale=None):
    """cc_simulate_infxn_2_death: daily death, seroconversion (and, with the Weibull parameters, seroreversion) counts
    [n_rep][T][A] and the means [T][A] of the first two."""
    inf = np.ascontiguousarray(infections, dtype=np.float64)
    ifr, rho, popN = (np.ascontiguousarray(v, dtype=np.float64) for v in (ifr, rho, popN))
    T, A = inf.size, ifr.size
    if rho.size != A or popN.size != A:
        raise ValueError("ifr, rho and popN must have one entry per stratum")
    rev = sero_rev_shape is not None or sero_rev_scale is not None
    deaths, sero = np.empty((n_rep, T, A)), np.empty((n_rep, T, A))
    srev = np.empty((n_rep, T, A)) if rev else None
    md, ms = np.empty((T, A)), np.empty((T, A))
    check(load().cc_simulate_infxn_2_death(device, _dp(inf), T, _dp(ifr), _dp(rho), _dp(popN), A, float(m_od), float(s_od),
                                           float(sero_delay_rate), float(smplfrac), float(sero_rev_shape or 0.0) if rev else 0.0,
                                           float(sero_rev_scale or 0.0) if rev else 0.0, int(seed), int(n_rep), _dp(deaths), _dp(sero),
                                           _dp(srev) if rev else None, _dp(md), _dp(ms)))
    out = dict(deaths=deaths, serocon=sero, mean_deaths=md, mean_serocon=ms)
    if rev:
        out["serorev"] = srev
    return out
