"""Developer aid: stage-by-stage comparison of saved cfg3 parameter vectors (tools/data/worst_rows_r2.npy), tile kernel vs
oracle vs mpmath (run on a GPU box with CC_DEBUG_DUMP=1)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import Oracle, lib
import mpmath as mp
mp.mp.dps = 40
spec, truth = synth.make_spec("cfg3")
rows = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "worst_rows_r2.npy"))
m, o = _lib.Model(spec), Oracle(spec)
L = _lib.load(); L.cc_debug_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_double)]; L.cc_debug_fetch.restype = C.c_int
ol = lib()
for th in rows[: int(sys.argv[1]) if len(sys.argv) > 1 else 1]:
    sod, mod = th[spec.index("sod")], th[spec.index("mod")]
    got = m.loglike(th[None, :])[0]
    buf = np.zeros(4 * spec.T)
    assert L.cc_debug_fetch(m.h, buf.ctypes.data_as(C.POINTER(C.c_double))) == 0
    dg, infxn, auc, ch = buf.reshape(4, spec.T)
    tr = o.trace(th)
    a, sc = 1 / sod ** 2, mod * sod ** 2
    pg = np.array([ol.cco_pgamma(float(k), a, sc, 1, 0) for k in range(spec.T + 1)])
    dgo = np.diff(pg)
    ex = [mp.gammainc(a, 0, k / sc, regularized=True) for k in range(spec.T + 1)]
    dge = np.array([float(ex[k + 1] - ex[k]) for k in range(spec.T)])
    print("sod %.6f mod %.4f alph %.2f h %.3f  gpu ll %.15g oracle %.15g rel %.2e  L1 %.6g L2 %.6g L3 %.6g" % (
        sod, mod, a, 1 / sc, got, tr["loglike"], abs(got - tr["loglike"]) / abs(tr["loglike"]), tr["L1"], tr["L2"], tr["L3"]))
    with np.errstate(all="ignore"):
        for name, x, y in (("dg gpu-vs-oracle", dg, dgo), ("dg gpu-vs-exact", dg, dge), ("dg oracle-vs-exact", dgo, dge),
                           ("infxn", infxn, tr["infxn"]), ("auc", auc, tr["auc"])):
            rel = np.where(np.abs(y) > 1e-300, np.abs(x - y) / np.abs(y), 0)
            k = int(np.nanargmax(rel))
            print("  %-20s max rel %.3e at day %d (%.17g vs %.17g)" % (name, rel[k], k, x[k], y[k]))
        rel = np.where(dgo > 1e-300, np.abs(dg - dgo) / dgo, 0)
        for k in np.argsort(-rel)[:6]:
            print("     day %d gpu %.15e oracle %.15e exact %.15e" % (k, dg[k], dgo[k], dge[k]))
        # contribution of each day to the L1 difference
        obs = spec.obs_deaths
        lam_g, lam_o = auc, tr["auc"]
        dl = obs * (np.log(lam_g) - np.log(lam_o)) * (lam_o > 0)
        k = np.argsort(-np.abs(np.nan_to_num(dl)))[:5]
        print("  largest per-day L1 differences:", [(int(i), float(dl[i]), float(obs[i]), float(lam_o[i])) for i in k])
