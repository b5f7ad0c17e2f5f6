"""Developer aid: stage comparison of one row of the bench workload (cfg3, seed 20202). usage: CC_DEBUG_DUMP=1 python tools/debug_row2.py row"""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import Oracle, lib
spec, truth = synth.make_spec("cfg3")
r = int(sys.argv[1])
th = synth.random_theta(spec, 262144, seed=synth.SEED_THETA, bad_frac=0.01)[r]
m, o = _lib.Model(spec), Oracle(spec)
got = m.loglike(th[None, :])[0]
buf = np.zeros(4 * spec.T)
L = _lib.load(); L.cc_debug_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_double)]; L.cc_debug_fetch.restype = C.c_int
assert L.cc_debug_fetch(m.h, buf.ctypes.data_as(C.POINTER(C.c_double))) == 0
dg, infxn, auc, ch = buf.reshape(4, spec.T)
tr = o.trace(th)
sod, mod = th[spec.index("sod")], th[spec.index("mod")]
ol = lib()
a, sc = 1 / sod ** 2, mod * sod ** 2
pg = np.array([ol.cco_pgamma(float(k), a, sc, 1, 0) for k in range(spec.T + 1)])
dgo = np.diff(pg)
print("sod", sod, "mod", mod, "a", a, "h", 1 / sc, "gpu ll", got, "oracle", tr["loglike"], "status", tr["status"])
with np.errstate(all="ignore"):
    rel = np.where(dgo != 0, np.abs(dg - dgo) / np.abs(dgo), np.abs(dg))
    for k in np.argsort(-rel)[:10]:
        print("   dg day %d x %.1f gpu %.12e oracle %.12e rel %.2e  obs %g" % (k, (k + 1) / sc, dg[k], dgo[k], rel[k], spec.obs_deaths[k]))
    rel = np.where(tr["auc"] != 0, np.abs(auc - tr["auc"]) / np.abs(tr["auc"]), np.abs(auc))
    for k in np.argsort(-rel)[:6]:
        print("   auc day %d gpu %.12e oracle %.12e rel %.2e  obs %g" % (k, auc[k], tr["auc"][k], rel[k], spec.obs_deaths[k]))
