"""Developer aid: stage-by-stage comparison of ONE parameter vector, tile kernel vs oracle (run on a GPU box).
usage: CC_DEBUG_DUMP=1 python tools/debug_row.py sod mod"""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib
from oracle.oracle import Oracle, lib
from tests.golden import specs
spec, g = specs.modfit()
th = g["theta"][g["loglike"].argmax()].copy()
th[spec.index("sod")] = float(sys.argv[1]); th[spec.index("mod")] = float(sys.argv[2])
m, o = _lib.Model(spec), Oracle(spec)
got = m.loglike(th[None, :])[0]
buf = np.zeros(4 * spec.T)
L = _lib.load(); L.cc_debug_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_double)]; L.cc_debug_fetch.restype = C.c_int
assert L.cc_debug_fetch(m.h, buf.ctypes.data_as(C.POINTER(C.c_double))) == 0
dg, infxn, auc, ch = buf.reshape(4, spec.T)
tr = o.trace(th)
sod, mod = float(sys.argv[1]), float(sys.argv[2])
ol = lib()
pg = np.array([ol.cco_pgamma(float(k), 1 / sod ** 2, mod * sod ** 2, 1, 0) for k in range(spec.T + 1)])
dgo = np.diff(pg)
print("gpu ll", got, "oracle", tr["loglike"], "L1 L2 L3", tr["L1"], tr["L2"], tr["L3"])
with np.errstate(all="ignore"):
    for name, a, b in (("dg", dg, dgo), ("infxn", infxn, tr["infxn"]), ("auc", auc, tr["auc"])):
        rel = np.where(b != 0, np.abs(a - b) / np.abs(b), np.abs(a))
        k = int(np.nanargmax(rel))
        print("%-6s max rel %.3e at day %d (gpu %.17g oracle %.17g); max abs %.3e" % (name, rel[k], k, a[k], b[k], np.max(np.abs(a - b))))
    rel = np.where(dgo > 1e-300, np.abs(dg - dgo) / dgo, 0)
    for k in np.argsort(-rel)[:8]:
        print("   dg day %d gpu %.12e oracle %.12e rel %.2e" % (k, dg[k], dgo[k], rel[k]))
