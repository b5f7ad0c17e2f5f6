"""Developer aid (needs a -DCC_DEBUG_RAWQ build named by CCB200_LIB and CC_DEBUG_DUMP=1): accuracy of the unrounded upper-tail
values Q(x_k) of tile_gamma against mpmath, and whether fl(1 - Q) equals the oracle's (R's) lower-tail value bit for bit."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import lib
import mpmath as mp
mp.mp.dps = 50
spec, truth = synth.make_spec("cfg3")
rows = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "worst_rows_r2.npy"))
m = _lib.Model(spec)
L = _lib.load(); L.cc_debug_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_double)]; L.cc_debug_fetch.restype = C.c_int
ol = lib()
for th in rows:
    sod, mod = th[spec.index("sod")], th[spec.index("mod")]
    m.loglike(th[None, :])
    buf = np.zeros(4 * spec.T)
    assert L.cc_debug_fetch(m.h, buf.ctypes.data_as(C.POINTER(C.c_double))) == 0
    q = buf[3 * spec.T:]
    a, sc = 1 / sod ** 2, mod * sod ** 2
    print("sod %.6f alph %.2f h %.3f" % (sod, a, 1 / sc))
    for k in range(spec.T):
        if q[k] > 0 and q[k] > 1e-17:
            ex = mp.gammainc(a, k / sc, mp.inf, regularized=True)
            rel = float((mp.mpf(q[k]) - ex) / ex)
            pr = ol.cco_pgamma(float(k), a, sc, 1, 0)
            qr = ol.cco_pgamma(float(k), a, sc, 0, 0)
            flip = (1.0 - q[k]) != pr
            if k % 3 == 0 or flip:
                print("  day %3d Q %.6e  rel err gpu %.2e  R upper-tail rel err %.2e  %s" % (k, q[k], rel, float((mp.mpf(qr) - ex) / ex), "FLIP" if flip else ""))
