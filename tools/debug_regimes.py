import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib
from oracle.oracle import Oracle
from tests.golden import specs
spec, g = specs.modfit()
base = g["theta"][g["loglike"].argmax()].copy()
i_sod, i_mod = spec.index("sod"), spec.index("mod")
sods = np.concatenate([np.geomspace(0.01, 0.999, 160), [0.05, 0.1, 0.2, 0.5, 1.0, 1.3]])
th = np.tile(base, (len(sods) * 3, 1))
th[:, i_sod] = np.tile(sods, 3)
th[:, i_mod] = np.repeat([5.0, 19.8, 60.0], len(sods))
got, want = _lib.Model(spec).loglike(th), Oracle(spec).loglike(th, nthreads=8)
rel = np.abs(got - want) / np.abs(want)
bad = np.nonzero(~(rel <= 1e-10))[0]
print("bad", len(bad))
for r in bad[:40]:
    sod, mod = th[r, i_sod], th[r, i_mod]
    print("sod %.5f mod %.1f a %.2f h %.3f  gpu %.10g cpu %.10g rel %.2e" % (sod, mod, 1 / sod ** 2, 1 / (mod * sod ** 2), got[r], want[r], rel[r]))
