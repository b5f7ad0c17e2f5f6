import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib
from tests.golden import specs
spec, g = specs.modfit()
m = _lib.Model(spec)
ll = m.loglike(g["theta"][:64])
print(np.max(np.abs(ll - g["loglike"][:64]) / np.abs(g["loglike"][:64])))
