"""Developer aid: log-likelihoods of the cfg5 box + posterior-like vectors (2 x 131072) for the library named by CCB200_LIB, saved to
argv[1]; `ll_dump.py a.npy b.npy --cmp` compares two dumps (bit-identical?  max relative difference)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--cmp" in sys.argv:
    a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
    bad = -np.finfo(float).max / 100
    ok = (a != bad) & (b != bad)
    print("CMP %s %s: identical %s, failure values equal %s, max rel diff %.3e" % (sys.argv[1], sys.argv[2], np.array_equal(a, b),
          np.array_equal(a == bad, b == bad), np.max(np.abs(a[ok] - b[ok]) / np.abs(a[ok])) if ok.any() else 0.0))
    sys.exit(0)
from covidcurve_b200 import _lib, synth
spec, truth = synth.make_spec(os.environ.get("LL_CFG", "cfg3"), serorev=True if os.environ.get("LL_SEROREV") else None)
model = _lib.Model(spec, device=0)
n = 1 << 17
box = synth.random_theta(spec, n, seed=synth.SEED_THETA, bad_frac=0.01)
near = synth.random_theta(spec, n, seed=synth.SEED_THETA + 1000, bad_frac=0.0, near_truth=synth.true_theta(spec, truth), jitter=0.02)
np.save(sys.argv[1], np.concatenate([model.loglike(box), model.loglike(near)]))
