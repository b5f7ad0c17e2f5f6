"""Developer aid: GPU vs oracle on box-uniform vectors of every synthetic configuration (run on a GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import Oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
clamp = _lib.CC_OVERFLO_LOGLIKE
for cfg, kw in (("cfg2", {}), ("cfg2", {"reparam": True}), ("cfg3", {"serorev": True}), ("cfg3", {"binomial": False}), ("cfg4", {}), ("cfg4", {"serorev": True})):
    spec, truth = synth.make_spec(cfg, **kw)
    th = synth.random_theta(spec, n, seed=11, bad_frac=0.01)
    got, want = _lib.Model(spec).loglike(th), Oracle(spec).loglike(th, nthreads=os.cpu_count())
    ok = (got != clamp) & (want != clamp)
    rel = np.zeros(n); rel[ok] = np.abs(got[ok] - want[ok]) / np.abs(want[ok])
    mis = int(np.sum((got == clamp) != (want == clamp)))
    print("%s %-18s n=%d  clamp mismatches %d  rel>1e-10: %d  p99.9 %.2e  max %.2e  (full evaluations %.1f%%)"
          % (cfg, kw, n, mis, int((rel > 1e-10).sum()), np.quantile(rel[ok], 0.999), rel.max(), 100 * ok.mean()))
    big = np.nonzero(rel > 1e-10)[0][:5]
    for r in big:
        print("    row %d sod %.5f mod %.3f gpu %.12g cpu %.12g rel %.2e" % (r, th[r, spec.index("sod")], th[r, spec.index("mod")], got[r], want[r], rel[r]))
