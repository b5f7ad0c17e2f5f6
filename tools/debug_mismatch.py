"""Developer aid: list the bench-workload rows where GPU and oracle disagree (run on a GPU box)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import Oracle
spec, truth = synth.make_spec("cfg3")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
th = synth.random_theta(spec, 262144, seed=synth.SEED_THETA, bad_frac=0.01)[:n]
got, want = _lib.Model(spec).loglike(th), Oracle(spec).loglike(th, nthreads=os.cpu_count())
clamp = _lib.CC_OVERFLO_LOGLIKE
ok = (got != clamp) & (want != clamp)
rel = np.zeros(n); rel[ok] = np.abs(got[ok] - want[ok]) / np.abs(want[ok])
mis = np.nonzero((got == clamp) != (want == clamp))[0]
big = np.nonzero(rel > 1e-10)[0]
i_sod, i_mod = spec.index("sod"), spec.index("mod")
print("rows", n, "clamp mismatches", len(mis), "rel>1e-10", len(big), "p99.9 rel", np.quantile(rel[ok], 0.999), "max", rel.max())
for r in list(mis) + list(big[np.argsort(-rel[big])][:12]):
    sod, mod = th[r, i_sod], th[r, i_mod]
    print("row %d sod %.6f mod %.3f a %.1f h %.2f gpu %.12g cpu %.12g rel %.2e" % (r, sod, mod, 1 / sod ** 2, 1 / (mod * sod ** 2), got[r], want[r], rel[r]))
