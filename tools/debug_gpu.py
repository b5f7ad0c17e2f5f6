"""Developer aid: locate a GPU/oracle mismatch stage by stage (run on a GPU box)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import Oracle, lib

spec, truth = synth.make_spec("cfg3")
m, o = _lib.Model(spec), Oracle(spec)
th = synth.random_theta(spec, 262144, seed=synth.SEED_THETA, bad_frac=0.01)[:24000]
got, want = m.loglike(th), o.loglike(th, nthreads=8)
clamp = _lib.CC_OVERFLO_LOGLIKE
mis = np.nonzero((got == clamp) != (want == clamp))[0]
print("clamp-set mismatches:", len(mis), "of", len(th))
ix = {n: i for i, n in enumerate(spec.names)}
L = lib()
for r in mis[:6]:
    t = th[r]
    sod, mod = t[ix["sod"]], t[ix["mod"]]
    tr = o.trace(t)
    cv = m.curves(t[None, :])
    x = np.arange(spec.T + 1, dtype=float)
    pg_gpu = _lib.nmath("pgamma", x, 1 / sod ** 2, mod * sod ** 2)
    pg_cpu = np.array([L.cco_pgamma(v, 1 / sod ** 2, mod * sod ** 2, 1, 0) for v in x])
    bad = np.nonzero(pg_gpu != pg_cpu)[0]
    print("row", r, "gpu", got[r], "cpu", want[r], "sod", sod, "mod", mod, "status", tr["status"])
    print("   auc zero days cpu", np.nonzero(tr["auc"] == 0)[0][-3:], "gpu", np.nonzero(cv["auc"][0] == 0)[0][-3:])
    for k in bad[:6]:
        print("   pgamma day", k, "gpu", pg_gpu[k], "cpu", pg_cpu[k])
