"""Developer aid: locate GPU/oracle mismatches (run on a GPU box).  usage: debug_gpu.py cfg [n] [serorev] [binomial]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import Oracle, lib

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
kw = {}
if len(sys.argv) > 3 and sys.argv[3] != "-":
    kw["serorev"] = bool(int(sys.argv[3]))
if len(sys.argv) > 4 and sys.argv[4] != "-":
    kw["binomial"] = bool(int(sys.argv[4]))
spec, truth = synth.make_spec(cfg, **kw)
m, o = _lib.Model(spec), Oracle(spec)
th = synth.random_theta(spec, n, seed=11, bad_frac=0.02)
got, want = m.loglike(th), o.loglike(th, nthreads=16)
clamp = _lib.CC_OVERFLO_LOGLIKE
mis = np.nonzero((got == clamp) != (want == clamp))[0]
ok = (got != clamp) & (want != clamp)
rel = np.zeros(len(th)); rel[ok] = np.abs(got[ok] - want[ok]) / np.abs(want[ok])
big = np.nonzero(rel > 1e-10)[0]
print(cfg, kw, "clamp-set mismatches:", len(mis), " rel>1e-10:", len(big), "of", len(th), " max rel", rel.max())
ix = {nm: i for i, nm in enumerate(spec.names)}
isod = ix["sod"]
if len(big):
    print("sod of big-rel rows: min %.4g max %.4g" % (th[big, isod].min(), th[big, isod].max()))
if len(mis):
    print("sod of clamp-mismatch rows: min %.4g max %.4g" % (th[mis, isod].min(), th[mis, isod].max()))
L = lib()
for r in list(mis[:4]) + list(big[np.argsort(-rel[big])][:4]):
    t = th[r]
    sod, mod = t[ix["sod"]], t[ix["mod"]]
    tr = o.trace(t)
    print("row", r, "gpu", got[r], "cpu", want[r], "rel", rel[r], "sod", sod, "mod", mod, "status", tr["status"], "L1 L2 L3", tr["L1"], tr["L2"], tr["L3"])
    x = np.arange(spec.T + 1, dtype=float)
    pg_gpu = _lib.nmath("pgamma", x, 1 / sod ** 2, mod * sod ** 2)
    pg_cpu = np.array([L.cco_pgamma(v, 1 / sod ** 2, mod * sod ** 2, 1, 0) for v in x])
    d = np.abs(pg_gpu - pg_cpu)
    print("   generic device pgamma vs cpu: max abs diff %.3e at day %d" % (d.max(), d.argmax()))
    cv = m.curves(t[None, :])
    with np.errstate(all="ignore"):
        e = np.abs(cv["auc"][0] - tr["auc"]) / np.abs(tr["auc"])
    print("   block-kernel auc vs cpu max rel %.3e ; auc zero days cpu %d gpu %d" % (np.nanmax(e), (tr["auc"] == 0).sum(), (cv["auc"][0] == 0).sum()))
