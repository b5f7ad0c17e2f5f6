"""Developer aid: device pgamma (general algorithm) against the oracle's nmath port where the result is subnormal."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib
from oracle.oracle import lib
L = lib()
rng = np.random.default_rng(0)
n = 400000
a = rng.uniform(1800, 2300, n)
# x such that log P(a, x) ~ U(-744.4, -708.4) (the subnormal range): bisection on the leading-order lower-tail expansion
from scipy.special import gammaln
u = rng.uniform(-744.4, -706.0, n)
f = lambda xx: a * np.log(xx) - xx - gammaln(a + 1) - np.log1p(-xx / (a + 1)) - u
lo_, hi_ = np.full(n, 1.0), a - 1.0
for _ in range(80):
    mid = 0.5 * (lo_ + hi_)
    neg = f(mid) < 0
    lo_ = np.where(neg, mid, lo_); hi_ = np.where(neg, hi_, mid)
x = 0.5 * (lo_ + hi_)
want = np.array([L.cco_pgamma(xi, ai, 1.0, 1, 0) for xi, ai in zip(x[:60000], a[:60000])])
got = _lib.nmath("pgamma", x[:60000], a[:60000], 1.0)
sub = (want > 0) & (want < 2.3e-308)
print("subnormal results:", sub.sum(), " mismatches:", int((got[sub] != want[sub]).sum()), " zero-vs-nonzero:", int(((got == 0) != (want == 0)).sum()))
Q = 4.9406564584124654e-324
q = want[sub] / Q
for lo, hi in ((1, 8), (8, 64), (64, 512), (512, 4096), (4096, 1e6), (1e6, 1e9), (1e9, 1e12), (1e12, 1e16)):
    m = (q >= lo) & (q < hi)
    if m.any():
        dd = np.abs(got[sub][m] - want[sub][m]) / want[sub][m]
        print("  value in [%g, %g) quanta: n=%d mismatch=%d  median rel diff %.2e max %.2e" % (lo, hi, m.sum(), int((got[sub][m] != want[sub][m]).sum()), np.median(dd), dd.max()))
import mpmath
mpmath.mp.dps = 60
idx = np.nonzero(sub)[0]
small = [i for i in idx if want[i] / Q < 64 and got[i] != want[i]][:12]
for i in small:
    tru = mpmath.gammainc(mpmath.mpf(a[i]), 0, mpmath.mpf(x[i]), regularized=True)
    print("a %.6f x %.6f  gpu %d q  cpu %d q  exact %.3f q" % (a[i], x[i], round(got[i] / Q), round(want[i] / Q), float(tru / mpmath.mpf(Q))))
