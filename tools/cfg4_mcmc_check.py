"""Sanity + timing of the Metropolis sweep at the cfg4 shape (17 strata x 400 days, d = 48): final states re-evaluated by the oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from oracle.oracle import Oracle

spec, truth = synth.make_spec("cfg4")
model = _lib.Model(spec, device=0)
chains, rungs = 64, 16
init = np.tile(synth.true_theta(spec, truth), (chains, 1))
mc = _lib.Mcmc(model, burnin=60, samples=0, rungs=rungs, chains=chains, seed=5, record_rungs=True, init=init)
mc.advance(20)
t0 = time.perf_counter(); mc.advance(30); dt = time.perf_counter() - t0
out = mc.fetch()
th = out["theta"][:8, -1, -1, :]
chk = Oracle(spec).loglike(th)
rel = np.max(np.abs(out["loglike"][:8, -1, -1] - chk) / np.abs(chk))
print("cfg4: d=%d T=%d  %d particles  %.3f ms/iteration  %.3f M chain-rung iterations/s  final-state rel err vs oracle %.2e"
      % (spec.d, spec.T, chains * rungs, dt / 30 * 1e3, chains * rungs * 30 / dt / 1e6, rel))
assert rel < 1e-10
