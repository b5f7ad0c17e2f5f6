"""Developer aid: the statistics of tests/test_gpu_mcmc.py::test_posterior_agrees_with_the_oracle_cpu_sampler for one config, printed
instead of asserted (thresholds are chosen from this).  usage: posterior_vs_oracle.py cfg4"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
from covidcurve_b200.summaries import rhat_per_parameter
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_mcmc_%s.npz" % cfg))
spec, truth = synth.make_spec(cfg)
model = _lib.Model(spec, device=0)
chains, rungs, burn, samp = 32, 8, 4000, 12000
init = np.tile(synth.true_theta(spec, truth), (chains, 1))
mc = _lib.Mcmc(model, burnin=burn, samples=samp, rungs=rungs, chains=chains, seed=90210, record_rungs=False, GTI_pow=3.0, init=init, thin=10)
mc.advance(burn + samp - 1)
print("device ms", mc.last_advance_ms)
out = mc.fetch()
keep = out["iteration"] > burn
gpu = out["theta"][:, 0][:, keep, :]
ora = g["draws"]
def split(x):
    h = x.shape[1] // 2
    return np.concatenate([x[:, :h], x[:, h:2 * h]], axis=0)
r_gpu, r_ora = rhat_per_parameter(split(gpu)), rhat_per_parameter(split(ora))
q_gpu = np.quantile(gpu.reshape(-1, spec.d), [0.025, 0.5, 0.975], axis=0)
q_ora = np.quantile(ora.reshape(-1, spec.d), [0.025, 0.5, 0.975], axis=0)
n = min(gpu.shape[1], ora.shape[1]) // 2 * 2
print("max rhat gpu %.3f; oracle mixed %d of %d" % (r_gpu.max(), (r_ora < 1.1).sum(), spec.d))
for i, name in enumerate(spec.names):
    width = q_ora[2, i] - q_ora[0, i]
    pooled = np.concatenate([split(gpu[:8, -n:, i:i + 1]), split(ora[:, -n:, i:i + 1])], axis=0)
    print("%-14s rhat gpu %.3f ora %.3f  dmed/width %6.3f  width ratio %5.2f  pooled rhat %.3f  inside %s" % (
        name, r_gpu[i], r_ora[i], (q_gpu[1, i] - q_ora[1, i]) / width, (q_gpu[2, i] - q_gpu[0, i]) / width, rhat_per_parameter(pooled)[0],
        q_ora[0, i] <= q_gpu[1, i] <= q_ora[2, i]))
