import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from tests.golden import specs
from covidcurve_b200 import _lib
from covidcurve_b200.summaries import rhat_per_parameter
spec, g = specs.modfit()
model = _lib.Model(spec, device=0)
q_ref, names = g["post_q025_q50_q975"], list(g["param_names"])
mixed = [i for i, r in enumerate(g["stored_rhat"]) if r < 1.06]
for seed in (1, 2, 3, 2024, 77):
    burn, samp, chains, rungs = 3000, 3000, 8, 8
    mc = _lib.Mcmc(model, burnin=burn, samples=samp, rungs=rungs, chains=chains, seed=seed, record_rungs=False, GTI_pow=3.0)
    mc.advance(burn + samp - 1)
    out, diag = mc.fetch(), mc.diagnostics()
    draws = out["theta"][:, 0, burn:, :]
    q = np.quantile(draws.reshape(-1, spec.d), [0.025, 0.5, 0.975], axis=0)
    rh = rhat_per_parameter(draws)
    w = q_ref[2] - q_ref[0]
    dev = np.abs(q[1] - q_ref[1]) / w
    wr = (q[2] - q[0]) / w
    acc = diag["move_accept"][:, -1, :] / (burn + samp - 1)
    print("seed %d: max |dmedian|/width %.3f (%s)  width ratio %.2f..%.2f  max rhat %.3f  acc %.3f" % (
        seed, dev[mixed].max(), names[mixed[int(np.argmax(dev[mixed]))]], wr[mixed].min(), wr[mixed].max(), rh[mixed].max(), acc.mean()))
