"""Developer micro-bench of the Metropolis sweep: M chain-rung iterations/s at the cfg3 (64 x 50) and a cfg4 shard (512 x 64)
shape for the library named by CCB200_LIB, after the bandwidths have settled."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
ap = argparse.ArgumentParser()
ap.add_argument("--cfgs", nargs="+", default=["cfg3", "cfg4"])
ap.add_argument("--tag", default="")
a = ap.parse_args()
res = []
for cfg in a.cfgs:
    spec, truth = synth.make_spec(cfg)
    model = _lib.Model(spec, device=0)
    chains, rungs = (64, 50) if cfg == "cfg3" else (512, 64)
    settle, timed = (150, 400) if cfg == "cfg3" else (40, 40)
    init = np.tile(synth.true_theta(spec, truth), (chains, 1))
    mc = _lib.Mcmc(model, burnin=settle + timed + 2, samples=0, rungs=rungs, chains=chains, seed=synth.SEED_MCMC, record_rungs=False,
                   init=init, thin=50, GTI_pow=1.0 if cfg == "cfg3" else 3.0)
    mc.advance(settle)
    mc.advance(timed)
    ms = mc.last_advance_ms
    res.append("%s %dx%d %.3f M it/s (%.3f ms/it)" % (cfg, chains, rungs, chains * rungs * timed / ms / 1e3, ms / timed))
    mc.close(); model.close()
print("MCBENCH %s: %s" % (a.tag or os.path.basename(os.environ.get("CCB200_LIB", "in-tree")), "  ".join(res)))
