"""Generates tests/golden/oracle_mcmc_{cfg2,cfg3,cfg4}.npz: cold-rung posterior draws of the CPU oracle's drjacoby-equivalent sampler
(oracle/cc_oracle_mcmc.cpp driving oracle/cc_oracle.cpp, i.e. the restated reference likelihood) on the synthetic data of
BASELINE configs[1], configs[2] and configs[3] (covidcurve_b200/synth.py, seed 20201).  tests/test_gpu_mcmc.py compares the GPU engine's
posterior (different seed, different ladder) with these draws: quantile agreement within Monte-Carlo error and pooled
split-R-hat < 1.1 for every parameter (north_star correctness check #2; SURVEY.md 8d "Value check").

Run from the repo root (minutes of CPU time; deterministic given the seeds):   python tests/golden/make_mcmc_golden.py [threads] [cfg ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from covidcurve_b200 import synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

# one (cold) rung: the CPU budget goes into chain length, which is what the slowly mixing knot / curve-height parameters need
SETTINGS = {"cfg2": dict(chains=8, rungs=1, burnin=3000, samples=17000, seed=777),
            "cfg3": dict(chains=8, rungs=1, burnin=3000, samples=17000, seed=778),
            "cfg4": dict(chains=8, rungs=1, burnin=3000, samples=12000, seed=779)}
THIN = 20


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    only = sys.argv[2:]
    for cfg, kw in SETTINGS.items():
        if only and cfg not in only:
            continue
        spec, truth = synth.make_spec(cfg)
        init = np.tile(synth.true_theta(spec, truth), (kw["chains"], 1))
        t0 = time.time()
        r = Oracle(spec).mcmc(init=init, nthreads=threads, **kw)
        cold = r["theta"][:, -1, kw["burnin"]:, :]                       # [chains, samples, d], cold rung = last rung position
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_mcmc_%s.npz" % cfg)
        np.savez_compressed(out, draws=cold[:, ::THIN, :].astype(np.float64), names=np.array(spec.names),
                            q=np.quantile(cold.reshape(-1, spec.d), [0.025, 0.5, 0.975], axis=0),
                            loglike_cold=r["loglike"][:, -1, kw["burnin"]::THIN], mc_accept=r["mc_accept"], beta=r["beta"],
                            settings=np.array(sorted("%s=%s" % kv for kv in kw.items())), thin=THIN)
        print("%s: %d chains x %d rungs x %d iterations in %.0f s -> %s" % (cfg, kw["chains"], kw["rungs"], kw["burnin"] + kw["samples"],
                                                                             time.time() - t0, out))


if __name__ == "__main__":
    main()
