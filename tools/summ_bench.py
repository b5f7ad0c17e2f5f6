"""Developer timing of the device-side summaries at the cfg4 shape (one GPU, no process group): [chains, iterations, d] draws."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from covidcurve_b200 import _lib, dist, summaries
C, n, P = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (2048, 50, 48)))
g = torch.Generator(device="cuda").manual_seed(1)
big = torch.randn((C, 1, n + 40, P), dtype=torch.float64, device="cuda", generator=g)
draws = big[:, 0, 40:, :]
def t(name, fn, reps=3):
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        print("%-28s %8.2f ms" % (name, (time.perf_counter() - t0) * 1e3))
    return r
t("chain_moments", lambda: _lib.chain_moments(draws, first_iter=n // 2))
m, v = _lib.chain_moments(draws, first_iter=n // 2)
t("gelman_from_moments", lambda: _lib.gelman_from_moments(m, v, n - n // 2))
t("gelman_from_chain_moments", lambda: summaries.gelman_from_chain_moments(m, v, n - n // 2))
t("gelman_allgather", lambda: dist.gelman_allgather(draws, C))
t("summarize by_chain", lambda: _lib.summarize_draws(draws, by_chain=True))
t("summarize pooled", lambda: _lib.summarize_draws(draws, by_chain=False))
