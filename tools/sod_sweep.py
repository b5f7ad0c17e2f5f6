"""Developer aid: throughput of the batched likelihood as a function of sod (which gamma-table regime a vector is in)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib, synth
spec, truth = synth.make_spec("cfg3")
m = _lib.Model(spec)
V = 131072
base = synth.random_theta(spec, V, seed=5, bad_frac=0.0, near_truth=synth.true_theta(spec, truth), jitter=0.02)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for sod in (0.015, 0.025, 0.035, 0.05, 0.07, 0.1, 0.12, 0.15, 0.2, 0.3, 0.5, 0.85, 0.99):
    th = base.copy(); th[:, spec.index("sod")] = sod * (1 + 0.01 * np.random.default_rng(1).standard_normal(V))
    d_th = torch.from_numpy(np.ascontiguousarray(th.T)).cuda(); d_out = torch.empty(V, dtype=torch.float64, device="cuda")
    for _ in range(2): m.loglike_ptr(d_th.data_ptr(), V, V, d_out.data_ptr(), _lib.CC_MEM_DEVICE, st.cuda_stream)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m.loglike_ptr(d_th.data_ptr(), V, V, d_out.data_ptr(), _lib.CC_MEM_DEVICE, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    print("sod %.3f  a %8.1f  h %7.2f   %.1f M evals/s" % (sod, 1 / sod ** 2, 1 / (19.8 * sod ** 2), 3 * V / e0.elapsed_time(e1) / 1e3))
