"""Developer micro-bench of the batched likelihood (device-resident, cfg5 shape): M evals/s at several batch sizes for the
library named by CCB200_LIB (default: the in-tree build).  Box-uniform and posterior-like vectors; 2^18 distinct vectors
tiled up to the batch size (generating 2^20 conditioned draws takes half a minute of host time)."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from covidcurve_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--vectors", type=int, nargs="+", default=[1 << 17, 1 << 20])
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--cfg", default="cfg3")
ap.add_argument("--serorev", type=int, default=-1)
ap.add_argument("--tag", default="")
a = ap.parse_args()
spec, truth = synth.make_spec(a.cfg, serorev=None if a.serorev < 0 else bool(a.serorev))
model = _lib.Model(spec, device=0)
base = 1 << 18
box = synth.random_theta(spec, base, seed=synth.SEED_THETA, bad_frac=0.01)
near = synth.random_theta(spec, base, seed=synth.SEED_THETA + 1000, bad_frac=0.0, near_truth=synth.true_theta(spec, truth), jitter=0.02)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
res = []
for name, th in (("box", box), ("post", near)):
    for V in a.vectors:
        reps = (V + base - 1) // base
        soa = np.ascontiguousarray(np.tile(th, (reps, 1))[:V].T)
        dth = torch.from_numpy(soa).to(dev); dout = torch.empty(V, dtype=torch.float64, device=dev)
        for _ in range(3):
            model.loglike_ptr(dth.data_ptr(), V, V, dout.data_ptr(), _lib.CC_MEM_DEVICE, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            model.loglike_ptr(dth.data_ptr(), V, V, dout.data_ptr(), _lib.CC_MEM_DEVICE, st.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        res.append("%s@%dk %.2f" % (name, V >> 10, V * a.steps / (e0.elapsed_time(e1) * 1e-3) / 1e6))
        del dth, dout
print("LLBENCH %s %s: %s  (M evals/s)" % (a.tag or os.path.basename(os.environ.get("CCB200_LIB", "in-tree")), a.cfg, "  ".join(res)))
