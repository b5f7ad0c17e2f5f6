"""Developer aid for ncu captures: a few device-resident cc_loglike_batch launches on the cfg5 workload (box-uniform or
posterior-like vectors), nothing else.  usage: ll_once.py [box|post] [vectors] [launches]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from covidcurve_b200 import _lib, synth
kind = sys.argv[1] if len(sys.argv) > 1 else "box"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
spec, truth = synth.make_spec(os.environ.get("LL_CFG", "cfg3"), serorev=True if os.environ.get("LL_SEROREV") else None)
model = _lib.Model(spec, device=0)
if kind == "box":
    th = synth.random_theta(spec, V, seed=synth.SEED_THETA, bad_frac=0.01)
else:
    th = synth.random_theta(spec, V, seed=synth.SEED_THETA + 1000, bad_frac=0.0, near_truth=synth.true_theta(spec, truth), jitter=0.02)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
dth = torch.from_numpy(np.ascontiguousarray(th.T)).to(dev); dout = torch.empty(V, dtype=torch.float64, device=dev)
for _ in range(n):
    model.loglike_ptr(dth.data_ptr(), V, V, dout.data_ptr(), _lib.CC_MEM_DEVICE, st.cuda_stream)
torch.cuda.synchronize()
print("ok", kind, V, float(dout[:8].sum()))
