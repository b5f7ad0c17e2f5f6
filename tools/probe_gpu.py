"""Developer aid: FP64 vector / tensor-core throughput probes (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200 import _lib
for it in (2048, 8192, 32768):
    print("DFMA probe iters", it, _lib.fp64_peak_probe(0, it))
for f in (0, 1, 2, 4, 8):
    print("DMMA probe fma_per_mma", f, _lib.dmma_probe(0, 8192, f))
