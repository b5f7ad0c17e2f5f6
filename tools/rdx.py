"""Reader for R's serialisation (.rda / .rds) without R - now lives in the package (covidcurve_b200/rdata.py) next to the
writer; this module keeps the old import path used by tests/golden/make_golden.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covidcurve_b200.rdata import RDX, env_items, get, load, names, read_rds  # noqa: E402,F401
