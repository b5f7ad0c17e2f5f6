#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/check_tests.log 2>&1; echo "tests rc=$?" ; tail -4 $O/check_tests.log
CC_VERBOSE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/check_bench.json 2> $O/check_bench.err; echo "bench rc=$?"; grep covidcurve_b200 $O/check_bench.err | head -3
python - $O/check_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.2fM post %.2fM e2e %.2fM frac %.3f"%(d['value']/1e6,d['posterior_like']['value']/1e6,d['e2e']['value']/1e6,d['roofline']['frac']))
print(d['mcmc'])
PY
