#!/bin/bash
# dev run: GPU tests, then stagger / host-chunk experiments
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r3_tests.log 2>&1; echo "tests rc=$?" ; tail -4 $O/r3_tests.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --mcmc-iters 4 --mcmc-settle 6"
run() { # name, env...
  name=$1; shift
  env "$@" $B > $O/r3_$name.json 2> $O/r3_$name.err; rc=$?
  python - $O/r3_$name.json $rc <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2fM post %.2fM e2e %.2fM frac %.3f"%(d['value']/1e6,d['posterior_like']['value']/1e6,d['e2e']['value']/1e6,d['roofline']['frac']))
except Exception as e: print(sys.argv[1], "rc", sys.argv[2], "ERR", e)
PY
}
run base A=1
run stag5k CC_STAGGER=5000
run stag13k CC_STAGGER=13000
run stag30k CC_STAGGER=30000
run chunk128k CC_HOST_CHUNK=131072
run chunk512k CC_HOST_CHUNK=524288
