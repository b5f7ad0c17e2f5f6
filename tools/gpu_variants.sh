#!/bin/bash
# dev run: compare alternative builds of the library (build_variants/*.so) on the likelihood legs of the bench
set -u
O=gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-iters 0"
for lib in build_variants/*.so; do
  n=$(basename $lib .so)
  for rep in 1 2; do
    CC_VERBOSE=1 CCB200_LIB=$PWD/$lib $B > $O/var_$n.json 2> $O/var_$n.err; rc=$?; grep "per-day pass" $O/var_$n.err | head -1
    python - $O/var_$n.json $rc <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2fM post %.2fM e2e %.2fM frac %.3f"%(d['value']/1e6,d['posterior_like']['value']/1e6,d['e2e']['value']/1e6,d['roofline']['frac']))
except Exception as e: print(sys.argv[1], "rc", sys.argv[2], "ERR", e)
PY
  done
done
