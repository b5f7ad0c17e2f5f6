#!/bin/bash
# dev run: compare alternative builds (build_variants/*.so) on the MCMC leg of the bench
set -u
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --vectors 65536 --no-cpu-baseline --mcmc-iters 30 --mcmc-settle 100"
for lib in build_variants/*.so; do
  n=$(basename $lib .so)
  for rep in 1 2; do
    CC_VERBOSE=1 CCB200_LIB=$PWD/$lib $B > $O/var_$n.json 2> $O/var_$n.err; rc=$?; grep "sweep kernel" $O/var_$n.err | head -1
    python - $O/var_$n.json $rc <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "mcmc %.3fM iter/s  %.3f ms/iter"%(d['mcmc']['iter_per_sec_all_chains_x_rungs']/1e6, d['mcmc']['ms_per_iteration']))
except Exception as e: print(sys.argv[1], "rc", sys.argv[2], "ERR", e)
PY
  done
done
