#!/bin/bash
set -u
O=gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-iters 0"
for c in 131072 262144 524288 1048576; do
  CC_HOST_CHUNK=$c $B > $O/chunk_$c.json 2> $O/chunk_$c.err
  python - $O/chunk_$c.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.2fM e2e %.2fM"%(d['value']/1e6,d['e2e']['value']/1e6))
PY
done
