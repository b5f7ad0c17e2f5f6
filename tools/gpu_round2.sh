#!/bin/bash
# dev run: GPU tests, then the bench with the shipped build and with the 5-CTAs/SM variant of pass C
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_tests.log 2>&1; echo "tests rc=$?" ; tail -5 $O/r2_tests.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --mcmc-iters 10 --mcmc-settle 20"
$B --vectors 262144 > $O/r2_q4_256k.json 2> $O/r2_q4_256k.err; echo "q4 256k rc=$?"
$B > $O/r2_q4_1m.json 2> $O/r2_q4_1m.err; echo "q4 1m rc=$?"
CCB200_LIB=$PWD/build_variants/lib_p5.so $B --vectors 262144 > $O/r2_p5_256k.json 2> $O/r2_p5_256k.err; echo "p5 256k rc=$?"
CCB200_LIB=$PWD/build_variants/lib_p5.so $B > $O/r2_p5_1m.json 2> $O/r2_p5_1m.err; echo "p5 1m rc=$?"
for f in q4_256k q4_1m p5_256k p5_1m; do python - $O/r2_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2fM post %.2fM e2e %.2fM mcmc %.3fM frac %.3f"%(d['value']/1e6,d['posterior_like']['value']/1e6,d['e2e']['value']/1e6,d['mcmc']['iter_per_sec_all_chains_x_rungs']/1e6,d['roofline']['frac']))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
