#!/bin/bash
# dev run: A/B of library builds on the likelihood micro-bench (tools/ll_bench.py), alternating, 2 rounds: in-tree + build_variants/*.so
set -u
for rep in 1 2; do
  python tools/ll_bench.py --tag in-tree 2>&1 | grep LLBENCH
  for lib in build_variants/*.so; do CCB200_LIB=$PWD/$lib python tools/ll_bench.py 2>&1 | grep LLBENCH; done
done
