#!/bin/bash
# dev run: A/B of library builds on the MCMC micro-bench (tools/mcmc_bench.py): in-tree + build_variants/*.so, 2 rounds
set -u
for rep in 1 2; do
  python tools/mcmc_bench.py --tag in-tree 2>&1 | grep MCBENCH
  for lib in build_variants/*.so; do CCB200_LIB=$PWD/$lib python tools/mcmc_bench.py --cfgs cfg3 2>&1 | grep MCBENCH; done
done
