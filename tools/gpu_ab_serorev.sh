#!/bin/bash
# dev run: seroreversion shapes (cfg2, cfg3 + serorev): values of build_variants/*.so against the in-tree build, then throughput
set -u
for cfg in cfg2 cfg3; do
  LL_CFG=$cfg LL_SEROREV=1 python tools/ll_dump.py /tmp/ll_base.npy
  for lib in build_variants/*.so; do LL_CFG=$cfg LL_SEROREV=1 CCB200_LIB=$PWD/$lib python tools/ll_dump.py /tmp/ll_v.npy && python tools/ll_dump.py /tmp/ll_base.npy /tmp/ll_v.npy --cmp | sed "s|/tmp/ll_v.npy|$cfg $(basename $lib)|"; done
done
for rep in 1 2; do
  for cfg in cfg2 cfg3; do
    python tools/ll_bench.py --cfg $cfg --serorev 1 --vectors 262144 --tag in-tree 2>&1 | grep LLBENCH
    for lib in build_variants/*.so; do CCB200_LIB=$PWD/$lib python tools/ll_bench.py --cfg $cfg --serorev 1 --vectors 262144 2>&1 | grep LLBENCH; done
  done
done
