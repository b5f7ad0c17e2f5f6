#!/bin/bash
# dev run: A/B of library builds on the likelihood micro-bench (tools/ll_bench.py), alternating, 2 rounds: in-tree + build_variants/*.so;
# first the values of every variant against the in-tree build
set -u
python tools/ll_dump.py /tmp/ll_base.npy
for lib in build_variants/*.so; do CCB200_LIB=$PWD/$lib python tools/ll_dump.py /tmp/ll_v.npy && python tools/ll_dump.py /tmp/ll_base.npy /tmp/ll_v.npy --cmp | sed "s|/tmp/ll_v.npy|$(basename $lib)|"; done
for rep in 1 2; do
  python tools/ll_bench.py --tag in-tree 2>&1 | grep LLBENCH
  for lib in build_variants/*.so; do CCB200_LIB=$PWD/$lib python tools/ll_bench.py 2>&1 | grep LLBENCH; done
done
