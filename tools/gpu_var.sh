#!/bin/bash
# dev run: tools/ll_bench.py for every build_variants/*.so (and the in-tree build), results appended to gpurun_out/var.log
set -u
O=gpurun_out/var.log
: > $O
python tools/ll_bench.py --tag in-tree "$@" 2>&1 | grep LLBENCH | tee -a $O
for lib in build_variants/*.so; do
  [ -f "$lib" ] || continue
  CCB200_LIB=$PWD/$lib python tools/ll_bench.py "$@" 2>&1 | grep -E "LLBENCH|Error|error" | tee -a $O
done
