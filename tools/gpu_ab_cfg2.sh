for rep in 1 2; do for lib in build_variants/*.so; do CCB200_LIB=$PWD/$lib python tools/ll_bench.py --cfg cfg2 --vectors 262144 2>&1 | grep LLBENCH; done; done
