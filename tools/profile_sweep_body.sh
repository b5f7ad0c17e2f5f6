#!/bin/bash
# dev run: steady-state capture of the Metropolis sweep with every instruction attributed to the statement of the kernel body it was
# inlined into (NCU_LINES_OUTER=2); usage: profile_sweep_body.sh [tag]
set -u
TAG=${1:-swb}; O=gpurun_out; mkdir -p $O; LIB=covidcurve_b200/libcovidcurve_b200.so
CMD="python tools/mcmc_bench.py --cfgs cfg3"
$CMD > $O/${TAG}_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:mcmc_sweep -s 104 -c 1 -o /tmp/$TAG $CMD > $O/${TAG}_ncu.log 2>&1
NCU_LINES_OUTER=2 python tools/ncu_lines.py /tmp/$TAG.ncu-rep $LIB mcmc_sweep_cached_kernelILi5 80 > $O/${TAG}_body.txt 2>&1
head -60 $O/${TAG}_body.txt | cut -c1-170
