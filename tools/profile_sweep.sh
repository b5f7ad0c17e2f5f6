#!/bin/bash
# steady-state capture of the Metropolis sweep kernel (after 100 settle iterations), exported to text on the box
set -u
TAG=${1:-r1b}
CMD="python bench.py --steps 1 --warmup 3 --vectors 65536 --no-cpu-baseline --mcmc-settle 100 --mcmc-iters 10"
LIB=covidcurve_b200/libcovidcurve_b200.so
O=gpurun_out
$CMD > $O/${TAG}_sweep_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_sweep -s 104 -c 1 -o /tmp/sweep $CMD > $O/${TAG}_sweep_ncu.log 2>&1
ncu -i /tmp/sweep.ncu-rep --page details > $O/${TAG}_mcmc_sweep_details.txt 2>/dev/null
ncu -i /tmp/sweep.ncu-rep --page raw --csv > $O/${TAG}_mcmc_sweep_raw.csv 2>/dev/null
NCU_LINES_OUTER=1 python tools/ncu_lines.py /tmp/sweep.ncu-rep $LIB mcmc_sweep_cached_kernelILi5 60 > $O/${TAG}_mcmc_sweep_stages.txt 2>&1
python tools/ncu_lines.py /tmp/sweep.ncu-rep $LIB mcmc_sweep_cached_kernelILi5 90 > $O/${TAG}_mcmc_sweep_lines.txt 2>&1
rm -f /tmp/sweep.ncu-rep
tail -2 $O/${TAG}_sweep_ncu.log; head -30 $O/${TAG}_mcmc_sweep_stages.txt | cut -c1-160
