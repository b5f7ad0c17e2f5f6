#!/bin/bash
# ncu --set full capture of the per-day pass (loglike_phase2_kernel<5>) on the cfg5 workload, exported to text + the SASS page
# as csv (per-instruction executed counts and stall samples); usage: tools/profile_p2.sh <tag> [box|post] [lib.so]
set -u
TAG=${1:-p2}; KIND=${2:-box}; LIB=${3:-covidcurve_b200/libcovidcurve_b200.so}
O=gpurun_out
export CCB200_LIB=$PWD/$LIB
CMD="python tools/ll_once.py $KIND 262144 5"
$CMD > $O/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:loglike_phase2 -s 3 -c 1 -o /tmp/$TAG $CMD > $O/${TAG}_ncu.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page details > $O/${TAG}_details.txt 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
NCU_LINES_OUTER=1 python tools/ncu_lines.py /tmp/$TAG.ncu-rep $LIB loglike_phase2_kernelILi5 40 > $O/${TAG}_stages.txt 2>&1
python tools/ncu_lines.py /tmp/$TAG.ncu-rep $LIB loglike_phase2_kernelILi5 90 > $O/${TAG}_lines.txt 2>&1
ls -la $O | grep ${TAG}_
