#!/bin/bash
# dev run: ncu --set full of one of the small kernels of the likelihood pipeline; usage: profile_small.sh <kernel regex> <tag>
set -u
RX=${1:-loglike_phase3}; TAG=${2:-p3}; O=gpurun_out; mkdir -p $O
CMD="python tools/ll_once.py box 262144 5"
$CMD > $O/${TAG}_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$RX -s 3 -c 1 -o /tmp/$TAG $CMD > $O/${TAG}_ncu.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page details > $O/${TAG}_details.txt 2>/dev/null
python tools/ncu_metrics.py $O/${TAG}_raw.csv > $O/${TAG}_metrics.txt
ncu -i /tmp/$TAG.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
grep -E "time_duration|registers|inst_executed.sum|issue_active|fp64|warps_active|per_cycle_active|dram__bytes|hit_rate" $O/${TAG}_metrics.txt
