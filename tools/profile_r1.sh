#!/bin/bash
# Round-1 profile recipe (run under gpurun from the repo root): launch list + three full captures, exported to text on the box.
set -u
CMD="python bench.py --steps 2 --warmup 3 --vectors 262144 --mcmc-iters 4 --mcmc-settle 6 --no-cpu-baseline"
LIB=covidcurve_b200/libcovidcurve_b200.so
O=gpurun_out
$CMD > $O/r1_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r1_launches.csv $CMD > $O/r1_ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, mangled substring
  $CMD > $O/r1_plain_$1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 $CMD > $O/r1_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > $O/r1_$1_details.txt 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/r1_$1_raw.csv 2>/dev/null
  NCU_LINES_OUTER=1 python tools/ncu_lines.py /tmp/$1.ncu-rep $LIB $4 40 > $O/r1_$1_stages.txt 2>&1
  python tools/ncu_lines.py /tmp/$1.ncu-rep $LIB $4 70 > $O/r1_$1_lines.txt 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap phase2_box loglike_phase2 2 loglike_phase2_kernelILi5
cap phase2_post loglike_phase2 9 loglike_phase2_kernelILi5
cap mcmc_sweep mcmc_sweep 8 mcmc_sweep_cached_kernelILi5
ls -la $O | tail -20
