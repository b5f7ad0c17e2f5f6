#!/bin/bash
# Round-1 profile recipe (run under gpurun from the repo root): default bench line, launch list, three full captures
# (per-day pass on the box-uniform and on the posterior-like workload, Metropolis sweep in steady state), exported to text
# on the box (the .ncu-rep files stay there: gpurun_out/ is capped at 64 MiB).
set -u
O=gpurun_out
LIB=covidcurve_b200/libcovidcurve_b200.so
python bench.py > $O/r1_bench_default.json 2> $O/r1_bench_default.err; echo "default bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --vectors 262144 --mcmc-iters 4 --mcmc-settle 6 --no-cpu-baseline"
$CMD > $O/r1_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r1_launches.csv $CMD > $O/r1_ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, mangled substring, command
  local name=$1 rx=$2 skip=$3 sym=$4; shift 4
  "$@" > $O/r1_plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/$name "$@" > $O/r1_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page details > $O/r1_${name}_details.txt 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/r1_${name}_raw.csv 2>/dev/null
  NCU_LINES_OUTER=1 python tools/ncu_lines.py /tmp/$name.ncu-rep $LIB $sym 40 > $O/r1_${name}_stages.txt 2>&1
  python tools/ncu_lines.py /tmp/$name.ncu-rep $LIB $sym 70 > $O/r1_${name}_lines.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap phase2_box loglike_phase2 2 loglike_phase2_kernelILi5 $CMD
cap phase2_post loglike_phase2 9 loglike_phase2_kernelILi5 $CMD
cap mcmc_sweep mcmc_sweep 104 mcmc_sweep_cached_kernelILi5 python bench.py --steps 1 --warmup 3 --vectors 65536 --no-cpu-baseline --mcmc-settle 100 --mcmc-iters 10
ls -la $O | grep r1_ | tail -30
