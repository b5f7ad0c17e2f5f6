#!/bin/bash
# Round-2 profile recipe (run under gpurun from the repo root): GPU tests, default bench line, ncu launch list of a short bench,
# `ncu --set full` captures of the per-day pass (box workload) and of the Metropolis sweep in steady state, exported to text on
# the box (the .ncu-rep files stay there: gpurun_out/ is capped at 64 MiB), and profiles/r2_traffic.json's ingredients.
# usage: tools/profile_r2.sh [tag] [skip-tests | caps-only]
set -u
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
LIB=covidcurve_b200/libcovidcurve_b200.so
if [ "${2:-}" != "skip-tests" ] && [ "${2:-}" != "caps-only" ]; then
  python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 $O/${TAG}_pytest_gpu.log
fi
if [ "${2:-}" != "caps-only" ]; then
python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "default bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --vectors 262144 --mcmc-iters 4 --mcmc-settle 6 --mcmc-chains 512 --cfg3-seconds 0.05 --no-cpu-baseline --no-other-shapes"
$CMD > $O/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_l.log 2>&1
fi
cap() {  # name, kernel regex, skip, mangled substring, command
  local name=$1 rx=$2 skip=$3 sym=$4; shift 4
  "$@" > $O/${TAG}_plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/$name "$@" > $O/${TAG}_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page details > $O/${TAG}_${name}_details.txt 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${TAG}_${name}_raw.csv 2>/dev/null
  python tools/ncu_metrics.py $O/${TAG}_${name}_raw.csv > $O/${TAG}_${name}_metrics.txt 2>&1
  NCU_LINES_OUTER=1 python tools/ncu_lines.py /tmp/$name.ncu-rep $LIB $sym 40 > $O/${TAG}_${name}_stages.txt 2>&1
  python tools/ncu_lines.py /tmp/$name.ncu-rep $LIB $sym 90 > $O/${TAG}_${name}_lines.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap phase2_box loglike_phase2 3 loglike_phase2_kernelILi5 python tools/ll_once.py box 262144 5
python tools/traffic_json.py $O/${TAG}_phase2_box_raw.csv 262144 $O/${TAG}_traffic.json > /dev/null 2>&1
cap phase2_post loglike_phase2 3 loglike_phase2_kernelILi5 python tools/ll_once.py post 262144 5
cap mcmc_sweep mcmc_sweep 104 mcmc_sweep_cached_kernelILi5 python tools/mcmc_bench.py --cfgs cfg3
ls -la $O | grep ${TAG}_ | tail -40
