#!/bin/bash
# ncu --set full of the per-day pass at the other shapes: cfg4 (17 x 400: loglike_phase2_kernel<7>) and cfg3 with seroreversion
set -u
O=gpurun_out; mkdir -p $O
cap() {  # tag, env...
  local tag=$1; shift
  env "$@" python tools/ll_once.py box 262144 5 > $O/${tag}_plain.log 2>&1 || { echo plain failed; return; }
  env "$@" ncu --set full --clock-control none -k regex:loglike_phase2 -s 3 -c 1 -o /tmp/$tag python tools/ll_once.py box 262144 5 > $O/${tag}_ncu.log 2>&1
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/${tag}_raw.csv 2>/dev/null
  python tools/ncu_metrics.py $O/${tag}_raw.csv > $O/${tag}_metrics.txt
  grep -E "kernel|time_duration|registers_per|inst_executed.sum|issue_active.avg|fp64_cycles|dmma_cycles|warps_active|block_size|shared_mem_per_block" $O/${tag}_metrics.txt
}
cap r2_phase2_cfg4 LL_CFG=cfg4
cap r2_phase2_cfg3rev LL_CFG=cfg3 LL_SEROREV=1
