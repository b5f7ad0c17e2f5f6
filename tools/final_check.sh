set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in cfg2 cfg3 cfg4; do CC_VERBOSE=1 python tools/ll_bench.py --cfg $cfg --vectors 262144 1048576 2>&1 | grep -E "LLBENCH|per-day pass" | tail -2; done
python tools/ll_once.py box 262144 5 > $O/r2_plain_phase2_box.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:loglike_phase2 -s 3 -c 1 -o /tmp/p2box python tools/ll_once.py box 262144 5 > $O/r2_ncu_phase2_box.log 2>&1
ncu -i /tmp/p2box.ncu-rep --page raw --csv > $O/r2_phase2_box_raw.csv 2>/dev/null
ncu -i /tmp/p2box.ncu-rep --page details > $O/r2_phase2_box_details.txt 2>/dev/null
python tools/ncu_metrics.py $O/r2_phase2_box_raw.csv > $O/r2_phase2_box_metrics.txt
NCU_LINES_OUTER=1 python tools/ncu_lines.py /tmp/p2box.ncu-rep covidcurve_b200/libcovidcurve_b200.so loglike_phase2_kernelILi5 40 > $O/r2_phase2_box_stages.txt 2>&1
python tools/ncu_lines.py /tmp/p2box.ncu-rep covidcurve_b200/libcovidcurve_b200.so loglike_phase2_kernelILi5 90 > $O/r2_phase2_box_lines.txt 2>&1
python tools/traffic_json.py $O/r2_phase2_box_raw.csv 262144 $O/r2_traffic.json | cut -c1-300
