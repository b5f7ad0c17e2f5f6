"""Selected metrics of an `ncu --page raw --csv` export (one kernel per file) as `name [unit] value` lines.
usage: ncu_metrics.py raw.csv [regex ...]"""
import csv, re, sys
DEFAULT = [r"^dram__bytes_(read|write)\.sum$", r"^gpu__time_duration\.sum$", r"^launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_.*)$",
           r"^sm__inst_executed\.avg\.per_cycle_active$", r"^sm__inst_executed_pipe_(fp64|tensor_subpipe_dmma|lsu|alu|fma|xu)\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
           r"^sm__pipe_(fp64|tensor_subpipe_dmma)_cycles_active\.avg\.pct_of_peak_sustained_active$", r"^sm__throughput\.avg\.pct", r"^sm__warps_active\.avg\.pct",
           r"^sm__cycles_active\.avg$", r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
           r"^lts__t_sector_hit_rate\.pct$", r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", r"^smsp__issue_active\.avg\.pct", r"^dram__throughput\.avg\.pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
pats = [re.compile(p) for p in (sys.argv[2:] or DEFAULT)]
print("# kernel:", vals[names.index("Kernel Name")][:150])
for n, u, v in sorted(zip(names, units, vals)):
    if any(p.search(n) for p in pats):
        print("%s [%s] %s" % (n, u, v))
