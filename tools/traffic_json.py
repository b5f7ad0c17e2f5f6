"""profiles/r2_traffic.json from an `ncu --set full` raw csv of the per-day pass: dram bytes per evaluation, tied to the sources the
capture was taken on (bench.py refuses the figure when the digest differs).  usage: traffic_json.py raw.csv vectors out.json"""
import csv, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
def get(n):
    i = names.index(n)
    v = float(vals[i].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3, "Gbyte/s": 1e9, "Mbyte/s": 1e6,
                "Kbyte/s": 1e3, "byte/s": 1}.get(units[i], 1)
V = int(sys.argv[2])
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
def pct(n):
    return float(vals[names.index(n)].replace(",", "")) if n in names else None
ncu = {"fp64_pipe_pct": pct("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
       "dmma_pipe_pct": pct("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "dram_GBps": get("dram__bytes.sum.per_second") / 1e9 if "dram__bytes.sum.per_second" in names else None,
       "warp_instructions_per_eval": (pct("smsp__inst_executed.sum") or 0.0) / V}
out = {"kernel": vals[names.index("Kernel Name")].split("(")[0], "vectors": V, "dram_bytes_read": rd, "dram_bytes_write": wr, "ncu": ncu,
       "bytes_per_eval": (rd + wr) / V, "gpu_time_ms_under_ncu": get("gpu__time_duration.sum"), "source_digest": bench.source_digest(),
       "how": "ncu --set full --clock-control none, one launch of the cfg5 box workload (tools/profile_r2.sh)"}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(out)
