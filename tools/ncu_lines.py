"""Developer aid: per-source-line cost of one kernel from an .ncu-rep (SASS page) + nvdisasm -gi line/inline info.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [top N]
Attributes every SASS instruction to the OUTERMOST line in our own sources (so that inlined libdevice math shows up at
its call site) and prints instructions executed / stall samples per line.
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
start = None
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):   # one cubin per translation unit of the library
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    start = next((i for i, l in enumerate(dis) if l.startswith("_Z") and kern in l and l.rstrip().endswith(":")), None)
    if start is not None:
        break
if start is None:
    raise SystemExit("kernel %s not found in %s" % (kern, lib))
lines = []  # per instruction: attributed (file, line)
cur = None
chain = []
fresh = True
ours = lambda f: "/csrc/" in f
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or (l.startswith("_Z") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((m.group(1), int(m.group(2))))
        mine = [c for c in chain if ours(c[0])]
        if os.environ.get("NCU_LINES_OUTER") == "2":  # attribute to the statement of the kernel body (covidcurve_b200.cu) it was inlined into
            body = [c for c in chain if c[0].endswith("covidcurve_b200.cu")]
            cur = body[-1] if body else chain[-1]
        elif os.environ.get("NCU_LINES_OUTER"):  # attribute to the line of the top-level evaluation function (stage view)
            body = [c for c in chain if c[0].endswith("cc_eval_tile.cuh")]
            cur = body[-1] if body else chain[-1]
        else:
            cur = mine[0] if mine else chain[-1]
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        lines.append(cur)
        fresh = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + ("loglike_phase2" if "phase2" in kern else "mcmc_sweep" if "mcmc" in kern else "loglike")],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
ni, ii, si = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
print("sass instructions: disasm %d, ncu %d" % (len(lines), len(body)))
agg = collections.defaultdict(lambda: [0, 0, 0])
pipes = collections.Counter()
for (fl, r) in zip(lines, body):
    a = agg[fl]
    a[0] += int(r[ni] or 0)
    a[1] += int(r[ii] or 0)
    a[2] += 1
    op = r[si].split()[0] if not r[si].strip().startswith("@") else r[si].split()[1]
    pipes[op.split(".")[0]] += int(r[ii] or 0)
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("total samples %d, warp instructions %d" % (ts, ti))
src_cache = {}
def src(f, n):
    if f not in src_cache:
        try: src_cache[f] = open(f).read().splitlines()
        except OSError: src_cache[f] = []
    s = src_cache[f]
    return s[n - 1].strip()[:90] if 0 < n <= len(s) else ""
for (fl, a) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, n = fl if fl else ("?", 0)
    print("%5.1f%% smp %5.1f%% ins %5d sass  %s:%d  %s" % (100 * a[0] / ts, 100 * a[1] / ti, a[2], os.path.basename(f), n, src(f, n)))
print("opcode mix (warp instructions):", ", ".join("%s %.1f%%" % (k, 100 * v / ti) for k, v in pipes.most_common(22)))
