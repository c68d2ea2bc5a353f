#!/usr/bin/env python
"""Turn the ncu captures merged back under gpurun_out/ into the tracked summaries under profiles/.
usage: tools/make_profiles.py <tag>   (reads gpurun_out/launches_<tag>.csv, prof_<tag>_cv_raw.csv, prof_<tag>_warp_raw.csv)"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
rnd = "r01"

def short(n):
    n = n.replace("void ", "").replace("b2f::<unnamed>::", "").replace("unnamed>::", "")
    return n.split("(")[0]

# ---- launch list of the bench command ------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
def us(r):
    v = float(r[vi].replace(",", ""))
    return v / 1000 if r[ui].startswith("n") else v
ours = [(short(r[ki]), r[gi], us(r)) for r in data if "b2f::" in r[ki] or "unnamed>::" in r[ki]]
# one step = the last 56 kernels of ours (10 cv fwd, 18 warp fwd, 18 warp bwd, 10 cv bwd)
step = ours[-56:]
tot = sum(t for _, _, t in step)
with open(os.path.join(P, "%s_ncu_launch_list_bench.txt" % rnd), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:costvol|warp_|... -c 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-criterions\n")
    f.write("# last captured step (56 launches; cold-cache, serialised times).  sum = %.1f us\n" % tot)
    f.write("%-52s %-16s %9s %7s\n" % ("kernel", "grid", "us", "share"))
    for n, g, t in step:
        f.write("%-52s %-16s %9.2f %6.1f%%\n" % (n[:52], g, t, 100 * t / tot))
    agg = collections.OrderedDict()
    for n, g, t in step:
        k = n.split("<")[0] + (" " + g if "costvol" in n else "")
        agg[k] = agg.get(k, 0) + t
    f.write("\n# by kernel (cost volume split by grid = pyramid level)\n")
    for k, t in sorted(agg.items(), key=lambda a: -a[1]):
        f.write("%-52s %9.2f us %6.1f%%\n" % (k[:52], t, 100 * t / tot))
print("step sum %.1f us" % tot)

# ---- full-set summaries -------------------------------------------------------------------
traffic = {}
for part in ("cv", "warp", "crit"):
    raw = os.path.join(G, "prof_%s_%s_raw.csv" % (tag, part))
    if not os.path.exists(raw):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw], capture_output=True, text=True).stdout
    name = {"cv": "costvol", "warp": "warp", "crit": "criterions"}[part]
    with open(os.path.join(P, "%s_ncu_%s_summary_%s.txt" % (rnd, name, tag)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, python tools/prof_target.py %s 1 (BASELINE config sizes)\n" % {"cv": "cv3", "warp": "warp", "crit": "-> tools/prof_crit.py"}[part])
        f.write(out)
    r = list(csv.reader(open(raw)))
    h, u, d = r[0], r[1], r[2:]
    def mb(row, k):
        v = float(row[h.index(k)])
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[h.index(k)]]
    for row in d:
        n = short(row[h.index("Kernel Name")])
        key = None
        if "costvol_bwd_tma" in n: key = "costvol_bwd_L3"
        elif "costvol_fwd_tma" in n: key = "costvol_fwd_L3"
        if key and key not in traffic:
            traffic[key] = int(mb(row, "dram__bytes_read.sum") + mb(row, "dram__bytes_write.sum"))
json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
print(traffic)
