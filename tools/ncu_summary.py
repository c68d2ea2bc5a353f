#!/usr/bin/env python
"""One line per kernel launch from `ncu --page raw --csv` output: duration, DRAM traffic, pipe use, stalls."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[ix[k]])
    except Exception: return d
def unit(k): return units[ix[k]] if k in ix else ""
def to_mb(r, k):
    v, u = g(r, k), unit(k)
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
def to_us(r, k):
    v, u = g(r, k), unit(k)
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
print("%-44s %6s %9s %8s %8s %6s %6s %6s %6s %5s  top stalls" % ("kernel", "grid", "us", "rdMB", "wrMB", "dram%", "fma%", "lsu%", "issue", "occ%"))
for r in data:
    name = r[ix["Kernel Name"]].replace("void ", "").replace("unnamed>::", "").replace("b2f::<", "")[:44]
    stalls = sorted(((g(r, h), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in hdr
                     if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")), reverse=True)
    tot = sum(v for v, _ in stalls) or 1
    st = " ".join("%s=%.0f%%" % (n, 100 * v / tot) for v, n in stalls[:4])
    print("%-44s %6d %9.2f %8.1f %8.1f %6.1f %6.1f %6.1f %6.2f %5.1f  %s" % (
        name, g(r, "launch__grid_size"), to_us(r, "gpu__time_duration.sum"), to_mb(r, "dram__bytes_read.sum"),
        to_mb(r, "dram__bytes_write.sum"), g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        g(r, "smsp__issue_active.avg.per_cycle_active"), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), st))
