#!/usr/bin/env python
"""List the hottest SASS instructions of a kernel from an ncu report (needs --import-source on).
usage: tools/ncu_hot.py report.ncu-rep kernel_regex [launch_skip] [top_n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
key = "# Samples"
def f(x):
    try: return float(x)
    except Exception: return 0.0
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[ci["Address"]] != "Address"]
tot = sum(f(r[ci[key]]) for r in data) or 1
print("kernel", rx, "instructions", len(data), "samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") or "Stall" in h and "Sampling" not in h]
print("#samp   pct  cum   executed  idx source")
cum = 0
order = sorted(range(len(data)), key=lambda i: -f(data[i][ci[key]]))[:topn]
for i in order:
    r = data[i]
    cum += f(r[ci[key]])
    reasons = sorted(((f(r[ci[h]]), h) for h in hdr if h.startswith("stall_")), reverse=True)[:2]
    rs = " ".join("%s=%d" % (h.replace("stall_", ""), v) for v, h in reasons if v > 0)
    print("%6d %5.1f %5.1f %9s %5d %-90s %s" % (f(r[ci[key]]), 100 * f(r[ci[key]]) / tot, 100 * cum / tot,
          r[ci["Instructions Executed"]], i, r[ci["Source"]][:90], rs))
