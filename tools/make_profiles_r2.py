#!/usr/bin/env python
"""Round-2 summaries under profiles/ from the ncu captures merged back under gpurun_out/:
   launches_r2_net.csv (whole-network forward, tools/time_pwc.py --tc --B 8 --once), launches_r2a.csv (bench.py hot-path
   step), r2_prof_tc2.ncu-rep (--set full capture of the tensor-core convolution)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(n):
    n = n.replace("void ", "").replace("b2f::<unnamed>::", "").replace("unnamed>::", "")
    return n.split("(")[0]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, data = rows[0], rows[1:]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    out = []
    for r in data:
        v = float(r[vi].replace(",", ""))
        out.append((short(r[ki]), r[gi], v / 1000 if r[ui].startswith("n") else v, r[ki]))
    return out


def write_list(name, header, items):
    tot = sum(t for _, _, t, _ in items)
    with open(os.path.join(P, name), "w") as f:
        f.write(header)
        f.write("# %d launches, sum = %.1f us (cold-cache, serialised: compare shares, not absolutes)\n" % (len(items), tot))
        agg = collections.OrderedDict()
        for n, g, t, _ in items:
            k = n.split("<")[0] if "conv3x3_tc" not in n else n
            a = agg.setdefault(k, [0.0, 0])
            a[0] += t
            a[1] += 1
        f.write("\n# by kernel\n%-64s %6s %10s %7s\n" % ("kernel", "calls", "us", "share"))
        for k, (t, c) in sorted(agg.items(), key=lambda a: -a[1][0]):
            f.write("%-64s %6d %10.2f %6.1f%%\n" % (k[:64], c, t, 100 * t / tot))
        f.write("\n# every launch, in order\n%-64s %-16s %9s\n" % ("kernel", "grid", "us"))
        for n, g, t, _ in items:
            f.write("%-64s %-16s %9.2f\n" % (n[:64], g, t))
    return tot


net = launches(os.path.join(G, "launches_r2_net.csv"))
ours = [x for x in net if "b2f::" in x[3] or "unnamed>::" in x[3]]
n_per = 0
for line in open(os.path.join(G, "r2_net_once.log")):
    if line.startswith("launches per forward:"):
        n_per = int(line.split(":")[1])
# memsets / memcpys are not kernels: the last forward = the last K kernel launches of ours, K <= launches per forward
kernels_per = sum(1 for x in ours[len(ours) // 2:])
step = ours[len(ours) - len(ours) // 2:]
t1 = write_list("r02_ncu_launch_list_network.txt",
                "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tools/time_pwc.py --tc --B 8 --once\n"
                "# second of two eager whole-network forwards (Ours-Hard, 8 x 9 x 448 x 1024, decoders on tcgen05)\n", step)
b = launches(os.path.join(G, "launches_r2a.csv"))
ours_b = [x for x in b if "b2f::" in x[3] or "unnamed>::" in x[3]]
t2 = write_list("r02_ncu_launch_list_bench.txt",
                "# ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 2 --warmup 3 --no-cpu "
                "--no-e2e --no-criterions --no-training --no-network --no-parity\n# last captured hot-path step (56 launches)\n",
                ours_b[-56:])
print("network forward %.1f us over %d kernels; hot-path step %.1f us" % (t1, len(step), t2))

# ---- tensor-core convolution: raw metrics of the --set full capture ----------------------------------------------
rep = os.path.join(G, "r2_prof_tc2.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_uniform.sum", "sm__cycles_elapsed.max"]
with open(os.path.join(P, "r02_ncu_conv_tc_summary.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 python tools/prof_tc.py 8\n")
    f.write("# conv3x3_tc_kernel<128>: decoder layer 128 -> 128 channels at 112 x 256, B = 8 (2048 tiles of 7 x 16 pixels)\n")
    f.write("# algorithmic: 2 * 8*112*256 * 128*128*9 = 67.6 GFLOP fp32-equivalent = 202.9 GFLOP of TF32 MMA (three passes);\n")
    f.write("#              input (hi, lo) 235 MB + output (hi, lo) 235 MB + weights 1.2 MB\n")
    for h, u, v in zip(hdr, units, vals):
        if any(h.endswith(w) or h == w for w in want):
            f.write("%-86s %16s %s\n" % (h, v, u))
print(open(os.path.join(P, "r02_ncu_conv_tc_summary.txt")).read())
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "conv3x3_tc", "0", "30"], capture_output=True, text=True).stdout
open(os.path.join(P, "r02_ncu_conv_tc_hotspots.txt"), "w").write(
    "# hottest SASS instructions of conv3x3_tc_kernel<128> (sampling; 126 of 128 threads spin on the accumulator barrier while\n"
    "# one thread issues TMA and one issues UTCHMMA, so the samples show the waits, the tensor pipe's own activity is in the summary)\n" + hot)
