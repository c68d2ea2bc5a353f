#!/usr/bin/env python
"""profiles/r02_sass_hot_loops.txt: SASS LISTINGS (not counts) of the hot loops -- for every kernel below the longest
smallest backward-branch loop that holds at least half of its FMA / MMA / reduction instructions (the slab / channel /
tile loop), abridged to its first lines plus a mnemonic histogram, as
`cuobjdump -sass` of the objects built by `make` prints it."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("costvol", "costvol_fwd_tma<8, 1, true>", "cost-volume forward, FFMA2 channel loop"),
        ("costvol", "costvol_bwd_tma<1, 32, 2>", "cost-volume backward, slab loop"),
        ("costvol_tc", "costvol_fwd_tc<1>", "tensor-core cost volume: MMA issue loop / epilogue"),
        ("conv_tc", "conv3x3_tc_kernel<128>", "tensor-core convolution: MMA issue loop"),
        ("conv", "conv3x3_tma<8, 1, 8>", "FFMA2 convolution, channel loop"),
        ("warp", "warp_fwd_c3_win", "windowed C = 3 sampler forward"),
        ("warp", "warp_bwd_c3_lean<false>", "lean C = 3 sampler backward")]
HEAD = 28


def functions(obj):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "build", obj + ".o")], capture_output=True, text=True).stdout
    fn, body = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            body[fn] = []
        elif fn:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                body[fn].append((int(m.group(1), 16), m.group(2).strip()))
    return body


out = ["# cuobjdump -sass build/*.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)",
       "# per kernel: every backward branch = a loop; the smallest loop holding >= half of the kernel's FMA / MMA / RED instructions is listed (first %d lines from its first such instruction)" % HEAD, ""]
cache = {}
for obj, want, what in WANT:
    if obj not in cache:
        cache[obj] = functions(obj)
    for fn, ins in cache[obj].items():
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"b2f::\(anonymous namespace\)::", "", dem).replace("void ", "")
        dem = re.sub(r"\(int\)", "", dem)
        if want not in dem:
            continue
        addr = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
                loops.append((i - addr[int(m.group(1), 16)] + 1, addr[int(m.group(1), 16)], i))
        out.append("== %s  [%s]  %d instructions, %d loops" % (dem.split("(")[0], what, len(ins), len(loops)))
        if not loops:
            out.append("   (no loop)\n")
            break
        # the SMALLEST loop that still holds at least half of the kernel's arithmetic / tensor / reduction instructions
        keyset = ("FFMA2", "FFMA", "UTCHMMA", "RED.", "LDTM")
        nkey = lambda a, b: sum(any(k in t for k in keyset) for _, t in ins[a:b + 1])
        total = nkey(0, len(ins) - 1)
        good = [l for l in loops if nkey(l[1], l[2]) * 2 >= total] or [max(loops)]
        n, lo, hi = min(good)
        body = [t for _, t in ins[lo:hi + 1]]
        hist = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0] for t in body)
        out.append("   hot loop: /*%04x*/ .. /*%04x*/, %d instructions: %s" % (
            ins[lo][0], ins[hi][0], n, ", ".join("%s %d" % kv for kv in hist.most_common(12))))
        # skip to the first arithmetic-dense window so that the listing shows the inner pattern
        keys = ("FFMA2", "FFMA", "UTCHMMA", "RED", "LDS")
        start = next((i for i, t in enumerate(body) if any(k in t for k in keys)), 0)
        for t in body[start:start + HEAD]:
            out.append("      " + t)
        out.append("      ...\n")
        break
open(os.path.join(ROOT, "profiles", "r02_sass_hot_loops.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:6000])
