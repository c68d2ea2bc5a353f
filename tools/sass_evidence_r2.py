#!/usr/bin/env python
"""profiles/r02_sass_evidence_trunk.txt: per kernel of the conv trunk / training path (conv.o, conv_tc.o, wgrad_tc.o,
train.o), registers and the SASS mnemonics that show which hardware paths it uses: UTCHMMA = tcgen05.mma, LDTM =
tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA loads / stores, SYNCS = mbarrier, LDGSTS = cp.async,
FFMA2 = packed fp32 FMA, REDG = red.global.add.  Reads the objects built by `make` (cuobjdump -sass / -res-usage)."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "FFMA2", "FFMA", "LDS.128", "LDS", "STS.128",
        "LDG", "STG", "REDG", "SHFL", "BAR"]
out = ["# cuobjdump -sass of build/{conv,conv_tc,wgrad_tc,train}.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo): static counts",
       "%-66s %5s %5s | %s" % ("kernel (demangled prefix)", "regs", "instr", " ".join(KEYS))]
for obj in ("conv", "conv_tc", "wgrad_tc", "train"):
    path = os.path.join(ROOT, "build", obj + ".o")
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    regs, cur = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    fn, body = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            body[fn] = []
        elif fn and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", line)
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            body[fn].append(ins.split(";")[0].strip())
    rows = []
    for fn, ins in body.items():
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"b2f::\(anonymous namespace\)::", "", dem).split("(")[0].replace("void ", "")
        ops = [i.split()[0] if i else "" for i in ins]
        cnt = [sum(1 for o in ops if (o.startswith(k) if "." in k else o.split(".")[0] == k)) for k in KEYS]
        rows.append("%-66s %5s %5d | %s" % (dem[:66], regs.get(fn, "?"), len(ins), " ".join("%*d" % (len(k), c) for k, c in zip(KEYS, cnt))))
    out += sorted(rows)
open(os.path.join(ROOT, "profiles", "r02_sass_evidence_trunk.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
