#!/usr/bin/env python
"""profiles/r01_sass_evidence.txt: per hot kernel, registers and the SASS mnemonics that show which hardware
paths it uses (UTMALDG/UTMASTG = TMA loads/stores, SYNCS = mbarrier, FFMA2 = packed fp32 FMA, LDS.128, RED = global
reductions, MUFU = SFU approximations).  Reads the objects built by `make` (cuobjdump -sass / -res-usage)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ["costvol_fwd_tma", "costvol_bwd_tma", "warp_fwd_vec4", "warp_bwd_vec4", "warp_fwd_c3_lean", "warp_bwd_c3_lean",
       "warp_bdhw_fwd", "warp_bdhw_bwd", "ob_kernel", "smooth1_lean_kernel", "smooth2_lean_kernel", "constvel_c2_kernel",
       "occprior_c2_kernel"]
KEYS = ["UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FFMA", "FMUL", "FADD", "LDS.128", "LDS", "STS.128", "STS", "LDG.E.128",
        "LDG", "STG.E.128", "STG", "RED", "MUFU", "SHFL", "BAR", "FSEL", "IMAD"]
out = ["# cuobjdump -sass of build/*.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo): static counts",
       "%-58s %5s %5s | %s" % ("kernel (demangled prefix)", "regs", "instr", " ".join(KEYS))]
for obj in ("costvol", "warp", "criterions"):
    path = os.path.join(ROOT, "build", obj + ".o")
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    regs = {}
    cur = None
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
    for fn, ins in body.items():
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"b2f::\(anonymous namespace\)::", "", dem).split("(")[0].replace("void ", "")
        if not any(h in dem for h in HOT):
            continue
        ops = [i.split()[0] if i else "" for i in ins]
        cnt = []
        for k in KEYS:
            if "." in k:
                cnt.append(sum(1 for o in ops if o.startswith(k)))
            else:
                cnt.append(sum(1 for o in ops if o.split(".")[0] == k))
        out.append("%-58s %5s %5d | %s" % (dem[:58], regs.get(fn, "?"), len(ins),
                                           " ".join("%*d" % (len(k), c) for k, c in zip(KEYS, cnt))))
open(os.path.join(ROOT, "profiles", "r01_sass_evidence.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
