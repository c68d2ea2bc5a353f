#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs under build/: kernel, registers, spills, smem."""
import re, sys, glob, subprocess
for log in sorted(glob.glob('build/*.ptxas.log')):
    txt = open(log).read()
    blocks = re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]
    for b in blocks:
        name = b.split("'")[0]
        try:
            name = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
        except Exception:
            pass
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        name = name.split('(')[0].replace('void b2f::', '')
        regs = re.search(r'Used (\d+) registers', b)
        spill = re.search(r'(\d+) bytes spill stores, (\d+) bytes spill loads', b)
        stack = re.search(r'(\d+) bytes stack frame', b)
        smem = re.search(r'(\d+) bytes smem', b)
        print(f"{name:60s} regs={regs.group(1) if regs else '?':>4} stack={stack.group(1) if stack else 0:>5} "
              f"spill_st={spill.group(1) if spill else 0:>5} spill_ld={spill.group(2) if spill else 0:>5} smem={smem.group(1) if smem else 0}")
