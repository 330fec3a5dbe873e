#!/usr/bin/env python
"""Registers / spills / smem per kernel from an nvcc -Xptxas -v log.  usage: ptxas_summary.py LOG [substring]"""
import re, subprocess, sys
log = open(sys.argv[1]).read().split("\n")
pat = sys.argv[2] if len(sys.argv) > 2 else ""
cur = None
for ln in log:
    m = re.search(r"Compiling entry function '([^']+)'", ln)
    if m:
        cur = m.group(1); continue
    if cur and "Used" in ln:
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        if pat in name:
            regs = re.search(r"Used (\d+) registers", ln).group(1)
            print("%-70s regs %s  %s" % (name[-70:], regs, spill))
        cur = None
    elif cur and "spill" in ln:
        spill = ln.strip()
