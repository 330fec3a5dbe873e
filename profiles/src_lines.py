#!/usr/bin/env python
"""Per CUDA source line: executed warp-instructions, stall samples, avg active threads, from an .ncu-rep captured with
--import-source on.  usage: src_lines.py REP [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out = []; fname = None; hdr = None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): fname = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); th = hdr.index("Thread Instructions Executed"); continue
    if hdr and r and r[0].isdigit():
        def num(x):
            return int(x) if x.strip().isdigit() else 0
        n = num(r[ie])
        if n: out.append((n, num(r[sm]), num(r[th]), fname, int(r[0]), r[1]))
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total warp-instructions %d, stall samples %d" % (tot, tots))
for n, s, t, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  thr %4.1f  %s:%d  %s" % (100.0 * n / tot, 100.0 * s / max(tots, 1), t / n, f, ln, src.strip()[:110]))
