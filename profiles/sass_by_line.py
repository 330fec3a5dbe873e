#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the
same kernel and aggregate executed warp-instructions / stall samples per CUDA source line.

usage: sass_by_line.py <ncu_source.csv> <nvdisasm_kernel.sass> [source file] [top N]
"""
import csv
import re
import sys

src_csv, sass = sys.argv[1], sys.argv[2]
srcfile = sys.argv[3] if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ie, smp, ath = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
insts = [(r[1].strip(), int(r[ie]), int(r[smp]), float(r[ath] or 0)) for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
lines, cur, inl = [], None, None
for ln in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
assert len(lines) == len(insts), (len(lines), len(insts))
agg = {}
tot = sum(i[1] for i in insts); tots = sum(i[2] for i in insts)
for (txt, n, s, a), loc in zip(insts, lines):
    d = agg.setdefault(loc, [0, 0, 0.0])
    d[0] += n; d[1] += s; d[2] += n * a
print("total warp-instructions %d, samples %d" % (tot, tots))
text = {}
if srcfile:
    for i, l in enumerate(open(srcfile), 1):
        text[i] = l.rstrip()
for loc, d in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall-samples  thr/inst %4.1f  %s:%s  %s" % (100 * d[0] / tot, 100 * d[1] / max(tots, 1), d[2] / max(d[0], 1), loc[0] if loc else "?", loc[1] if loc else "?",
                                                                          text.get(loc[1], "")[:100].strip() if loc else ""))
