#!/usr/bin/env python
"""Bucket the executed warp-instructions of k_window_cpl by kernel phase.  usage: phase_cpl.py REP [chunk_steps]"""
import re, subprocess, sys
rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 3.35e6
FILES = {"window_cpl.cuh": None, "win_common.cuh": None, "window_tpr.cuh": None}
for f in FILES: FILES[f] = open("epialleler_b200/csrc/" + f).read().split("\n")
def find(f, pat):
    for i, l in enumerate(FILES[f], 1):
        if pat in l: return i
    raise KeyError(pat)
C, K = "win_common.cuh", "window_cpl.cuh"
marks = {C: sorted([(find(C, "void k_thr_tables"), "thr tables"), (find(C, "bool eval_pass_tables"), "eval_pass"), (find(C, "void red_shared"), "red/lds helpers"),
         (find(C, "uint32_t block_exscan_u32"), "block scan"), (find(C, "uint32_t count4"), "count4"), (find(C, "void want_masks"), "want_masks"),
         (find(C, "void chunk_events"), "events: E/runs"), (find(C, "uint32_t ev = m[0]"), "events: loop"), (find(C, "void push_meth"), "push_meth"),
         (find(C, "void window_finalize"), "finalize")]),
         K: sorted([(find(K, "k_window_cpl(Csr"), "prologue/zero"), (find(K, "for (int base = my_lo"), "row phase 1"), (find(K, "const uint8_t* ptr = csr.data"), "chunk body"),
         (find(K, "rows: class counts -> pass flag"), "row phase 2"), (find(K, "for (int d = 16; d >= 1; d >>= 1) npass_w"), "npass/finalize call")])}
def phase(f, l):
    if f in marks:
        name = f
        for ln, nm in marks[f]:
            if l >= ln: name = nm
        return name
    if f == "kernels.cuh":
        if 70 <= l <= 82: return "make_planes"
        if 84 <= l <= 118: return "cls_*"
        if 128 <= l <= 142: return "eval_pass"
        return "ldg"
    return f
out = subprocess.run([sys.executable, "profiles/src_lines.py", rep, "100000"], capture_output=True, text=True).stdout
agg = {}; total = 0
for ln in out.split("\n"):
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% smp\s+thr\s+([\d.]+)\s+(\S+):(\d+)", ln)
    if ln.startswith("total"): total = int(ln.split()[2].rstrip(","))
    if not m: continue
    d = agg.setdefault(phase(m.group(4), int(m.group(5))), [0.0, 0.0]); d[0] += float(m.group(1)); d[1] += float(m.group(2))
print("total warp-instructions %d  (%.0f per chunk-step)" % (total, total / steps))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %5.1f%% inst %5.1f%% samples  ~%.0f instr/chunk-step" % (k, v[0], v[1], v[0] / 100 * total / steps))
