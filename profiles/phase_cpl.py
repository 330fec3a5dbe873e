#!/usr/bin/env python
"""Bucket the executed warp-instructions of k_window_cpl by kernel phase.  usage: phase_cpl.py REP [chunk_steps]"""
import re, subprocess, sys
rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 3.35e6
src = open("epialleler_b200/csrc/window_cpl.cuh").read().split("\n")
def find(pat):
    for i, l in enumerate(src, 1):
        if pat in l: return i
    raise KeyError(pat)
marks = sorted([(find("k_thr_tables(ThrParams"), "thr tables"), (find("bool eval_pass_tables"), "eval_pass"), (find("void red_shared"), "red/lds helpers"),
         (find("uint32_t block_exscan_u32"), "block scan"), (find("uint32_t count4"), "count4"), (find("void want_masks"), "want_masks"),
         (find("void chunk_events"), "events: E/runs"), (find("uint32_t ev = m[0]"), "events: loop"), (find("void push_meth"), "push_meth"),
         (find("k_window_cpl(Csr"), "prologue/zero"), (find("for (int base = lo"), "row phase 1"), (find("const uint8_t* ptr = csr.data"), "chunk body"),
         (find("if (FUSED) {\n".strip()) if False else find("rows: class counts -> pass flag"), "row phase 2"), (find("for (int d = 16; d >= 1; d >>= 1) npass_w"), "npass"),
         (find("const int q0 = threadIdx.x * PER"), "finalize")])
def phase(f, l):
    if f == "window_cpl.cuh":
        name = "?"
        for ln, nm in marks:
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
