#!/usr/bin/env python
"""Warp-instruction and stall-sample shares of k_window_pl by phase (source-line ranges of planar.cuh / win_common.cuh).
usage: buckets_pl.py REP"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
src = {}
for name in ("planar.cuh", "win_common.cuh"):
    src[name] = open("/root/repo/epialleler_b200/csrc/" + name).read().split("\n")
def find(name, needle, after=0):
    for i, l in enumerate(src[name]):
        if i + 1 > after and needle in l: return i + 1
    raise KeyError(needle)
P = "planar.cuh"
kp = find(P, "k_window_pl(Planar pl")
marks = [(find(P, "__device__ __forceinline__ uint32_t pl_dash"), "dash edges"),
         (find(P, "__device__ __forceinline__ void pl_events"), "event loops"),
         (find(P, "__device__ __forceinline__ void pl_push_meth"), "meth push loop"),
         (find(P, "__device__ __forceinline__ void window_finalize_pk"), "finalize"),
         (kp, "kernel prologue / hoisted address arithmetic"),
         (find(P, "for (int win = blockIdx.x", kp), "window/tile setup + row metadata + coverage deltas"),
         (find(P, "if (KEEP) {", kp), "unit loads"),
         (find(P, "if (FUSED) {", find(P, "if (KEEP) {", kp)), "class counts + pass chain + flag"),
         (find(P, "const uint32_t passm"), "event sweep: masks, dash logic, loop control"),
         (find(P, "} else {", find(P, "const uint32_t passm")), "streaming variant (MAXU=0)"),
         (find(P, "if (FUSED) {", find(P, "const bool pass_known")), "npass reduce + barrier + finalize call"),
         (find(P, "k_threshold_pl") - 3, "other")]
fin0 = find("win_common.cuh", "window_finalize_rt(")
out = {}; fname = None; hdr = None
tot = tots = 0
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): fname = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); th = hdr.index("Thread Instructions Executed"); continue
    if hdr and r and r[0].isdigit():
        num = lambda x: int(x) if x.strip().isdigit() else 0
        n, s, t, ln = num(r[ie]), num(r[sm]), num(r[th]), int(r[0])
        if not n and not s: continue
        if fname == P:
            b = "other"
            for m, nm in marks:
                if ln >= m: b = nm
        elif fname == "win_common.cuh":
            b = "finalize" if ln >= fin0 else "win_common helpers (red_shared, tables)"
        else:
            b = "intrinsics (" + (fname or "?") + ")"
        o = out.setdefault(b, [0, 0, 0]); o[0] += n; o[1] += s; o[2] += t
        tot += n; tots += s
print("total warp-instructions %d, stall samples %d" % (tot, tots))
for b, (n, s, t) in sorted(out.items(), key=lambda kv: -kv[1][0]):
    print("%5.1f%% inst %5.1f%% smp  thr %4.1f  %s" % (100.0 * n / tot, 100.0 * s / max(tots, 1), t / max(n, 1), b))
