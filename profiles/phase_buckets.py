#!/usr/bin/env python
"""Bucket executed warp-instructions of k_window_tpr by kernel phase (inlined-at chains resolved through
nvdisasm -g 'inlined at' info when present, else by the line of the innermost function)."""
import csv, re, sys
src_csv, sass = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
insts = [(r[1].strip(), int(r[ie]), int(r[smp])) for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
locs, cur = [], None
for ln in open(sass):
    m = re.findall(r'File "([^"]+)", line (\d+)', ln)
    if ln.strip().startswith("//##") and m:
        cur = [(f.split("/")[-1], int(l)) for f, l in m]   # innermost first, then "inlined at" callers
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        locs.append(cur)
assert len(locs) == len(insts), (len(locs), len(insts))
# phase by the OUTERMOST window_tpr.cuh line of k_window_tpr's body
text = open("epialleler_b200/csrc/window_tpr.cuh").read().split("\n")
def find(pat):
    for i, l in enumerate(text, 1):
        if pat in l: return i
    raise KeyError(pat)
marks = [("prologue", find("k_window_tpr(Csr csr")), ("zero+window setup", find("for (int win = blockIdx.x")), ("tile: metadata+TMA", find("for (int base = lo + warp * 32")),
         ("row sweep", find("const bool mine = valid")), ("meth fix-up", find("uint32_t todo = __ballot_sync")), ("finalize: coverage scan", find("const int q0 = threadIdx.x * PER")),
         ("finalize: row test", find("uint32_t flags = 0;")), ("finalize: compaction+emit", find("int nrows_total;")), ("end", 10**9)]
def fn_ranges(path):
    """(start_line, name) of every __device__/__global__ function in a header"""
    out = []
    for i, l in enumerate(open(path), 1):
        m = re.search(r"(?:__device__|__global__)[^(]*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m: out.append((i, m.group(1)))
    return out
FN = {"window_tpr.cuh": fn_ranges("epialleler_b200/csrc/window_tpr.cuh"), "kernels.cuh": fn_ranges("epialleler_b200/csrc/kernels.cuh")}
def phase(chain):
    if not chain: return "?"
    outer = [l for f, l in chain if f == "window_tpr.cuh" and l >= marks[0][1]]
    if not outer:
        f, l = chain[0]
        if f in FN:
            name = "?"
            for a, nm in FN[f]:
                if a <= l: name = nm
            return "inl:" + name
        return "inl:" + f
    l = outer[-1]
    for (n, a), (_, b) in zip(marks, marks[1:]):
        if a <= l < b: return n
    return "?"
def inner(chain):
    f, l = chain[0]
    if f == "window_tpr.cuh":
        t = text[l - 1]
    else:
        t = ""
    return f, l
agg, agg2 = {}, {}
tot = sum(i[1] for i in insts); tots = sum(i[2] for i in insts)
for (t, n, s), ch in zip(insts, locs):
    p = phase(ch)
    d = agg.setdefault(p, [0, 0]); d[0] += n; d[1] += s
    if p == "row sweep":
        f, l = inner(ch)
        sub = "events(push_events)" if (f == "window_tpr.cuh" and find("__device__ __forceinline__ uint32_t push_events") <= l < find("template <int NWARPS>\n".split("\n")[0]) and l < marks[0][1]) else ("planes/count (kernels.cuh)" if f == "kernels.cuh" else ("atomics" if "atomic" in f else "sweep body"))
        d = agg2.setdefault(sub, [0, 0]); d[0] += n; d[1] += s
print("total warp-instructions %d, stall samples %d" % (tot, tots))
for k, d in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %5.1f%% inst  %5.1f%% samples" % (k, 100 * d[0] / tot, 100 * d[1] / tots))
print("-- row sweep split --")
for k, d in sorted(agg2.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %5.1f%% inst  %5.1f%% samples" % (k, 100 * d[0] / tot, 100 * d[1] / tots))
