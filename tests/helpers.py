"""Shared test helpers: golden fixtures, KAT inputs, comparison utilities."""
import glob
import os

import numpy as np

from epialleler_b200 import BamData

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COLS = ("rname", "strand", "pos", "context", "meth", "unmeth")

# must match tests/golden/make_golden.py
CASES = [
    ("cg_default", True, "CG", 2, 0.5, 0.1, "CG"),
    ("cx_nothr", False, "CG", 2, 0.5, 0.1, "CX"),
    ("cg_thr_cx_rep", True, "CG", 2, 0.5, 0.1, "CX"),
    ("chg_thr_cxg_rep", True, "CHG", 1, 0.3, 0.5, "CxG"),
    ("chh_thr_chh_rep", True, "CHH", 3, 0.1, 0.9, "CHH"),
    ("cxg_thr_chg_rep", True, "CxG", 2, 0.25, 0.2, "CHG"),
    ("cx_thr_cg_rep", True, "CX", 5, 0.05, 0.1, "CG"),
    ("cg_zero_thresholds", True, "CG", 0, 0.0, 1.0, "CG"),
]
BETAS = [("beta_cg", "Z", "z"), ("beta_cxg", "ZX", "zx"), ("beta_cx", "ZXH", "zxh"), ("beta_oo_cg", "XH", "xh")]
MHL_CASES = [("mhl_default", "CG", 0, 0, 0.1), ("mhl_win1", "CG", 1, 0, 0.1), ("mhl_nofilter", "CG", 0, 0, 1.0), ("mhl_cx_win3_len2", "CX", 3, 2, 0.5),
             ("mhl_chh_len5", "CHH", 0, 5, 0.2), ("mhl_cxg_win2", "CxG", 2, 1, 0.05)]
MHL_COLS = ("rname", "strand", "pos", "context", "coverage", "length", "lmhl")
# fixture -> [(case, bed.type, match.tolerance / match.min.overlap)]; must match tests/golden/make_golden.py
BED_CASES = {
    "amplicon010meth": [("ampl_tol1", "amplicon", 1), ("ampl_tol0", "amplicon", 0), ("ampl_tol25", "amplicon", 25), ("ampl_as_capture", "capture", 1)],
    "amplicon100meth": [("ampl_tol1", "amplicon", 1)],
    "capture": [("capt_ov1", "capture", 1), ("capt_ov100", "capture", 100), ("capt_ov_neg", "capture", -200), ("capt_as_amplicon", "amplicon", 300)],
    "capture_q30_bq20": [("capt_ov1", "capture", 1)],
}


def golden_names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    bam = BamData(z["rname"], z["strand"], z["start"], z["templid"], z["blob"], z["off"], [str(s) for s in z["levels"]])
    return bam, z


def ctx_to_idx(c):
    return ((ord(c) + 2) >> 2) & 15


def bam_from_xm(rows, high_nibble=0x10):
    """rows = [(rname, strand, start, "XM text")] in file order -> BamData sorted like R/internal.R:195."""
    rn = np.asarray([r[0] for r in rows], dtype=np.int32).reshape(-1)
    st = np.asarray([r[1] for r in rows], dtype=np.int32).reshape(-1)
    sa = np.asarray([r[2] for r in rows], dtype=np.int32).reshape(-1)
    seqxm = [bytes((high_nibble & 0xF0) | ctx_to_idx(ch) for ch in r[3]) for r in rows]
    order = np.lexsort((sa, rn)) if len(rows) else np.zeros(0, np.int64)
    return BamData.from_strings(rn[order], st[order], sa[order], np.arange(len(rows), dtype=np.int32)[order], seqxm)


def report_tuples(rep):
    return list(zip(*[rep[k].tolist() for k in COLS]))


def assert_report_equal(got, want, what=""):
    for k in COLS:
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, "%s: %s has %d rows, expected %d" % (what, k, g.shape[0], w.shape[0])
        if not np.array_equal(g, w):
            i = int(np.nonzero(g != w)[0][0])
            raise AssertionError("%s: column %s differs first at row %d: got %s want %s (row got=%s want=%s)" % (
                what, k, i, g[i], w[i], [int(got[c][i]) for c in COLS], [int(want[c][i]) for c in COLS]))


# ---- BAM-free known-answer tests (C-1: R/generateCytosineReport.R:16-54,
#      C-2: inst/unitTests/test_simulateBam.R:55-72, C-4: test_simulateBam.R:74-89) ------------------
KAT_C1_ROWS = [(1, 1, 1, "...Z..x.h..x..h."), (1, 1, 1, "...Z..z.h..x..h."), (1, 1, 1, "...Z..z.h..X..h."), (1, 1, 1, "...Z..z.h..z-.h.")]
KAT_C1_PASS = [0, 1, 0, 0]
KAT_C1_BETA = [1.0, 0.5, 0.5, 1.0 / 3.0]
KAT_C1_THR_ZXH = [(1, 1, 4, 7, 1, 3), (1, 1, 7, 7, 0, 3), (1, 1, 9, 2, 0, 4), (1, 1, 12, 6, 0, 3), (1, 1, 15, 2, 0, 4)]
KAT_C1_NOTHR_ZXH = [(1, 1, 4, 7, 4, 0), (1, 1, 7, 7, 0, 3), (1, 1, 9, 2, 0, 4), (1, 1, 12, 6, 1, 2), (1, 1, 15, 2, 0, 4)]

KAT_C2_ROWS = [(1, 1 if i % 2 else 2, i, "ZZZzzZZZ" if i % 2 else "ZZzzzzZZ") for i in range(1, 7)]
KAT_C2_Z = [(1, 1, 1, 7, 1, 0), (1, 1, 2, 7, 1, 0), (1, 2, 2, 7, 1, 0), (1, 1, 3, 7, 2, 0), (1, 2, 3, 7, 1, 0), (1, 1, 4, 7, 1, 1),
            (1, 2, 4, 7, 1, 1), (1, 1, 5, 7, 2, 1), (1, 2, 5, 7, 1, 1), (1, 1, 6, 7, 2, 1), (1, 2, 6, 7, 1, 2), (1, 1, 7, 7, 2, 1),
            (1, 2, 7, 7, 1, 2), (1, 1, 8, 7, 2, 1), (1, 2, 8, 7, 1, 2), (1, 1, 9, 7, 1, 1), (1, 2, 9, 7, 1, 2), (1, 1, 10, 7, 2, 0),
            (1, 2, 10, 7, 1, 1), (1, 1, 11, 7, 1, 0), (1, 2, 11, 7, 1, 1), (1, 1, 12, 7, 1, 0), (1, 2, 12, 7, 1, 0), (1, 2, 13, 7, 1, 0)]

KAT_C3_ROWS = [(1, 1, 10, "Zz"), (1, 1, 11, "-Z"), (1, 2, 11, "hH"), (2, 1, 5, "-----"), (2, 1, 6, "z.z"), (2, 1, 7, "U.Z"),
               (2, 2, 7, "xXx"), (2, 2, 8, "Xh")]
KAT_C3_PASS = [1, 0, 0, 0, 0, 0, 0, 0]
KAT_C3_NOTHR_ZXH = [(1, 1, 10, 7, 1, 0), (1, 1, 11, 7, 0, 1), (1, 2, 11, 2, 0, 1), (1, 1, 12, 7, 1, 0), (1, 2, 12, 2, 1, 0),
                    (2, 1, 6, 7, 0, 1), (2, 2, 7, 6, 0, 1), (2, 2, 8, 6, 2, 0), (2, 1, 9, 7, 1, 0)]
KAT_C3_THR_ZX = [(1, 1, 10, 7, 1, 0), (1, 1, 11, 7, 0, 1), (1, 1, 12, 7, 0, 1), (2, 1, 6, 7, 0, 1), (2, 2, 7, 6, 0, 1),
                 (2, 2, 8, 6, 0, 2), (2, 1, 9, 7, 0, 1)]


def kat_c4_rows():
    rows = [(1, 1, 1, "ZZZZZZZZZZ")]
    for i in range(999):
        s = ["z"] * 10
        s[i % 10] = "Z"
        rows.append((1, 1, 1, "".join(s)))
    return rows


def assert_mhl_equal(got, want, what=""):
    """Bit-exact: integer columns equal, double columns equal including NaN payload positions."""
    for k in MHL_COLS:
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, "%s: %s has %d rows, expected %d" % (what, k, g.shape[0], w.shape[0])
        if not np.array_equal(g, w, equal_nan=(g.dtype.kind == "f")):
            i = int(np.nonzero(~((g == w) | ((g != g) & (w != w))))[0][0])
            raise AssertionError("%s: column %s differs first at row %d: got %r want %r (pos %s)" % (what, k, i, g[i], w[i], want["pos"][i]))

# VCF-assisted report: fixture -> [(case, threshold.reads)]
VCF_CASES = {"amplicon010meth": [("vcf_thr", True), ("vcf_nothr", False)], "capture": [("vcf_thr", True)]}
VCF_COLS = ("M+Ref", "U+Ref", "M-Ref", "U-Ref", "M+Alt", "U+Alt", "M-Alt", "U-Alt", "SumRef", "SumAlt")


def golden_vcf(z):
    return {"seqnames": z["vcf.chr"], "start": z["vcf.start"], "REF": [str(x) for x in z["vcf.ref"]], "ALT": [str(x) for x in z["vcf.alt"]],
            "name": ["v%d" % i for i in range(z["vcf.chr"].shape[0])]}


def random_vcf(rng, bam, n, na_frac=0.05):
    """Sorted random SNV positions around the reads' own coordinates (duplicates included), a few NA chromosomes sprinkled in."""
    i = rng.integers(0, bam.nrow, n)
    pos = bam.start[i].astype(np.int64) + rng.integers(-20, 400, n)
    chr_ = bam.rname[i].astype(np.int64)
    order = np.lexsort((pos, chr_))
    chr_, pos = chr_[order].astype(np.int32), np.maximum(pos[order], 0).astype(np.int32)
    chr_[rng.random(n) < na_frac] = -2147483648
    letters = np.array(list("ACGT"))
    ref = letters[rng.integers(0, 4, n)]
    alt = letters[(np.searchsorted(letters, ref) + rng.integers(1, 4, n)) % 4]
    return {"seqnames": chr_, "start": pos, "REF": list(ref), "ALT": list(alt), "name": ["r%d" % k for k in range(n)]}


def format_report_text(rep, rname_levels, header=True):
    """What data.table::fwrite(report, quote=FALSE, sep="\\t", col.names=TRUE) writes for a cytosine report
    (/root/reference/R/internal.R:273-287): factor labels for rname / strand / context, plain integers."""
    ctx = {2: "CHH", 6: "CHG", 7: "CG"}
    lines = ["rname\tstrand\tpos\tcontext\tmeth\tunmeth\n"] if header else []
    for i in range(rep["pos"].shape[0]):
        lines.append("%s\t%s\t%d\t%s\t%d\t%d\n" % (rname_levels[rep["rname"][i] - 1], "+-"[rep["strand"][i] - 1], rep["pos"][i],
                                                    ctx[int(rep["context"][i])], rep["meth"][i], rep["unmeth"][i]))
    return "".join(lines).encode()
