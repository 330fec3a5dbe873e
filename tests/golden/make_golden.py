"""Regenerates tests/golden/*.npz.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Inputs : the reference's bundled fixture BAMs (/root/reference/inst/extdata), decoded to the
         reference's ``bam.data`` layout by oracle/bamdata.py (no HTSlib/R here).
Outputs: for every case in CASES, the per-read pass flags / betas and the six report columns
         produced by the REFERENCE'S OWN, UNMODIFIED C++ (oracle/_ref/libepiref.so, compiled by
         oracle/Makefile from /root/reference/src).  These are the committed golden vectors the
         CPU and GPU parity tests compare against on machines where /root/reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bamdata import read_bam, read_bed, read_vcf  # noqa: E402
from oracle.oracle import CONTEXT_TO_BASES, RefCpu, build  # noqa: E402

EXT = "/root/reference/inst/extdata/"
OUT = os.path.dirname(os.path.abspath(__file__))

FIXTURES = {
    "amplicon000meth": (EXT + "amplicon000meth.bam", {}),
    "amplicon010meth": (EXT + "amplicon010meth.bam", {}),
    "amplicon100meth": (EXT + "amplicon100meth.bam", {}),
    "capture": (EXT + "capture.bam", {}),
    "capture_q30_bq20": (EXT + "capture.bam", dict(min_mapq=30, min_baseq=20)),
    "capture_trim": (EXT + "capture.bam", dict(trim=(3, 5))),
    "dragen_se": (EXT + "test/dragen-se-unsort-xg-xm.bam", {}),
    "dragen_pe": (EXT + "test/dragen-pe-namesort-xg-xm.bam", {}),
}

# (name, threshold_reads, threshold_context, min_sites, min_beta, max_oo_beta, report_context)
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
# lMHL report: (case, haplotype.context, max.haplotype.window, min.haplotype.length, max.outofcontext.beta)
MHL_CASES = [("mhl_default", "CG", 0, 0, 0.1), ("mhl_win1", "CG", 1, 0, 0.1), ("mhl_nofilter", "CG", 0, 0, 1.0), ("mhl_cx_win3_len2", "CX", 3, 2, 0.5),
             ("mhl_chh_len5", "CHH", 0, 5, 0.2), ("mhl_cxg_win2", "CxG", 2, 1, 0.05)]
# BED-assisted report: fixture -> (bed file, [(case, bed.type, match.tolerance / match.min.overlap)])
BEDS = {
    "amplicon010meth": (EXT + "amplicon.bed", [("ampl_tol1", "amplicon", 1), ("ampl_tol0", "amplicon", 0), ("ampl_tol25", "amplicon", 25),
                                               ("ampl_as_capture", "capture", 1)]),
    "amplicon100meth": (EXT + "amplicon.bed", [("ampl_tol1", "amplicon", 1)]),
    "capture": (EXT + "capture.bed", [("capt_ov1", "capture", 1), ("capt_ov100", "capture", 100), ("capt_ov_neg", "capture", -200),
                                      ("capt_as_amplicon", "amplicon", 300)]),
    "capture_q30_bq20": (EXT + "capture.bed", [("capt_ov1", "capture", 1)]),
}


# VCF-assisted report: fixture -> (vcf file, bed file or None, vcf.style, [(case, threshold.reads)])
VCFS = {
    "amplicon010meth": (EXT + "amplicon.vcf.gz", EXT + "amplicon.bed", "NCBI", [("vcf_thr", True), ("vcf_nothr", False)]),
    "capture": (EXT + "capture.vcf.gz", None, None, [("vcf_thr", True)]),  # test_generateVcfReport.R:9-17: a pre-read VCF, BED has no effect -> 26292 rows
}


def main():
    build(("ref",))
    for name, (path, kw) in FIXTURES.items():
        bam = read_bam(path, **kw)
        blob, off = bam.blob()
        out = dict(rname=bam.rname, strand=bam.strand, start=bam.start, templid=bam.templid, blob=blob, off=off,
                   levels=np.array(bam.rname_levels))
        ref = RefCpu(bam)
        for cname, thr, tctx, ms, mb, mo, rctx in CASES:
            p, rep = ref.generateCytosineReport(thr, tctx, ms, mb, mo, rctx)
            out[cname + ".pass"] = p.astype(np.uint8)
            for k, v in rep.items():
                out[cname + "." + k] = v
        for bname, m, u in BETAS:
            out[bname] = ref.rcpp_get_xm_beta(m, u)
        for cname, hctx, hmax, hmin, moo in MHL_CASES:
            rep = ref.generateMhlReport(hctx, hmax, hmin, moo)
            for k, v in rep.items():
                out[cname + "." + k] = v
        if name in BEDS:
            bed_path, cases = BEDS[name]
            chr_, st, en, _ = read_bed(bed_path, bam.rname_levels)
            out["bed.chr"], out["bed.start"], out["bed.end"] = chr_, st, en
            for cname, typ, par in cases:
                match, rep = ref.generateBedReport((chr_, st, en), typ, match_tolerance=par, match_min_overlap=par)
                out["bed." + cname + ".match"] = match
                out["bed." + cname + ".counts"] = rep["counts"]
                out["bed." + cname + ".vef"] = rep["VEF"]
                print("  bed", cname, "matched", int((match != -2147483648).sum()), "rows", rep["VEF"].shape[0])
        if name in VCFS:
            vcf_path, bed_path, style, cases = VCFS[name]
            vcf = read_vcf(vcf_path, bam.rname_levels, bed_path, style)
            out["vcf.chr"], out["vcf.start"] = vcf["seqnames"], vcf["start"]
            out["vcf.ref"], out["vcf.alt"] = np.array(vcf["REF"]), np.array(vcf["ALT"])
            for cname, thr in cases:
                freqs, rep = ref.generateVcfReport(vcf, threshold_reads=thr)
                out["vcf." + cname + ".freqs"] = freqs
                for k in ("M+Ref", "U+Ref", "M-Ref", "U-Ref", "M+Alt", "U+Alt", "M-Alt", "U-Alt", "SumRef", "SumAlt", "FEp+", "FEp-"):
                    out["vcf." + cname + "." + k] = rep[k]
                print("  vcf", cname, "entries", freqs.shape[0], "covered", int((freqs.sum(axis=1) > 0).sum()), "SumRef", rep["SumRef"].sum(), "SumAlt", rep["SumAlt"].sum())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, bam.nrow, int(off[-1]), {c[0]: int(out[c[0] + ".pos"].shape[0]) for c in CASES})


if __name__ == "__main__":
    main()
