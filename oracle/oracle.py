"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

ctypes front-ends for the two CPU checkers:

* ``RefCpu``  -- the reference's own three hot-path C++ sources compiled unmodified
  (oracle/_ref/libepiref.so, built by oracle/Makefile from /root/reference/src).
* ``PortCpu`` -- the independent plain-C restatement oracle/cx_oracle.c
  (oracle/_build/liboracle.so).

Both expose the same three calls, named like the reference's Rcpp exports
(/root/reference/R/RcppExports.R:12-14,28-30,64-66):
``rcpp_threshold_reads``, ``rcpp_get_xm_beta``, ``rcpp_cx_report``.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libepiref.so")
PORT_SO = os.path.join(HERE, "_build", "liboracle.so")

# context name -> the four letter sets (/root/reference/R/internal.R:54-65)
CONTEXT_TO_BASES = {
    "CG": ("Z", "z", "XH", "xh"),
    "CHG": ("X", "x", "ZH", "zh"),
    "CHH": ("H", "h", "ZX", "zx"),
    "CxG": ("ZX", "zx", "H", "h"),
    "CX": ("ZXH", "zxh", "", ""),
}


def build(which=("port", "ref")):
    """(Re)build the checkers.  'ref' needs /root/reference; absent there, a prebuilt .so is used."""
    for w in which:
        if w == "ref" and not os.path.isdir("/root/reference/src"):
            continue
        subprocess.run(["make", "-s", "-C", HERE, w], check=True)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


NA_INT = -2147483648  # R's NA_integer_


def bed_report_table(strand, pass_, match, nbed):
    """The table R/internal.R:529-561 builds with data.table: rows = BED rows 1..nbed in file order, then one NA row
    if any read is unmatched; columns nreads+ / nreads- (reads of strand 1 / 2 matched to the row, passing or not) and
    VEF = passing / all.  BED rows that no read matched hold NaN (NA after the all=TRUE merge)."""
    strand = np.asarray(strand); pass_ = np.asarray(pass_).astype(bool); match = np.asarray(match)
    na = match == NA_INT
    idx = np.where(na, nbed, match - 1).astype(np.int64)          # NA bucket last
    n_rows = nbed + (1 if na.any() else 0)
    cnt = np.zeros((nbed + 1, 2, 2), np.int64)                      # [row][pass][strand-1]
    np.add.at(cnt, (idx, pass_.astype(np.int64), strand.astype(np.int64) - 1), 1)
    cnt = cnt[:n_rows]
    tot = cnt.sum(axis=(1, 2))
    plus = (cnt[:, 0, 0] + cnt[:, 1, 0]).astype(np.float64)
    minus = (cnt[:, 0, 1] + cnt[:, 1, 1]).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        vef = (cnt[:, 1, 0] + cnt[:, 1, 1]) / tot.astype(np.float64)
    empty = tot == 0
    plus[empty] = np.nan; minus[empty] = np.nan; vef[empty] = np.nan
    return {"bedrow": np.concatenate([np.arange(1, nbed + 1), [NA_INT] if n_rows > nbed else []]).astype(np.int64),
            "nreads+": plus, "nreads-": minus, "VEF": vef, "counts": cnt}


# .getBaseFreqReport's REF/ALT table (R/internal.R:642-665), restated as data: for every (REF, ALT) the base columns that
# are summed into the Ref and the Alt count on the '+' and on the '-' strand; None = NA (that strand cannot tell the two
# alleles apart after bisulfite conversion: C>T on '+', G>A on '-').
BASE_FREQ_RULES = {
    ("A", "C"): (("A",), ("C", "T"), ("A",), ("C",)),
    ("A", "T"): (("A",), ("T",), ("A",), ("T",)),
    ("A", "G"): (("A",), ("G",), None, None),
    ("C", "A"): (("C", "T"), ("A",), ("C",), ("A",)),
    ("C", "T"): (None, None, ("C",), ("T",)),
    ("C", "G"): (("C", "T"), ("G",), ("C",), ("A", "G")),
    ("T", "A"): (("T",), ("A",), ("T",), ("A",)),
    ("T", "C"): (None, None, ("T",), ("C",)),
    ("T", "G"): (("T",), ("G",), ("T",), ("A", "G")),
    ("G", "A"): (("G",), ("A",), None, None),
    ("G", "C"): (("G",), ("C", "T"), ("A", "G"), ("C",)),
    ("G", "T"): (("G",), ("T",), ("A", "G"), ("T",)),
}
_BASE_COL = {"A": 0, "C": 1, "G": 2, "T": 3}


def base_freq_table(freqs, ref, alt):
    """freqs: (n, 20) from rcpp_get_base_freqs; ref/alt: sequences of single letters.  Returns the eight Ref/Alt columns
    (NaN = NA), SumRef and SumAlt (rowSums with na.rm=TRUE).  Rows whose (REF, ALT) is not in the table stay NA."""
    n = freqs.shape[0]
    cols = {k: np.full(n, np.nan) for k in ("M+Ref", "U+Ref", "M-Ref", "U-Ref", "M+Alt", "U+Alt", "M-Alt", "U-Alt")}
    block = {"U+": 0, "U-": 5, "M+": 10, "M-": 15}
    for i in range(n):
        rule = BASE_FREQ_RULES.get((ref[i], alt[i]))
        if rule is None:
            continue
        for strand, (r, a) in (("+", rule[0:2]), ("-", rule[2:4])):
            for m in ("M", "U"):
                if r is not None:
                    cols[m + strand + "Ref"][i] = sum(freqs[i, block[m + strand] + _BASE_COL[x]] for x in r)
                    cols[m + strand + "Alt"][i] = sum(freqs[i, block[m + strand] + _BASE_COL[x]] for x in a)
    cols["SumRef"] = np.nansum([cols["M+Ref"], cols["U+Ref"], cols["M-Ref"], cols["U-Ref"]], axis=0)
    cols["SumAlt"] = np.nansum([cols["M+Alt"], cols["U+Alt"], cols["M-Alt"], cols["U-Alt"]], axis=0)
    return cols


class _Cpu:
    PREFIX = ""
    SO = ""

    def __init__(self, bam):
        """bam: anything with rname/strand/start/templid int32 arrays and .blob() (file-order strings)."""
        if not os.path.exists(self.SO):
            build((self.KIND,))
        self.lib = C.CDLL(self.SO)
        L, px = self.lib, self.PREFIX
        getattr(L, px + "bam_new").restype = C.c_void_p
        getattr(L, px + "bam_new").argtypes = [C.c_long] + [C.POINTER(C.c_int)] * 4 + [C.c_long, C.POINTER(C.c_uint8), C.POINTER(C.c_int64)]
        getattr(L, px + "bam_free").argtypes = [C.c_void_p]
        getattr(L, px + "threshold_reads").restype = C.c_double
        getattr(L, px + "threshold_reads").argtypes = [C.c_void_p] + [C.c_char_p] * 4 + [C.c_uint, C.c_double, C.c_double, C.POINTER(C.c_uint8)]
        getattr(L, px + "get_xm_beta").restype = C.c_double
        getattr(L, px + "get_xm_beta").argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]
        getattr(L, px + "cx_report").restype = C.c_long
        getattr(L, px + "cx_report").argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_char_p, C.POINTER(C.c_double)]
        getattr(L, px + "cx_fetch").argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 6
        getattr(L, px + "match_target").restype = C.c_double
        getattr(L, px + "match_target").argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.POINTER(C.c_int)] * 3 + [C.c_int, C.POINTER(C.c_int)]
        self.strand = np.ascontiguousarray(bam.strand, dtype=np.int32)
        getattr(L, px + "mhl_report").restype = C.c_long
        getattr(L, px + "mhl_report").argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
        getattr(L, px + "mhl_fetch").argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_double)] * 2
        getattr(L, px + "base_freqs").restype = C.c_double
        getattr(L, px + "base_freqs").argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        getattr(L, px + "fep").argtypes = [C.c_long] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_double)]
        self.n = int(bam.rname.shape[0])
        data, off = bam.blob()
        self.nstr = int(off.shape[0] - 1)
        cols = [np.ascontiguousarray(getattr(bam, k), dtype=np.int32) for k in ("rname", "strand", "start", "templid")]
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if data.shape[0] == 0:
            data = np.zeros(1, np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        self.h = getattr(L, px + "bam_new")(self.n, *[_p(c, C.c_int) for c in cols], self.nstr, _p(data, C.c_uint8), _p(off, C.c_int64))
        self.last_seconds = 0.0

    def close(self):
        if getattr(self, "h", None):
            getattr(self.lib, self.PREFIX + "bam_free")(self.h)
            self.h = None

    __del__ = close

    def _check_per_read(self):
        # the reference loops over the strings and indexes templid by the same counter
        # (src/rcpp_threshold_reads.cpp:27-33): nrow != nstr is out of contract (and out of bounds)
        if self.n != self.nstr:
            raise ValueError("nrow(bam) must equal the number of seqxm strings")

    def rcpp_threshold_reads(self, ctx_meth, ctx_unmeth, ooctx_meth, ooctx_unmeth, min_n_ctx, min_ctx_meth_frac, max_ooctx_meth_frac):
        self._check_per_read()
        out = np.zeros(max(self.nstr, 1), np.uint8)
        self.last_seconds = getattr(self.lib, self.PREFIX + "threshold_reads")(
            self.h, ctx_meth.encode(), ctx_unmeth.encode(), ooctx_meth.encode(), ooctx_unmeth.encode(),
            C.c_uint(min_n_ctx & 0xFFFFFFFF), float(min_ctx_meth_frac), float(max_ooctx_meth_frac), _p(out, C.c_uint8))
        return out[:self.nstr].astype(bool)

    def rcpp_get_xm_beta(self, ctx_meth, ctx_unmeth):
        self._check_per_read()
        out = np.zeros(max(self.nstr, 1), np.float64)
        self.last_seconds = getattr(self.lib, self.PREFIX + "get_xm_beta")(self.h, ctx_meth.encode(), ctx_unmeth.encode(), _p(out, C.c_double))
        return out[:self.nstr]

    def rcpp_cx_report(self, pass_, ctx):
        """pass_: bool/int array of nrow (None = all TRUE).  Returns dict of six int32 columns."""
        if pass_ is None:
            pv = np.ones(max(self.n, 1), np.int32)
        else:
            pv = np.ascontiguousarray(np.asarray(pass_).astype(np.int32))
            if pv.shape[0] == 0:
                pv = np.ones(1, np.int32)
        sec = C.c_double(0)
        nrow = getattr(self.lib, self.PREFIX + "cx_report")(self.h, _p(pv, C.c_int), ctx.encode(), C.byref(sec))
        self.last_seconds = sec.value
        cols = [np.zeros(max(nrow, 1), np.int32) for _ in range(6)]
        getattr(self.lib, self.PREFIX + "cx_fetch")(self.h, *[_p(c, C.c_int) for c in cols])
        names = ("rname", "strand", "pos", "context", "meth", "unmeth")
        return {k: c[:nrow] for k, c in zip(names, cols)}

    # ---- lMHL report ----
    def rcpp_mhl_report(self, ctx, hmax, hmin, max_ooctx_meth_frac):
        """src/rcpp_mhl_report.cpp:46 -> dict rname, strand, pos, context, coverage (int32), length, lmhl (float64)."""
        sec = C.c_double(0)
        nrow = getattr(self.lib, self.PREFIX + "mhl_report")(self.h, ctx.encode(), int(hmax), int(hmin), float(max_ooctx_meth_frac), C.byref(sec))
        self.last_seconds = sec.value
        ic = [np.zeros(max(nrow, 1), np.int32) for _ in range(5)]
        dc = [np.zeros(max(nrow, 1), np.float64) for _ in range(2)]
        getattr(self.lib, self.PREFIX + "mhl_fetch")(self.h, *[_p(c, C.c_int) for c in ic], *[_p(c, C.c_double) for c in dc])
        names = ("rname", "strand", "pos", "context", "coverage", "length", "lmhl")
        return {k: c[:nrow] for k, c in zip(names, ic + dc)}

    def generateMhlReport(self, haplotype_context="CG", max_haplotype_window=0, min_haplotype_length=0, max_outofcontext_beta=0.1):
        """R/generateMhlReport.R:170-197: ctx = ctx.meth + ctx.unmeth letters of the haplotype context."""
        b = CONTEXT_TO_BASES[haplotype_context]
        return self.rcpp_mhl_report(b[0] + b[1], max_haplotype_window, min_haplotype_length, max_outofcontext_beta)

    # ---- BED-assisted report ----
    def _match(self, mode, bed, param):
        """bed: (seqnames, start, end) int arrays, seqnames = factor code on the BAM's rname levels (NA = INT_MIN)."""
        chr_, st, en = (np.ascontiguousarray(a, dtype=np.int32) for a in bed)
        nbed = int(chr_.shape[0])
        pad = lambda a: a if a.shape[0] else np.zeros(1, np.int32)
        out = np.zeros(max(self.n, 1), np.int32)
        self.last_seconds = getattr(self.lib, self.PREFIX + "match_target")(self.h, mode, nbed, _p(pad(chr_), C.c_int), _p(pad(st), C.c_int),
                                                                            _p(pad(en), C.c_int), int(param), _p(out, C.c_int))
        return out[:self.n]

    def rcpp_match_amplicon(self, bed, tolerance):
        """src/rcpp_match_target.cpp:16 -> 1-based BED row per read, INT_MIN = NA."""
        return self._match(0, bed, tolerance)

    def rcpp_match_capture(self, bed, min_overlap):
        """src/rcpp_match_target.cpp:51."""
        return self._match(1, bed, min_overlap)

    def getBedReport(self, pass_, match, nbed):
        """R/internal.R:529-561 restated with numpy: per BED row (file order) and, when any read is unmatched, one NA row
        at the end: nreads+, nreads-, VEF = (TRUE+ + TRUE-) / all.  BED rows without reads are NaN (the merge's NA)."""
        return bed_report_table(self.strand, pass_, match, nbed)

    def generateBedReport(self, bed, bed_type="amplicon", match_tolerance=1, match_min_overlap=1, threshold_reads=True,
                          threshold_context="CG", min_context_sites=2, min_context_beta=0.5, max_outofcontext_beta=0.1):
        """R/generateBedReport.R:219-273."""
        if threshold_reads:
            b = CONTEXT_TO_BASES[threshold_context]
            pass_ = self.rcpp_threshold_reads(b[0], b[1], b[2], b[3], min_context_sites, min_context_beta, max_outofcontext_beta)
        else:
            pass_ = np.ones(self.n, bool)
        match = self.rcpp_match_amplicon(bed, match_tolerance) if bed_type == "amplicon" else self.rcpp_match_capture(bed, match_min_overlap)
        rep = self.getBedReport(pass_, match, int(np.asarray(bed[0]).shape[0]))
        if not threshold_reads:
            rep["VEF"] = np.full(rep["VEF"].shape, np.nan)
        return match, rep

    # ---- VCF-assisted report ----
    def rcpp_get_base_freqs(self, pass_, vcf):
        """src/rcpp_get_base_freqs.cpp:15 -> float64 matrix (n_vcf, 20): U+ACGTN, U-ACGTN, M+ACGTN, M-ACGTN.
        vcf: (seqnames, start) int arrays sorted by (seqnames, start), seqnames = factor code on the BAM's rname
        levels (NA = INT_MIN)."""
        chr_, pos = (np.ascontiguousarray(a, dtype=np.int32) for a in vcf)
        nv = int(chr_.shape[0])
        pv = np.ascontiguousarray(np.asarray(pass_).astype(bool).astype(np.uint8))
        if pv.shape[0] != self.n:
            raise ValueError("pass must have one entry per read")
        pad = lambda a, t: a if a.shape[0] else np.zeros(1, t)
        out = np.zeros((20, max(nv, 1)), np.float64)  # column-major nv x 20
        self.last_seconds = getattr(self.lib, self.PREFIX + "base_freqs")(self.h, _p(pad(pv, np.uint8), C.c_uint8), nv, _p(pad(chr_, np.int32), C.c_int),
                                                                          _p(pad(pos, np.int32), C.c_int), _p(out, C.c_double))
        return np.ascontiguousarray(out[:, :nv].T)

    def rcpp_fep(self, a, b, c, d):
        """src/rcpp_fep.cpp:9: two-sided Fisher exact p per row of four int columns (INT_MIN = NA -> NaN)."""
        cols = [np.ascontiguousarray(x, dtype=np.int32) for x in (a, b, c, d)]
        n = int(cols[0].shape[0])
        out = np.zeros(max(n, 1), np.float64)
        if n:
            getattr(self.lib, self.PREFIX + "fep")(n, *[_p(x, C.c_int) for x in cols], _p(out, C.c_double))
        return out[:n]

    def generateVcfReport(self, vcf, threshold_reads=True, threshold_context="CG", min_context_sites=2, min_context_beta=0.5,
                          max_outofcontext_beta=0.1):
        """R/generateVcfReport.R:196-231 + .getBaseFreqReport (R/internal.R:611-675).  vcf: dict with seqnames (factor
        codes), start, REF, ALT (single letters), already expanded / sorted / filtered (oracle.bamdata.read_vcf)."""
        if threshold_reads:
            b = CONTEXT_TO_BASES[threshold_context]
            pass_ = self.rcpp_threshold_reads(b[0], b[1], b[2], b[3], min_context_sites, min_context_beta, max_outofcontext_beta)
        else:
            pass_ = np.ones(self.n, bool)
        freqs = self.rcpp_get_base_freqs(pass_, (vcf["seqnames"], vcf["start"]))
        rep = base_freq_table(freqs, vcf["REF"], vcf["ALT"])
        na = lambda x: np.where(np.isnan(x), NA_INT, np.nan_to_num(x)).astype(np.int32)
        rep["FEp+"] = self.rcpp_fep(na(rep["M+Ref"]), na(rep["U+Ref"]), na(rep["M+Alt"]), na(rep["U+Alt"]))
        rep["FEp-"] = self.rcpp_fep(na(rep["M-Ref"]), na(rep["U-Ref"]), na(rep["M-Alt"]), na(rep["U-Alt"]))
        return freqs, rep

    # reference-level glue (R/generateCytosineReport.R:164-208, R/internal.R:439-456,486-500)
    def generateCytosineReport(self, threshold_reads=True, threshold_context="CG", min_context_sites=2,
                               min_context_beta=0.5, max_outofcontext_beta=0.1, report_context=None):
        report_context = report_context or threshold_context
        seconds = 0.0
        if threshold_reads:
            b = CONTEXT_TO_BASES[threshold_context]
            pass_ = self.rcpp_threshold_reads(b[0], b[1], b[2], b[3], min_context_sites, min_context_beta, max_outofcontext_beta)
            seconds += self.last_seconds
        else:
            pass_ = np.ones(self.n, bool)
        rep = self.rcpp_cx_report(pass_, CONTEXT_TO_BASES[report_context][0])
        self.last_seconds += seconds
        return pass_, rep


class RefCpu(_Cpu):
    PREFIX = "ref_"
    SO = REF_SO
    KIND = "ref"


class PortCpu(_Cpu):
    PREFIX = "port_"
    SO = PORT_SO
    KIND = "port"


# ---- the synthetic workload's rows, regenerated on the host WITHOUT the product library ------------------------------
SYNTH_SO = os.path.join(HERE, "_build", "libsynth.so")
_SYNTH_MAX_CHR = 32


class SynthCfg(C.Structure):
    """ea_synth_cfg of include/epialleler_b200.h (a plain C struct; restated so that nothing of the product is imported)."""
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_int64), ("n_chr", C.c_int32), ("len_min", C.c_int32),
                ("len_max", C.c_int32), ("n_islands", C.c_int32), ("island_len", C.c_int32), ("reserved", C.c_int32),
                ("chr_len", C.c_int64 * _SYNTH_MAX_CHR)]


class HostRows:
    """What the checkers take: int32 columns + the file-order strings as (blob, off)."""

    def __init__(self, rname, strand, start, blob, off):
        self.rname, self.strand, self.start = rname, strand, start
        self.templid = np.arange(rname.shape[0], dtype=np.int32)
        self._blob, self._off = blob, off
        self.nrow = int(rname.shape[0])

    def blob(self):
        return self._blob, self._off


def synth_rows(seed, n_reads, chr_len, len_min, len_max=None, n_islands=0, island_len=0, row_begin=0, row_end=None) -> HostRows:
    """Rows [row_begin, row_end) of the synthetic data set from oracle/_build/libsynth.so, the host build
    of the repo's generator header.  Bit-identical to ea_synth_host / ea_synth_device (tests/test_oracle.py)."""
    if not os.path.exists(SYNTH_SO):
        build(("port",))
    L = C.CDLL(SYNTH_SO)
    L.oracle_synth_host.restype = C.c_int
    L.oracle_synth_host.argtypes = [C.POINTER(SynthCfg), C.c_int64, C.c_int64] + [C.POINTER(C.c_int32)] * 3 + [C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]
    cfg = SynthCfg()
    cfg.seed, cfg.n_reads, cfg.n_chr, cfg.len_min = seed, n_reads, len(chr_len), len_min
    cfg.len_max = len_max if len_max is not None else len_min
    cfg.n_islands, cfg.island_len = n_islands, island_len
    for i, v in enumerate(chr_len):
        cfg.chr_len[i] = v
    row_end = n_reads if row_end is None else row_end
    n = row_end - row_begin
    rn, st, sa = (np.zeros(max(n, 1), np.int32) for _ in range(3))
    off = np.zeros(n + 1, np.int64)
    if L.oracle_synth_host(C.byref(cfg), row_begin, row_end, _p(rn, C.c_int32), _p(st, C.c_int32), _p(sa, C.c_int32), _p(off, C.c_int64), None):
        raise ValueError("bad synthetic configuration / row range")
    blob = np.zeros(max(int(off[-1]), 1), np.uint8)
    L.oracle_synth_host(C.byref(cfg), row_begin, row_end, None, None, None, _p(off, C.c_int64), _p(blob, C.c_uint8))
    return HostRows(rn[:n], st[:n], sa[:n], blob[:int(off[-1])], off)
