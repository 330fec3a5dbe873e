"""Host-side mirror of the reference's R/Rcpp interface for the cytosine-report path.

Names, argument order/meaning and error behaviour follow the reference (file:line cited per
function); the bodies only marshal numpy arrays across the C ABI (epialleler_b200/_native.py).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
import gzip as _gzip
from typing import Optional, Sequence

import numpy as np

from . import _native as N

# R/internal.R:54-65 -- context name -> (ctx.meth, ctx.unmeth, ooctx.meth, ooctx.unmeth)
CONTEXT_TO_BASES = {
    "CG": ("Z", "z", "XH", "xh"),
    "CHG": ("X", "x", "ZH", "zh"),
    "CHH": ("H", "h", "ZX", "zx"),
    "CxG": ("ZX", "zx", "H", "h"),
    "CX": ("ZXH", "zxh", "", ""),
}
CONTEXT_LEVELS = ("NA1", "CHH", "NA3", "NA4", "NA5", "CHG", "CG")  # src/rcpp_cx_report.cpp:150-152
STRAND_LEVELS = ("+", "-")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class BamData:
    """The reference's ``bam.data``: four int columns in (rname,start) order plus the packed SEQ|XM
    strings in file order (src/rcpp_read_bam.cpp:166-189, R/internal.R:193-195).  The device-resident
    packed copy is created on first use and kept on the object, like the reference keeps its
    ``seqxm_xptr`` external pointer on the data.table (pack once, many reports)."""

    def __init__(self, rname, strand, start, templid, seqxm_blob, seqxm_off, rname_levels: Sequence[str] = (), device: int = 0):
        self.rname = np.ascontiguousarray(rname, dtype=np.int32)
        self.strand = np.ascontiguousarray(strand, dtype=np.int32)
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.templid = np.ascontiguousarray(templid, dtype=np.int32)
        self._blob = np.ascontiguousarray(seqxm_blob, dtype=np.uint8)
        self._off = np.ascontiguousarray(seqxm_off, dtype=np.int64)
        self.rname_levels = list(rname_levels)
        self.device = device
        self._h = None
        self._has_bases = True
        self._owned = None
        n = self.rname.shape[0]
        if not (self.strand.shape[0] == self.start.shape[0] == self.templid.shape[0] == n):
            raise ValueError("rname/strand/start/templid must have the same length")

    @classmethod
    def from_strings(cls, rname, strand, start, templid, seqxm: Sequence[bytes], rname_levels=(), device=0):
        lens = np.fromiter((len(s) for s in seqxm), dtype=np.int64, count=len(seqxm))
        off = np.zeros(len(seqxm) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        blob = np.frombuffer(b"".join(seqxm), dtype=np.uint8) if len(seqxm) else np.zeros(0, np.uint8)
        return cls(rname, strand, start, templid, blob, off, rname_levels, device)

    @property
    def nrow(self) -> int:
        return int(self.rname.shape[0])

    @property
    def nstr(self) -> int:
        return int(self._off.shape[0] - 1)

    def blob(self):
        return self._blob, self._off

    # -- device handle ------------------------------------------------------------------------
    def handle(self, need_bases: bool = True):
        """The packed device copy.  need_bases=False (the cytosine-report path: thresholds, betas, reports, BED matching) lets
        the library pack the call codes only (ea_pack_blob_calls: the host transposes the strings into the bit-sliced call
        planes while they cross PCIe, half the bytes); a later call that needs the byte rows (VCF, lMHL reports) repacks."""
        L = N.lib()
        if self._h is not None and need_bases and not self._has_bases:
            self.close()
        if self._h is None:
            h = C.c_void_p()
            blob = self._blob if self._blob.shape[0] else np.zeros(1, np.uint8)
            calls_only = (not need_bases and os.environ.get("EA_HOST_PACK", "1") != "0" and L.ea_set_kernel_policy(-1) == 0
                          and self.nrow == self.nstr and self.nrow > 0)
            pack = L.ea_pack_blob_calls if calls_only else L.ea_pack_blob
            N.check(pack(_p(self.rname, C.c_int32), _p(self.strand, C.c_int32), _p(self.start, C.c_int32),
                         _p(self.templid, C.c_int32), self.nrow, _p(blob, C.c_uint8), _p(self._off, C.c_int64),
                         self.nstr, self.device, C.byref(h)))
            self._h = h
            self._has_bases = bool(L.ea_has_bytes(h))
            if self._owned is not None:
                N.check(L.ea_set_owned_range(h, *self._owned))
        return self._h

    def set_owned(self, key_lo, key_hi, own_row_lo):
        """These rows are one range shard (epialleler_b200/shard.py): reports keep the rows with key rname<<32|pos in
        [key_lo, key_hi) - filtered on the device - and pass flags are returned for rows [own_row_lo, nrow)."""
        self._owned = (int(key_lo), int(key_hi), int(own_row_lo))
        if self._h is not None:
            N.check(N.lib().ea_set_owned_range(self._h, *self._owned))

    @property
    def n_owned(self) -> int:
        return self.nrow - (self._owned[2] if self._owned else 0)

    def close(self):
        if self._h is not None:
            N.lib().ea_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBam:
    """A bam.data that only exists on the device (synthetic reads); same calls as BamData."""

    def __init__(self, h, device):
        self._h = h
        self.device = device
        self.rname_levels = []
        self._owned = None

    def handle(self, need_bases: bool = True):
        return self._h

    def set_owned(self, key_lo, key_hi, own_row_lo):
        self._owned = (int(key_lo), int(key_hi), int(own_row_lo))
        N.check(N.lib().ea_set_owned_range(self._h, *self._owned))

    @property
    def n_owned(self):
        return self.nrow - (self._owned[2] if self._owned else 0)

    @property
    def nrow(self):
        return int(N.lib().ea_nrow(self._h))

    @property
    def nstr(self):
        return self.nrow

    @property
    def nbytes(self):
        return int(N.lib().ea_nbytes(self._h))

    def to_host(self) -> BamData:
        L = N.lib()
        n = self.nrow
        rn, st, sa = (np.zeros(max(n, 1), np.int32) for _ in range(3))
        off = np.zeros(n + 1, np.int64)
        blob = np.zeros(max(self.nbytes, 1), np.uint8)
        N.check(L.ea_export_host(self._h, _p(rn, C.c_int32), _p(st, C.c_int32), _p(sa, C.c_int32), _p(off, C.c_int64), _p(blob, C.c_uint8)))
        return BamData(rn[:n], st[:n], sa[:n], np.arange(n, dtype=np.int32), blob[:self.nbytes], off, device=self.device)

    def close(self):
        if self._h is not None:
            N.lib().ea_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def preprocessBam(bam_file, **kwargs):
    """R/preprocessBam.R:197-237.  Already preprocessed reads pass through unchanged; decoding a BAM
    file stays on the host with HTSlib (BASELINE.json north_star) and is not part of this package."""
    if isinstance(bam_file, (BamData, DeviceBam, MultiBam)):
        return bam_file
    raise NotImplementedError("BAM decoding (HTSlib) is outside the B200 hot path: pass preprocessed reads "
                              "(BamData) as the reference allows (R/preprocessBam.R:226-235)")


def _on_calls(df, fn):
    """Run fn(handle) on a handle that may hold the call planes only; a request the planes cannot serve (unusual letter
    sets, more than 65535 rows over a window, a non-default kernel policy) makes the library answer EA_ERR_NEEDS_BYTES:
    pack again with the byte rows and repeat."""
    try:
        return fn(df.handle(need_bases=False))
    except N.EaError as e:
        if e.code != N.EA_ERR_NEEDS_BYTES:
            raise
        return fn(df.handle(need_bases=True))


def _check_nrow(df):
    # the reference sizes its result by the number of strings and indexes templid by row
    # (src/rcpp_threshold_reads.cpp:27-33): anything else is out of contract
    if df.nrow != df.nstr:
        raise ValueError("nrow(bam) must equal the number of seqxm strings")


def rcpp_threshold_reads(df, ctx_meth, ctx_unmeth, ooctx_meth, ooctx_unmeth, min_n_ctx, min_ctx_meth_frac, max_ooctx_meth_frac):
    """src/rcpp_threshold_reads.cpp:15-74 -> logical vector, TRUE for reads passing the thresholds."""
    _check_nrow(df)
    L = N.lib()
    out = np.zeros(max(df.nrow, 1), np.uint8)
    _on_calls(df, lambda h: N.check(L.ea_threshold_reads(h, ctx_meth.encode(), ctx_unmeth.encode(), ooctx_meth.encode(), ooctx_unmeth.encode(),
                                                         int(min_n_ctx) & 0xFFFFFFFF, float(min_ctx_meth_frac), float(max_ooctx_meth_frac), _p(out, C.c_uint8))))
    return out[:df.nrow].astype(bool)


def rcpp_get_xm_beta(df, ctx_meth, ctx_unmeth):
    """src/rcpp_get_xm_beta.cpp:10-43 -> per-read average beta (double)."""
    _check_nrow(df)
    L = N.lib()
    out = np.zeros(max(df.nrow, 1), np.float64)
    _on_calls(df, lambda h: N.check(L.ea_get_xm_beta(h, ctx_meth.encode(), ctx_unmeth.encode(), _p(out, C.c_double))))
    return out[:df.nrow]


_PINNED_MIN_BYTES = 1 << 20


def _host_out(n, dtype):
    """Result buffer for a device->host copy.  Large results live in page-locked blocks from the library's cache
    (ea_host_alloc): the copy runs at PCIe speed instead of through the driver's pageable staging, and the block goes
    back to the cache when the last numpy view of it is garbage collected.  Small results are plain numpy arrays."""
    n = max(int(n), 1)
    dt = np.dtype(dtype)
    nbytes = n * dt.itemsize
    if nbytes < _PINNED_MIN_BYTES:
        return np.empty(n, dt)
    L = N.lib()
    p = L.ea_host_alloc(nbytes)
    if not p:
        return np.empty(n, dt)
    buf = (C.c_uint8 * nbytes).from_address(p)
    weakref.finalize(buf, L.ea_host_free, p)
    return np.frombuffer(buf, dtype=dt, count=n)


def _fetch_report(df, nrow):
    L = N.lib()
    cols = [_host_out(nrow, np.int32) for _ in range(6)]
    N.check(L.ea_cx_report_fetch(df.handle(need_bases=False), *[_p(c, C.c_int32) for c in cols]))
    names = ("rname", "strand", "pos", "context", "meth", "unmeth")
    rep = {k: c[:nrow] for k, c in zip(names, cols)}
    rep["levels"] = {"rname": list(df.rname_levels), "strand": list(STRAND_LEVELS), "context": list(CONTEXT_LEVELS)}
    return rep


def rcpp_cx_report(df, pass_, ctx):
    """src/rcpp_cx_report.cpp:34-159 -> six integer columns (rname, strand, pos, context, meth, unmeth),
    factor levels under "levels".  pass_: logical vector of nrow (None/NA = TRUE)."""
    L = N.lib()
    nrow = C.c_int64(0)
    if pass_ is None:
        pp = None
    else:
        pv = np.ascontiguousarray(np.asarray(pass_) != 0, dtype=np.uint8)
        if pv.shape[0] != df.nrow:
            raise ValueError("length(pass) must equal nrow(bam)")
        pp = _p(pv if pv.shape[0] else np.zeros(1, np.uint8), C.c_uint8)
    _on_calls(df, lambda h: N.check(L.ea_cx_report(h, pp, ctx.encode(), C.byref(nrow))))
    return _fetch_report(df, nrow.value)


def _match_arg(value, choices, name):
    if isinstance(value, (list, tuple)):
        value = value[0]
    if value not in choices:
        raise ValueError("'%s' should be one of %s" % (name, ", ".join('"%s"' % c for c in choices)))
    return value


def _params(threshold_reads, threshold_context, min_context_sites, min_context_beta, max_outofcontext_beta, report_context):
    b = CONTEXT_TO_BASES[threshold_context]
    keep = [s.encode() for s in b] + [CONTEXT_TO_BASES[report_context][0].encode()]
    p = N.ReportParams(1 if threshold_reads else 0, keep[0], keep[1], keep[2], keep[3], int(min_context_sites) & 0xFFFFFFFF,
                       float(min_context_beta), float(max_outofcontext_beta), keep[4])
    p._keep = keep
    return p


def generateCytosineReport(bam, report_file=None, threshold_reads=True, threshold_context="CG", min_context_sites=2,
                           min_context_beta=0.5, max_outofcontext_beta=0.1, report_context=None, gzip=False,
                           verbose=False, return_pass=False, **kwargs):
    """R/generateCytosineReport.R:164-208.  Thresholding and aggregation run as ONE fused pass over the
    XM bytes (ea_cytosine_report).  Returns the report (dict of columns), or writes a TSV like the
    reference's .writeReport (R/internal.R:274-287) when report_file is given."""
    contexts = tuple(CONTEXT_TO_BASES)
    threshold_context = _match_arg(threshold_context, contexts, "threshold.context")
    report_context = _match_arg(report_context if report_context is not None else threshold_context, contexts, "report.context")
    bam = preprocessBam(bam, **kwargs)
    L = N.lib()
    p = _params(threshold_reads, threshold_context, min_context_sites, min_context_beta, max_outofcontext_beta, report_context)
    nrow = C.c_int64(0)
    if isinstance(bam, MultiBam):
        return bam._report(p, (report_file, gzip) if report_file is not None else None, return_pass)
    n_flags = getattr(bam, "n_owned", bam.nrow)   # a range shard reports the flags of the rows it owns
    pass_out = _host_out(n_flags, np.uint8) if return_pass else None
    _on_calls(bam, lambda h: N.check(L.ea_cytosine_report(h, C.byref(p), _p(pass_out, C.c_uint8) if return_pass else None, C.byref(nrow))))
    if report_file is not None:  # formatted on the device, the rows never become Python objects
        _write_report_text(report_text(bam), report_file, gzip)
        return None
    rep = _fetch_report(bam, nrow.value)
    if return_pass:
        return pass_out[:n_flags].view(np.bool_), rep  # the library writes 0/1 bytes
    return rep


class MultiBam:
    """One process, several GPUs (ea_pack_multi): the host `bam` is cut into contiguous genomic ranges with halo rows,
    shard i is packed on devices[i] by its own host thread, generateCytosineReport runs on all devices at once and the
    owned rows come back concatenated in genomic order - the result equals the single-GPU report bit for bit."""

    def __init__(self, bam: BamData, devices: Sequence[int], calls_only: bool = True):
        L = N.lib()
        self.bam = bam
        self.rname_levels = bam.rname_levels
        self.devices = list(devices)
        dv = (C.c_int * len(self.devices))(*self.devices)
        m = C.c_void_p()
        blob, off = bam.blob()
        blob = blob if blob.shape[0] else np.zeros(1, np.uint8)
        N.check(L.ea_pack_multi(_p(bam.rname, C.c_int32), _p(bam.strand, C.c_int32), _p(bam.start, C.c_int32), _p(bam.templid, C.c_int32),
                                bam.nrow, _p(blob, C.c_uint8), _p(off, C.c_int64), bam.nstr, dv, len(self.devices), 1 if calls_only else 0, C.byref(m)))
        self._m = m

    @classmethod
    def from_handle(cls, m, nrow, devices, rname_levels=()):
        """Wraps an ea_multi* packed elsewhere (ea_pack_multi_records through the shim / integration/harness)."""
        self = cls.__new__(cls)
        self.bam, self._nrow, self.rname_levels, self.devices, self._m = None, int(nrow), list(rname_levels), list(devices), m
        return self

    @property
    def nrow(self):
        return self.bam.nrow if self.bam is not None else self._nrow

    def shards(self):
        L = N.lib()
        out = []
        for i in range(L.ea_multi_ndev(self._m)):
            s = N.Shard()
            N.check(L.ea_multi_shard(self._m, i, C.byref(s)))
            out.append(s)
        return out

    def _report(self, p, report_file, return_pass):
        L = N.lib()
        nrow = C.c_int64(0)
        pass_out = _host_out(self.nrow, np.uint8) if return_pass else None
        N.check(L.ea_multi_cytosine_report(self._m, C.byref(p), _p(pass_out, C.c_uint8) if return_pass else None, C.byref(nrow)))
        if report_file is not None:   # every shard formats the rows it owns on its device; the pieces are the file, in genomic order
            _write_report_text(self.report_text(), report_file[0], report_file[1])
            return None
        n = nrow.value
        cols = [_host_out(n, np.int32) for _ in range(6)]
        N.check(L.ea_multi_report_fetch(self._m, *[_p(c, C.c_int32) for c in cols]))
        rep = {k: c[:n] for k, c in zip(("rname", "strand", "pos", "context", "meth", "unmeth"), cols)}
        rep["levels"] = {"rname": list(self.rname_levels), "strand": list(STRAND_LEVELS), "context": list(CONTEXT_LEVELS)}
        if return_pass:
            return pass_out[:self.nrow].view(np.bool_), rep
        return rep

    def report_text(self, header=True):
        """The last report as the bytes .writeReport would put into the file: shard i's owned rows formatted on device i
        (ea_cx_report_text honours the ownership filter), concatenated in genomic order."""
        L = N.lib()
        names = [str(x).encode() for x in self.rname_levels]
        arr = (C.c_char_p * max(len(names), 1))(*names)
        parts = []
        for i in range(L.ea_multi_ndev(self._m)):
            h = L.ea_multi_handle(self._m, i)
            nbytes = C.c_int64(0)
            N.check(L.ea_cx_report_text(h, arr, len(names), 1 if (header and i == 0) else 0, C.byref(nbytes)))
            out = np.empty(max(nbytes.value, 1), np.uint8)
            if nbytes.value:
                N.check(L.ea_cx_report_text_fetch(h, out.ctypes.data_as(C.c_void_p)))
            parts.append(out[:nbytes.value])
        return np.concatenate(parts) if parts else np.zeros(0, np.uint8)

    def close(self):
        if self._m is not None:
            N.lib().ea_multi_free(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass



class Ingest:
    """Streaming ingestion (ea_ingest_*): what the reader loop (src/rcpp_read_bam.cpp:61-69, push_template) would call
    instead of seqxm->emplace_back.  append() takes one template in FILE order; finish() sorts the rows by (rname,start)
    on the device (stable, like R's setorder) and returns (DeviceBam, columns) - columns = the sorted bam.data columns
    rname/strand/start/templid the reference's data.table would hold."""

    def __init__(self, device=0, expected_templates=0):
        g = C.c_void_p()
        N.check(N.lib().ea_ingest_begin(device, expected_templates, C.byref(g)))
        self._g, self.device, self.n = g, device, 0

    def append(self, rname, strand, start, seqxm: bytes):
        buf = np.frombuffer(seqxm, dtype=np.uint8) if len(seqxm) else np.zeros(1, np.uint8)
        r = N.lib().ea_ingest_append(self._g, int(rname), int(strand), int(start), buf.ctypes.data, len(seqxm))
        if r < 0:
            N.check(int(r))
        self.n += 1
        return int(r)

    def finish(self):
        n = self.n
        cols = [np.zeros(max(n, 1), np.int32) for _ in range(4)]
        h = C.c_void_p()
        g, self._g = self._g, None
        N.check(N.lib().ea_ingest_finish(g, *[_p(c, C.c_int32) for c in cols], C.byref(h)))
        return DeviceBam(h, self.device), dict(zip(("rname", "strand", "start", "templid"), (c[:n] for c in cols)))

    def abort(self):
        if self._g is not None:
            N.lib().ea_ingest_abort(self._g)
            self._g = None

    __del__ = abort


# ---- lMHL report ---------------------------------------------------------------
def rcpp_mhl_report(df, ctx, hmax, hmin, max_ooctx_meth_frac):
    """src/rcpp_mhl_report.cpp:46-228 -> rname, strand, pos, context, coverage (int32), length, lmhl (float64)."""
    L = N.lib()
    nrow = C.c_int64(0)
    # haplotype contexts made of letter pairs run on the call planes; anything else makes the library ask for the byte rows
    _on_calls(df, lambda h: N.check(L.ea_mhl_report(h, ctx.encode(), int(hmax), int(hmin), float(max_ooctx_meth_frac), C.byref(nrow))))
    n = nrow.value
    ic = [_host_out(n, np.int32) for _ in range(5)]
    dc = [_host_out(n, np.float64) for _ in range(2)]
    N.check(L.ea_mhl_report_fetch(df.handle(need_bases=False), *[_p(c, C.c_int32) for c in ic], *[_p(c, C.c_double) for c in dc]))
    rep = {k: c[:n] for k, c in zip(("rname", "strand", "pos", "context", "coverage", "length", "lmhl"), ic + dc)}
    rep["levels"] = {"rname": list(df.rname_levels), "strand": list(STRAND_LEVELS), "context": list(CONTEXT_LEVELS)}
    return rep


def generateMhlReport(bam, report_file=None, haplotype_context="CG", max_haplotype_window=0, min_haplotype_length=0,
                      max_outofcontext_beta=0.1, **kwargs):
    """R/generateMhlReport.R:170-197."""
    haplotype_context = _match_arg(haplotype_context, tuple(CONTEXT_TO_BASES), "haplotype.context")
    bam = preprocessBam(bam, **kwargs)
    b = CONTEXT_TO_BASES[haplotype_context]
    rep = rcpp_mhl_report(bam, b[0] + b[1], max_haplotype_window, min_haplotype_length, max_outofcontext_beta)
    if report_file is not None:
        lv = rep["levels"]
        with open(report_file, "w") as fh:
            fh.write("rname\tstrand\tpos\tcontext\tcoverage\tlength\tlmhl\n")
            for i in range(rep["pos"].shape[0]):
                rn = rep["rname"][i]
                name = lv["rname"][rn - 1] if 0 < rn <= len(lv["rname"]) else str(rn)
                fh.write("%s\t%s\t%d\t%s\t%d\t%r\t%r\n" % (name, STRAND_LEVELS[rep["strand"][i] - 1], rep["pos"][i], CONTEXT_LEVELS[rep["context"][i] - 1],
                                                              rep["coverage"][i], float(rep["length"][i]), float(rep["lmhl"][i])))
        return None
    return rep


# ---- BED-assisted reports ------------------------------------------------
NA_INT = -2147483648  # R's NA_integer_


def _bed_cols(bed):
    """bed: (seqnames, start, end[, ...]) - seqnames = factor codes on the BAM's rname levels (NA_INT = NA), as
    R/internal.R:467-468 hands them to the matchers; start/end 1-based closed."""
    c, s, e = (np.ascontiguousarray(a, dtype=np.int32) for a in bed[:3])
    if not (c.shape[0] == s.shape[0] == e.shape[0]):
        raise ValueError("BED columns must have the same length")
    return c, s, e


def _bedp(a):
    return _p(a if a.shape[0] else np.zeros(1, np.int32), C.c_int32)


def _match_target(df, mode, bed, param):
    c, s, e = _bed_cols(bed)
    out = np.empty(max(df.nrow, 1), np.int32)
    N.check(N.lib().ea_match_target(df.handle(need_bases=False), mode, _bedp(c), _bedp(s), _bedp(e), c.shape[0], int(param), _p(out, C.c_int32)))
    return out[:df.nrow]


def rcpp_match_amplicon(df, bed, tolerance):
    """src/rcpp_match_target.cpp:16-45 -> per read the 1-based BED row (first in file order) whose start is within
    `tolerance` of the read's start or whose end is within `tolerance` of the read's end; NA_INT if none."""
    return _match_target(df, 0, bed, tolerance)


def rcpp_match_capture(df, bed, min_overlap):
    """src/rcpp_match_target.cpp:51-81 -> first BED row overlapping the read by at least `min_overlap` bases."""
    return _match_target(df, 1, bed, min_overlap)


def _bed_table(counts, nbed):
    """R/internal.R:537-556: nreads+ / nreads- / VEF per BED row (+ one NA row when reads are unmatched); rows no read
    matched are NaN (NA after the all=TRUE merge).  counts[row] = {FALSE+, FALSE-, TRUE+, TRUE-}."""
    cnt = counts.reshape(nbed + 1, 2, 2)
    n_rows = nbed + (1 if cnt[nbed].sum() > 0 else 0)
    cnt = cnt[:n_rows]
    tot = cnt.sum(axis=(1, 2))
    plus = (cnt[:, 0, 0] + cnt[:, 1, 0]).astype(np.float64)
    minus = (cnt[:, 0, 1] + cnt[:, 1, 1]).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        vef = (cnt[:, 1, 0] + cnt[:, 1, 1]) / tot.astype(np.float64)
    empty = tot == 0
    plus[empty] = np.nan
    minus[empty] = np.nan
    vef[empty] = np.nan
    rows = np.concatenate([np.arange(1, nbed + 1), [NA_INT] if n_rows > nbed else []]).astype(np.int64)
    return {"bedrow": rows, "nreads+": plus, "nreads-": minus, "VEF": vef, "counts": cnt}


def generateBedReport(bam, bed, report_file=None, bed_type="amplicon", match_tolerance=1, match_min_overlap=1, threshold_reads=True,
                      threshold_context="CG", min_context_sites=2, min_context_beta=0.5, max_outofcontext_beta=0.1,
                      return_match=False, **kwargs):
    """R/generateBedReport.R:219-273: match reads to BED targets, threshold them, count reads per (target, strand) and
    the fraction of passing reads (VEF) - one device call (ea_bed_report)."""
    bed_type = _match_arg(bed_type, ("amplicon", "capture"), "bed.type")
    threshold_context = _match_arg(threshold_context, tuple(CONTEXT_TO_BASES), "threshold.context")
    bam = preprocessBam(bam, **kwargs)
    c, s, e = _bed_cols(bed)
    nbed = int(c.shape[0])
    p = _params(threshold_reads, threshold_context, min_context_sites, min_context_beta, max_outofcontext_beta, threshold_context)
    counts = np.zeros((nbed + 1) * 4, np.int64)
    match = np.empty(max(bam.nrow, 1), np.int32) if return_match else None
    _on_calls(bam, lambda h: N.check(N.lib().ea_bed_report(h, 0 if bed_type == "amplicon" else 1, _bedp(c), _bedp(s), _bedp(e), nbed,
                                                           int(match_tolerance if bed_type == "amplicon" else match_min_overlap), C.byref(p), None,
                                                           _p(match, C.c_int32) if return_match else None, None, _p(counts, C.c_int64))))
    rep = _bed_table(counts, nbed)
    if not threshold_reads:
        rep["VEF"] = np.full(rep["VEF"].shape, np.nan)  # R/generateBedReport.R:264
    if report_file is not None:
        with open(report_file, "w") as fh:
            fh.write("bedrow\tnreads+\tnreads-\tVEF\n")
            for i in range(rep["VEF"].shape[0]):
                fh.write("%s\t%s\t%s\t%s\n" % tuple("NA" if (isinstance(v, float) and np.isnan(v)) or v == NA_INT else ("%d" % v if k < 3 else repr(float(v)))
                                                   for k, v in enumerate((int(rep["bedrow"][i]), rep["nreads+"][i], rep["nreads-"][i], rep["VEF"][i]))))
        return None
    return (match[:bam.nrow], rep) if return_match else rep


def generateAmpliconReport(bam, bed, report_file=None, match_tolerance=1, **kw):
    """R/generateBedReport.R:187-199."""
    return generateBedReport(bam, bed, report_file, bed_type="amplicon", match_tolerance=match_tolerance, **kw)


def generateCaptureReport(bam, bed, report_file=None, match_min_overlap=1, **kw):
    """R/generateBedReport.R:203-215."""
    return generateBedReport(bam, bed, report_file, bed_type="capture", match_min_overlap=match_min_overlap, **kw)


def generateBedEcdf(bam, bed, bed_type="amplicon", bed_rows=(1,), match_tolerance=1, match_min_overlap=1, ecdf_context="CG", **kwargs):
    """R/generateBedEcdf.R + R/internal.R:568-604: per requested BED row the empirical CDFs of the per-read in-context
    and out-of-context betas (rcpp_get_xm_beta on the device).  bed_rows = None: every matched row (NA row last).
    Returns {row: {"context": sorted betas, "out.of.context": sorted betas}}; ecdf(x) = searchsorted(v, x, 'right') / len(v)."""
    bed_type = _match_arg(bed_type, ("amplicon", "capture"), "bed.type")
    ecdf_context = _match_arg(ecdf_context, tuple(CONTEXT_TO_BASES), "ecdf.context")
    bam = preprocessBam(bam, **kwargs)
    match = rcpp_match_amplicon(bam, bed, match_tolerance) if bed_type == "amplicon" else rcpp_match_capture(bam, bed, match_min_overlap)
    b = CONTEXT_TO_BASES[ecdf_context]
    ctx_beta = rcpp_get_xm_beta(bam, b[0], b[1])
    oo_beta = rcpp_get_xm_beta(bam, b[2], b[3])
    present = sorted(set(int(m) for m in np.unique(match) if m != NA_INT)) + ([NA_INT] if (match == NA_INT).any() else [])
    rows = present if bed_rows is None else [r for r in bed_rows if r in present]
    out = {}
    for r in rows:
        sel = match == r
        out[r] = {"context": np.sort(ctx_beta[sel]), "out.of.context": np.sort(oo_beta[sel])}
    return out


def ecdf_at(sorted_values, x):
    """stats::ecdf(v)(x)."""
    return float(np.searchsorted(sorted_values, x, side="right")) / float(sorted_values.shape[0])


# ---- VCF-assisted report -------------------------------------------------------------
def rcpp_get_base_freqs(df, pass_, vcf):
    """src/rcpp_get_base_freqs.cpp:15-61 -> float64 matrix (n_vcf, 20) with the reference's column order
    U+A U+C U+G U+T U+N, U-..., M+..., M-...  vcf: (seqnames, start) or a dict with those keys; seqnames = factor codes
    on the BAM's rname levels (NA_INT = NA), entries sorted by (seqnames, start).  pass_: logical vector of nrow
    (None = all TRUE)."""
    chr_, pos = (vcf["seqnames"], vcf["start"]) if isinstance(vcf, dict) else vcf
    chr_ = np.ascontiguousarray(chr_, dtype=np.int32)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    nv = int(chr_.shape[0])
    if pos.shape[0] != nv:
        raise ValueError("seqnames/start must have the same length")
    pp = None
    if pass_ is not None:
        pp = np.ascontiguousarray(np.asarray(pass_) != 0, dtype=np.uint8)
        if pp.shape[0] != df.nrow:
            raise ValueError("pass must have nrow(bam) entries")
    out = _host_out(20 * max(nv, 1), np.float64).reshape(20, max(nv, 1))  # column-major n_vcf x 20 (page-locked when large)
    if nv:
        N.check(N.lib().ea_base_freqs(df.handle(), _p(pp, C.c_uint8) if pp is not None and pp.shape[0] else None,
                                      _p(chr_, C.c_int32), _p(pos, C.c_int32), nv, _p(out, C.c_double)))
    return np.ascontiguousarray(out[:, :nv].T)


def rcpp_fep(a, b, c, d):
    """src/rcpp_fep.cpp:9-36: two-sided Fisher's exact p per row; NA_INT in any cell -> NaN."""
    cols = [np.ascontiguousarray(x, dtype=np.int32) for x in (a, b, c, d)]
    n = int(cols[0].shape[0])
    out = np.zeros(max(n, 1), np.float64)
    if n:
        N.check(N.lib().ea_fisher_exact(*[_p(x, C.c_int32) for x in cols], n, _p(out, C.c_double)))
    return out[:n]


# R/internal.R:642-665 as data: (REF, ALT) -> base columns summed into Ref / Alt on the '+' strand and on the '-' strand;
# None = NA (bisulfite conversion makes the alleles indistinguishable on that strand: C>T on '+', G>A on '-')
BASE_FREQ_RULES = {
    ("A", "C"): ("A", "CT", "A", "C"), ("A", "T"): ("A", "T", "A", "T"), ("A", "G"): ("A", "G", None, None),
    ("C", "A"): ("CT", "A", "C", "A"), ("C", "T"): (None, None, "C", "T"), ("C", "G"): ("CT", "G", "C", "AG"),
    ("T", "A"): ("T", "A", "T", "A"), ("T", "C"): (None, None, "T", "C"), ("T", "G"): ("T", "G", "T", "AG"),
    ("G", "A"): ("G", "A", None, None), ("G", "C"): ("G", "CT", "AG", "C"), ("G", "T"): ("G", "T", "AG", "T"),
}
_FREQ_BLOCK = {"U+": 0, "U-": 5, "M+": 10, "M-": 15}
_FREQ_BASE = {"A": 0, "C": 1, "G": 2, "T": 3}


def _base_freq_report(freqs, ref, alt):
    """.getBaseFreqReport (R/internal.R:636-672) after rcpp_get_base_freqs: Ref/Alt counts per strand and methylation
    status (NaN = NA), SumRef / SumAlt (rowSums, na.rm=TRUE), FEp+ / FEp- (Fisher's exact test on {M Ref, U Ref; M Alt, U Alt})."""
    n = freqs.shape[0]
    names = ("M+Ref", "U+Ref", "M-Ref", "U-Ref", "M+Alt", "U+Alt", "M-Alt", "U-Alt")
    rep = {k: np.full(n, np.nan) for k in names}
    ref = np.asarray(ref, dtype=object); alt = np.asarray(alt, dtype=object)
    for (r, a), rule in BASE_FREQ_RULES.items():
        rows = np.nonzero((ref == r) & (alt == a))[0]
        if rows.shape[0] == 0:
            continue
        for strand, (rb, ab) in (("+", rule[0:2]), ("-", rule[2:4])):
            if rb is None:
                continue
            for m in "MU":
                blk = _FREQ_BLOCK[m + strand]
                rep[m + strand + "Ref"][rows] = sum(freqs[rows, blk + _FREQ_BASE[x]] for x in rb)
                rep[m + strand + "Alt"][rows] = sum(freqs[rows, blk + _FREQ_BASE[x]] for x in ab)
    rep["SumRef"] = np.nansum([rep[k] for k in ("M+Ref", "U+Ref", "M-Ref", "U-Ref")], axis=0)
    rep["SumAlt"] = np.nansum([rep[k] for k in ("M+Alt", "U+Alt", "M-Alt", "U-Alt")], axis=0)
    as_int = lambda x: np.where(np.isnan(x), NA_INT, np.nan_to_num(x)).astype(np.int32)
    rep["FEp+"] = rcpp_fep(*[as_int(rep[k]) for k in ("M+Ref", "U+Ref", "M+Alt", "U+Alt")])
    rep["FEp-"] = rcpp_fep(*[as_int(rep[k]) for k in ("M-Ref", "U-Ref", "M-Alt", "U-Alt")])
    return rep


def _ucsc_name(name):
    if name.startswith("chr"):
        return name
    return "chrM" if name in ("MT", "M") else "chr" + name


def read_vcf(vcf_file, rname_levels, bed=None, vcf_style=None, zero_based_bed=False):
    """.readVcf + the filters of .getBaseFreqReport (R/internal.R:230-267,617-625) without VariantAnnotation: records of a
    (gzipped) VCF - only those overlapping the reduced BED ranges when `bed` (a BED file) is given - expanded to one row
    per ALT allele, sorted by (chromosome, start, end), then REF of width 1 and single-letter ALT only.  vcf_style="NCBI":
    the VCF names chromosomes 1..22,X,Y,MT and they are renamed to the BED/BAM style (chr1..).  seqnames are factor codes on
    `rname_levels` (NA_INT when absent).  Host-side parsing, like the reference."""
    levels = {str(nm): i + 1 for i, nm in enumerate(rname_levels)}
    conv = _ucsc_name if (vcf_style or "").upper() == "NCBI" else (lambda x: x)
    ranges = None
    if bed is not None:
        ranges = {}
        with open(bed) as fh:
            fh.readline()
            for ln in fh:
                f = ln.rstrip("\n").split("\t")
                if len(f) >= 3 and f[0]:
                    ranges.setdefault(f[0], []).append([int(f[1]) + (1 if zero_based_bed else 0), int(f[2])])
        for k, v in ranges.items():  # GenomicRanges::reduce
            v.sort()
            merged = [v[0]]
            for a, b in v[1:]:
                if a <= merged[-1][1] + 1:
                    merged[-1][1] = max(merged[-1][1], b)
                else:
                    merged.append([a, b])
            ranges[k] = merged
    rows, order = [], {}
    opener = _gzip.open if str(vcf_file).endswith(".gz") else open
    with opener(vcf_file, "rt") as fh:
        for ln in fh:
            if ln.startswith("#"):
                continue
            f = ln.rstrip("\n").split("\t")
            if len(f) < 5:
                continue
            chrom, pos, vid, ref = conv(f[0]), int(f[1]), f[2], f[3]
            end = pos + len(ref) - 1
            if ranges is not None and not any(a <= end and pos <= b for a, b in ranges.get(chrom, ())):
                continue
            order.setdefault(chrom, len(order))
            for alt in f[4].split(","):
                rows.append((order[chrom], pos, end, chrom, ref, alt, vid if vid != "." else "%s:%d_%s/%s" % (chrom, pos, ref, alt)))
    rows.sort(key=lambda r: (r[0], r[1], r[2]))
    rows = [r for r in rows if len(r[4]) == 1 and len(r[5]) == 1]
    return {"seqnames": np.asarray([levels.get(r[3], NA_INT) for r in rows], np.int32), "start": np.asarray([r[1] for r in rows], np.int32),
            "REF": [r[4] for r in rows], "ALT": [r[5] for r in rows], "name": [r[6] for r in rows], "chrom": [r[3] for r in rows]}


def generateVcfReport(bam, vcf, vcf_style=None, bed=None, report_file=None, zero_based_bed=False, threshold_reads=True,
                      threshold_context="CG", min_context_sites=2, min_context_beta=0.5, max_outofcontext_beta=0.1,
                      gzip=False, **kwargs):
    """R/generateVcfReport.R:169-231: per SNV of the VCF the reads carrying the reference / alternative base among the
    passing ("M") and failing ("U") reads of each strand, and Fisher's exact p per strand.  vcf: a VCF file (read_vcf) or
    the dict read_vcf returns.  Returns a dict of columns: name, seqnames, range, REF, ALT, the eight Ref/Alt counts,
    SumRef, SumAlt, FEp+, FEp- (NaN = NA)."""
    threshold_context = _match_arg(threshold_context, tuple(CONTEXT_TO_BASES), "threshold.context")
    bam = preprocessBam(bam, **kwargs)
    if not isinstance(vcf, dict):
        vcf = read_vcf(vcf, bam.rname_levels, bed=bed, vcf_style=vcf_style, zero_based_bed=zero_based_bed)
    if len(vcf["REF"]) and np.all(np.asarray(vcf["seqnames"]) == NA_INT):
        raise ValueError("Looks like seqlevels styles of BAM and VCF don't match. Please provide VCF as an object with correct seqlevels.")
    if threshold_reads:
        b = CONTEXT_TO_BASES[threshold_context]
        pass_ = rcpp_threshold_reads(bam, b[0], b[1], b[2], b[3], min_context_sites, min_context_beta, max_outofcontext_beta)
    else:
        pass_ = None
    freqs = rcpp_get_base_freqs(bam, pass_, vcf)
    rep = {"name": list(vcf["name"]), "seqnames": np.asarray(vcf["seqnames"], np.int32), "range": np.asarray(vcf["start"], np.int32),
           "REF": list(vcf["REF"]), "ALT": list(vcf["ALT"])}
    rep.update(_base_freq_report(freqs, vcf["REF"], vcf["ALT"]))
    if report_file is None:
        return rep
    cols = ("name", "seqnames", "range", "REF", "ALT", "M+Ref", "U+Ref", "M-Ref", "U-Ref", "M+Alt", "U+Alt", "M-Alt", "U-Alt", "SumRef", "SumAlt", "FEp+", "FEp-")
    fmt = lambda v: "NA" if (isinstance(v, float) and np.isnan(v)) else (repr(float(v)) if isinstance(v, float) and v != int(v) else str(int(v)) if isinstance(v, (float, np.floating, np.integer, int)) else str(v))
    chrom = vcf.get("chrom") or [str(x) for x in rep["seqnames"]]
    with (_gzip.open(report_file, "wt") if gzip else open(report_file, "w")) as fh:
        fh.write("\t".join(cols) + "\n")
        for i in range(len(rep["name"])):
            fh.write("\t".join(fmt(chrom[i] if k == "seqnames" else (rep[k][i] if not isinstance(rep[k][i], np.floating) else float(rep[k][i]))) for k in cols) + "\n")
    return None


def report_text(df, header=True):
    """The last cytosine report computed on `df` as the bytes .writeReport / data.table::fwrite would put into the file
    (R/internal.R:273-287), formatted on the device (ea_cx_report_text).  Returns a uint8 numpy array."""
    L = N.lib()
    names = [str(x).encode() for x in df.rname_levels]
    arr = (C.c_char_p * max(len(names), 1))(*names)
    nbytes = C.c_int64(0)
    N.check(L.ea_cx_report_text(df.handle(need_bases=False), arr, len(names), 1 if header else 0, C.byref(nbytes)))
    out = _host_out(nbytes.value, np.uint8)
    if nbytes.value:
        N.check(L.ea_cx_report_text_fetch(df.handle(need_bases=False), out.ctypes.data_as(C.c_void_p)))
    return out[:nbytes.value]


def _write_report_text(text, path, gz):
    with (_gzip.open(path, "wb") if gz else open(path, "wb")) as fh:
        fh.write(memoryview(np.ascontiguousarray(text)))


# ---- deterministic synthetic reads (bench / parity at BASELINE.json sizes) -----------------------
def synth_config(seed, n_reads, chr_len, len_min, len_max=None, n_islands=0, island_len=0):
    cfg = N.SynthCfg()
    cfg.seed = seed
    cfg.n_reads = n_reads
    cfg.n_chr = len(chr_len)
    cfg.len_min = len_min
    cfg.len_max = len_max if len_max is not None else len_min
    cfg.n_islands = n_islands
    cfg.island_len = island_len
    for i, v in enumerate(chr_len):
        cfg.chr_len[i] = v
    return cfg


def synth_device(cfg, row_begin=0, row_end=None, device=0) -> DeviceBam:
    L = N.lib()
    h = C.c_void_p()
    N.check(L.ea_synth_device(C.byref(cfg), row_begin, cfg.n_reads if row_end is None else row_end, device, C.byref(h)))
    return DeviceBam(h, device)


def synth_host(cfg, row_begin=0, row_end=None) -> BamData:
    L = N.lib()
    row_end = cfg.n_reads if row_end is None else row_end
    n = row_end - row_begin
    rn, st, sa = (np.zeros(max(n, 1), np.int32) for _ in range(3))
    off = np.zeros(n + 1, np.int64)
    N.check(L.ea_synth_host(C.byref(cfg), row_begin, row_end, _p(rn, C.c_int32), _p(st, C.c_int32), _p(sa, C.c_int32), _p(off, C.c_int64), None))
    blob = np.zeros(max(int(off[-1]), 1), np.uint8)
    N.check(L.ea_synth_host(C.byref(cfg), row_begin, row_end, None, None, None, _p(off, C.c_int64), _p(blob, C.c_uint8)))
    return BamData(rn[:n], st[:n], sa[:n], np.arange(n, dtype=np.int32), blob[:int(off[-1])], off)
