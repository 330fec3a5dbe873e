"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Stdlib-only BAM -> ``bam.data`` decoder used to turn the reference's bundled
fixture BAMs into hot-path inputs (HTSlib, R and pysam are absent here).  BAM
decoding itself is OUT OF SCOPE for the product (it stays on the host with
HTSlib); this file only exists so the parity tests can start from the same
inputs as the reference's RUnit tests.

Behaviour restated (not copied) from the reference:
  * flag/mapq/tag filters, template detection by QNAME, CIGAR walk, mate merge
    by base quality, packing ``(seq4bit << 4) | ctx_to_idx(XM)``:
    /root/reference/src/rcpp_read_bam.cpp:72-152 (paired), :236-308 (single)
  * ``ctx_to_idx``: /root/reference/src/epialleleR.h:15-28
  * skip flags, templid, stable order by (rname, start):
    /root/reference/R/internal.R:173-195
  * paired-end detection from the first 1024 records:
    /root/reference/src/rcpp_check_bam.cpp:40-50, R/internal.R:83
"""
from __future__ import annotations

import gzip
import struct
from dataclasses import dataclass, field
from typing import List

import numpy as np

FILL = 0xFB  # 'N' (15) in the high nibble, '-' (11) in the low one


def ctx_to_idx(c: int) -> int:
    """ASCII methylation-call character -> 4-bit code (epialleleR.h:28)."""
    return ((c + 2) >> 2) & 15


_CTX_LUT = np.array([ctx_to_idx(c) for c in range(256)], dtype=np.uint8)


@dataclass
class BamData:
    """The reference's ``bam.data`` restricted to what the hot path reads."""
    rname: np.ndarray          # int32, 1-based factor codes, row order
    strand: np.ndarray         # int32, 1 '+', 2 '-'
    start: np.ndarray          # int32, 1-based
    templid: np.ndarray        # int32, 0-based index into seqxm (file order)
    seqxm: List[bytes]         # packed strings, file order
    rname_levels: List[str] = field(default_factory=list)

    @property
    def nrow(self) -> int:
        return int(self.rname.shape[0])

    def blob(self):
        """Strings as one byte blob + offsets (file order)."""
        lens = np.fromiter((len(s) for s in self.seqxm), dtype=np.int64, count=len(self.seqxm))
        off = np.zeros(len(self.seqxm) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        data = np.frombuffer(b"".join(self.seqxm), dtype=np.uint8).copy() if len(self.seqxm) else np.zeros(0, np.uint8)
        return data, off


def _parse_aux(buf: bytes, p: int, end: int):
    """Return {'XM': bytes, 'XG': bytes} for the Z-typed tags present."""
    tags = {}
    sizes = {ord('c'): 1, ord('C'): 1, ord('s'): 2, ord('S'): 2, ord('i'): 4, ord('I'): 4, ord('f'): 4, ord('A'): 1}
    while p + 3 <= end:
        tag = buf[p:p + 2]
        typ = buf[p + 2]
        p += 3
        if typ in sizes:
            p += sizes[typ]
        elif typ in (ord('Z'), ord('H')):
            q = buf.index(b"\0", p)
            if tag in (b"XM", b"XG"):
                tags[tag.decode()] = buf[p:q]
            p = q + 1
        elif typ == ord('B'):
            sub = buf[p]
            cnt = struct.unpack_from("<I", buf, p + 1)[0]
            p += 5 + cnt * sizes[sub]
        else:
            raise ValueError("unknown aux type %r" % chr(typ))
    return tags


def _records(path: str):
    with gzip.open(path, "rb") as fh:
        raw = fh.read()
    if raw[:4] != b"BAM\1":
        raise ValueError("not a BAM stream")
    l_text = struct.unpack_from("<i", raw, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, p)[0]
    p += 4
    names = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, p)[0]
        names.append(raw[p + 4:p + 4 + l_name - 1].decode())
        p += 4 + l_name + 4
    recs = []
    n = len(raw)
    while p + 4 <= n:
        bs = struct.unpack_from("<i", raw, p)[0]
        q = p + 4
        (ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nref, mpos, tlen) = struct.unpack_from("<iiBBHHHiiii", raw, q)
        q += 32
        qname = raw[q:q + l_rn - 1]
        q += l_rn
        cigar = struct.unpack_from("<%dI" % n_cig, raw, q) if n_cig else ()
        q += 4 * n_cig
        seq = raw[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        qual = raw[q:q + l_seq]
        q += l_seq
        tags = _parse_aux(raw, q, p + 4 + bs)
        recs.append((ref_id, pos, mapq, flag, mpos, tlen, qname, cigar, seq, qual, l_seq, tags))
        p += 4 + bs
    return names, recs


def _seq4(seq: bytes, l_seq: int) -> np.ndarray:
    a = np.frombuffer(seq, dtype=np.uint8)
    out = np.empty(l_seq + (l_seq & 1), dtype=np.uint8)
    out[0::2] = a & 0xF0
    out[1::2] = (a << 4) & 0xF0
    return out[:l_seq]


def read_bam(path: str, min_mapq: int = 0, min_baseq: int = 0, skip_duplicates: bool = False,
             skip_secondary: bool = True, skip_qcfail: bool = True, skip_supplementary: bool = True,
             trim=(0, 0), paired=None) -> BamData:
    """preprocessBam(file) for XM/XG-tagged short-read BAMs (defaults = R/preprocessBam.R:197-209)."""
    if isinstance(trim, int):
        trim = (trim, trim)
    trim5, trim3 = trim
    names, recs = _records(path)
    head = recs[:1024]
    if not head:
        raise ValueError("Empty file provided! Exiting")
    if paired is None:
        paired = sum(1 for r in head if r[3] & 2) > len(head) / 2
    skip = 4 | (256 if skip_secondary else 0) | (512 if skip_qcfail else 0) | \
        (1024 if skip_duplicates else 0) | (2048 if skip_supplementary else 0)
    if paired:
        skip |= 8

    rname, strand, start, seqxm = [], [], [], []

    def walk(rec, buf, qbuf, dest, pe):
        (_ref_id, _pos, _mapq, _flag, _mpos, _tlen, _qn, cigar, seq, qual, l_seq, tags) = rec
        xm = _CTX_LUT[np.frombuffer(tags["XM"], dtype=np.uint8)]
        packed = _seq4(seq, l_seq) | xm[:l_seq]
        q = np.frombuffer(qual, dtype=np.uint8)
        qp = 0
        for c in cigar:
            op, ln = c & 15, c >> 4
            if op in (0, 7, 8):
                if dest + ln > buf.shape[0]:
                    grow = dest + ln - buf.shape[0]
                    buf = np.concatenate([buf, np.full(grow, FILL, np.uint8)])
                    if qbuf is not None:
                        qbuf = np.concatenate([qbuf, np.full(grow, qbuf_fill[0], np.int32)])
                if pe:
                    sel = q[qp:qp + ln].astype(np.int32) > qbuf[dest:dest + ln]
                    qbuf[dest:dest + ln][sel] = q[qp:qp + ln][sel]
                else:
                    sel = q[qp:qp + ln] >= min_baseq
                buf[dest:dest + ln][sel] = packed[qp:qp + ln][sel]
                qp += ln
                dest += ln
            elif op in (1, 4):
                qp += ln
            elif op in (2, 3):
                dest += ln
            elif op in (5, 6, 9):
                pass
            else:
                raise ValueError("Unknown CIGAR operation")
        return buf, qbuf, dest

    qbuf_fill = [min_baseq - (1 if min_baseq > 0 else 0)]
    if paired:
        cur = None  # [qname, rname, start0, strand, width, buf, qbuf]

        def push(t):
            rname.append(t[1] + 1)
            strand.append(t[3])
            start.append(t[2] + trim5 + 1)
            w = t[4]
            seqxm.append(bytes(t[5][:w][trim5:w - trim3 if trim3 else w]))

        for rec in recs:
            (ref_id, pos, mapq, flag, mpos, tlen, qname, _c, _s, _q, _l, tags) = rec
            if (flag & skip) or not (flag & 2) or mapq < min_mapq:
                continue
            if "XG" not in tags or "XM" not in tags:
                continue
            if cur is None or cur[0] != qname:
                if cur is not None:
                    push(cur)
                width = abs(tlen)
                cur = [qname, ref_id, min(pos, mpos), 1 if tags["XG"][:1] == b"C" else 2, width,
                       np.full(max(width, 1), FILL, np.uint8), np.full(max(width, 1), qbuf_fill[0], np.int32)]
            dest0 = pos - cur[2]
            cur[5], cur[6], dest = walk(rec, cur[5], cur[6], dest0, True)
            if cur[4] < dest:
                cur[4] = dest
        if cur is not None:
            push(cur)
    else:
        for rec in recs:
            (ref_id, pos, mapq, flag, _mpos, _tlen, _qn, _c, _s, _q, _l, tags) = rec
            if (flag & skip) or mapq < min_mapq:
                continue
            if "XG" not in tags or "XM" not in tags:
                continue
            buf, _, dest = walk(rec, np.full(0, FILL, np.uint8), None, 0, False)
            rname.append(ref_id + 1)
            strand.append(1 if tags["XG"][:1] == b"C" else 2)
            start.append(pos + trim5 + 1)
            seqxm.append(bytes(buf[:dest][trim5:dest - trim3 if trim3 else dest]))

    rn = np.asarray(rname, dtype=np.int32)
    st = np.asarray(strand, dtype=np.int32)
    sa = np.asarray(start, dtype=np.int32)
    tid = np.arange(rn.shape[0], dtype=np.int32)
    order = np.lexsort((sa, rn))  # stable, keys: rname then start (R/internal.R:195)
    return BamData(rn[order], st[order], sa[order], tid[order], seqxm, names)


def from_xm_text(rows, high_nibble: int = 0x10) -> BamData:
    """BAM-free inputs: rows = [(rname, strand, start, "XM text"), ...] in FILE order."""
    rn = np.asarray([r[0] for r in rows], dtype=np.int32)
    st = np.asarray([r[1] for r in rows], dtype=np.int32)
    sa = np.asarray([r[2] for r in rows], dtype=np.int32)
    seqxm = [bytes(((high_nibble & 0xF0) | ctx_to_idx(ord(ch))) for ch in r[3]) for r in rows]
    tid = np.arange(len(rows), dtype=np.int32)
    order = np.lexsort((sa, rn))
    return BamData(rn[order], st[order], sa[order], tid[order], seqxm, [])


def read_bed(path: str, rname_levels, zero_based: bool = False):
    """The reference's .readBed + .matchTarget factor step (R/internal.R:206-223,467-468): tab-separated file with a header
    line, first three columns chr / start / end (1-based closed unless zero_based); seqnames become factor codes on the
    BAM's rname levels (1-based; INT_MIN = NA when the chromosome is not among them).  Returns (seqnames, start, end, names)."""
    levels = {str(n): i + 1 for i, n in enumerate(rname_levels)}
    chr_, st, en, nm = [], [], [], []
    with open(path) as fh:
        header = fh.readline()
        assert header, "empty BED file"
        for ln in fh:
            f = ln.rstrip("\n").split("\t")
            if len(f) < 3 or not f[0]:
                continue
            chr_.append(levels.get(f[0], -2147483648))
            st.append(int(f[1]) + (1 if zero_based else 0))
            en.append(int(f[2]))
            nm.append(f[3] if len(f) > 3 else "")
    return np.asarray(chr_, np.int32), np.asarray(st, np.int32), np.asarray(en, np.int32), nm


def _ucsc_name(name: str) -> str:
    """NCBI -> UCSC sequence name, the part of GenomeInfoDb::seqlevelsStyle<- the fixtures need."""
    if name.startswith("chr"):
        return name
    return "chrM" if name in ("MT", "M") else "chr" + name


def read_vcf(path: str, rname_levels, bed_path: str = None, vcf_style: str = None, zero_based_bed: bool = False):
    """The reference's .readVcf + the filtering of .getBaseFreqReport (R/internal.R:230-267,617-625), restated without
    VariantAnnotation: records of a (gzipped) VCF, restricted to those overlapping the reduced BED ranges when a BED file is
    given, expanded to one row per ALT allele, sorted by (seqnames, start, end), then only REF of width 1 and single-letter
    ALT kept.  seqnames become factor codes on the BAM's rname levels (INT_MIN = NA).  vcf_style "NCBI" means the VCF
    names chromosomes 1..22,X,Y,MT while the BED/BAM use chr1..: the names are converted like seqlevelsStyle<- does.
    Returns dict(seqnames, start, REF, ALT, name)."""
    import gzip
    levels = {str(n): i + 1 for i, n in enumerate(rname_levels)}
    conv = _ucsc_name if (vcf_style or "").upper() == "NCBI" else (lambda x: x)
    ranges = None
    if bed_path is not None:
        ranges = {}
        with open(bed_path) as fh:
            fh.readline()
            for ln in fh:
                f = ln.rstrip("\n").split("\t")
                if len(f) >= 3 and f[0]:
                    ranges.setdefault(f[0], []).append((int(f[1]) + (1 if zero_based_bed else 0), int(f[2])))
        for k, v in ranges.items():  # GenomicRanges::reduce
            v.sort()
            merged = [list(v[0])]
            for a, b in v[1:]:
                if a <= merged[-1][1] + 1:
                    merged[-1][1] = max(merged[-1][1], b)
                else:
                    merged.append([a, b])
            ranges[k] = merged
    rows = []
    order = {}
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        for ln in fh:
            if ln.startswith("#"):
                continue
            f = ln.rstrip("\n").split("\t")
            if len(f) < 5:
                continue
            chrom, pos, vid, ref, alts = conv(f[0]), int(f[1]), f[2], f[3], f[4]
            end = pos + len(ref) - 1
            if ranges is not None and not any(a <= end and pos <= b for a, b in ranges.get(chrom, ())):
                continue
            order.setdefault(chrom, len(order))
            for alt in alts.split(","):
                rows.append((order[chrom], pos, end, chrom, ref, alt, vid if vid != "." else "%s:%d_%s/%s" % (chrom, pos, ref, alt)))
    rows.sort(key=lambda r: (r[0], r[1], r[2]))  # stable, like BiocGenerics::sort on GRanges
    rows = [r for r in rows if len(r[4]) == 1 and len(r[5]) == 1]
    return {"seqnames": np.asarray([levels.get(r[3], -2147483648) for r in rows], np.int32),
            "start": np.asarray([r[1] for r in rows], np.int32),
            "REF": [r[4] for r in rows], "ALT": [r[5] for r in rows], "name": [r[6] for r in rows]}
