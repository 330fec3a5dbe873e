"""Range partitioning of coordinate-sorted reads over ranks (one process per GPU).

The path shards with NO data-path collective: thresholding is per read, aggregation
is per (rname,pos), and rows are sorted by (rname,start), so a contiguous genomic range is a
contiguous row range.  A cut that falls inside continuous coverage is handled by a HALO: the few
rows that start before the cut but reach past it are packed on both neighbours, and every rank keeps
only the report rows whose (rname,pos) lies in the range it owns.  The only exchange is the
all-gather of per-rank row counts that places each rank's rows in the concatenated report.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class Shard:
    rank: int
    row_lo: int      # first row packed on this rank (halo included)
    own_lo: int      # rows [own_lo, own_hi) are the ones whose pass flags this rank reports
    own_hi: int
    key_lo: int      # report rows with key_lo <= (rname<<32|pos) < key_hi belong to this rank
    key_hi: int


KEY_MIN, KEY_MAX = 0, (1 << 62)


def _key(rname, pos):
    return (np.asarray(rname, dtype=np.int64) << 32) | np.asarray(pos, dtype=np.int64)


def plan_shards(rname, start, length, world: int, weights=None) -> List[Shard]:
    """Split sorted rows into `world` contiguous ranges of about equal weight (bytes by default).

    rname/start/length: per-row arrays (length = string length).  Cuts are snapped to the first row
    of a run of equal (rname,start) so that all rows starting at or after a cut position are on the
    right of it."""
    rname = np.asarray(rname, dtype=np.int64)
    start = np.asarray(start, dtype=np.int64)
    length = np.asarray(length, dtype=np.int64)
    n = rname.shape[0]
    keys = _key(rname, start)
    w = length if weights is None else np.asarray(weights, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(w)])
    ends = start + length
    # furthest end so far within the chromosome (rows are sorted by rname, so the combined key is monotone)
    maxend = np.maximum.accumulate((rname << 32) | np.where(length > 0, ends, 0))
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r // world
        c = int(np.searchsorted(cum, target, side="left"))
        c = min(max(c, cuts[-1]), n)
        if c < n:
            c = int(np.searchsorted(keys, keys[c], side="left"))  # snap to the first row with that (rname,start)
        cuts.append(max(c, cuts[-1]))
    cuts.append(n)
    shards = []
    for r in range(world):
        own_lo, own_hi = cuts[r], cuts[r + 1]
        key_lo = KEY_MIN if r == 0 or own_lo >= n else int(keys[own_lo])
        key_hi = KEY_MAX if r == world - 1 or own_hi >= n else int(keys[own_hi])
        if r > 0 and own_lo < n:
            # halo: earlier rows of the same chromosome that reach past the cut position
            row_lo = int(np.searchsorted(maxend[:own_lo], key_lo, side="right"))
        else:
            row_lo = own_lo
        if own_lo >= n:  # nothing left for this rank
            row_lo, key_lo, key_hi = n, KEY_MAX, KEY_MAX
        shards.append(Shard(r, row_lo, own_lo, own_hi, key_lo, key_hi))
    # later empty shards must not steal the open-ended range of the last non-empty one
    for r in range(world - 1):
        if shards[r + 1].own_lo >= n and shards[r].own_lo < n:
            shards[r].key_hi = KEY_MAX
    return shards


def plan_shards_native(bam, world: int) -> List[Shard]:
    """The same plan from the library's planner (ea_plan_shards, csrc/multi.inc): no per-row Python temporaries - the cuts
    are binary searches over the blob offsets and the halo is found by walking back from the cut, so it runs on a
    billion rows.  Pure host code; identical to plan_shards (tests/test_shard.py)."""
    import ctypes as C

    from . import _native as N
    L = N.lib()
    out = (N.Shard * world)()
    off = bam.blob()[1]
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    N.check(L.ea_plan_shards(p(bam.rname, C.c_int32), p(bam.start, C.c_int32), p(bam.templid, C.c_int32), bam.nrow, p(off, C.c_int64),
                             int(off.shape[0] - 1), world, out))
    return [Shard(r, s.row_lo, s.own_lo, s.own_hi, s.key_lo, s.key_hi) for r, s in enumerate(out)]


def plan_shards_windowed(n: int, world: int, fetch, max_len: int, window: int = 1 << 16) -> List[Shard]:
    """The plan for rows that are NOT materialised anywhere as whole columns (a billion-row data set of which every rank
    loads only its own range): cuts at equal ROW counts, and only `window` rows around every cut are looked at.
    fetch(i0, i1) -> (rname, start, length) arrays of rows [i0, i1).  max_len bounds the string lengths (the halo walk
    stops at the first row that cannot reach the cut).  Equal to plan_shards(..., weights=ones) whenever the run of equal
    (rname,start) at a cut and the halo fit into the window (asserted)."""
    cuts = [0]
    info = {}
    for r in range(1, world):
        c = min(max(n * r // world, cuts[-1]), n)
        if c < n:
            i0 = max(0, c - window)
            rn, st, ln = (np.asarray(a, dtype=np.int64) for a in fetch(i0, c + 1))
            k = (rn << 32) | st
            first = int(np.searchsorted(k, k[-1], side="left"))
            assert first > 0 or i0 == 0, "run of equal (rname,start) longer than the window"
            c = max(i0 + first, cuts[-1])
            # halo: rows before the cut, same chromosome, reaching past the cut position
            cut_rn, cut_pos = rn[c - i0], st[c - i0]
            cand = (rn[:c - i0] == cut_rn) & (st[:c - i0] + ln[:c - i0] > cut_pos)
            lo = i0 + int(np.argmax(cand)) if cand.any() else c
            could = (rn[:c - i0] == cut_rn) & (st[:c - i0] + max_len > cut_pos)
            assert i0 == 0 or not could[0], "halo longer than the window"
            info[c] = (lo, int(k[c - i0]))
        cuts.append(c)
    cuts.append(n)
    shards = []
    for r in range(world):
        own_lo, own_hi = cuts[r], cuts[r + 1]
        key_lo = KEY_MIN if r == 0 or own_lo >= n else info[own_lo][1]
        key_hi = KEY_MAX if r == world - 1 or own_hi >= n else info[own_hi][1]
        row_lo = info[own_lo][0] if (r > 0 and own_lo < n) else own_lo
        if own_lo >= n:
            row_lo, key_lo, key_hi = n, KEY_MAX, KEY_MAX
        shards.append(Shard(r, row_lo, own_lo, own_hi, key_lo, key_hi))
    for r in range(world - 1):
        if shards[r + 1].own_lo >= n and shards[r].own_lo < n:
            shards[r].key_hi = KEY_MAX
    return shards


def filter_report(rep: dict, shard: Shard) -> dict:
    """Keep the report rows this rank owns (host-side mask: used where the per-shard engine is not the library - the CPU
    tests; the library filters on the device, ea_set_owned_range)."""
    k = _key(rep["rname"], rep["pos"])
    m = (k >= shard.key_lo) & (k < shard.key_hi)
    out = {c: (v[m] if isinstance(v, np.ndarray) else v) for c, v in rep.items()}
    return out


def merge_reports(parts: List[dict]) -> dict:
    """Concatenate per-rank reports in rank order (= genomic order)."""
    cols = ("rname", "strand", "pos", "context", "meth", "unmeth")
    return {c: np.concatenate([p[c] for p in parts]) if parts else np.zeros(0, np.int32) for c in cols}


def slice_bam(bam, shard: Shard, chunk_rows: int = 1 << 16):
    """Rows [row_lo, own_hi) of a host BamData as a self-contained BamData: its strings in row order
    (templid = 0..n-1), so nrow equals the number of strings as the per-read routines require.

    When the rows are the strings in file order (what preprocessBam produces for sorted BAMs) the shard's
    strings are one contiguous piece of the blob: the slice is a VIEW plus rebased offsets, no copy and
    no per-byte temporaries.  Otherwise the strings are gathered `chunk_rows` rows at a time, so the
    temporaries stay bounded (a few tens of MB) however large the shard is."""
    from .api import BamData
    sl = slice(shard.row_lo, shard.own_hi)
    blob, off = bam.blob()
    tid = bam.templid[sl].astype(np.int64)
    n = tid.shape[0]
    if n == 0:
        return BamData(bam.rname[sl], bam.strand[sl], bam.start[sl], np.zeros(0, np.int32), np.zeros(0, np.uint8),
                       np.zeros(1, np.int64), bam.rname_levels, bam.device)
    if tid[-1] - tid[0] == n - 1 and (n == 1 or bool(np.all(np.diff(tid) == 1))):
        b0 = int(off[tid[0]])
        new_off = off[tid[0]:tid[-1] + 2] - b0
        new_blob = blob[b0:int(off[tid[-1] + 1])]
    else:
        lens = off[tid + 1] - off[tid]
        new_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=new_off[1:])
        new_blob = np.empty(int(new_off[-1]), dtype=np.uint8)
        for c0 in range(0, n, chunk_rows):
            c1 = min(c0 + chunk_rows, n)
            ln = lens[c0:c1]
            tot = int(new_off[c1] - new_off[c0])
            if tot == 0:
                continue
            # byte i of the chunk comes from off[tid[row]] + (i - chunk-relative start of the row)
            row_of = np.repeat(np.arange(c1 - c0, dtype=np.int64), ln)
            src = (off[tid[c0:c1]] - (new_off[c0:c1] - new_off[c0]))[row_of] + np.arange(tot, dtype=np.int64)
            new_blob[new_off[c0]:new_off[c1]] = blob[src]
    return BamData(bam.rname[sl], bam.strand[sl], bam.start[sl], np.arange(n, dtype=np.int32), new_blob, new_off,
                   bam.rname_levels, bam.device)
