"""GPU parity of the round-2 boundary additions, all through the C ABI and bit-exact against the CPU oracle:
range sharding with the ownership filter ON THE DEVICE (ea_set_owned_range), the one-process multi-device entry
(ea_pack_multi; several shards may share one device, so it runs on a 1-GPU box), the scattered-string packer the R shim
calls (ea_pack_calls, through integration/harness = the reader's std::vector<std::string>), streaming ingestion
(ea_ingest_*), and string offsets that do not start at 0."""
import ctypes as C
import os

import numpy as np
import pytest

import epialleler_b200 as ea
import integration.harness as H
from epialleler_b200 import _native as N
from epialleler_b200.distributed import local_report, shard_report
from epialleler_b200.shard import merge_reports, plan_shards_native, slice_bam
from helpers import COLS, assert_report_equal, load_golden
from oracle.oracle import PortCpu

pytestmark = pytest.mark.gpu

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622,
        133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
KW_SETS = [dict(), dict(threshold_reads=False, report_context="CX"), dict(threshold_context="CG", report_context="CX")]


def oracle(bam, **kw):
    return PortCpu(bam).generateCytosineReport(kw.get("threshold_reads", True), kw.get("threshold_context", "CG"), kw.get("min_context_sites", 2),
                                               kw.get("min_context_beta", 0.5), kw.get("max_outofcontext_beta", 0.1),
                                               kw.get("report_context") or kw.get("threshold_context", "CG"))


def shapes():
    yield "wgbs150", ea.synth_host(ea.synth_config(41, 30000, [400000], 150))
    # config 5's layout in small: 24 chromosomes with hg38's proportions, ~40x
    yield "genome24", ea.synth_host(ea.synth_config(20261021, 60000, [c // 14000 for c in HG38], 150))
    yield "pe_merged", ea.synth_host(ea.synth_config(42, 12000, [90000, 50000], 150, 450))
    yield "islands_deep", ea.synth_host(ea.synth_config(43, 20000, [c // 20000 for c in HG38[:6]], 100, 417, n_islands=6, island_len=1500))
    yield "long_reads", ea.synth_host(ea.synth_config(44, 300, [600000, 300000], 10000, 30000))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_range_shards_device_filter_equal_whole(world):
    for name, bam in shapes():
        for kw in KW_SETS:
            want_pass, want = oracle(bam, **kw)
            parts, passes = [], []
            for r in range(world):
                sh, own_pass, rep = shard_report(bam, r, world, **kw)
                assert own_pass.shape[0] == sh.own_hi - sh.own_lo
                parts.append(rep)
                passes.append(own_pass)
            assert np.array_equal(np.concatenate(passes), want_pass), (name, kw)
            assert_report_equal(merge_reports(parts), want, "%s world=%d %s" % (name, world, kw))
        bam.close()


def test_owned_range_filter_is_on_the_device():
    """The trimmed row count comes from the library (k_owned_rows), the fetch returns only owned rows, clearing the filter
    restores the whole report, and the formatted text honours the filter too."""
    bam = ea.synth_host(ea.synth_config(45, 20000, [150000, 80000], 150))
    whole = ea.generateCytosineReport(bam)
    key = (whole["rname"].astype(np.int64) << 32) | whole["pos"].astype(np.int64)
    lo, hi = int(key[len(key) // 3]), int(key[2 * len(key) // 3])
    bam.set_owned(lo, hi, 0)
    part = ea.generateCytosineReport(bam)
    m = (key >= lo) & (key < hi)
    assert part["pos"].shape[0] == int(m.sum())
    assert_report_equal(part, {k: whole[k][m] for k in COLS}, "owned range")
    text = ea.report_text(bam, header=False).tobytes()
    assert text.count(b"\n") == int(m.sum())
    N.check(N.lib().ea_set_owned_range(bam.handle(need_bases=False), -1, 0, 0))
    bam._owned = None
    assert_report_equal(ea.generateCytosineReport(bam), whole, "filter cleared")
    bam.close()


@pytest.mark.parametrize("ndev_shards", [1, 2, 5])
def test_pack_multi_equals_single(ndev_shards):
    devices = [i % max(1, N.lib().ea_device_count()) for i in range(ndev_shards)]
    for name, bam in shapes():
        for kw in KW_SETS[:2] if name != "genome24" else KW_SETS:
            want_pass, want = oracle(bam, **kw)
            for calls_only in (True, False):
                mb = ea.MultiBam(bam, devices, calls_only=calls_only)
                sh = mb.shards()
                assert sh[0].own_lo == 0 and sh[-1].own_hi == bam.nrow
                got_pass, got = ea.generateCytosineReport(mb, return_pass=True, **kw)
                assert np.array_equal(got_pass, want_pass), (name, kw, calls_only)
                assert_report_equal(got, want, "%s multi x%d %s" % (name, ndev_shards, kw))
                mb.close()
        bam.close()


def test_pack_multi_from_reader_strings(monkeypatch):
    """ea_pack_multi_records: the reader's heap strings cut into range shards, one device (here: one shared device) each -
    rows in file order and rows that reach their strings through templid (paired-end fixtures)."""
    monkeypatch.setenv("EA_HOST_PACK_FORCE", "1")
    cases = [ea.synth_host(ea.synth_config(54, 30000, [c // 14000 for c in HG38], 150)), ea.synth_host(ea.synth_config(55, 9000, [70000, 30000], 60, 300)),
             load_golden("capture")[0], load_golden("amplicon010meth")[0]]
    for bam in cases:
        rs = H.ReaderStrings(bam)
        for nsh in (1, 3, 7):
            mb = rs.pack_multi([0] * nsh)
            sh = mb.shards()
            assert sh[0].own_lo == 0 and sh[-1].own_hi == bam.nrow and all(a.own_hi == b.own_lo for a, b in zip(sh, sh[1:]))
            for kw in KW_SETS[:2]:
                want_pass, want = oracle(bam, **kw)
                got_pass, got = ea.generateCytosineReport(mb, return_pass=True, **kw)
                assert np.array_equal(got_pass, want_pass), (nsh, kw)
                assert_report_equal(got, want, "multi from reader strings x%d %s" % (nsh, kw))
            mb.close()
        rs.close(); bam.close()


def test_pack_multi_report_file(tmp_path):
    """report_file with several shards: every device formats the rows it owns; the concatenation is the single-device file."""
    bam = ea.synth_host(ea.synth_config(51, 20000, [90000, 60000, 30000], 150))
    bam.rname_levels = ["chr1", "chr2", "chrX"]
    one, many = str(tmp_path / "one.tsv"), str(tmp_path / "many.tsv.gz")
    ea.generateCytosineReport(bam, report_file=one, report_context="CX")
    mb = ea.MultiBam(bam, [0, 0, 0, 0])
    ea.generateCytosineReport(mb, report_file=many, report_context="CX", gzip=True)
    import gzip
    assert gzip.open(many, "rb").read() == open(one, "rb").read()
    assert open(one, "rb").read().count(b"\n") > 1000
    mb.close(); bam.close()


def test_pack_multi_shuffled_templid_and_errors():
    bam0 = ea.synth_host(ea.synth_config(46, 9000, [70000, 30000], 60, 300))
    blob, off = bam0.blob()
    rng = np.random.default_rng(3)
    perm = rng.permutation(bam0.nrow)
    lens = np.diff(off)
    new_off = np.zeros(bam0.nrow + 1, np.int64)
    np.cumsum(lens[perm], out=new_off[1:])
    new_blob = np.concatenate([blob[off[t]:off[t + 1]] for t in perm])
    inv = np.empty(bam0.nrow, np.int32)
    inv[perm] = np.arange(bam0.nrow, dtype=np.int32)
    bam = ea.BamData(bam0.rname, bam0.strand, bam0.start, inv, new_blob, new_off)   # row r -> string inv[r] = its own bytes
    want_pass, want = oracle(bam0)
    mb = ea.MultiBam(bam, [0, 0, 0])
    got_pass, got = ea.generateCytosineReport(mb, return_pass=True)
    assert np.array_equal(got_pass, want_pass)
    assert_report_equal(got, want, "multi, shuffled templid")
    mb.close()
    with pytest.raises(ea.EaError):
        ea.MultiBam(bam0, [99])
    bad = ea.BamData(bam0.rname[::-1].copy(), bam0.strand, bam0.start[::-1].copy(), bam0.templid, blob, off)
    with pytest.raises(ea.EaError) as e:
        ea.MultiBam(bad, [0, 0])
    assert e.value.code == N.EA_ERR_UNSORTED


def test_blob_offsets_need_not_start_at_zero():
    """A shard is a slice of a larger blob: offsets are used where they lie (both blob packers)."""
    bam = ea.synth_host(ea.synth_config(47, 6000, [50000], 40, 260))
    blob, off = bam.blob()
    lo, hi = 1777, 5333
    want_pass, want = oracle(slice_bam(bam, plan_like(lo, hi)))
    L = N.lib()
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    for pack in (L.ea_pack_blob, L.ea_pack_blob_calls):
        h = C.c_void_p()
        N.check(pack(p(bam.rname[lo:], C.c_int32), p(bam.strand[lo:], C.c_int32), p(bam.start[lo:], C.c_int32), None, hi - lo,
                     p(blob, C.c_uint8), p(off[lo:], C.c_int64), hi - lo, 0, C.byref(h)))
        dev = ea.api.DeviceBam(h, 0)
        assert dev.nbytes == int(off[hi] - off[lo])
        got_pass, got = ea.generateCytosineReport(dev, return_pass=True)
        assert np.array_equal(got_pass, want_pass)
        assert_report_equal(got, want, "offset base")
        dev.close()


def plan_like(lo, hi):
    from epialleler_b200.shard import Shard
    return Shard(0, lo, lo, hi, 0, 1 << 62)


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_pack_calls_from_reader_strings(threads, monkeypatch):
    """ea_pack_calls on the reader's heap strings == the blob packers, for short, ragged and long reads, many small chunks."""
    monkeypatch.setenv("EA_HOST_PACK_THREADS", threads)
    monkeypatch.setenv("EA_HOST_PACK_FORCE", "1")
    monkeypatch.setenv("EA_PACK_CHUNK_MB", "1")
    for name, bam in shapes():
        rs = H.ReaderStrings(bam)
        for records in (False, True):   # pointer tables / the table-free call on the std::string objects themselves
            dev = rs.pack(calls_only=True, records=records)
            assert not N.lib().ea_has_bytes(dev.handle()) and dev.nrow == bam.nrow and dev.nbytes == int(bam.blob()[1][-1])
            for kw in KW_SETS[:2]:
                want_pass, want = oracle(bam, **kw)
                got_pass, got = ea.generateCytosineReport(dev, return_pass=True, **kw)
                assert np.array_equal(got_pass, want_pass), (name, kw)
                assert_report_equal(got, want, "%s ea_pack_calls %s records=%s" % (name, kw, records))
            assert np.array_equal(ea.rcpp_get_xm_beta(dev, "Z", "z"), PortCpu(bam).rcpp_get_xm_beta("Z", "z"))
            dev.close()
        dev2 = rs.pack(calls_only=False)   # ea_pack: byte rows too
        assert N.lib().ea_has_bytes(dev2.handle())
        want_pass, want = oracle(bam)
        got_pass, got = ea.generateCytosineReport(dev2, return_pass=True)
        assert np.array_equal(got_pass, want_pass)
        assert_report_equal(got, want, "%s ea_pack" % name)
        dev2.close(); rs.close(); bam.close()


def test_pack_calls_rejects_bad_input_and_falls_back(monkeypatch):
    monkeypatch.setenv("EA_HOST_PACK_FORCE", "1")
    bam = ea.synth_host(ea.synth_config(48, 2000, [20000], 150))
    blob, off = bam.blob()
    bad = blob.copy()
    bad[int(off[1234]) + 63] = 0x13   # code 3 at a unit border
    rs = H.ReaderStrings(ea.BamData(bam.rname, bam.strand, bam.start, bam.templid, bad, off))
    with pytest.raises(ea.EaError) as e:
        rs.pack(calls_only=True)
    assert e.value.code == N.EA_ERR_DOMAIN
    rs.close()
    # rows that are not the strings in file order (paired-end data: templid is a permutation) stay on the fast path
    for name in ("amplicon010meth", "capture"):
        gold, _ = load_golden(name)
        rs = H.ReaderStrings(gold)
        for records in (False, True):
            dev = rs.pack(calls_only=True, records=records)
            assert not N.lib().ea_has_bytes(dev.handle())
            for kw in KW_SETS:
                want_pass, want = oracle(gold, **kw)
                got_pass, got = ea.generateCytosineReport(dev, return_pass=True, **kw)
                assert np.array_equal(got_pass, want_pass), (name, kw)
                assert_report_equal(got, want, "%s through ea_pack_calls %s" % (name, kw))
            dev.close()
        rs.close()
    # a templid outside the strings is rejected like the reference's seqxm->at() would throw
    bad_tid = bam.templid.copy()
    bad_tid[17] = bam.nrow + 5
    rs = H.ReaderStrings(ea.BamData(bam.rname, bam.strand, bam.start, bad_tid, blob, off))
    with pytest.raises(ea.EaError) as e:
        rs.pack(calls_only=True)
    assert e.value.code == N.EA_ERR_ARG
    rs.close()


@pytest.mark.parametrize("name", ["amplicon010meth", "capture"])
def test_ingest_fixture_templates_one_at_a_time(name, monkeypatch):
    """Templates appended in FILE order, as push_template would; finish() must reproduce the sorted bam.data and the reports."""
    monkeypatch.setenv("EA_PACK_CHUNK_MB", "1")
    gold, z = load_golden(name)
    blob, off = gold.blob()
    nstr = off.shape[0] - 1
    # per-template metadata in file order (templid maps rows -> templates)
    rn, st, sa = (np.zeros(nstr, np.int32) for _ in range(3))
    rn[gold.templid], st[gold.templid], sa[gold.templid] = gold.rname, gold.strand, gold.start
    g = ea.Ingest(0, expected_templates=nstr)
    for t in range(nstr):
        assert g.append(rn[t], st[t], sa[t], blob[off[t]:off[t + 1]].tobytes()) == t
    dev, cols = g.finish()
    assert np.array_equal(cols["rname"], gold.rname) and np.array_equal(cols["start"], gold.start)
    assert np.array_equal(cols["strand"], gold.strand)
    assert np.array_equal(np.diff(off)[cols["templid"]], np.diff(off)[gold.templid])
    want_pass, want = oracle(gold)
    got_pass, got = ea.generateCytosineReport(dev, return_pass=True)
    assert np.array_equal(got_pass, want_pass) if np.array_equal(cols["templid"], gold.templid) else got_pass.sum() == want_pass.sum()
    assert_report_equal(got, want, "%s ingest" % name)
    assert_report_equal(ea.generateCytosineReport(dev, threshold_reads=False, report_context="CX"), oracle(gold, threshold_reads=False, report_context="CX")[1], "ingest CX")
    dev.close()


def test_ingest_synthetic_shuffled_and_spanning_blocks(monkeypatch):
    monkeypatch.setenv("EA_PACK_CHUNK_MB", "1")
    bam = ea.synth_host(ea.synth_config(49, 300, [400000], 10000, 30000))   # templates larger than a ring block's remainder
    blob, off = bam.blob()
    order = np.random.default_rng(9).permutation(bam.nrow)
    g = ea.Ingest(0)
    for t in order:
        g.append(bam.rname[t], bam.strand[t], bam.start[t], blob[off[t]:off[t + 1]].tobytes())
    dev, cols = g.finish()
    assert np.array_equal(cols["start"], bam.start) and np.array_equal(cols["rname"], bam.rname)
    want_pass, want = oracle(bam)
    got_pass, got = ea.generateCytosineReport(dev, return_pass=True)
    if np.unique(bam.start).shape[0] == bam.nrow:   # distinct starts: the sorted order is unique
        assert np.array_equal(got_pass, want_pass)
    assert_report_equal(got, want, "ingest shuffled long reads")
    dev.close()
    with pytest.raises(ea.EaError):   # invalid strand is rejected at finish
        g = ea.Ingest(0)
        g.append(1, 3, 5, b"\xfb\xfc")
        g.finish()
    g = ea.Ingest(0)
    g.append(1, 1, 5, b"\xfc")
    g.abort()
    dev, cols = ea.Ingest(0).finish()   # no templates at all
    assert dev.nrow == 0 and cols["rname"].shape[0] == 0
    assert ea.generateCytosineReport(dev)["pos"].shape[0] == 0
    dev.close()


def test_large_results_into_pageable_memory(monkeypatch):
    """Results of more than 32 MB fetched into PAGEABLE caller memory (what every binding passes) go through the staged copy
    (two page-locked blocks, host threads copying out): the bytes must equal the direct copy and the oracle's values."""
    cfg = ea.synth_config(50, 4_500_000, [45_000_000], 150)
    dev = ea.synth_device(cfg)
    beta = ea.rcpp_get_xm_beta(dev, "Z", "z")          # 36 MB of doubles into a plain numpy array
    monkeypatch.setenv("EA_NO_STAGED_D2H", "1")
    beta_direct = ea.rcpp_get_xm_beta(dev, "Z", "z")
    assert np.array_equal(beta, beta_direct, equal_nan=True)
    host = ea.synth_host(cfg, 4_400_000, 4_500_000)     # the last rows against the oracle (tail piece of the staged copy)
    assert np.array_equal(beta[4_400_000:], PortCpu(host).rcpp_get_xm_beta("Z", "z"))
    monkeypatch.delenv("EA_NO_STAGED_D2H")
    # report columns and flags into pageable arrays through the C ABI directly
    from epialleler_b200.api import _params
    L = N.lib()
    prm = _params(False, "CG", 2, 0.5, 0.1, "CX")
    nrow = C.c_int64(0)
    N.check(L.ea_cytosine_report(dev.handle(), C.byref(prm), None, C.byref(nrow)))
    assert nrow.value * 4 > (32 << 20)
    cols = [np.empty(nrow.value, np.int32) for _ in range(6)]
    N.check(L.ea_cx_report_fetch(dev.handle(), *[c.ctypes.data_as(C.POINTER(C.c_int32)) for c in cols]))
    rep = ea.generateCytosineReport(dev, threshold_reads=False, report_context="CX")   # page-locked result blocks
    for c, k in zip(cols, COLS):
        assert np.array_equal(c, rep[k]), k
    dev.close()


def test_interrupt_poll_process_wide_and_per_handle():
    """The poll callback (the reference polls Rcpp::checkUserInterrupt, src/rcpp_cx_report.cpp:110): a non-zero answer aborts
    the report with EA_ERR_INTERRUPT; a handle's own callback wins over the process-wide one and leaves other handles alone."""
    L = N.lib()
    POLL = C.CFUNCTYPE(C.c_int)
    stop, go = POLL(lambda: 1), POLL(lambda: 0)
    a = ea.synth_host(ea.synth_config(52, 3000, [30000], 150))
    b = ea.synth_host(ea.synth_config(53, 3000, [30000], 150))
    want = ea.generateCytosineReport(a)
    try:
        L.ea_set_handle_poll(a.handle(need_bases=False), C.cast(stop, C.c_void_p))
        with pytest.raises(ea.EaError) as e:
            ea.generateCytosineReport(a)
        assert e.value.code == N.EA_ERR_INTERRUPT
        assert ea.generateCytosineReport(b)["pos"].shape[0] > 0          # another handle is not affected
        L.ea_set_poll(C.cast(stop, C.c_void_p))                          # process-wide: b stops, a's own callback decides for a
        L.ea_set_handle_poll(a.handle(need_bases=False), C.cast(go, C.c_void_p))
        with pytest.raises(ea.EaError):
            ea.generateCytosineReport(b)
        assert_report_equal(ea.generateCytosineReport(a), want, "per-handle poll wins")
    finally:
        L.ea_set_poll(None)
        L.ea_set_handle_poll(a.handle(need_bases=False), None)
    assert_report_equal(ea.generateCytosineReport(a), want, "poll cleared")
    a.close(); b.close()
