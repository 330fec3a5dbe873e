"""GPU parity tests: the CUDA path through the C ABI against the committed golden vectors (outputs of
the reference's own compiled C++) and against the CPU oracle on seeded synthetic reads.  Bit-exact:
identical pass flags, identical betas (IEEE double), identical integer report rows in identical order."""
import ctypes as C

import numpy as np
import pytest

import epialleler_b200 as ea
from epialleler_b200 import _native as N
from helpers import (BETAS, CASES, COLS, KAT_C1_BETA, KAT_C1_NOTHR_ZXH, KAT_C1_PASS, KAT_C1_ROWS, KAT_C1_THR_ZXH, KAT_C2_ROWS,
                     KAT_C2_Z, KAT_C3_NOTHR_ZXH, KAT_C3_PASS, KAT_C3_ROWS, KAT_C3_THR_ZX, assert_report_equal, bam_from_xm,
                     golden_names, kat_c4_rows, load_golden, report_tuples)
from oracle.oracle import CONTEXT_TO_BASES, PortCpu

pytestmark = pytest.mark.gpu
CG = ("Z", "z", "XH", "xh", 2, 0.5, 0.1)


@pytest.fixture(params=["auto", "vertical", "byte-auto", "lane-group", "thread-per-row-tma", "chunk-per-lane", "long-read"], autouse=True)
def kernel_policy(request):
    """Every parity test runs with each window kernel: the automatic choice (row-per-lane kernel for one-context
    reports over short 32-byte-multiple rows, else the chunk-per-lane kernel, else the lane-group kernel), the general
    lane-group kernel forced, the thread-per-row TMA kernel, and the chunk-per-lane kernel."""
    old = N.lib().ea_set_kernel_policy({"auto": 0, "lane-group": 1, "thread-per-row-tma": 2, "chunk-per-lane": 3, "long-read": 4, "byte-auto": 5, "vertical": 6}[request.param])
    yield request.param
    N.lib().ea_set_kernel_policy(old)


def two_pass(bam, thr, tctx, ms, mb, mo, rctx):
    """The reference's own call sequence: rcpp_threshold_reads then rcpp_cx_report."""
    if thr:
        p = ea.rcpp_threshold_reads(bam, *CONTEXT_TO_BASES[tctx], ms, mb, mo)
    else:
        p = np.ones(bam.nrow, bool)
    return p, ea.rcpp_cx_report(bam, p, CONTEXT_TO_BASES[rctx][0])


@pytest.mark.parametrize("name", golden_names())
def test_golden_vectors_two_pass_and_fused(name):
    bam, z = load_golden(name)
    for cname, thr, tctx, ms, mb, mo, rctx in CASES:
        want = {k: z[cname + "." + k] for k in COLS}
        p, rep = two_pass(bam, thr, tctx, ms, mb, mo, rctx)
        assert np.array_equal(p.astype(np.uint8), z[cname + ".pass"]), (name, cname, "two-pass flags")
        assert_report_equal(rep, want, "%s/%s two-pass" % (name, cname))
        p2, rep2 = ea.generateCytosineReport(bam, threshold_reads=thr, threshold_context=tctx, min_context_sites=ms,
                                             min_context_beta=mb, max_outofcontext_beta=mo, report_context=rctx, return_pass=True)
        assert np.array_equal(p2.astype(np.uint8), z[cname + ".pass"]), (name, cname, "fused flags")
        assert_report_equal(rep2, want, "%s/%s fused" % (name, cname))
    for bname, m, u in BETAS:
        assert np.array_equal(ea.rcpp_get_xm_beta(bam, m, u), z[bname]), (name, bname)
    bam.close()


def test_known_answer_tests():
    b = bam_from_xm(KAT_C1_ROWS)
    p = ea.rcpp_threshold_reads(b, *CG)
    assert p.astype(int).tolist() == KAT_C1_PASS
    assert ea.rcpp_get_xm_beta(b, "Z", "z").tolist() == KAT_C1_BETA
    assert report_tuples(ea.rcpp_cx_report(b, p, "ZXH")) == KAT_C1_THR_ZXH
    assert report_tuples(ea.rcpp_cx_report(b, None, "ZXH")) == KAT_C1_NOTHR_ZXH
    assert report_tuples(ea.generateCytosineReport(b, report_context="CX")) == KAT_C1_THR_ZXH

    b = bam_from_xm(KAT_C2_ROWS)
    p = ea.rcpp_threshold_reads(b, *CG)
    assert p.all()
    assert report_tuples(ea.rcpp_cx_report(b, p, "Z")) == KAT_C2_Z
    assert report_tuples(ea.generateCytosineReport(b)) == KAT_C2_Z

    b = bam_from_xm(KAT_C3_ROWS)
    p = ea.rcpp_threshold_reads(b, *CG)
    assert p.astype(int).tolist() == KAT_C3_PASS
    assert report_tuples(ea.rcpp_cx_report(b, None, "ZXH")) == KAT_C3_NOTHR_ZXH
    assert report_tuples(ea.rcpp_cx_report(b, p, "ZX")) == KAT_C3_THR_ZX
    assert report_tuples(ea.generateCytosineReport(b, report_context="CxG")) == KAT_C3_THR_ZX

    b = bam_from_xm(kat_c4_rows())
    p, rep = ea.generateCytosineReport(b, return_pass=True)
    assert p.sum() == 1 and p[0]
    assert (rep["pos"].shape[0], rep["meth"].sum(), rep["unmeth"].sum()) == (10, 10, 9990)


SYNTH = {
    "wgbs150": dict(seed=20261018, n_reads=40000, chr_len=[200000], len_min=150),
    "pe_merged": dict(seed=5, n_reads=20000, chr_len=[150000, 90000, 40000], len_min=150, len_max=450),
    "capture_islands": dict(seed=20261019, n_reads=60000, chr_len=[400000, 300000], len_min=100, len_max=417, n_islands=12, island_len=2000),
    "long_reads": dict(seed=20261020, n_reads=300, chr_len=[900000, 500000], len_min=10000, len_max=30000),
    "tiny_reads": dict(seed=9, n_reads=30000, chr_len=[30000], len_min=1, len_max=40),
    "deep_amplicon": dict(seed=11, n_reads=70000, chr_len=[20000], len_min=120, len_max=260, n_islands=1, island_len=600),
    # 32768..65535 rows over one 2048-position window: fused thresholds on the lane-group kernel's 16|16-bit counters
    "amplicon_45k": dict(seed=12, n_reads=45000, chr_len=[20000], len_min=120, len_max=260, n_islands=1, island_len=600),
    # long reads at ~400x: k_window_vsr with 9 planes per counter; at ~1000x: more rows per window than it takes -> k_window_plr
    "long_reads_400x": dict(seed=13, n_reads=2000, chr_len=[100000], len_min=10000, len_max=30000),
    "long_reads_1000x": dict(seed=14, n_reads=1500, chr_len=[40000], len_min=10000, len_max=30000),
    # BASELINE configs[4]'s layout in small: 24 chromosomes in hg38's proportions, 150 bp reads at ~40x
    "genome24": dict(seed=20261021, n_reads=60000, chr_len=[c // 14000 for c in (248956422, 242193529, 198295559, 190214555, 181538259, 170805979,
                     159345973, 145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                     58617616, 64444167, 46709983, 50818468, 156040895, 57227415)], len_min=150),
}
SYNTH_CASES = [(True, "CG", 2, 0.5, 0.1, "CG"), (False, "CG", 2, 0.5, 0.1, "CX"), (True, "CHH", 1, 0.02, 0.9, "CxG"),
               (True, "CX", 3, 0.04, 0.1, "CHG"),
               (True, "CG", 2, 0.5, 0.1, "CX")]   # BASELINE configs[2]'s exact pair: CG thresholds fused into a three-context report


@pytest.mark.parametrize("shape", sorted(SYNTH))
def test_synthetic_reads_against_oracle(shape):
    cfg = ea.synth_config(**SYNTH[shape])
    host = ea.synth_host(cfg)
    dev = ea.synth_device(cfg)
    # the device generator and the host generator are the same function
    back = dev.to_host()
    assert np.array_equal(back.blob()[1], host.blob()[1]) and np.array_equal(back.blob()[0], host.blob()[0])
    assert np.array_equal(back.start, host.start) and np.array_equal(back.strand, host.strand) and np.array_equal(back.rname, host.rname)
    oracle = PortCpu(host)
    for thr, tctx, ms, mb, mo, rctx in SYNTH_CASES:
        want_p, want = oracle.generateCytosineReport(thr, tctx, ms, mb, mo, rctx)
        assert want["pos"].shape[0] > 0
        for bam in (host, dev):  # packed from host buffers / generated in HBM
            p, rep = ea.generateCytosineReport(bam, threshold_reads=thr, threshold_context=tctx, min_context_sites=ms, min_context_beta=mb,
                                               max_outofcontext_beta=mo, report_context=rctx, return_pass=True)
            assert np.array_equal(p, want_p), (shape, tctx, "fused flags")
            assert_report_equal(rep, want, "%s fused %s/%s" % (shape, tctx, rctx))
        p, rep = two_pass(host, thr, tctx, ms, mb, mo, rctx)
        assert np.array_equal(p, want_p)
        assert_report_equal(rep, want, "%s two-pass %s/%s" % (shape, tctx, rctx))
    assert np.array_equal(ea.rcpp_get_xm_beta(host, "ZX", "zx"), oracle.rcpp_get_xm_beta("ZX", "zx"))
    host.close(); dev.close()


def test_unusual_letter_sets_use_the_generic_kernel():
    cfg = ea.synth_config(3, 5000, [40000], 90, 200)
    host = ea.synth_host(cfg)
    o = PortCpu(host)
    for letters in (("ZZ", "z", "X", "x"), ("Z", "zx", "", ""), ("ZU", "zu", "H", "h."), ("Z-", "z", "XH", "xh"), ("z", "Z", "xh", "XH")):
        assert np.array_equal(ea.rcpp_threshold_reads(host, *letters, 2, 0.5, 0.1), o.rcpp_threshold_reads(*letters, 2, 0.5, 0.1)), letters
        assert np.array_equal(ea.rcpp_get_xm_beta(host, letters[0], letters[1]), o.rcpp_get_xm_beta(letters[0], letters[1])), letters
    # report letters other than Z/X/H select nothing; lower-case never selects (src/rcpp_cx_report.cpp:88-91)
    for ctx in ("", "zxh", "U.", "Zz", "hX"):
        assert_report_equal(ea.rcpp_cx_report(host, None, ctx), o.rcpp_cx_report(None, ctx), ctx)


def test_threshold_edge_values():
    cfg = ea.synth_config(4, 8000, [50000], 150)
    host = ea.synth_host(cfg)
    o = PortCpu(host)
    for ms, mb, mo in ((0, 0.0, 0.0), (1, 1.0, 1.0), (2, 0.3, 0.3), (4, 1.0 / 3.0, 0.1), (2, float("nan"), 0.1), (2, 0.5, float("nan")),
                       (2**32 - 1, 0.5, 0.1), (2, -1.0, 2.0), (3, 0.1 + 0.2, 1e-300)):
        got = ea.rcpp_threshold_reads(host, "Z", "z", "XH", "xh", ms, mb, mo)
        assert np.array_equal(got, o.rcpp_threshold_reads("Z", "z", "XH", "xh", ms, mb, mo)), (ms, mb, mo)
        p, _ = ea.generateCytosineReport(host, min_context_sites=ms, min_context_beta=mb, max_outofcontext_beta=mo, return_pass=True)
        assert np.array_equal(p, got), (ms, mb, mo, "fused")


def test_empty_and_degenerate_inputs():
    empty = ea.BamData.from_strings([], [], [], [], [])
    assert ea.rcpp_threshold_reads(empty, *CG).shape == (0,)
    assert ea.rcpp_cx_report(empty, None, "ZXH")["pos"].shape == (0,)
    assert ea.generateCytosineReport(empty)["pos"].shape == (0,)
    # empty strings, a single byte, an all-'-' read, NA-like pass values
    rows = [(1, 1, 5, ""), (1, 2, 5, "Z"), (1, 1, 6, "----"), (2, 1, 1, "zZ"), (2, 1, 4000, "Z" * 33), (2, 2, 8191, "xXh"), (3, 1, 4096, "ZzZ")]
    b = bam_from_xm(rows)
    o = PortCpu(b)
    assert np.array_equal(ea.rcpp_threshold_reads(b, *CG), o.rcpp_threshold_reads(*CG))
    for ctx in ("Z", "ZXH", "X"):
        assert_report_equal(ea.rcpp_cx_report(b, None, ctx), o.rcpp_cx_report(None, ctx), ctx)
        pv = np.array([1, 0, 2, -2147483648, 0, 1, 1])
        assert_report_equal(ea.rcpp_cx_report(b, pv, ctx), o.rcpp_cx_report(pv, ctx), ctx + " pass")
    for rctx in ("CG", "CX"):
        p, rep = ea.generateCytosineReport(b, report_context=rctx, min_context_sites=1, return_pass=True)
        wp, want = o.generateCytosineReport(True, "CG", 1, 0.5, 0.1, rctx)
        assert np.array_equal(p, wp)
        assert_report_equal(rep, want, rctx)


def test_invalid_inputs_are_rejected():
    with pytest.raises(ea.EaError) as e:  # unsorted rows
        ea.BamData.from_strings([1, 1], [1, 1], [10, 5], [0, 1], [b"\x17", b"\x17"]).handle()
    assert e.value.code == -3
    with pytest.raises(ea.EaError) as e:  # low nibble 0 is not a methylation-call code
        ea.BamData.from_strings([1], [1], [1], [0], [b"\x10"]).handle()
    assert e.value.code == -4
    with pytest.raises(ea.EaError) as e:  # templid out of range
        ea.BamData.from_strings([1], [1], [1], [3], [b"\x17"]).handle()
    assert e.value.code == -1


def test_string_pointer_packer_matches_blob_packer():
    bam, z = load_golden("amplicon010meth")
    L = N.lib()
    blob, off = bam.blob()
    strs = [blob[off[i]:off[i + 1]].tobytes() for i in range(bam.nstr)]
    keep = [C.create_string_buffer(s, len(s)) for s in strs]
    ptrs = (C.c_void_p * len(strs))(*[C.addressof(k) for k in keep])
    lens = np.array([len(s) for s in strs], np.int64)
    h = C.c_void_p()
    N.check(L.ea_pack(bam.rname.ctypes.data_as(C.POINTER(C.c_int32)), bam.strand.ctypes.data_as(C.POINTER(C.c_int32)),
                      bam.start.ctypes.data_as(C.POINTER(C.c_int32)), bam.templid.ctypes.data_as(C.POINTER(C.c_int32)), bam.nrow,
                      ptrs, lens.ctypes.data_as(C.POINTER(C.c_int64)), len(strs), 0, C.byref(h)))
    out = np.zeros(bam.nrow, np.uint8)
    N.check(L.ea_threshold_reads(h, b"Z", b"z", b"XH", b"xh", 2, 0.5, 0.1, out.ctypes.data_as(C.POINTER(C.c_uint8))))
    assert np.array_equal(out, z["cg_default.pass"])
    L.ea_free(h)


def test_pack_pipeline_chunks_and_shuffled_strings(monkeypatch):
    """ea_pack_blob: (a) the blob travels in several chunks, each repacked behind the next one's transfer (identity
    templid); (b) templid is a permutation and there are more strings than rows - the one-shot repack after the last
    chunk.  Both must give the same bytes as the device generator / the oracle.  Odd lengths exercise every source
    alignment and tail length of the funnel-shift repack."""
    monkeypatch.setenv("EA_PACK_CHUNK_MB", "1")
    cfg = ea.synth_config(77, 60000, [250000, 100000], 1, 333)
    host = ea.synth_host(cfg)
    blob, off = host.blob()
    assert blob.nbytes > 8 << 20  # at least 8 chunks of 1 MB
    oracle = PortCpu(host)
    want_p, want = oracle.generateCytosineReport(True, "CG", 2, 0.5, 0.1, "CX")
    # (a) chunked, identity
    a = ea.BamData(host.rname, host.strand, host.start, np.arange(host.nrow, dtype=np.int32), blob, off)
    view = ea.api.DeviceBam(a.handle(), 0)  # export the packed rows again: the repack must be lossless
    back = view.to_host()
    view._h = None                           # the handle stays owned by `a`
    assert np.array_equal(back.blob()[0], blob) and np.array_equal(back.blob()[1], off)
    p, rep = ea.generateCytosineReport(a, report_context="CX", return_pass=True)
    assert np.array_equal(p, want_p)
    assert_report_equal(rep, want, "chunked pack")
    # (b) strings stored in a shuffled order, with unreferenced decoys in between
    rng = np.random.default_rng(5)
    n = host.nrow
    nstr = n + 1000
    slot = rng.permutation(nstr)[:n].astype(np.int32)          # row r's string lives in slot[r]
    lens = np.zeros(nstr, np.int64)
    lens[slot] = np.diff(off)
    decoy = np.setdiff1d(np.arange(nstr), slot)
    lens[decoy] = rng.integers(0, 50, decoy.shape[0])
    off2 = np.zeros(nstr + 1, np.int64)
    np.cumsum(lens, out=off2[1:])
    blob2 = np.full(int(off2[-1]), 0x1C, np.uint8)              # decoys: valid '.' bytes
    for r in range(n):
        blob2[off2[slot[r]]:off2[slot[r]] + lens[slot[r]]] = blob[off[r]:off[r + 1]]
    b = ea.BamData(host.rname, host.strand, host.start, slot, blob2, off2)
    p, rep = ea.generateCytosineReport(b, report_context="CX", return_pass=True)
    assert np.array_equal(p, want_p)
    assert_report_equal(rep, want, "shuffled strings")
    assert N.lib().ea_nbytes(b.handle()) == int(np.diff(off).sum())
    a.close(); b.close(); host.close()


def test_calls_only_packer(monkeypatch, kernel_policy):
    """ea_pack_blob_calls: the host transposes chunks from the back (AVX2), the device transposes raw chunks from the front,
    both into the same call planes.  Many small chunks, 1 / 3 / 16 host threads, short ragged reads and long reads: reports,
    flags and betas equal the oracle; what the planes cannot serve repacks with the byte rows; bad inputs are rejected."""
    if kernel_policy != "auto":
        pytest.skip("the calls-only packer is used with the default kernel policy")
    L = N.lib()
    monkeypatch.setenv("EA_PACK_CHUNK_MB", "1")
    monkeypatch.setenv("EA_HOST_PACK_FORCE", "1")
    shapes = [ea.synth_config(77, 60000, [250000, 100000], 1, 333), ea.synth_config(20261020, 120, [900000, 500000], 10000, 30000)]
    for cfg in shapes:
        host = ea.synth_host(cfg)
        blob, off = host.blob()
        oracle = PortCpu(host)
        want_p, want = oracle.generateCytosineReport(True, "CG", 2, 0.5, 0.1, "CX")
        for nthreads in ("1", "3", "16"):
            monkeypatch.setenv("EA_HOST_PACK_THREADS", nthreads)
            b = ea.BamData(host.rname, host.strand, host.start, np.arange(host.nrow, dtype=np.int32), blob, off)
            assert L.ea_has_bytes(b.handle(need_bases=False)) == 0
            assert 0 < L.ea_pack_h2d_bytes(b.handle(need_bases=False)) < blob.nbytes + 24 * host.nrow + 65536
            p, rep = ea.generateCytosineReport(b, report_context="CX", return_pass=True)
            assert np.array_equal(p, want_p)
            assert_report_equal(rep, want, "calls-only pack, %s threads" % nthreads)
            assert np.array_equal(ea.rcpp_get_xm_beta(b, "ZX", "zx"), oracle.rcpp_get_xm_beta("ZX", "zx"))
            assert L.ea_has_bytes(b.handle(need_bases=False)) == 0
            # a letter set outside the named contexts needs the byte rows: transparent repack
            assert np.array_equal(ea.rcpp_threshold_reads(b, "ZU", "z", "X", "x", 1, 0.2, 0.5), oracle.rcpp_threshold_reads("ZU", "z", "X", "x", 1, 0.2, 0.5))
            assert L.ea_has_bytes(b.handle(need_bases=False)) == 1
            b.close()
        host.close()
    # invalid inputs through the calls-only packer
    def pack(rn, st, sa, strs, tid=None):
        b = ea.BamData.from_strings(rn, st, sa, tid if tid is not None else list(range(len(strs))), strs)
        try:
            return b, L.ea_has_bytes(b.handle(need_bases=False))
        finally:
            b.close()
    ok = bytes([0x1C]) * 70
    for bad_at in (0, 31, 32, 69):   # first unit, unit border, last (partial) unit
        s = bytearray(ok); s[bad_at] = 0x13
        with pytest.raises(ea.EaError) as e:
            pack([1, 1], [1, 2], [5, 6], [ok, bytes(s)])
        assert e.value.code == -4, bad_at
    with pytest.raises(ea.EaError) as e:
        pack([1, 1], [1, 1], [10, 5], [ok, ok])
    assert e.value.code == -3
    with pytest.raises(ea.EaError) as e:
        pack([1, 1], [1, 3], [5, 6], [ok, ok])
    assert e.value.code == -1
    assert pack([1, 1], [1, 2], [5, 6], [ok, ok])[1] == 0
    assert pack([1, 1], [1, 2], [5, 6], [ok, ok], tid=[1, 0])[1] == 1   # rows are not the strings in file order: the general packer


def test_row_capacity_retry_on_sparse_coverage():
    """A CX report over ~1x coverage has about one row per five bytes of calls, far above the first capacity guess:
    the window kernel only counts, the pipeline re-runs with the exact capacity, and the rows are still bit-exact."""
    cfg = ea.synth_config(78, 60000, [9000000], 150)
    host = ea.synth_host(cfg)
    want_p, want = PortCpu(host).generateCytosineReport(False, "CG", 2, 0.5, 0.1, "CX")
    assert want["pos"].shape[0] > host.blob()[0].nbytes // 64 + 4096
    for bam in (host, ea.synth_device(cfg)):
        rep = ea.generateCytosineReport(bam, threshold_reads=False, report_context="CX")
        assert_report_equal(rep, want, "capacity retry")
        rep = ea.generateCytosineReport(bam, threshold_reads=False, report_context="CG")  # smaller report, same buffers
        assert_report_equal(rep, PortCpu(host).generateCytosineReport(False, "CG", 2, 0.5, 0.1, "CG")[1], "after retry")
        bam.close()


def test_range_sharded_gpu_equals_whole():
    """The N>1 data path on one GPU: every rank's shard (halo included) through the CUDA kernels, ownership
    filter, concatenation == the unsharded oracle report."""
    from epialleler_b200.distributed import shard_report
    from epialleler_b200.shard import merge_reports
    for cfg, kw in ((ea.synth_config(41, 30000, [120000], 150), dict()),
                    (ea.synth_config(42, 20000, [60000, 50000], 80, 420), dict(threshold_reads=False, report_context="CX")),
                    (ea.synth_config(43, 30000, [80000], 100, 300, n_islands=3, island_len=1500), dict(threshold_context="CxG", report_context="CHG"))):
        host = ea.synth_host(cfg)
        want_pass, want = PortCpu(host).generateCytosineReport(kw.get("threshold_reads", True), kw.get("threshold_context", "CG"), 2, 0.5, 0.1,
                                                               kw.get("report_context") or kw.get("threshold_context", "CG"))
        for world in (2, 3, 8):
            parts, passes = [], []
            for r in range(world):
                _, own_pass, rep = shard_report(host, r, world, **kw)
                parts.append(rep); passes.append(own_pass)
            assert np.array_equal(np.concatenate(passes), want_pass), (world, kw)
            assert_report_equal(merge_reports(parts), want, "world=%d" % world)


# ---- BED-assisted report: ea_match_target / ea_bed_report ------------------
from helpers import BED_CASES  # noqa: E402
from test_oracle import golden_bed, random_bed  # noqa: E402


@pytest.mark.parametrize("name", sorted(BED_CASES))
def test_bed_report_golden_vectors(name):
    bam, z = load_golden(name)
    bed = golden_bed(z)
    for cname, typ, par in BED_CASES[name]:
        match, rep = ea.generateBedReport(bam, bed, bed_type=typ, match_tolerance=par, match_min_overlap=par, return_match=True)
        assert np.array_equal(match, z["bed." + cname + ".match"]), (name, cname)
        assert np.array_equal(rep["counts"], z["bed." + cname + ".counts"]), (name, cname)
        assert np.array_equal(rep["VEF"], z["bed." + cname + ".vef"], equal_nan=True), (name, cname)
        m2 = (ea.rcpp_match_amplicon if typ == "amplicon" else ea.rcpp_match_capture)(bam, bed, par)
        assert np.array_equal(m2, z["bed." + cname + ".match"])
    bam.close()


def test_bed_report_runit_constants_and_ecdf():
    # inst/unitTests/test_generateBedReport.R:12-54, test_generateBedEcdf.R:8-12, through the public mirror
    bam, z = load_golden("amplicon010meth")
    rep = ea.generateAmpliconReport(bam, golden_bed(z))
    assert rep["VEF"].shape[0] == 5 and np.nansum(rep["nreads-"]) == 440 and np.nansum(rep["nreads+"]) + np.nansum(rep["nreads-"]) == 500
    assert np.allclose(rep["VEF"], [0.08333333333, 0.11475409836, 0.05376344086, 0.10714285714, 0.13207547170], rtol=0, atol=1e-10)
    assert np.isnan(ea.generateAmpliconReport(bam, golden_bed(z), threshold_reads=False)["VEF"]).all()
    ecdfs = ea.generateBedEcdf(bam, golden_bed(z), bed_rows=(1, 2))
    got = [ea.ecdf_at(ecdfs[r][k], 0.5) for r in (1, 2) for k in ("context", "out.of.context")]
    assert np.allclose(got, [0.916666666667, 1, 0.885245901639, 1], rtol=0, atol=1e-8)
    cap, zc = load_golden("capture")
    rep = ea.generateCaptureReport(cap, golden_bed(zc))
    assert rep["VEF"].shape[0] == 565 and np.nansum(rep["nreads-"]) == 1472 and np.nansum(rep["nreads+"]) + np.nansum(rep["nreads-"]) == 2968
    bam.close(); cap.close()


def test_bed_matching_random_targets_against_oracle():
    rng = np.random.default_rng(11)
    cases = [load_golden("capture")[0], ea.synth_host(ea.synth_config(3, 20000, [60000, 50000], 80, 420)),
             ea.synth_host(ea.synth_config(4, 30000, [80000], 100, 300, n_islands=3, island_len=1500))]
    for bam in cases:
        port = PortCpu(bam)
        for nbed in (0, 1, 50, 3000):
            bed = random_bed(rng, bam, nbed)
            for par in (-3, 0, 1, 60):
                assert np.array_equal(ea.rcpp_match_amplicon(bam, bed, par), port.rcpp_match_amplicon(bed, par)), (nbed, par, "amplicon")
                assert np.array_equal(ea.rcpp_match_capture(bam, bed, par), port.rcpp_match_capture(bed, par)), (nbed, par, "capture")
            want_m, want = port.generateBedReport(bed, "capture", match_min_overlap=20, threshold_context="CxG", min_context_sites=1,
                                                  min_context_beta=0.2, max_outofcontext_beta=0.6)
            got_m, got = ea.generateBedReport(bam, bed, bed_type="capture", match_min_overlap=20, threshold_context="CxG", min_context_sites=1,
                                              min_context_beta=0.2, max_outofcontext_beta=0.6, return_match=True)
            assert np.array_equal(got_m, want_m) and np.array_equal(got["counts"], want["counts"])
            assert np.array_equal(got["VEF"], want["VEF"], equal_nan=True)
        bam.close()


# ---- lMHL report: ea_mhl_report ------------------------------------------------
from helpers import MHL_CASES, assert_mhl_equal  # noqa: E402


@pytest.mark.parametrize("name", golden_names())
def test_mhl_report_golden_vectors(name, kernel_policy):
    if kernel_policy not in ("auto", "byte-auto"):
        pytest.skip("the lMHL report has two kernels: planar call planes (auto) and byte rows (byte-auto)")
    bam, z = load_golden(name)
    for cname, hctx, hmax, hmin, moo in MHL_CASES:
        rep = ea.generateMhlReport(bam, haplotype_context=hctx, max_haplotype_window=hmax, min_haplotype_length=hmin, max_outofcontext_beta=moo)
        assert_mhl_equal(rep, {k: z[cname + "." + k] for k in ("rname", "strand", "pos", "context", "coverage", "length", "lmhl")}, "%s/%s" % (name, cname))
    bam.close()


def test_mhl_report_identity_and_synthetic(kernel_policy):
    if kernel_policy not in ("auto", "byte-auto"):
        pytest.skip("the lMHL report has two kernels: planar call planes (auto) and byte rows (byte-auto)")
    # inst/unitTests/test_generateMhlReport.R:7-12: window 1 == meth/(meth+unmeth) of the unthresholded CG report, identical()
    bam, _ = load_golden("capture")
    m1 = ea.generateMhlReport(bam, max_haplotype_window=1)
    cg = ea.generateCytosineReport(bam, threshold_reads=False)
    assert np.array_equal(m1["pos"], cg["pos"]) and np.array_equal(m1["lmhl"], cg["meth"] / (cg["meth"] + cg["unmeth"]))
    bam.close()
    for shape in ("wgbs150", "pe_merged", "capture_islands", "tiny_reads", "long_reads"):
        host = ea.synth_host(ea.synth_config(**SYNTH[shape]))
        oracle = PortCpu(host)
        for hctx, hmax, hmin, moo in (("CG", 0, 0, 0.1), ("CX", 2, 3, 0.6), ("CHG", 0, 1, 1.0), ("CxG", 5, 0, 0.4), ("CHH", 0, 2, 0.9)):
            want = oracle.generateMhlReport(hctx, hmax, hmin, moo)
            got = ea.generateMhlReport(host, haplotype_context=hctx, max_haplotype_window=hmax, min_haplotype_length=hmin, max_outofcontext_beta=moo)
            assert want["pos"].shape[0] > 0
            assert_mhl_equal(got, want, "%s %s" % (shape, hctx))
        # letter sets that are not whole Hh/Xx/Zz pairs take the byte-serial sweep (rcpp_mhl_report is called directly)
        for ctx, hmax, hmin, moo in (("Z", 0, 0, 0.5), ("zX", 3, 0, 1.0), ("ZzU", 0, 2, 0.3)):
            want = oracle.rcpp_mhl_report(ctx, hmax, hmin, moo)
            got = ea.rcpp_mhl_report(host, ctx, hmax, hmin, moo)
            assert_mhl_equal(got, want, "%s ctx=%s" % (shape, ctx))
        host.close()


# ---- VCF-assisted report: ea_base_freqs / ea_fisher_exact --------------------------
from helpers import VCF_CASES, VCF_COLS, golden_vcf, random_vcf  # noqa: E402


@pytest.mark.parametrize("name", sorted(VCF_CASES))
def test_vcf_report_golden_vectors(name):
    bam, z = load_golden(name)
    vcf = golden_vcf(z)
    for cname, thr in VCF_CASES[name]:
        rep = ea.generateVcfReport(bam, vcf, threshold_reads=thr)
        for k in VCF_COLS:
            assert np.array_equal(rep[k], z["vcf." + cname + "." + k], equal_nan=True), (name, cname, k)
        for k in ("FEp+", "FEp-"):  # same walk as the reference build's kt_fisher_exact stand-in
            assert np.array_equal(rep[k], z["vcf." + cname + "." + k], equal_nan=True), (name, cname, k)
        pass_ = z["cg_default.pass"] if thr else None
        assert np.array_equal(ea.rcpp_get_base_freqs(bam, pass_, vcf), z["vcf." + cname + ".freqs"]), (name, cname)
    if name == "amplicon010meth":  # inst/unitTests/test_generateVcfReport.R:56-74
        rep = ea.generateVcfReport(bam, vcf)
        assert np.nansum(rep["FEp+"]) == 40 and abs(np.nansum(rep["FEp-"]) - 40.15024191) < 1e-8
        assert np.nansum(rep["SumRef"]) == 5282 and np.nansum(rep["SumAlt"]) == 14


def test_base_freqs_random_vcfs_against_oracle():
    """Sorted random SNV tables with duplicates and NA chromosomes (which hide earlier entries from later reads in the
    reference's cursor walk, rcpp_get_base_freqs.cpp:44-47) on fixtures and synthetic reads, random pass flags."""
    rng = np.random.default_rng(23)
    cases = [load_golden("capture")[0], load_golden("amplicon010meth")[0],
             ea.synth_host(ea.synth_config(5, 20000, [150000, 90000, 40000], 150, 450)),
             ea.synth_host(ea.synth_config(9, 30000, [30000], 1, 40)),
             ea.synth_host(ea.synth_config(11, 70000, [20000], 120, 260, n_islands=1, island_len=600))]
    for bam in cases:
        o = PortCpu(bam)
        for n, na in ((0, 0.0), (1, 0.0), (300, 0.0), (5000, 0.05), (2000, 0.5)):
            vcf = random_vcf(rng, bam, n, na_frac=na)
            pass_ = rng.random(bam.nrow) < 0.4
            want = o.rcpp_get_base_freqs(pass_, (vcf["seqnames"], vcf["start"]))
            got = ea.rcpp_get_base_freqs(bam, pass_, vcf)
            assert got.shape == (n, 20) and np.array_equal(got, want), (bam.nrow, n, na)
        want = o.rcpp_get_base_freqs(np.ones(bam.nrow, bool), (vcf["seqnames"], vcf["start"]))
        assert np.array_equal(ea.rcpp_get_base_freqs(bam, None, vcf), want)  # pass = NULL: every read passes
    with pytest.raises(ea.EaError) as e:  # unsorted entries: the reference's stated precondition
        ea.rcpp_get_base_freqs(cases[0], None, ([1, 1], [500, 400]))
    assert e.value.code == -3


# ---- report writer: ea_cx_report_text ---------------------------------------------
from helpers import format_report_text  # noqa: E402


def test_report_text_matches_fwrite_format(tmp_path):
    """The device-formatted report is byte for byte what the reference's .writeReport (data.table::fwrite, tab separated,
    factor labels, header) writes for the oracle's rows; through the file path of generateCytosineReport too (plain and gzip)."""
    import gzip
    for name, kw in (("amplicon010meth", dict()), ("capture", dict(threshold_reads=False, report_context="CX")), ("dragen_se", dict(report_context="CHG"))):
        bam, z = load_golden(name)
        want_p, want = PortCpu(bam).generateCytosineReport(kw.get("threshold_reads", True), "CG", 2, 0.5, 0.1, kw.get("report_context", "CG"))
        expect = format_report_text(want, bam.rname_levels)
        ea.generateCytosineReport(bam, **kw)
        assert bytes(ea.report_text(bam)) == expect, name
        assert bytes(ea.report_text(bam, header=False)) == format_report_text(want, bam.rname_levels, header=False), name
        f1, f2 = tmp_path / (name + ".tsv"), tmp_path / (name + ".tsv.gz")
        assert ea.generateCytosineReport(bam, report_file=str(f1), **kw) is None
        assert ea.generateCytosineReport(bam, report_file=str(f2), gzip=True, **kw) is None
        assert f1.read_bytes() == expect and gzip.open(f2, "rb").read() == expect, name
        bam.close()
    # an empty report is the header line; large positions and counts keep all their digits
    empty = ea.BamData.from_strings([], [], [], [], [], rname_levels=["chr1"])
    ea.generateCytosineReport(empty)
    assert bytes(ea.report_text(empty)) == b"rname\tstrand\tpos\tcontext\tmeth\tunmeth\n"
    rows = [(2, 1, 2000000000, "Z"), (2, 2, 2000000001, "zZ" * 20)]
    b = bam_from_xm(rows)
    b.rname_levels = ["chrA", "a-very-long-contig-name|with|pipes"]
    rep = ea.generateCytosineReport(b, threshold_reads=False)
    assert bytes(ea.report_text(b)) == format_report_text(rep, b.rname_levels)
