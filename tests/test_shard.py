"""CPU-only: the N>1 path.  (1) shard planning + halo + ownership filter reproduce the unsharded report
for any world size; (2) the same through torch.distributed (gloo, world_size 2) including the row-count
all-gather.  The per-shard engine here is the CPU oracle; the GPU engine is covered by test_gpu_*."""
import os
import subprocess
import sys

import numpy as np
import pytest

import epialleler_b200 as ea
from epialleler_b200.distributed import shard_report
from epialleler_b200.shard import merge_reports, plan_shards
from helpers import assert_report_equal, load_golden
from oracle.oracle import PortCpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_engine(sub, threshold_reads=True, threshold_context="CG", min_context_sites=2, min_context_beta=0.5,
                  max_outofcontext_beta=0.1, report_context=None):
    return PortCpu(sub).generateCytosineReport(threshold_reads, threshold_context, min_context_sites, min_context_beta,
                                               max_outofcontext_beta, report_context or threshold_context)


def check(bam, world, **kw):
    want_pass, want = oracle_engine(bam, **kw)
    parts, passes = [], []
    for r in range(world):
        sh, own_pass, rep = shard_report(bam, r, world, engine=oracle_engine, **kw)
        assert sh.row_lo <= sh.own_lo <= sh.own_hi
        parts.append(rep)
        passes.append(own_pass)
    assert np.array_equal(np.concatenate(passes), want_pass)
    assert_report_equal(merge_reports(parts), want, "world=%d" % world)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_sharded_equals_whole_synthetic(world):
    check(ea.synth_host(ea.synth_config(31, 5000, [60000], 150)), world)
    check(ea.synth_host(ea.synth_config(32, 4000, [30000, 30000, 9000], 30, 500)), world, threshold_reads=False, report_context="CX")
    check(ea.synth_host(ea.synth_config(33, 6000, [50000], 100, 417, n_islands=2, island_len=1200)), world, threshold_context="CxG", report_context="CX")


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_equals_whole_fixtures(world):
    for name in ("amplicon010meth", "capture"):
        bam, _ = load_golden(name)
        check(bam, world)
        check(bam, world, threshold_reads=False, report_context="CX")


def test_shard_plan_properties():
    bam = ea.synth_host(ea.synth_config(34, 3000, [20000], 100, 200))
    lens = np.diff(bam.blob()[1])
    for world in (1, 2, 7, 64):
        sh = plan_shards(bam.rname, bam.start, lens, world)
        assert sh[0].own_lo == 0 and sh[-1].own_hi == bam.nrow
        assert all(a.own_hi == b.own_lo for a, b in zip(sh, sh[1:]))
        assert all(a.key_hi <= b.key_lo or b.own_lo >= bam.nrow for a, b in zip(sh, sh[1:]))
    # more ranks than rows, and an empty input
    tiny = ea.synth_host(ea.synth_config(35, 3, [5000], 50))
    check(tiny, 8)
    assert len(plan_shards(np.zeros(0), np.zeros(0), np.zeros(0), 4)) == 4


def test_native_planner_equals_python_planner():
    """ea_plan_shards (binary searches over the offsets + a backward halo walk, no per-row temporaries) == plan_shards."""
    from epialleler_b200.shard import plan_shards_native
    rng = np.random.default_rng(0)
    for trial in range(200):
        n, world = int(rng.integers(0, 400)), int(rng.integers(1, 9))
        rn = rng.integers(1, 4, n).astype(np.int32)
        st = rng.integers(1, 300, n).astype(np.int32)
        o = np.lexsort((st, rn))
        rn, st = rn[o], st[o]
        ln = rng.integers(0, 60, n).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)
        tid = np.arange(n, dtype=np.int32) if trial % 2 == 0 else rng.permutation(n).astype(np.int32)
        bam = ea.BamData(rn, np.ones(n, np.int32), st, tid, np.zeros(int(off[-1]), np.uint8), off)
        want = plan_shards(rn, st, ln[tid], world)
        got = plan_shards_native(bam, world)
        assert [(w.row_lo, w.own_lo, w.own_hi, w.key_lo, w.key_hi) for w in want] == [(g.row_lo, g.own_lo, g.own_hi, g.key_lo, g.key_hi) for g in got], (trial, n, world)


def test_native_planner_many_pieces():
    """Rows that reach their strings through templid: the native planner sums PIECES of the rows and scans inside the piece a
    cut falls into (no per-row prefix array) - several pieces, zero-length strings, ties at the cuts."""
    from epialleler_b200.shard import plan_shards_native
    rng = np.random.default_rng(1)
    n = 300_000
    rn = np.sort(rng.integers(1, 4, n)).astype(np.int32)
    st = rng.integers(1, 5000, n).astype(np.int32)
    o = np.lexsort((st, rn))
    rn, st = rn[o], st[o]
    ln = rng.integers(0, 40, n).astype(np.int64)
    ln[rng.random(n) < 0.3] = 0
    off = np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)
    tid = rng.permutation(n).astype(np.int32)
    bam = ea.BamData(rn, np.ones(n, np.int32), st, tid, np.zeros(int(off[-1]), np.uint8), off)
    for world in (2, 5, 8):
        want = plan_shards(rn, st, ln[tid], world)
        got = plan_shards_native(bam, world)
        assert [(w.row_lo, w.own_lo, w.own_hi, w.key_lo, w.key_hi) for w in want] == [(g.row_lo, g.own_lo, g.own_hi, g.key_lo, g.key_hi) for g in got], world


def test_native_planner_on_a_slice_of_a_larger_blob_and_degenerate_inputs():
    """Offsets that do not start at 0 (a shard of a larger blob is planned where it lies), more shards than rows, one shard,
    no rows; a templid outside the strings is an error."""
    import ctypes as C

    from epialleler_b200 import _native as N
    L = N.lib()
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rng = np.random.default_rng(2)
    n = 5000
    rn = np.ones(n, np.int32)
    st = np.sort(rng.integers(1, 100000, n)).astype(np.int32)
    ln = rng.integers(1, 200, n).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)
    lo, hi = 1200, 4100
    for world in (1, 2, 6):
        want = plan_shards(rn[lo:hi], st[lo:hi], ln[lo:hi], world)
        out = (N.Shard * world)()
        N.check(L.ea_plan_shards(p(rn[lo:], C.c_int32), p(st[lo:], C.c_int32), None, hi - lo, p(off[lo:], C.c_int64), hi - lo, world, out))
        assert [(w.row_lo, w.own_lo, w.own_hi, w.key_lo, w.key_hi) for w in want] == [(g.row_lo, g.own_lo, g.own_hi, g.key_lo, g.key_hi) for g in out], world
    for nrow, world in ((3, 8), (0, 4), (1, 1)):
        want = plan_shards(rn[:nrow], st[:nrow], ln[:nrow], world)
        out = (N.Shard * world)()
        N.check(L.ea_plan_shards(p(rn, C.c_int32), p(st, C.c_int32), None, nrow, p(off, C.c_int64), nrow, world, out))
        assert [(w.row_lo, w.own_lo, w.own_hi, w.key_lo, w.key_hi) for w in want] == [(g.row_lo, g.own_lo, g.own_hi, g.key_lo, g.key_hi) for g in out], (nrow, world)
    bad = np.arange(n, dtype=np.int32)
    bad[10] = n + 3
    out = (N.Shard * 2)()
    assert L.ea_plan_shards(p(rn, C.c_int32), p(st, C.c_int32), p(bad, C.c_int32), n, p(off, C.c_int64), n, 2, out) == N.EA_ERR_ARG


def test_windowed_planner_equals_python_planner():
    from epialleler_b200.shard import plan_shards_windowed
    for cfg in (ea.synth_config(36, 20000, [300000, 100000], 150), ea.synth_config(37, 5000, [c // 30000 for c in (248956422, 242193529, 198295559, 190214555)], 150)):
        bam = ea.synth_host(cfg)
        lens = np.diff(bam.blob()[1])
        fetch = lambda a, b: (bam.rname[a:b], bam.start[a:b], lens[a:b])
        for world in (1, 2, 3, 8):
            want = plan_shards(bam.rname, bam.start, lens, world, weights=np.ones(bam.nrow, np.int64))
            got = plan_shards_windowed(bam.nrow, world, fetch, 150, window=512)
            assert [(w.row_lo, w.own_lo, w.own_hi, w.key_lo, w.key_hi) for w in want] == [(g.row_lo, g.own_lo, g.own_hi, g.key_lo, g.key_hi) for g in got], world


def test_gloo_world_size_2():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_slice_bam_large_and_shuffled():
    """slice_bam must not build per-byte index arrays for a whole shard: 10 M rows go through the zero-copy view
    (rows == strings in file order), and a shuffled templid goes through the chunked gather; both equal a direct gather."""
    import tracemalloc
    from epialleler_b200.shard import Shard, slice_bam
    n = 10_000_000
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 12, n).astype(np.int64)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    blob = rng.integers(0, 256, int(off[-1]), dtype=np.uint8)
    start = np.arange(n, dtype=np.int32)
    ones = np.ones(n, np.int32)
    bam = ea.BamData(ones, ones, start, np.arange(n, dtype=np.int32), blob, off)
    sh = Shard(1, 3_000_001, 3_000_100, 9_999_999, 0, 0)
    tracemalloc.start()
    sub = slice_bam(bam, sh)
    peak = tracemalloc.get_traced_memory()[1]
    tracemalloc.stop()
    b2, o2 = sub.blob()
    assert peak < 4 * (sh.own_hi - sh.row_lo) * 8 + (64 << 20), "slice of contiguous strings must not copy per byte (peak %d)" % peak
    assert np.shares_memory(b2, blob)
    assert o2[0] == 0 and np.array_equal(np.diff(o2), lens[sh.row_lo:sh.own_hi])
    assert np.array_equal(b2, blob[off[sh.row_lo]:off[sh.own_hi]])
    assert np.array_equal(sub.templid, np.arange(sh.own_hi - sh.row_lo))
    # shuffled templid: strings are not in row order
    m = 2_000_000
    perm = rng.permutation(m).astype(np.int32)
    bam2 = ea.BamData(ones[:m], ones[:m], start[:m], perm, blob[:off[m]], off[:m + 1])
    sh2 = Shard(0, 100_003, 100_003, 1_900_001, 0, 0)
    sub2 = slice_bam(bam2, sh2)
    b3, o3 = sub2.blob()
    tid = perm[sh2.row_lo:sh2.own_hi].astype(np.int64)
    assert np.array_equal(np.diff(o3), lens[tid])
    for i in (0, 1, 777, 1_234_567, tid.shape[0] - 1):
        assert np.array_equal(b3[o3[i]:o3[i + 1]], blob[off[tid[i]]:off[tid[i] + 1]])
    import zlib
    want = zlib.crc32(np.concatenate([blob[off[t]:off[t + 1]] for t in tid[:50_000]]))
    assert zlib.crc32(b3[:o3[50_000]]) == want
