"""GPU tests at BASELINE.json's full size (configs[1]: 50 M reads x 150 bp, one chromosome, CpG report, CG thresholds).
The CPU oracle cannot finish this size in seconds, so the checks are the size-independent properties the domain offers:
  * the fused single-pass report equals the reference's own two-call sequence (rcpp_threshold_reads, then rcpp_cx_report)
    - a second, independently scheduled computation of every count on the device;
  * every window kernel (kernel policy fixture) produces the same bytes;
  * rows are strictly ordered by (pos, strand) with no duplicates, as the reference's flat_map iteration emits them
    (src/rcpp_cx_report.cpp:59-60), meth + unmeth > floor(cov/2) >= 0 implies meth + unmeth >= 1, context is the requested one;
  * a PREFIX of the report is bit-exact against the CPU oracle: rows whose position lies before the start of read K only
    depend on reads 0..K-1 (rows are sorted by start), so the oracle on the first K reads pins them;
  * the pass-flag count returned by the device pipeline equals the number of TRUE flags fetched."""
import ctypes as C
import zlib

import numpy as np
import pytest

import epialleler_b200 as ea
from epialleler_b200 import _native as N
from epialleler_b200.api import _params
import os

from oracle.oracle import REF_SO, PortCpu, RefCpu

# the compiled reference itself whenever it travelled with the snapshot (oracle/_ref is git-ignored, not gpurun-ignored)
Checker = RefCpu if os.path.exists(REF_SO) else PortCpu

pytestmark = pytest.mark.gpu

CHR1 = 248_956_422
READS = 50_000_000
SEED = 20261018
THRESH = dict(threshold_reads=True, threshold_context="CG", min_context_sites=2, min_context_beta=0.5, max_outofcontext_beta=0.1,
              report_context="CG")
PREFIX = 300_000
_cache = {}


def crc(rep, pass_flags):
    c = zlib.crc32(np.ascontiguousarray(pass_flags).view(np.uint8))
    for k in ("rname", "strand", "pos", "context", "meth", "unmeth"):
        c = zlib.crc32(np.ascontiguousarray(rep[k]).view(np.uint8), c)
    return c


@pytest.fixture(params=["auto", "vertical", "byte-auto", "lane-group", "thread-per-row-tma", "chunk-per-lane", "long-read"])
def kernel_policy(request):
    old = N.lib().ea_set_kernel_policy({"auto": 0, "lane-group": 1, "thread-per-row-tma": 2, "chunk-per-lane": 3, "long-read": 4, "byte-auto": 5, "vertical": 6}[request.param])
    yield request.param
    N.lib().ea_set_kernel_policy(old)


def test_config2_full_size_properties(kernel_policy):
    cfg = ea.synth_config(SEED, READS, [CHR1], 150)
    dev = ea.synth_device(cfg)
    assert dev.nrow == READS and dev.nbytes == READS * 150
    p, rep = ea.generateCytosineReport(dev, return_pass=True, **THRESH)
    n = rep["pos"].shape[0]
    assert n > 9_000_000
    # the device-side counters agree with what was fetched
    tm = N.Timing()
    prm = _params(**THRESH)
    N.check(N.lib().ea_cytosine_report_device(dev.handle(), C.byref(prm), 0, C.byref(tm)))
    assert int(tm.n_rows) == n and int(tm.n_pass) == int(p.sum())
    assert int(tm.alg_bytes) == READS * 150 + 14 * READS + 24 * n

    # ordering / uniqueness / value ranges
    key = rep["pos"].astype(np.int64) * 2 + (rep["strand"].astype(np.int64) - 1)
    assert np.all(np.diff(key) > 0), "rows must be strictly ordered by (pos, strand)"
    assert np.all(rep["rname"] == 1) and np.all(rep["context"] == 7)
    assert np.all((rep["strand"] == 1) | (rep["strand"] == 2))
    assert np.all(rep["meth"] >= 0) and np.all(rep["unmeth"] >= 0) and np.all(rep["meth"] + rep["unmeth"] >= 1)
    assert rep["pos"].min() >= 1 and rep["pos"].max() <= CHR1

    # fused == the reference's two-call sequence, bit for bit
    p2 = ea.rcpp_threshold_reads(dev, "Z", "z", "XH", "xh", 2, 0.5, 0.1)
    assert np.array_equal(p, p2)
    rep2 = ea.rcpp_cx_report(dev, p2, "Z")
    for k in ("rname", "strand", "pos", "context", "meth", "unmeth"):
        assert np.array_equal(rep[k], rep2[k]), k

    # every kernel policy yields the same bytes
    _cache[kernel_policy] = crc(rep, p)
    assert len(set(_cache.values())) == 1, _cache

    # prefix against the CPU oracle
    host = ea.synth_host(cfg, 0, PREFIX)
    want_p, want = Checker(host).generateCytosineReport(True, "CG", 2, 0.5, 0.1, "CG")
    assert np.array_equal(p[:PREFIX], want_p)
    limit = int(host.start[-1])  # rows at positions < start of read PREFIX-1 are complete in the prefix
    k_dev = int(np.searchsorted(rep["pos"], limit, side="left"))
    k_cpu = int(np.searchsorted(want["pos"], limit, side="left"))
    assert k_dev == k_cpu and k_dev > 10_000
    for k in ("rname", "strand", "pos", "context", "meth", "unmeth"):
        assert np.array_equal(rep[k][:k_dev], want[k][:k_cpu]), k
    host.close()

    # three more windows against the oracle: the middle of the chromosome, a stretch straddling a multiple of every window
    # size the kernels use (2^24), and the last rows (int32 positions near 2.49e8, the last windows, `hi` clamping).  The
    # sub-range generator gives rows [a, b); report rows at positions covered ONLY by those rows - from the furthest end of
    # the rows before a (< start[a] + 150) up to the start of row b-1 - must equal the oracle's on the sub-range.
    def window(a, b):
        sub = ea.synth_host(cfg, a, b)
        wp, w = Checker(sub).generateCytosineReport(True, "CG", 2, 0.5, 0.1, "CG")
        assert np.array_equal(p[a:b], wp)
        lo_pos = int(sub.start[0]) + 150 if a > 0 else 0
        hi_pos = int(sub.start[-1]) if b < READS else CHR1 + 1
        sel_d = (rep["pos"] >= lo_pos) & (rep["pos"] < hi_pos)
        sel_c = (w["pos"] >= lo_pos) & (w["pos"] < hi_pos)
        assert int(sel_c.sum()) > 10_000
        for k in ("rname", "strand", "pos", "context", "meth", "unmeth"):
            assert np.array_equal(rep[k][sel_d], w[k][sel_c]), (a, b, k)
        sub.close()
    window(READS // 2 - PREFIX // 2, READS // 2 + PREFIX // 2)
    r_edge = int(np.searchsorted(ea.synth_host(cfg, 3_300_000, 3_500_000).start, 1 << 24)) + 3_300_000   # first row starting beyond position 2^24
    window(r_edge - PREFIX // 2, r_edge + PREFIX // 2)
    window(READS - PREFIX, READS)
    dev.close()
