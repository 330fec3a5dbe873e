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


def test_gloo_world_size_2():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
