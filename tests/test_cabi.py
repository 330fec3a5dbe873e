"""CPU-only: the C-ABI library builds, loads and exports every symbol include/epialleler_b200.h declares;
without a GPU every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import epialleler_b200 as ea
from epialleler_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "epialleler_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ea_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = N.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "library lacks %s" % s
        assert s in N.SIGNATURES, "python binding lacks %s" % s
    assert L.ea_version() >= 100


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import, link or execute it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "epialleler_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle/|liboracle|libepiref)", text), f


def test_no_cpu_fallback_without_device():
    if N.lib().ea_device_count() > 0:
        pytest.skip("a GPU is present")
    bam = ea.BamData.from_strings([1], [1], [1], [0], [bytes([0x17, 0x1F])])
    with pytest.raises(ea.EaError) as e:
        ea.rcpp_threshold_reads(bam, "Z", "z", "XH", "xh", 2, 0.5, 0.1)
    assert e.value.code == -5
    with pytest.raises(ea.EaError):
        ea.generateCytosineReport(bam)
    with pytest.raises(ea.EaError):
        ea.synth_device(ea.synth_config(1, 10, [10000], 100))


def test_argument_checks_mirror_the_reference():
    with pytest.raises(ValueError):  # match.arg failure (R/generateCytosineReport.R:176-177)
        ea.generateCytosineReport(ea.BamData.from_strings([], [], [], [], []), threshold_context="CpG")
    with pytest.raises(NotImplementedError):  # BAM decoding stays with HTSlib
        ea.preprocessBam("/nonexistent.bam")
    L = N.lib()
    assert L.ea_cx_report_fetch(None, None, None, None, None, None, None) == -1
    assert b"null" in L.ea_last_error()


def test_synth_host_is_deterministic_sorted_and_in_domain():
    cfg = ea.synth_config(20261018, 5000, [200000, 100000], 100, 417, n_islands=10, island_len=2000)
    a, b = ea.synth_host(cfg), ea.synth_host(cfg)
    assert np.array_equal(a.blob()[0], b.blob()[0]) and np.array_equal(a.start, b.start)
    key = a.rname.astype(np.int64) << 32 | a.start
    assert (np.diff(key) >= 0).all() and set(np.unique(a.rname)) == {1, 2}
    lens = np.diff(a.blob()[1])
    assert lens.min() >= 100 and lens.max() <= 417
    assert set(np.unique(a.blob()[0] & 15)) <= {2, 5, 6, 7, 10, 11, 12, 13, 14, 15}
    # a sub-range regenerates the same rows
    sub = ea.synth_host(cfg, 1000, 1100)
    o = a.blob()[1]
    assert np.array_equal(sub.blob()[0], a.blob()[0][o[1000]:o[1100]]) and np.array_equal(sub.start, a.start[1000:1100])


def test_reference_arm_runs_without_the_product_library():
    """bench.py --impl reference: the compiled reference (or the C port) on rows from the oracle's own generator build; the
    line carries the contract's keys and the process maps no library of the product."""
    import json
    import subprocess
    import sys
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--ref-rows', '20000'];"
            "runpy.run_path('bench.py', run_name='__main__');"
            "print('MAPPED', [l.split()[-1] for l in open('/proc/self/maps') if 'libepialleler_b200' in l], file=sys.stderr)")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["cores"] == 1
    assert "MAPPED []" in out.stderr, out.stderr[-500:]
