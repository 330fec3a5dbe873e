"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import epialleler_b200 as ea
from epialleler_b200 import _native as N
from oracle.oracle import PortCpu
from test_oracle import random_bed
shapes = [dict(seed=20261018, n_reads=6000, chr_len=[30000], len_min=150), dict(seed=5, n_reads=3000, chr_len=[20000, 9000], len_min=150, len_max=450),
          dict(seed=20261020, n_reads=40, chr_len=[200000], len_min=10000, len_max=30000)]
rng = np.random.default_rng(3)
for kw in shapes:
    host = ea.synth_host(ea.synth_config(**kw))
    o = PortCpu(host)
    for pol in (0, 1, 2, 3):
        N.lib().ea_set_kernel_policy(pol)
        for thr, rctx in ((True, "CG"), (False, "CX")):
            p, rep = ea.generateCytosineReport(host, threshold_reads=thr, report_context=rctx, return_pass=True)
            wp, want = o.generateCytosineReport(thr, "CG", 2, 0.5, 0.1, rctx)
            assert np.array_equal(p, wp) and all(np.array_equal(rep[k], want[k]) for k in want)
    N.lib().ea_set_kernel_policy(0)
    m = ea.generateMhlReport(host); wm = o.generateMhlReport()
    assert all(np.array_equal(m[k], wm[k], equal_nan=True) for k in wm)
    bed = random_bed(rng, host, 200)
    gm, g = ea.generateBedReport(host, bed, bed_type="capture", return_match=True); wm2, w = o.generateBedReport(bed, "capture")
    assert np.array_equal(gm, wm2) and np.array_equal(g["counts"], w["counts"])
    assert np.array_equal(ea.rcpp_get_xm_beta(host, "Z", "z"), o.rcpp_get_xm_beta("Z", "z"))
    host.close()
print("sanitize workload ok")
