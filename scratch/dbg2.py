import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import epialleler_b200 as ea
from epialleler_b200 import _native as N
from oracle.oracle import PortCpu
for n_reads, clen in ((40000, 200000), (400000, 2000000)):
    cfg = ea.synth_config(20261018, n_reads, [clen], 150)
    host = ea.synth_host(cfg)
    try:
        p, rep = ea.generateCytosineReport(host, return_pass=True)
        print(n_reads, "kernel:", N.lib().ea_report_kernel(host.handle()).decode(), "rows", rep["pos"].shape[0])
        wp, want = PortCpu(host).generateCytosineReport()
        print("  pass equal", np.array_equal(p, wp), "rows equal", all(np.array_equal(rep[k], want[k]) for k in want))
    except Exception as e:
        print(n_reads, "ERROR", e); break
