import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import epialleler_b200 as ea
from epialleler_b200 import _native as N
from helpers import load_golden, COLS
bam, z = load_golden("amplicon000meth")
p = np.ones(bam.nrow, bool)
want = {k: z["cx_nothr." + k] for k in COLS}
print("expected rows", want["pos"].shape[0], "ctx counts", np.unique(want["context"], return_counts=True), "strand", np.unique(want["strand"], return_counts=True))
for pol in (0, 1):
    N.lib().ea_set_kernel_policy(pol)
    for ctx in ("ZXH", "Z", "X", "H", "ZX", "XH"):
        rep = ea.rcpp_cx_report(bam, p, ctx)
        print("policy", pol, ctx, rep["pos"].shape[0], np.unique(rep["context"], return_counts=True), np.unique(rep["strand"], return_counts=True))
