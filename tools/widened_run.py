"""Runs the widened paths (lMHL, BED, base frequencies, report text) once on resident synthetic reads - for ncu launch lists."""
import ctypes as C
import sys

sys.path.insert(0, ".")
import numpy as np
import epialleler_b200 as ea
from epialleler_b200 import _native as N
from epialleler_b200.api import _params
import bench

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cfg = bench.synth_cfg(ea, 1, reads)
dev = ea.synth_device(cfg, 0, reads, device=0)
L = N.lib()
p = _params(**bench.THRESH)
i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
nrow = C.c_int64(0)
rng = np.random.default_rng(1)
chr_len = int(cfg.chr_len[0])
for _ in range(2):
    N.check(L.ea_mhl_report(dev.handle(), b"Zz", 0, 0, 0.1, C.byref(nrow)))
    nb = 20000
    bst = np.sort(rng.integers(1, chr_len - 3000, nb)).astype(np.int32)
    bed = (np.ones(nb, np.int32), bst, (bst + 2000).astype(np.int32))
    counts = np.zeros((nb + 1) * 4, np.int64)
    N.check(L.ea_bed_report(dev.handle(), 1, i32(bed[0]), i32(bed[1]), i32(bed[2]), nb, 1, C.byref(p), None, None, None, counts.ctypes.data_as(C.POINTER(C.c_int64))))
    nv = 1_000_000
    vpos = np.sort(rng.integers(1, chr_len, nv)).astype(np.int32)
    freqs = np.zeros(nv * 20, np.float64)
    N.check(L.ea_base_freqs(dev.handle(), None, i32(np.ones(nv, np.int32)), i32(vpos), nv, freqs.ctypes.data_as(C.POINTER(C.c_double))))
    N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 0, None))
    dev.rname_levels = ["chr1"]
    txt = ea.report_text(dev)
print("ok", nrow.value, int(counts.sum()), int(freqs.sum()), txt.shape[0])
