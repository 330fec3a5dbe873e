"""Where the end-to-end time of the shim path goes (reader's heap strings -> ea_pack_calls -> report -> fetch -> free).
EA_TRACE=1 adds the library's own stage marks.  Usage: python tools/shim_breakdown.py [reads]"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import epialleler_b200 as ea  # noqa: E402
import integration.harness as H  # noqa: E402
from epialleler_b200 import _native as N  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cfg = ea.synth_config(20261018, reads, [max(int(248956422 * reads / 50_000_000), 100000)], 150)
dev = ea.synth_device(cfg)
host = dev.to_host()
dev.close()
L = N.lib()
t0 = time.perf_counter()
rs = H.ReaderStrings(host)
print("build heap strings %.2f s" % (time.perf_counter() - t0), flush=True)
L.ea_trim(0)
for i in range(8):
    records = i >= 4   # calls 4-7: the table-free call (ea_pack_calls_records)
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    d = rs.pack(calls_only=True, records=records); t.append(time.perf_counter())
    p, rep = ea.generateCytosineReport(d, return_pass=True); t.append(time.perf_counter())
    d.close(); torch.cuda.synchronize(); t.append(time.perf_counter())
    del p, rep
    print("%s call %d: pack %.1f ms  report+fetch %.1f ms  free %.1f ms  total %.1f ms" % ("records" if records else "tables ", i, *(1e3 * (b - a) for a, b in zip(t, t[1:])), 1e3 * (t[-1] - t[0])), flush=True)

# ---- the same with results fetched into PAGEABLE arrays (what the R shim's wrap_report does) ----
import ctypes as C  # noqa: E402
from epialleler_b200.api import _params  # noqa: E402
prm = _params(True, "CG", 2, 0.5, 0.1, "CG")
i32p, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
for mode in ("fresh np.empty", "pre-touched np.zeros", "fresh np.empty"):
    for i in range(2):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        d = rs.pack(calls_only=True); t.append(time.perf_counter())
        alloc = np.empty if mode.startswith("fresh") else np.zeros
        pf = alloc(d.nrow, np.uint8); t.append(time.perf_counter())
        nrow = C.c_int64(0)
        N.check(L.ea_cytosine_report(d.handle(), C.byref(prm), pf.ctypes.data_as(u8p), C.byref(nrow))); t.append(time.perf_counter())
        cols = [alloc(max(nrow.value, 1), np.int32) for _ in range(6)]; t.append(time.perf_counter())
        N.check(L.ea_cx_report_fetch(d.handle(), *[c.ctypes.data_as(i32p) for c in cols])); t.append(time.perf_counter())
        d.close(); torch.cuda.synchronize(); t.append(time.perf_counter())
        del cols, pf
        print("%-22s pack %.1f  alloc-flags %.1f  report+flags %.1f  alloc-cols %.1f  fetch %.1f  free %.1f  total %.1f ms" % (
            mode, *(1e3 * (b - a) for a, b in zip(t, t[1:])), 1e3 * (t[-1] - t[0])), flush=True)
