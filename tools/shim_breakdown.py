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
for i in range(4):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    d = rs.pack(calls_only=True); t.append(time.perf_counter())
    p, rep = ea.generateCytosineReport(d, return_pass=True); t.append(time.perf_counter())
    d.close(); torch.cuda.synchronize(); t.append(time.perf_counter())
    del p, rep
    print("call %d: pack %.1f ms  report+fetch %.1f ms  free %.1f ms  total %.1f ms" % (i, *(1e3 * (b - a) for a, b in zip(t, t[1:])), 1e3 * (t[-1] - t[0])), flush=True)
