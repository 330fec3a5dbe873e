"""Are the sporadic ~1 s stalls of repeated pack/report/free calls caused by the device pool growing?  Runs the shim path
12 times (a) with the pool warm from a 12 GB handle that was freed but not trimmed, (b) after ea_trim()."""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
import epialleler_b200 as ea  # noqa: E402
import integration.harness as H  # noqa: E402
from epialleler_b200 import _native as N  # noqa: E402

reads = 50_000_000
cfg = ea.synth_config(**bench.headline_synth_kwargs(reads))
dev = ea.synth_device(cfg)
host = dev.to_host()
rs = H.ReaderStrings(host)
L = N.lib()
dev.close()   # 12 GB go back to the library's pool, not to the driver
for label in ("pool warm (12 GB cached)", "after ea_trim()"):
    if label.startswith("after"):
        L.ea_trim(0)
    ts = []
    for i in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d = rs.pack(calls_only=True)
        p, rep = ea.generateCytosineReport(d, return_pass=True)
        d.close(); torch.cuda.synchronize()
        ts.append(round((time.perf_counter() - t0) * 1e3))
        del p, rep
    print(label, ts, flush=True)
