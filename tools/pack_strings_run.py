"""Times the two scattered-string packers on the bench workload through integration/harness (the reader's heap
std::strings): ea_pack_calls (call planes only, what the shim calls) and ea_pack (byte rows too: VCF base frequencies).
The first call of each runs right after ea_trim() dropped the cached page-locked blocks.  Usage: python tools/pack_strings_run.py [reads]"""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
import epialleler_b200 as ea  # noqa: E402
import integration.harness as H  # noqa: E402
from epialleler_b200 import _native as N  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cfg = ea.synth_config(**bench.headline_synth_kwargs(reads))
dev = ea.synth_device(cfg, 0, reads, device=0)
host = dev.to_host(); dev.close()
rs = H.ReaderStrings(host)
L = N.lib()
for calls_only in (True, False):
    L.ea_trim(0)
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d = rs.pack(calls_only=calls_only)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print("%s call %d: %d strings, %.2f GB: %.1f ms (bytes over PCIe %.2f GB)" % ("ea_pack_calls" if calls_only else "ea_pack      ", it, host.nrow,
              host.blob()[0].nbytes / 1e9, (t1 - t0) * 1e3, L.ea_pack_h2d_bytes(d.handle()) / 1e9), flush=True)
        d.close()
