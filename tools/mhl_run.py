import sys, ctypes as C, time
sys.path.insert(0, ".")
import torch
import epialleler_b200 as ea
from epialleler_b200 import _native as N
import bench
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cfg = ea.synth_config(**bench.headline_synth_kwargs(reads))
dev = ea.synth_device(cfg, 0, reads, device=0)
L = N.lib(); nrow = C.c_int64(0)
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    N.check(L.ea_mhl_report(dev.handle(), b"Zz", 0, 0, 0.1, C.byref(nrow)))
    torch.cuda.synchronize(); print("mhl ms", (time.perf_counter() - t) * 1e3, nrow.value, (L.ea_report_kernel(dev.handle()) or b"").decode())
