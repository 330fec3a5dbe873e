"""Times ea_pack (the scattered-strings packer the R shim uses: std::vector<std::string> -> pinned blob -> device) on the
bench workload; the strings here are slices of one host blob, addressed individually like std::string buffers."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import epialleler_b200 as ea
from epialleler_b200 import _native as N
import bench

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cfg = bench.synth_cfg(ea, 1, reads)
dev = ea.synth_device(cfg, 0, reads, device=0)
host = dev.to_host(); dev.close()
blob, off = host.blob()
ptrs = (blob.ctypes.data + off[:-1]).astype(np.uint64)
lens = np.diff(off).astype(np.int64)
tid = np.arange(host.nrow, dtype=np.int32)
L = N.lib()
i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
for it in range(3):
    h = C.c_void_p()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    N.check(L.ea_pack(i32(host.rname), i32(host.strand), i32(host.start), i32(tid), host.nrow,
                      ptrs.ctypes.data_as(C.POINTER(C.c_void_p)), lens.ctypes.data_as(C.POINTER(C.c_int64)), host.nrow, 0, C.byref(h)))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("ea_pack %d strings, %.2f GB: %.1f ms" % (host.nrow, blob.nbytes / 1e9, (t1 - t0) * 1e3), flush=True)
    L.ea_free(h)
