import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import epialleler_b200 as ea
from epialleler_b200 import _native as N
import bench
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cfg = bench.synth_cfg(ea, 1, reads)
dev = ea.synth_device(cfg, 0, reads, device=0)
host = dev.to_host(); dev.close()
blob, off = host.blob()
src = (blob, off, host.rname, host.strand, host.start, np.arange(host.nrow, dtype=np.int32))
pin = [torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True) for a in src]
for t, a in zip(pin, src): t.numpy()[...] = a
blob, off, rn, st, sa, tid = [t.numpy() for t in pin]
# raw H2D bandwidth
d = torch.empty(blob.shape[0], dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pin[0], non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw H2D %.1f GB in %.1f ms = %.1f GB/s" % (blob.nbytes / 1e9, dt * 1e3, blob.nbytes / 1e9 / dt))
del d; torch.cuda.empty_cache()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = ea.BamData(rn, st, sa, tid, blob, off, device=0); b.handle(need_bases=False); torch.cuda.synchronize(); t1 = time.perf_counter()
    p, rep = ea.generateCytosineReport(b, return_pass=True, **bench.THRESH); torch.cuda.synchronize(); t2 = time.perf_counter()
    b.close(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("iter %d: pack %.1f ms, report+fetch %.1f ms, free %.1f ms, total %.1f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
