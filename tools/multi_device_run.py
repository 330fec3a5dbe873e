"""One process, several devices (ea_pack_multi / ea_multi_*): what the single-threaded R process would call to use every
GPU of the box.  Packs the first `reads` rows of the config-5 genome from ONE page-locked host blob onto 1..N devices,
runs the fused report, fetches the concatenated rows, and checks that every device count gives the same bytes.
Usage: python tools/multi_device_run.py [reads] > profiles/<name>.json"""
import json
import sys
import time
import zlib

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import epialleler_b200 as ea  # noqa: E402
from epialleler_b200 import _native as N  # noqa: E402

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622,
        133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cfg = ea.synth_config(20261021, 1_000_000_000, HG38, 150)
# rows straddling the chromosome 1 -> 2 boundary, so that cuts fall inside coverage and next to a chromosome change
mid = int(1_000_000_000 * HG38[0] / sum(HG38))
host = ea.synth_host(cfg, mid - reads // 3, mid - reads // 3 + reads)
blob, off = host.blob()
src = (blob, off, host.rname, host.strand, host.start, host.templid)
pin = [torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True) for a in src]
for t, a in zip(pin, src):
    t.numpy()[...] = a
blob, off, rn, st, sa, tid = [t.numpy() for t in pin]
bam = ea.BamData(rn, st, sa, tid, blob, off)
ndev = N.lib().ea_device_count()
out = {"reads": reads, "devices_present": ndev, "runs": []}
crcs = set()
for k in [d for d in (1, 2, 4, 8) if d <= ndev]:
    times = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mb = ea.MultiBam(bam, list(range(k)))
        t1 = time.perf_counter()
        p, rep = ea.generateCytosineReport(mb, return_pass=True)
        t2 = time.perf_counter()
        sh = mb.shards()
        mb.close()
        for d in range(k):
            torch.cuda.synchronize(d)
        t3 = time.perf_counter()
        if i:
            times.append((t1 - t0, t2 - t1, t3 - t0))
    c = zlib.crc32(np.ascontiguousarray(p).view(np.uint8))
    for col in ("rname", "strand", "pos", "context", "meth", "unmeth"):
        c = zlib.crc32(np.ascontiguousarray(rep[col]).view(np.uint8), c)
    crcs.add(c)
    med = np.median(np.asarray(times), axis=0)
    out["runs"].append({"devices": k, "pack_s": float(med[0]), "report_fetch_s": float(med[1]), "total_s": float(med[2]), "reads_per_s": reads / float(med[2]),
                        "rows": int(rep["pos"].shape[0]), "halo_rows": [int(s.own_lo - s.row_lo) for s in sh], "crc": c})
# the same through the shim's calls: the reader's heap std::strings (pageable), ea_pack_multi_records
import integration.harness as H  # noqa: E402
rs = H.ReaderStrings(bam)
out["from_reader_strings"] = []
for k in [d for d in (1, 2, 4, 8) if d <= ndev]:
    times = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mb = rs.pack_multi(list(range(k)))
        t1 = time.perf_counter()
        p, rep = ea.generateCytosineReport(mb, return_pass=True)
        t2 = time.perf_counter()
        mb.close()
        for d in range(k):
            torch.cuda.synchronize(d)
        t3 = time.perf_counter()
        if i:
            times.append((t1 - t0, t2 - t1, t3 - t0))
    c = zlib.crc32(np.ascontiguousarray(p).view(np.uint8))
    for col in ("rname", "strand", "pos", "context", "meth", "unmeth"):
        c = zlib.crc32(np.ascontiguousarray(rep[col]).view(np.uint8), c)
    crcs.add(c)
    med = np.median(np.asarray(times), axis=0)
    out["from_reader_strings"].append({"devices": k, "pack_s": float(med[0]), "report_fetch_s": float(med[1]), "total_s": float(med[2]),
                                       "reads_per_s": reads / float(med[2]), "rows": int(rep["pos"].shape[0]), "crc": c})
out["identical_across_device_counts"] = len(crcs) == 1
print(json.dumps(out, indent=1))
