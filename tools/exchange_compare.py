"""what a cut through continuous coverage costs under the three candidate schemes, measured on 2 GPUs.

  (A) halo      - the rows that start before the cut but reach past it are packed on both neighbours; no data-path exchange
                  (what the package does: shard.py).  Cost = processing the halo rows a second time.
  (B) exchange  - rank g sends its partial {cov, total, meth} counters of the spill span to rank g+1 (NCCL send/recv over
                  NVLink), which adds them before finalising its first window.  Cost = one send/recv + one add per cut.
  (C) host      - the same counters go device -> host, are merged by the host processes (gloo), and come back.

The counter block is sized as SURVEY 8(e) states it: (longest read - 1) positions x 2 strands x 3 int32 {cov, total, meth}
(one report context): 449 positions for short reads (config 2/5), 29,999 for long reads (config 4).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/exchange_compare.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import epialleler_b200 as ea  # noqa: E402
from epialleler_b200 import _native as N  # noqa: E402
from epialleler_b200.api import _params  # noqa: E402


def timed(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6  # us


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gloo = dist.new_group(backend="gloo")
    out = {}
    for name, positions in (("short reads (449 positions)", 449), ("long reads (29,999 positions)", 29999)):
        n_int = positions * 2 * 3
        block = torch.ones(n_int, dtype=torch.int32, device="cuda")
        acc = torch.zeros(n_int, dtype=torch.int32, device="cuda")
        pinned = torch.empty(n_int, dtype=torch.int32, pin_memory=True)
        host_in = torch.empty(n_int, dtype=torch.int32)

        def nccl_exchange():
            if rank == 0:
                dist.send(block, dst=1)
            else:
                dist.recv(acc, src=0)
                acc.add_(block)

        def host_merge():
            if rank == 0:
                pinned.copy_(block, non_blocking=True)
                torch.cuda.synchronize()
                dist.send(pinned, dst=1, group=gloo)
            else:
                dist.recv(host_in, src=0, group=gloo)
                pinned.copy_(acc, non_blocking=True)
                torch.cuda.synchronize()
                pinned.add_(host_in)
                acc.copy_(pinned, non_blocking=True)
                torch.cuda.synchronize()

        out[name] = {"bytes": n_int * 4, "B_nccl_send_recv_add_us": timed(nccl_exchange), "C_host_merge_us": timed(host_merge, iters=50, warm=5)}

    # (A): rank 1's shard of the 2-GPU benchmark workload with and without its halo rows
    L = N.lib()
    cfg = bench.synth_cfg(ea, world, bench.READS_PER_GPU)
    row_lo, own_lo, own_hi = bench.shard_rows(ea, cfg, rank, world)
    p = _params(**bench.THRESH)
    tm = N.Timing()
    res = {}
    for what, lo in (("with_halo", row_lo), ("without_halo", own_lo)):
        dev = ea.synth_device(cfg, lo, own_hi, device=local)
        ts = []
        for i in range(8):
            N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 0, C.byref(tm)))
            if i >= 3:
                ts.append(tm.total_ms)
        res[what] = {"rows": own_hi - lo, "bytes": int(dev.nbytes), "ms": float(np.median(ts))}
        dev.close()
    out["A_halo"] = {"halo_rows": own_lo - row_lo, "halo_bytes": res["with_halo"]["bytes"] - res["without_halo"]["bytes"],
                     "report_ms_with_halo": res["with_halo"]["ms"], "report_ms_without_halo": res["without_halo"]["ms"],
                     "note": "rank %d; the report without the halo rows is WRONG near the cut - it only prices the extra rows" % rank}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(json.dumps({"rank0": gathered[0], "rank1": gathered[1]}, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
