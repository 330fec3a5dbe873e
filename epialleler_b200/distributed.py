"""One process per GPU: range-sharded generateCytosineReport.

No data-path collective exists on this path: every rank reports the genomic range it owns from its
own rows plus a halo (epialleler_b200/shard.py).  The only exchange is the all-gather of per-rank row
counts (8 bytes per rank) that places each rank's rows in the concatenated report; gathering the
rows themselves on one rank is optional.
"""
from __future__ import annotations

import numpy as np

from .shard import Shard, filter_report, merge_reports, plan_shards, plan_shards_native, slice_bam


def _lengths(bam):
    off = bam.blob()[1]
    return (off[1:] - off[:-1])[bam.templid]


def local_report(sub, shard: Shard, **kw):
    """One shard whose rows [row_lo, own_hi) are ALREADY this rank's `sub` (a BamData or DeviceBam holding nothing else):
    the ownership filter runs on the device (ea_set_owned_range: two binary searches over the ordered report, only the
    owned rows cross PCIe) and the pass flags come back for the owned rows only.  This is the call a rank makes when it
    loaded just its own genomic range - no rank ever needs the whole data set."""
    from .api import generateCytosineReport
    sub.set_owned(shard.key_lo, shard.key_hi, shard.own_lo - shard.row_lo)
    own_pass, rep = generateCytosineReport(sub, return_pass=True, **kw)
    return np.asarray(own_pass), rep


def shard_report(bam, rank: int, world: int, engine=None, **kw):
    """This rank's part of a host `bam` every rank holds: (shard, pass flags of its own rows, the report rows it owns).
    engine=None runs the library with the device-side ownership filter; a custom engine (the CPU tests) gets the host mask."""
    if engine is None:
        sh = plan_shards_native(bam, world)[rank]
        sub = slice_bam(bam, sh)
        own_pass, rep = local_report(sub, sh, **kw)
        sub.close()
        return sh, own_pass, rep
    sh = plan_shards(bam.rname, bam.start, _lengths(bam), world)[rank]
    sub = slice_bam(bam, sh)
    pass_, rep = engine(sub, **kw)
    own_pass = np.asarray(pass_)[sh.own_lo - sh.row_lo:]
    if hasattr(sub, "close"):
        sub.close()
    return sh, own_pass, filter_report(rep, sh)


def row_offsets(n_rows: int):
    """all-gather of the per-rank row counts -> (first global row of this rank, total rows)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor([n_rows], dtype=torch.int64, device=dev)
    allc = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allc, mine)
    counts = [int(c.item()) for c in allc]
    return sum(counts[:rank]), sum(counts)


def generateCytosineReportSharded(bam, gather_to=0, engine=None, **kw):
    """Call under torchrun on every rank with the same host `bam`.  Returns (first_row, total_rows, report
    part) and, on rank `gather_to`, the full concatenated report and pass vector as 4th/5th values."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    sh, own_pass, rep = shard_report(bam, rank, world, engine=engine, **kw)
    first, total = row_offsets(int(rep["pos"].shape[0]))
    full = full_pass = None
    if gather_to is not None:
        cols = ("rname", "strand", "pos", "context", "meth", "unmeth")
        payload = ({c: rep[c] for c in cols}, own_pass)
        box = [None] * world if rank == gather_to else None
        dist.gather_object(payload, box, dst=gather_to)
        if rank == gather_to:
            full = merge_reports([b[0] for b in box])
            full_pass = np.concatenate([b[1] for b in box])
            assert full["pos"].shape[0] == total
    return first, total, rep, full, full_pass
