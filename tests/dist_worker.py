"""Worker for tests/test_shard.py: run under torchrun with the gloo backend (CPU).  The per-shard
compute engine is the CPU oracle (tests may use it); what is under test is the host-side sharding,
halo, ownership filter and row-count exchange of epialleler_b200.distributed."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import epialleler_b200 as ea  # noqa: E402
from epialleler_b200.distributed import generateCytosineReportSharded  # noqa: E402
from oracle.oracle import PortCpu  # noqa: E402


def oracle_engine(sub, **kw):
    return PortCpu(sub).generateCytosineReport(kw.get("threshold_reads", True), kw.get("threshold_context", "CG"), kw.get("min_context_sites", 2),
                                               kw.get("min_context_beta", 0.5), kw.get("max_outofcontext_beta", 0.1),
                                               kw.get("report_context") or kw.get("threshold_context", "CG"))


def main():
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    ok = True
    for cfg in (ea.synth_config(21, 6000, [70000], 150), ea.synth_config(22, 5000, [40000, 30000, 20000], 60, 400),
                ea.synth_config(23, 8000, [90000], 100, 300, n_islands=3, island_len=1500)):
        bam = ea.synth_host(cfg)
        for kw in (dict(), dict(threshold_reads=False, report_context="CX"), dict(threshold_context="CHH", report_context="CxG", min_context_beta=0.05)):
            first, total, part, full, full_pass = generateCytosineReportSharded(bam, gather_to=0, engine=oracle_engine, **kw)
            if rank == 0:
                want_pass, want = oracle_engine(bam, **kw)
                ok &= total == want["pos"].shape[0] and np.array_equal(full_pass, want_pass)
                ok &= all(np.array_equal(full[c], want[c]) for c in want)
            ok &= first >= 0 and part["pos"].shape[0] <= total
    if rank == 0:
        print("DIST_OK" if ok else "DIST_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
