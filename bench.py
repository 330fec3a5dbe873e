#!/usr/bin/env python
"""Benchmark of the cytosine-report hot path (BASELINE.json metric: generateCytosineReport reads/s and
XM GB/s vs HBM peak).

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU code on host cores

Workload at N=1 = BASELINE.json configs[1]: synthetic WGBS, 50 M reads x 150 bp on one chromosome
(248,956,422 bp), CpG report, threshold.reads=TRUE with the CG defaults; the other named shapes (configs 2b, 3, 4 and the
per-GPU shard of config 5) are timed at FULL size in the same run and reported under "other_configs".
At N>1 the workload is BASELINE.json configs[4]: the 1-billion-read genome-scale WGBS data set over the 24 hg38
chromosomes, 125 M reads per GPU (weak scaling: N GPUs take the first N/8 of the genome), range-sharded by the product's
planner (epialleler_b200/shard.py) into contiguous row ranges with halo rows, ownership filter on the device
(ea_set_owned_range), no data-path collective.
A "step" is one fused generateCytosineReport pass over the rank's resident reads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHR1 = 248_956_422
READS_PER_GPU = 50_000_000
READ_LEN = 150
SEED = 20261018
THRESH = dict(threshold_reads=True, threshold_context="CG", min_context_sites=2, min_context_beta=0.5,
              max_outofcontext_beta=0.1, report_context="CG")
METRIC = "generateCytosineReport_reads_per_s"
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622,
        133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
CFG5 = dict(seed=20261021, n_reads=1_000_000_000, chr_len=HG38, len_min=150)   # BASELINE.json configs[4] (SURVEY.md 8d "config 5")
CFG5_READS_PER_GPU = 125_000_000
CG_THR = dict(threshold_reads=True, threshold_context="CG", min_context_sites=2, min_context_beta=0.5, max_outofcontext_beta=0.1)
# name -> (workload text, generator arguments, row range or None, report arguments, rows of the bit-exact sample, where the sample starts)
OTHER_CONFIGS = {
    "cfg2b_pe_merged": ("configs[1] variant 2b: 25 M PE-merged templates of 150-450 bp (mate gap when > 300) on one chromosome, CpG report, thresholded",
                        dict(seed=20261018, n_reads=25_000_000, chr_len=[CHR1], len_min=150, len_max=450), None, dict(CG_THR, report_context="CG"), 200_000, 0.5),
    "cfg3_capture_cx": ("configs[2]: targeted capture, 200 M reads of 100-417 bp on 20,000 islands x 2 kb over 24 chromosomes (10,000 reads per island), CX report, threshold.context=CG",
                        dict(seed=20261019, n_reads=200_000_000, chr_len=HG38, len_min=100, len_max=417, n_islands=20000, island_len=2000), None,
                        dict(CG_THR, report_context="CX"), 200_000, 0.5),
    "cfg4_long_reads": ("configs[3]: 2 M long reads of 10-30 kb over 24 chromosomes (~13x), CpG report, threshold.reads=FALSE (the reference's advice for long reads)",
                        dict(seed=20261020, n_reads=2_000_000, chr_len=HG38, len_min=10000, len_max=30000), None,
                        dict(CG_THR, threshold_reads=False, report_context="CG"), 40_000, 0.5),
    "cfg4_long_reads_thr": ("configs[3], secondary: the same reads with threshold.reads=TRUE",
                            dict(seed=20261020, n_reads=2_000_000, chr_len=HG38, len_min=10000, len_max=30000), None, dict(CG_THR, report_context="CG"), 0, 0.5),
    "cfg5_shard": ("configs[4], one GPU's shard: rows [0, 125 M) of the 1-billion-read WGBS data set over the 24 hg38 chromosomes (chr1 and the start of chr2, ~48x), CpG report, thresholded",
                   CFG5, (0, CFG5_READS_PER_GPU), dict(CG_THR, report_context="CG"), 200_000, None),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def headline_synth_kwargs(reads):
    """configs[1] (or a proportionally shorter chromosome when --reads shrinks it for a quick run)."""
    return dict(seed=SEED, n_reads=reads, chr_len=[CHR1 if reads == READS_PER_GPU else max(int(CHR1 * reads / READS_PER_GPU), 100_000)], len_min=READ_LEN)


def bind_host_to_gpu(local):
    """Pin this rank's host threads (and with them the first-touch placement of its page-locked buffers) to the CPUs that
    share a NUMA node with its GPU.  Uses NVML's ideal CPU affinity; silently does nothing where that is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d CPUs local to GPU %d (NVML affinity)" % (len(cpus), local)
    except Exception:
        pass
    return "not set"


def timed_report(L, N, dev, prm, warm, steps):
    """`steps` device-timed passes of the fused report (CUDA events on the library's stream) -> mean ms (total, window kernel,
    scan+gather) and the last call's counters."""
    tm = N.Timing()
    for _ in range(warm):
        N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(prm), 0, C.byref(tm)))
    tot = win = gat = thr = 0.0
    for _ in range(steps):
        N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(prm), 0, C.byref(tm)))
        tot += tm.total_ms; thr += tm.kernel_ms[0]; win += tm.kernel_ms[1]; gat += tm.kernel_ms[2]
    return tot / steps, thr / steps, win / steps, gat / steps, tm


def sample_bit_exact(ea, synth_kw, r0, r1, report_kw):
    """Rows [r0, r1) of a workload through the product (host buffers -> C ABI) and through the compiled reference
    (oracle/_ref, else the C port) - regenerated by the oracle's own host build of the generator: True iff pass flags and
    every report column are identical."""
    from oracle.oracle import REF_SO, PortCpu, RefCpu, synth_rows
    rows = synth_rows(row_begin=r0, row_end=r1, **synth_kw)
    o = (RefCpu if os.path.exists(REF_SO) else PortCpu)(rows)
    t0 = time.perf_counter()
    p, rep = o.generateCytosineReport(report_kw["threshold_reads"], report_kw["threshold_context"], report_kw["min_context_sites"],
                                      report_kw["min_context_beta"], report_kw["max_outofcontext_beta"], report_kw["report_context"])
    cpu_s = time.perf_counter() - t0
    bam = ea.BamData(rows.rname, rows.strand, rows.start, rows.templid, *rows.blob())
    gp, grep = ea.generateCytosineReport(bam, return_pass=True, **report_kw)
    ok = bool(np.array_equal(gp, p)) and all(np.array_equal(grep[k], rep[k]) for k in ("rname", "strand", "pos", "context", "meth", "unmeth"))
    bam.close()
    return ok, int(rep["pos"].shape[0]), cpu_s, "reference" if os.path.exists(REF_SO) else "port"


def run_other_configs(ea, L, N, args, peak, device):
    """The other named shapes at full size, device-resident: which window kernel each takes, its time and fraction of the HBM
    roofline (algorithmic bytes as SURVEY.md 8d defines them), and a bit-exact check of a row sample against the reference."""
    from epialleler_b200.api import _params
    out = {}
    for name, (text, skw, rng, rkw, n_sample, where) in OTHER_CONFIGS.items():
        if args.only and name not in args.only:
            continue
        t_wall = time.perf_counter()
        cfg = ea.synth_config(**skw)
        r0, r1 = rng if rng else (0, cfg.n_reads)
        try:
            dev = ea.synth_device(cfg, r0, r1, device=device)
        except Exception as e:  # e.g. not enough free device memory on a shared box: say so instead of dying
            out[name] = {"workload": text, "error": str(e)[:200]}
            L.ea_trim(device)
            continue
        prm = _params(**rkw)
        tot, thr, win, gat, tm = timed_report(L, N, dev, prm, 2, 3)
        alg = int(tm.alg_bytes)
        entry = {"workload": text, "reads": int(dev.nrow), "xm_bytes": int(dev.nbytes), "rows": int(tm.n_rows), "reads_passing": int(tm.n_pass),
                 "kernel": (L.ea_report_kernel(dev.handle()) or b"").decode(), "ms_per_step": round(tot, 4), "threshold_ms": round(thr, 4),
                 "kernel_ms": round(win, 4), "scan_gather_ms": round(gat, 4), "algorithmic_bytes": alg,
                 "achieved_gbs": round(alg / win / 1e6, 1), "frac": round(alg / win / 1e6 / peak, 4),
                 "pipeline_frac": round(alg / tot / 1e6 / peak, 4), "reads_per_s": dev.nrow / (tot / 1e3), "xm_gbs": dev.nbytes / 1e9 / (tot / 1e3),
                 "steps": 3, "warmup": 2}
        dev.close()
        L.ea_trim(device)
        if n_sample and not args.no_cpu:
            n_all = r1 - r0
            if where is None:   # config 5: straddle the chromosome 1 -> 2 boundary
                mid = r0 + int(cfg.n_reads * HG38[0] / sum(HG38))
            else:
                mid = r0 + int(n_all * where)
            s0 = max(r0, min(mid - n_sample // 2, r1 - n_sample))
            ok, srows, cpu_s, kind = sample_bit_exact(ea, skw, s0, s0 + n_sample, rkw)
            entry["sample_check"] = {"rows": [s0, s0 + n_sample], "bit_exact": ok, "checker": kind, "report_rows": srows, "cpu_seconds": round(cpu_s, 2)}
        entry["wall_s"] = round(time.perf_counter() - t_wall, 1)
        out[name] = entry
    return out


def e2e_shim_path(ea, L, host, device, steps, warm=3):
    """End to end through the boundary R binds: the reader's std::vector<std::string> (one PAGEABLE heap string per
    template, integration/harness) + pageable int columns -> handle_of() of the shim (pointer tables, ea_pack_calls) ->
    ea_cytosine_report -> fetch into fresh host arrays -> ea_free.  Runs FIRST in the process: the first call is a real cold
    call (no cached page-locked block, empty device pool, kernels not loaded yet)."""
    import torch
    import integration.harness as H
    from epialleler_b200._native import check as N_check
    t0 = time.perf_counter()
    rs = H.ReaderStrings(host)
    build_s = time.perf_counter() - t0
    from epialleler_b200.api import _params
    prm = _params(**THRESH)
    i32p, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    times, rows, n_pass, h2d, calls_only = [], 0, 0, 0, True
    # call 0 = the cold call; then `warm` untimed calls (a GPU that sat idle while the host prepared the strings takes a few
    # seconds of work before its allocations and copies run at speed - measured: 0.6-1.4 s per call in that phase); then `steps`
    for i in range(1 + warm + steps):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        dev = rs.pack(device=device, calls_only=True, records=True)
        # results land in freshly allocated PAGEABLE arrays, like the R vectors the shim fetches into (wrap_report)
        pass_flags = np.empty(dev.nrow, np.uint8)
        nrow = C.c_int64(0)
        N_check(L.ea_cytosine_report(dev.handle(), C.byref(prm), pass_flags.ctypes.data_as(u8p), C.byref(nrow)))
        cols = [np.empty(max(nrow.value, 1), np.int32) for _ in range(6)]
        N_check(L.ea_cx_report_fetch(dev.handle(), *[c.ctypes.data_as(i32p) for c in cols]))
        h2d = int(L.ea_pack_h2d_bytes(dev.handle()))
        calls_only = not bool(L.ea_has_bytes(dev.handle()))
        dev.close()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t1)
        rows, n_pass = int(nrow.value), int(pass_flags.sum())
        del cols, pass_flags
    rs.close()
    return dict(cold_s=times[0], warm_s=float(np.median(times[1 + warm:])), warmup_s=[round(t, 4) for t in times[1:1 + warm]],
                all_s=[round(t, 4) for t in times[1 + warm:]], rows=rows, n_pass=n_pass, h2d=h2d, calls_only=calls_only, build_strings_s=build_s, steps=steps, warm=warm)


def e2e_pinned_blob(ea, L, host, device, steps, barrier, owned=None):
    """The same from one contiguous PAGE-LOCKED blob (ea_pack_blob_calls): what a caller that owns its buffers can reach."""
    import torch
    blob, off = host.blob()
    pinned = True
    try:
        src = (blob, off, host.rname, host.strand, host.start, np.arange(host.nrow, dtype=np.int32))
        pin = [torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True) for a in src]
        for t, a in zip(pin, src):
            t.numpy()[...] = a
        blob, off, rn, st, sa, tid = [t.numpy() for t in pin]
        del src
    except Exception:
        pinned = False
        rn, st, sa, tid = host.rname, host.strand, host.start, np.arange(host.nrow, dtype=np.int32)
    n_loc = host.nrow
    times, rows_e2e, n_pass, h2d_actual, calls_only = [], 0, 0, 0, False
    for i in range(steps + 2):  # two untimed passes: the two alternating sets of page-locked result blocks get allocated
        barrier()
        t1 = time.perf_counter()
        b = ea.BamData(rn, st, sa, tid, blob, off, device=device)
        if owned:
            b.set_owned(*owned)
        pass_flags, rep = ea.generateCytosineReport(b, return_pass=True, **THRESH)
        h2d_actual = int(L.ea_pack_h2d_bytes(b.handle(need_bases=False)))
        calls_only = not bool(L.ea_has_bytes(b.handle(need_bases=False)))
        b.close()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t1
        if i > 1:
            times.append(dt)
        rows_e2e, n_pass = int(rep["pos"].shape[0]), int(pass_flags.sum())
    h2d_input = int(blob.nbytes + off.nbytes + rn.nbytes + st.nbytes + sa.nbytes + tid.nbytes)
    return dict(seconds=float(np.median(times)), h2d=h2d_actual, h2d_input=h2d_input, calls_only=calls_only, d2h=int(pass_flags.shape[0] + 24 * rows_e2e),
                pinned=pinned, steps=steps, rows=rows_e2e, n_pass=n_pass, reads=n_loc)


def run_ours(args):
    import torch

    import epialleler_b200 as ea
    from epialleler_b200 import _native as N
    from epialleler_b200.api import _params
    from epialleler_b200.shard import plan_shards_windowed

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    host_affinity = bind_host_to_gpu(local) if world > 1 else "not set (single GPU)"
    if world > 1:  # the ranks share the host's cores: the host-side transposition only pays with enough threads per rank
        per_rank = max(1, (os.cpu_count() or 1) // world)
        os.environ.setdefault("EA_HOST_PACK_THREADS", str(per_rank))
        if per_rank < 8:
            os.environ.setdefault("EA_HOST_PACK", "0")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = N.lib()
    L.ea_set_kernel_policy(args.policy)
    peak, peak_src = measured_peak()
    p = _params(**THRESH)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shard = None
    aux_errors = {}   # failures of the auxiliary legs (they must not cost the line)
    if world == 1:
        skw = headline_synth_kwargs(args.reads)
        cfg = ea.synth_config(**skw)
        row_lo, own_lo, own_hi = 0, 0, cfg.n_reads
        workload = ("configs[1]: synthetic WGBS, %d reads x %d bp on one chromosome (%d bp), CpG report, threshold.reads=TRUE (CG: 2 / 0.5 / 0.1)"
                    % (args.reads, READ_LEN, cfg.chr_len[0]))
    else:
        # configs[4]: N GPUs take rows [0, N x 125 M) of the 1-billion-read genome (the first N/8 of it); the product's planner
        # cuts them into N ranges looking only at a window of rows around every cut (generated on the host, nothing else is)
        skw = dict(CFG5)
        per_gpu = args.reads if args.reads != READS_PER_GPU else CFG5_READS_PER_GPU
        cfg = ea.synth_config(**skw)
        n_job = min(cfg.n_reads, per_gpu * world)

        def fetch(i0, i1):
            h = ea.synth_host(cfg, i0, i1)
            return h.rname, h.start, np.diff(h.blob()[1])
        shard = plan_shards_windowed(n_job, world, fetch, READ_LEN)[rank]
        row_lo, own_lo, own_hi = shard.row_lo, shard.own_lo, shard.own_hi
        workload = ("configs[4]: synthetic genome-scale WGBS, 1 B reads x %d bp over the 24 hg38 chromosomes; %d GPUs take rows [0, %d) "
                    "(%d reads per GPU, the first %d/8 of the genome), CpG report, threshold.reads=TRUE (CG: 2 / 0.5 / 0.1)"
                    % (READ_LEN, world, n_job, per_gpu, world))
    # ---- N=1: end to end through the public API with HOST buffers FIRST (a fresh process: the first call is the cold call) ----
    e2e = shim = None
    if world == 1 and not args.no_e2e:
        host = ea.synth_host(cfg)
        if not args.no_shim:
            try:
                shim = e2e_shim_path(ea, L, host, local, max(1, min(args.steps, 5)))
            except Exception as e:   # e.g. the harness could not be built on this box: the blob figure below becomes e2e
                aux_errors["e2e_shim_path"] = repr(e)[:300]
        e2e = e2e_pinned_blob(ea, L, host, local, max(1, min(args.steps, 5)), barrier)
        del host
        L.ea_trim(local)
    dev = ea.synth_device(cfg, row_lo, own_hi, device=local)
    if shard is not None:
        dev.set_owned(shard.key_lo, shard.key_hi, own_lo - row_lo)
    own_reads = own_hi - own_lo
    tm = N.Timing()

    def step():
        N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 0, C.byref(tm)))
        return tm.total_ms, tm.kernel_ms[1], tm.kernel_ms[2]

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    tot = win = gat = 0.0
    for _ in range(args.steps):
        a, b, c = step()
        tot += a; win += b; gat += c
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    n_rows, n_pass, alg_bytes, launches = int(tm.n_rows), int(tm.n_pass), int(tm.alg_bytes), int(tm.n_launches)
    kernel_name = (L.ea_report_kernel(dev.handle()) or b"").decode()
    dev_nbytes = int(dev.nbytes)
    halo = None
    if shard is not None:
        halo = {"rows": int(own_lo - row_lo), "bytes": int((own_lo - row_lo) * READ_LEN)}

    # ---- two-pass variant (threshold kernel + report kernel), for the record ----
    two_pass = None
    if world == 1:
        N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 1, C.byref(tm)))
        N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 1, C.byref(tm)))
        thr_alg = int(dev.nbytes) + 9 * int(dev.nrow)  # rcpp_threshold_reads alone: every XM byte once, 8 B offset read + 1 B flag written per row
        two_pass = dict(total_ms=round(tm.total_ms, 4), threshold_ms=round(tm.kernel_ms[0], 4), window_ms=round(tm.kernel_ms[1], 4),
                        rows=int(tm.n_rows), n_pass=int(tm.n_pass),
                        threshold_kernel={"what": "rcpp_threshold_reads stand-alone (k_threshold_pl on the planar call planes + thr tables)", "algorithmic_bytes": thr_alg,
                                          "achieved_gbs": round(thr_alg / max(tm.kernel_ms[0], 1e-9) / 1e6, 1),
                                          "frac_of_hbm_peak": round(thr_alg / max(tm.kernel_ms[0], 1e-9) / 1e6 / peak, 4)})

    # ---- the widened paths on the same resident reads, for the record: device call + row-count read-back ----
    widened = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_widened:
        widened = {}
        nrow = C.c_int64(0)
        for _ in range(2):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            N.check(L.ea_mhl_report(dev.handle(), b"Zz", 0, 0, 0.1, C.byref(nrow)))
            torch.cuda.synchronize(); dt = time.perf_counter() - t1
        widened["mhl_report"] = {"seconds": dt, "reads_per_s": own_reads / dt, "rows": int(nrow.value), "what": "ea_mhl_report (rcpp_mhl_report), CG haplotypes, defaults"}
        rng = np.random.default_rng(1)
        nb = 20000
        bst = np.sort(rng.integers(1, CHR1 - 3000, nb)).astype(np.int32)
        bed = (np.ones(nb, np.int32), bst, (bst + 2000).astype(np.int32))
        counts = np.zeros((nb + 1) * 4, np.int64)
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        for _ in range(2):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            N.check(L.ea_bed_report(dev.handle(), 1, i32(bed[0]), i32(bed[1]), i32(bed[2]), nb, 1, C.byref(p), None, None, None, counts.ctypes.data_as(C.POINTER(C.c_int64))))
            torch.cuda.synchronize(); dt = time.perf_counter() - t1
        widened["bed_report"] = {"seconds": dt, "reads_per_s": own_reads / dt, "targets": nb, "reads_matched": int(counts[:-4].sum()),
                                 "what": "ea_bed_report (rcpp_match_capture + thresholds + per-target counts), 20000 targets of 2 kb"}

        nv = 1_000_000
        vpos = np.sort(rng.integers(1, CHR1, nv)).astype(np.int32)
        vchr = np.ones(nv, np.int32)
        freqs = np.zeros(nv * 20, np.float64)
        for _ in range(2):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            N.check(L.ea_base_freqs(dev.handle(), None, i32(vchr), i32(vpos), nv, freqs.ctypes.data_as(C.POINTER(C.c_double))))
            torch.cuda.synchronize(); dt = time.perf_counter() - t1
        widened["base_freqs"] = {"seconds": dt, "reads_per_s": own_reads / dt, "snvs": nv, "bases_counted": int(freqs.sum()),
                                 "what": "ea_base_freqs (rcpp_get_base_freqs), 1 M sorted SNV positions, all reads passing"}

    # ---- N>1: end to end from host buffers (pack + H2D + kernels + D2H) on a bounded part of every rank's shard ----
    e2e_reads = own_reads
    dev.close()
    L.ea_trim(local)
    if world > 1 and not args.no_e2e:
        # every rank regenerates the first `--e2e-reads` rows of its shard on the host (the whole shard is 18.75 GB per rank
        # and would need twice that again page-locked): rate per read is what is reported
        e2e_reads = min(own_hi - row_lo, args.e2e_reads)
        host = ea.synth_host(cfg, row_lo, row_lo + e2e_reads)
        own_first = own_lo - row_lo
        e2e = e2e_pinned_blob(ea, L, host, local, max(1, min(args.steps, 5)), barrier,
                              owned=(shard.key_lo, shard.key_hi, min(own_first, e2e_reads)))
        e2e_reads -= min(own_first, e2e_reads)
        del host

    ms_step = tot / args.steps
    stats = torch.tensor([ms_step, wall_ms / args.steps, win / args.steps, e2e["seconds"] if e2e else 0.0], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(own_reads), float(alg_bytes), float(dev_nbytes), float(n_rows), float(n_pass), float(e2e_reads),
                         float(e2e["h2d"] if e2e else 0), float(e2e["d2h"] if e2e else 0), float(halo["rows"] if halo else 0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_step, wall_step, win_step, e2e_s = [float(x) for x in stats.tolist()]
    tot_reads, tot_alg, tot_xm, tot_rows, tot_pass, tot_e2e_reads, tot_h2d, tot_d2h, tot_halo = [float(x) for x in sums.tolist()]

    other = cpu = None
    if rank == 0 and world == 1:
        if not args.no_other:
            try:
                other = run_other_configs(ea, L, N, args, peak, local)
            except Exception as e:
                aux_errors["other_configs"] = repr(e)[:300]
        if not args.no_cpu:
            try:
                cpu = cpu_baseline(ea, skw, min(args.cpu_rows, cfg.n_reads), check_gpu=True)
            except Exception as e:
                aux_errors["cpu_baseline"] = repr(e)[:300]

    if rank == 0:
        traffic, traffic_src = None, None
        try:  # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("reads_per_gpu") == args.reads and world == 1:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        achieved = (alg_bytes / 1e9) / (win / args.steps / 1e3)   # this rank's dominant kernel
        line = {
            "metric": METRIC, "value": tot_reads / (ms_step / 1e3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "reads_per_gpu": int(own_reads) if world == 1 else int(tot_reads / world), "read_len": READ_LEN, "report_context": "CG",
                       "threshold_context": "CG",
                       "sharding": ("contiguous genomic ranges planned by epialleler_b200.shard.plan_shards_windowed, halo rows packed on both neighbours, "
                                    "ownership filter on the device (ea_set_owned_range), no collective on the data path") if world > 1 else "single GPU",
                       "l2": "inputs (%.1f GB per GPU) are larger than L2 (126 MB); no flush needed" % (dev_nbytes / 1e9), "seed": int(skw["seed"]),
                       "report_rows": int(tot_rows), "reads_passing": int(tot_pass)},
            "xm_gbs": tot_xm / 1e9 / (ms_step / 1e3),
            "wall_ms_per_step": wall_step,
            "gpu_launches": launches * args.steps,
            "kernels_per_step": {"thr_tables": 1, "window(fused threshold+aggregate)": 1, "scan_windows (cub: init + scan)": 2, "gather_rows": 1,
                                 **({"owned_rows (ownership filter)": 1} if world > 1 else {})},
            "roofline": {"bound": "hbm", "kernel": kernel_name + ", fused", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": win_step, "formula": "sum(L) + 14 N + 24 R", "pipeline_frac": (tot_alg / world / 1e9) / (ms_step / 1e3) / peak,
                         "device_format": ("the kernel reads the packer's 4-bit bit-sliced call planes (16 bytes per 32 positions, tiles of 32 rows) plus 8 bytes "
                                           "of metadata per read, not the reference's byte per position: DRAM traffic per launch (`traffic`) is about half the "
                                           "algorithmic bytes, which are counted on the reference's format as SURVEY.md 8(d) defines them"),
                         "dram_gbs": (traffic / 1e9) / (win_step / 1e3) if traffic else None,
                         "dram_frac_of_peak": ((traffic / 1e9) / (win_step / 1e3) / peak) if traffic else None},
            "clocks": clocks,
        }
        if two_pass:
            line["two_pass_variant"] = two_pass
        if world > 1:
            line["scaling_note"] = ("weak scaling of configs[4] at %d reads per GPU; the 1-GPU point of THIS workload is other_configs.cfg5_shard of the N=1 line "
                                    "(rows [0, 125 M) of the same data set on one GPU) - the N=1 headline value is configs[1], a different shape" % int(tot_reads / world))
        if halo is not None:
            line["halo"] = {"rows_total": int(tot_halo), "bytes_total": int(tot_halo * READ_LEN), "cuts": world - 1,
                            "what": ("rows that start before a cut and reach past it, packed on both neighbours (zero communication); the measured alternatives "
                                     "(NCCL counter exchange 31 us per cut, host merge 56 us per cut, halo 4 us) are in profiles/r2_exchange_compare_2gpu.json")}
        if shim:  # the path the R shim calls is the headline e2e; the pinned-blob figure rides along
            line["e2e"] = {"value": tot_reads / shim["warm_s"], "unit": "reads/s", "h2d_bytes_per_step": shim["h2d"], "d2h_bytes_per_step": int(own_reads + 24 * shim["rows"]),
                           "seconds_per_step": shim["warm_s"], "cold_first_call_seconds": shim["cold_s"], "cold_first_call_reads_per_s": tot_reads / shim["cold_s"],
                           "steps": shim["steps"], "warmup": shim["warm"], "warmup_seconds": shim["warmup_s"], "step_seconds": shim["all_s"],
                           "pinned_host_memory": False, "calls_only": shim["calls_only"],
                           "host_threads": os.environ.get("EA_HOST_PACK_THREADS", str(os.cpu_count())),
                           "what": ("ea_pack_calls_records through integration/harness: the reader's std::vector<std::string> (one PAGEABLE heap string per template) and pageable int "
                                    "columns -> the shim's handle_of() (the vector's std::string objects are handed over as a record array, no pointer tables; host threads "
                                    "transpose the strings where they lie into a ring of 3 page-locked blocks, the blocks cross PCIe) -> ea_cytosine_report -> ea_cx_report_fetch of pass flags and rows into fresh PAGEABLE arrays (the R "
                                    "vectors of wrap_report) -> ea_free.  The leg runs first in the process: cold = the process's first call (no cached page-locked block, empty device "
                                    "pool, kernels not yet loaded, GPU idle since start-up); then `warmup` untimed calls; seconds_per_step = median of the `steps` calls after them"),
                           "from_pinned_blob": None}
        if e2e:
            blob_fig = {"value": tot_e2e_reads / e2e_s, "unit": "reads/s", "seconds_per_step": e2e_s, "h2d_bytes_per_step": int(tot_h2d), "d2h_bytes_per_step": int(tot_d2h),
                        "pinned_host_memory": e2e["pinned"], "steps": e2e["steps"], "reads_per_step": int(tot_e2e_reads),
                        "what": ("ea_pack_blob_calls from ONE contiguous page-locked blob (raw chunks from the front transposed on the device, host-transposed chunks from the back) "
                                 "+ fused report + D2H + free" if e2e["calls_only"] else "ea_pack_blob (H2D + device repack + validation + window index) + fused report + D2H + free")}
            if shim:
                line["e2e"]["from_pinned_blob"] = blob_fig
            else:
                line["e2e"] = dict(blob_fig, host_affinity=host_affinity, host_threads=os.environ.get("EA_HOST_PACK_THREADS", str(os.cpu_count())))
                if world > 1:
                    line["e2e"]["sample"] = "the first %d rows of every rank's shard, regenerated on the host (the whole 18.75 GB shard twice over does not fit every host)" % args.e2e_reads
        if cpu:
            line["cpu_baseline"] = cpu
        if other:
            line["other_configs"] = other
        if widened:
            line["widened_paths"] = widened
        if aux_errors:
            line["aux_errors"] = aux_errors
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _oracle_report(o):
    return o.generateCytosineReport(THRESH["threshold_reads"], "CG", 2, 0.5, 0.1, "CG")


def cpu_baseline(ea, skw, rows, check_gpu=False):
    """The reference's own three C++ files (oracle/_ref, else the C port) on `rows` leading rows of the workload, regenerated
    by the oracle's host build of the generator."""
    from oracle.oracle import REF_SO, PortCpu, RefCpu, synth_rows
    host = synth_rows(row_begin=0, row_end=rows, **skw)
    kind = "reference" if os.path.exists(REF_SO) else "port"
    o = (RefCpu if kind == "reference" else PortCpu)(host)
    t0 = time.perf_counter()
    p, rep = _oracle_report(o)
    wall = time.perf_counter() - t0
    sec = o.last_seconds
    out = {"value": rows / sec, "unit": "reads/s", "cores": 1, "kind": kind, "seconds": sec, "wall_seconds": wall,
           "sample": "first %d rows of the same synthetic workload (the reference is single-threaded; %d host cores present)" % (rows, os.cpu_count()),
           "xm_gbs": float(host.blob()[1][-1]) / 1e9 / sec,
           "note": "flat_map is a sorted-vector stand-in for Boost (not installed); function time only, marshalling excluded"}
    if check_gpu:
        bam = ea.BamData(host.rname, host.strand, host.start, host.templid, *host.blob())
        gp, grep = ea.generateCytosineReport(bam, return_pass=True, **THRESH)
        ok = bool(np.array_equal(gp, p)) and all(np.array_equal(grep[k], rep[k]) for k in ("rname", "strand", "pos", "context", "meth", "unmeth"))
        out["gpu_matches_bit_exact"] = ok
        bam.close()
    return out


def run_reference(args):
    """Reference arm: the reference's own CPU implementation, bounded sample per step, rank 0 only.  Nothing of the product
    is imported here: the rows come from oracle/_build/libsynth.so (the generator header built for the host)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.oracle import REF_SO, PortCpu, RefCpu, synth_rows
    world = args.gpus
    if world == 1:
        skw = headline_synth_kwargs(args.reads)
        text = "configs[1]: synthetic WGBS, %d reads x %d bp on one chromosome, CpG report, threshold.reads=TRUE" % (args.reads, READ_LEN)
    else:
        skw = dict(CFG5)
        text = "configs[4]: synthetic genome-scale WGBS, 1 B reads x %d bp over the 24 hg38 chromosomes, CpG report, threshold.reads=TRUE" % READ_LEN
    rows = min(args.ref_rows, skw["n_reads"])
    host = synth_rows(row_begin=0, row_end=rows, **skw)
    kind = "reference" if os.path.exists(REF_SO) else "port"
    o = (RefCpu if kind == "reference" else PortCpu)(host)
    secs = []
    for i in range(max(args.warmup, 1) + args.steps):
        _oracle_report(o)
        if i >= max(args.warmup, 1):
            secs.append(o.last_seconds)
    s = float(np.mean(secs))
    v = rows / s
    sample = ("each step = the first %d rows of the workload; the rate is what is reported, extrapolated to the full size (the reference is linear in the "
              "number of reads and single-threaded; %d host cores present)" % (rows, os.cpu_count()))
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": text + "; bounded sample, rate extrapolated", "sample_rows": rows},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "xm_gbs": float(host.blob()[1][-1]) / 1e9 / s})


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL banners, warnings)
    was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=READS_PER_GPU, help="reads per GPU (default: BASELINE configs[1])")
    ap.add_argument("--cpu-rows", type=int, default=2_000_000)
    ap.add_argument("--ref-rows", type=int, default=2_000_000, help="rows per step of the reference arm (same sample size as cpu_baseline)")
    ap.add_argument("--e2e-reads", type=int, default=25_000_000, help="N>1: rows of every rank's shard the end-to-end leg regenerates on the host")
    ap.add_argument("--no-other", action="store_true", help="skip the other named shapes (configs 2b, 3, 4, 5-shard)")
    ap.add_argument("--only", nargs="*", default=None, help="names of the other configs to run")
    ap.add_argument("--no-shim", action="store_true", help="skip the end-to-end leg through the reader's heap strings")
    ap.add_argument("--no-widened", action="store_true")
    ap.add_argument("--policy", type=int, default=0, help="kernel policy for A/B runs: 0 automatic (planar call planes), 5 automatic over the byte rows (round-1 kernels), 6 vertical counters, 1-4 single byte-row kernels")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:  # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        os.dup2(_REAL_STDOUT, 1)
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
