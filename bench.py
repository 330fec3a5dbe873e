#!/usr/bin/env python
"""Benchmark of the cytosine-report hot path (BASELINE.json metric: generateCytosineReport reads/s and
XM GB/s vs HBM peak).

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU code on host cores

Workload at N=1 = BASELINE.json configs[1]: synthetic WGBS, 50 M reads x 150 bp on one chromosome
(248,956,422 bp), CpG report, threshold.reads=TRUE with the CG defaults.  At N>1 the per-GPU work is
fixed (weak scaling): N x 50 M reads on one chromosome of N x 248,956,422 bp, range-sharded into N
contiguous row ranges with halo rows (epialleler_b200/shard.py), no data-path collective.
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


def synth_cfg(ea, world, reads_per_gpu):
    return ea.synth_config(SEED, reads_per_gpu * world, [CHR1 * world if reads_per_gpu == READS_PER_GPU else max(int(CHR1 * world * reads_per_gpu / READS_PER_GPU), 100_000)], READ_LEN)


def shard_rows(ea, cfg, rank, world):
    """Row range [row_lo, own_hi) packed on this rank and the rows [own_lo, own_hi) it owns."""
    from epialleler_b200 import _native as N
    n = cfg.n_reads
    own_lo, own_hi = n * rank // world, n * (rank + 1) // world
    if rank == 0:
        return 0, 0, own_hi
    back = min(own_lo, 65536)
    m = back + 1
    rn, st, sa = (np.zeros(m, np.int32) for _ in range(3))
    off = np.zeros(m + 1, np.int64)
    N.check(N.lib().ea_synth_host(C.byref(cfg), own_lo - back, own_lo + 1, rn.ctypes.data_as(C.POINTER(C.c_int32)), st.ctypes.data_as(C.POINTER(C.c_int32)),
                                  sa.ctypes.data_as(C.POINTER(C.c_int32)), off.ctypes.data_as(C.POINTER(C.c_int64)), None))
    ln = np.diff(off)
    cut_rn, cut_pos = rn[-1], sa[-1]
    same = (rn == cut_rn)
    own = int(np.argmax(same & (sa == cut_pos)))          # snap to the first row with the cut's (rname,start)
    reach = same & (sa + ln > cut_pos)
    halo = int(np.argmax(reach))                          # first earlier row reaching past the cut
    return own_lo - back + halo, own_lo - back + own, own_hi


def bind_host_to_gpu(local):
    """Pin this rank's host threads (and with them the first-touch placement of its page-locked buffers) to the CPUs that
    share a NUMA node with its GPU: with 8 ranks each pushing 8.7 GB per step over PCIe, host buffers on the far socket
    halve the end-to-end rate.  Uses NVML's ideal CPU affinity; silently does nothing where that is unavailable."""
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


def run_ours(args):
    import torch

    import epialleler_b200 as ea
    from epialleler_b200 import _native as N
    from epialleler_b200.api import _params

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
    cfg = synth_cfg(ea, world, args.reads)
    row_lo, own_lo, own_hi = shard_rows(ea, cfg, rank, world)
    dev = ea.synth_device(cfg, row_lo, own_hi, device=local)
    own_reads = own_hi - own_lo
    p = _params(**THRESH)
    tm = N.Timing()

    def step():
        N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 0, C.byref(tm)))
        return tm.total_ms, tm.kernel_ms[1], tm.kernel_ms[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    # ---- two-pass variant (threshold kernel + report kernel), for the record ----
    N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 1, C.byref(tm)))
    N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(p), 1, C.byref(tm)))
    thr_alg = int(dev.nbytes) + 9 * int(dev.nrow)  # rcpp_threshold_reads alone: every XM byte once, 8 B offset read + 1 B flag written per row
    two_pass = dict(total_ms=round(tm.total_ms, 4), threshold_ms=round(tm.kernel_ms[0], 4), window_ms=round(tm.kernel_ms[1], 4),
                    rows=int(tm.n_rows), n_pass=int(tm.n_pass),
                    threshold_kernel={"what": "rcpp_threshold_reads stand-alone (k_threshold_pl on the planar call planes + thr tables)", "algorithmic_bytes": thr_alg,
                                      "achieved_gbs": round(thr_alg / max(tm.kernel_ms[0], 1e-9) / 1e6, 1),
                                      "frac_of_hbm_peak": round(thr_alg / max(tm.kernel_ms[0], 1e-9) / 1e6 / measured_peak()[0], 4)})

    # ---- the widened paths on the same resident reads, for the record: device call + row-count read-back ----
    widened = None
    if rank == 0 and world == 1 and not args.no_cpu:
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

    # ---- end to end through the public API with HOST buffers (pack + H2D + kernels + D2H) ----
    e2e = None
    if not args.no_e2e:
        host = dev.to_host()
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
            rn, st, sa = host.rname, host.strand, host.start
            tid = np.arange(host.nrow, dtype=np.int32)
        n_loc = host.nrow
        del host
        e2e_steps = max(1, min(args.steps, 5))
        times = []
        rows_e2e = 0
        for i in range(e2e_steps + 2):  # two untimed passes: the two alternating sets of page-locked result blocks get allocated
            barrier()
            t1 = time.perf_counter()
            b = ea.BamData(rn, st, sa, tid, blob, off, device=local)
            pass_flags, rep = ea.generateCytosineReport(b, return_pass=True, **THRESH)
            h2d_actual = int(L.ea_pack_h2d_bytes(b.handle(need_bases=False)))
            calls_only = not bool(L.ea_has_bytes(b.handle(need_bases=False)))
            b.close()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t1
            if i > 1:
                times.append(dt)
            rows_e2e = rep["pos"].shape[0]
        e2e_s = float(np.median(times))  # a pack allocates and frees ~20 GB: single outliers of the allocator are not the path's speed
        h2d_input = int(blob.nbytes + off.nbytes + rn.nbytes + st.nbytes + sa.nbytes + tid.nbytes)
        d2h = int(n_loc + 24 * rows_e2e)
        e2e = dict(seconds=e2e_s, h2d=h2d_actual, h2d_input=h2d_input, calls_only=calls_only, d2h=d2h, pinned=pinned, steps=e2e_steps,
                   rows=rows_e2e, n_pass=int(pass_flags.sum()))

    ms_step = tot / args.steps
    stats = torch.tensor([ms_step, wall_ms / args.steps, win / args.steps, e2e["seconds"] if e2e else 0.0], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(own_reads), float(alg_bytes), float(dev.nbytes), float(n_rows), float(n_pass)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_step, wall_step, win_step, e2e_s = [float(x) for x in stats.tolist()]
    tot_reads, tot_alg, tot_xm, tot_rows, tot_pass = [float(x) for x in sums.tolist()]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(ea, cfg, min(args.cpu_rows, cfg.n_reads), check_gpu=True)

    kernel_name = (L.ea_report_kernel(dev.handle()) or b"").decode()
    if rank == 0:
        peak, peak_src = measured_peak()
        traffic, traffic_src = None, None
        try:  # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("reads_per_gpu") == args.reads:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        achieved = (alg_bytes / 1e9) / (win / args.steps / 1e3)   # this rank's dominant kernel
        line = {
            "metric": METRIC, "value": tot_reads / (ms_step / 1e3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic WGBS, %d reads x %d bp per GPU on one chromosome (%d bp x %d), CpG report, threshold.reads=TRUE (CG: 2 / 0.5 / 0.1)"
                       % (args.reads, READ_LEN, CHR1, world), "reads_per_gpu": args.reads, "read_len": READ_LEN, "report_context": "CG",
                       "threshold_context": "CG", "sharding": "contiguous genomic ranges with halo rows, no collective" if world > 1 else "single GPU",
                       "l2": "inputs (%.1f GB per GPU) are larger than L2 (126 MB); no flush needed" % (dev.nbytes / 1e9), "seed": SEED,
                       "report_rows": int(tot_rows), "reads_passing": int(tot_pass)},
            "xm_gbs": tot_xm / 1e9 / (ms_step / 1e3),
            "wall_ms_per_step": wall_step,
            "gpu_launches": launches * args.steps,
            "kernels_per_step": {"thr_tables": 1, "window(fused threshold+aggregate)": 1, "scan_windows (cub: init + scan)": 2, "gather_rows": 1},
            "roofline": {"bound": "hbm", "kernel": kernel_name + ", fused", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": win_step, "formula": "sum(L) + 14 N + 24 R",
                         "device_format": ("the kernel reads the packer's 4-bit bit-sliced call planes (16 bytes per 32 positions, tiles of 32 rows) plus 8 bytes "
                                           "of metadata per read, not the reference's byte per position: DRAM traffic per launch (`traffic`) is about half the "
                                           "algorithmic bytes, which are counted on the reference's format as SURVEY.md 8(d) defines them"),
                         "dram_gbs": (traffic / 1e9) / (win_step / 1e3) if traffic else None,
                         "dram_frac_of_peak": ((traffic / 1e9) / (win_step / 1e3) / peak) if traffic else None},
            "two_pass_variant": two_pass,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = {"value": tot_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "seconds_per_step": e2e_s, "pinned_host_memory": e2e["pinned"], "host_affinity": host_affinity, "steps": e2e["steps"],
                           "host_input_bytes_per_step": e2e["h2d_input"], "host_threads": os.environ.get("EA_HOST_PACK_THREADS", str(os.cpu_count())),
                           "what": ("ea_pack_blob_calls (host threads transpose the strings of the pinned input into 4-bit call planes chunk by chunk while the previous chunk "
                                    "crosses PCIe; validation + window index on the device) + fused report + D2H of pass flags and rows + free"
                                    if e2e["calls_only"] else
                                    "ea_pack_blob (H2D + device repack + validation + window index) + fused report + D2H of pass flags and rows + free")}
        if cpu:
            line["cpu_baseline"] = cpu
        if widened:
            line["widened_paths"] = widened
        emit(line)
    dev.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(ea, cfg, rows, check_gpu=False):
    """The reference's own three C++ files (oracle/_ref, else the C port) on `rows` leading rows of the workload."""
    from oracle.oracle import REF_SO, PortCpu, RefCpu
    host = ea.synth_host(cfg, 0, rows)
    kind = "reference" if os.path.exists(REF_SO) else "port"
    o = (RefCpu if kind == "reference" else PortCpu)(host)
    t0 = time.perf_counter()
    p, rep = o.generateCytosineReport(THRESH["threshold_reads"], "CG", 2, 0.5, 0.1, "CG")
    wall = time.perf_counter() - t0
    sec = o.last_seconds
    out = {"value": rows / sec, "unit": "reads/s", "cores": 1, "kind": kind, "seconds": sec, "wall_seconds": wall,
           "sample": "first %d rows of the same synthetic workload (the reference is single-threaded; %d host cores present)" % (rows, os.cpu_count()),
           "xm_gbs": float(host.blob()[1][-1]) / 1e9 / sec,
           "note": "flat_map is a sorted-vector stand-in for Boost (not installed); function time only, marshalling excluded"}
    if check_gpu:
        gp, grep = ea.generateCytosineReport(host, return_pass=True, **THRESH)
        ok = bool(np.array_equal(gp, p)) and all(np.array_equal(grep[k], rep[k]) for k in ("rname", "strand", "pos", "context", "meth", "unmeth"))
        out["gpu_matches_bit_exact"] = ok
        host.close()
    return out


def run_reference(args):
    """Reference arm: the reference's own CPU implementation, bounded sample per step, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import epialleler_b200 as ea
    from oracle.oracle import REF_SO, PortCpu, RefCpu
    world = args.gpus
    cfg = synth_cfg(ea, world, args.reads)
    rows = min(args.ref_rows, cfg.n_reads)
    host = ea.synth_host(cfg, 0, rows)
    kind = "reference" if os.path.exists(REF_SO) else "port"
    o = (RefCpu if kind == "reference" else PortCpu)(host)
    secs = []
    for i in range(max(args.warmup, 1) + args.steps):
        o.generateCytosineReport(True, "CG", 2, 0.5, 0.1, "CG")
        if i >= max(args.warmup, 1):
            secs.append(o.last_seconds)
    s = float(np.mean(secs))
    v = rows / s
    sample = "each step = first %d rows of the workload (single-threaded reference; %d host cores present)" % (rows, os.cpu_count())
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic WGBS, %d reads x %d bp per GPU, CpG report, threshold.reads=TRUE; bounded sample" % (args.reads, READ_LEN)},
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
    ap.add_argument("--ref-rows", type=int, default=500_000)
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
