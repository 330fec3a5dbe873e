"""Device-resident timing of the report pipeline on the OTHER BASELINE.json configurations (2b, 3, 4), scaled to one GPU.
Not a bench line (bench.py measures configs[1]); it shows which window kernel each data shape takes and how far from the
HBM roofline it runs.  Usage: python tools/bench_configs.py [name ...]"""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
import epialleler_b200 as ea  # noqa: E402
from epialleler_b200 import _native as N  # noqa: E402
from epialleler_b200.api import _params  # noqa: E402

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622,
        133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
PEAK = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6547.5) if False else 6547.5

CONFIGS = {
    # name: (synth_config kwargs, report kwargs)
    "cfg2_wgbs150": (dict(seed=20261018, n_reads=50_000_000, chr_len=[248956422], len_min=150),
                     dict(threshold_reads=True, threshold_context="CG", report_context="CG")),
    # the per-GPU shard of config 5 (1 B reads x 150 bp over hg38 on 8 GPUs): 125 M reads over an eighth of the genome, ~48x
    "cfg5_shard": (dict(seed=20261021, n_reads=125_000_000, chr_len=[248956422, 114364328, 25000000], len_min=150),
                   dict(threshold_reads=True, threshold_context="CG", report_context="CG")),
    "cfg2b_pe_merged": (dict(seed=20261018, n_reads=25_000_000, chr_len=[248956422], len_min=150, len_max=450),
                        dict(threshold_reads=True, threshold_context="CG", report_context="CG")),
    "cfg3_capture_cx": (dict(seed=20261019, n_reads=50_000_000, chr_len=HG38, len_min=100, len_max=417, n_islands=5000, island_len=2000),
                        dict(threshold_reads=True, threshold_context="CG", report_context="CX")),
    "deep150": (dict(seed=7, n_reads=10_000_000, chr_len=HG38, len_min=150, n_islands=1000, island_len=2000),
                dict(threshold_reads=True, threshold_context="CG", report_context="CG")),
    "cfg3_small": (dict(seed=20261019, n_reads=10_000_000, chr_len=HG38, len_min=100, len_max=417, n_islands=1000, island_len=2000),
                   dict(threshold_reads=True, threshold_context="CG", report_context="CX")),
    "cfg4_small": (dict(seed=20261020, n_reads=250_000, chr_len=HG38[:6], len_min=10000, len_max=30000),
                   dict(threshold_reads=False, threshold_context="CG", report_context="CG")),
    "cfg4_long_nothr": (dict(seed=20261020, n_reads=1_000_000, chr_len=HG38, len_min=10000, len_max=30000),
                        dict(threshold_reads=False, threshold_context="CG", report_context="CG")),
    "cfg4_long_thr": (dict(seed=20261020, n_reads=1_000_000, chr_len=HG38, len_min=10000, len_max=30000),
                      dict(threshold_reads=True, threshold_context="CG", report_context="CG")),
}


def main():
    names = sys.argv[1:] or list(CONFIGS)
    L = N.lib()
    import os
    L.ea_set_kernel_policy(int(os.environ.get("EA_POLICY", "0")))
    for name in names:
        skw, rkw = CONFIGS[name]
        cfg = ea.synth_config(**skw)
        dev = ea.synth_device(cfg)
        prm = _params(min_context_sites=2, min_context_beta=0.5, max_outofcontext_beta=0.1, **rkw)
        tm = N.Timing()
        best = None
        for i in range(6):
            N.check(L.ea_cytosine_report_device(dev.handle(), C.byref(prm), 0, C.byref(tm)))
            if i >= 2 and (best is None or tm.total_ms < best[0]):
                best = (tm.total_ms, tm.kernel_ms[0], tm.kernel_ms[1], tm.kernel_ms[2])
        kern = (L.ea_report_kernel(dev.handle()) or b"").decode()
        alg = int(tm.alg_bytes)
        print(json.dumps({"config": name, "reads": dev.nrow, "bytes": dev.nbytes, "rows": int(tm.n_rows), "n_pass": int(tm.n_pass),
                          "kernel": kern, "total_ms": round(best[0], 3), "threshold_ms": round(best[1], 3), "window_ms": round(best[2], 3),
                          "gather_ms": round(best[3], 3), "alg_gbs_window": round(alg / best[2] / 1e6, 1),
                          "frac_window": round(alg / best[2] / 1e6 / PEAK, 4), "greads_per_s": round(dev.nrow / best[0] / 1e6, 3)}), flush=True)
        dev.close()
        L.ea_trim(0)


if __name__ == "__main__":
    main()
