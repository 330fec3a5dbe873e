"""ctypes binding of the C-ABI shared library (include/epialleler_b200.h).

This module plays the role the `.Call` shim plays in R: it passes plain host
pointers and sizes across the C ABI and nothing else.  There is NO CPU fallback: if the library is
missing or no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EA_LIB_PATH") or os.path.join(HERE, "lib", "libepialleler_b200.so")  # override: A/B builds only

EA_SYNTH_MAX_CHR = 32

EA_ERR_NEEDS_BYTES = -8
EA_ERR_ARG, EA_ERR_CUDA, EA_ERR_UNSORTED, EA_ERR_DOMAIN, EA_ERR_NODEVICE, EA_ERR_INTERRUPT, EA_ERR_STATE = -1, -2, -3, -4, -5, -6, -7
ERRORS = {-8: "EA_ERR_NEEDS_BYTES", -1: "EA_ERR_ARG", -2: "EA_ERR_CUDA", -3: "EA_ERR_UNSORTED", -4: "EA_ERR_DOMAIN", -5: "EA_ERR_NODEVICE",
          -6: "EA_ERR_INTERRUPT", -7: "EA_ERR_STATE"}


class EaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERRORS.get(code, code), msg))
        self.code = code


class ReportParams(C.Structure):
    _fields_ = [("threshold_reads", C.c_int32), ("ctx_meth", C.c_char_p), ("ctx_unmeth", C.c_char_p),
                ("ooctx_meth", C.c_char_p), ("ooctx_unmeth", C.c_char_p), ("min_n_ctx", C.c_uint32),
                ("min_ctx_meth_frac", C.c_double), ("max_ooctx_meth_frac", C.c_double), ("report_ctx", C.c_char_p)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("kernel_ms", C.c_float * 8), ("n_launches", C.c_int32),
                ("n_rows", C.c_int64), ("n_pass", C.c_int64), ("alg_bytes", C.c_int64)]


class SynthCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_int64), ("n_chr", C.c_int32), ("len_min", C.c_int32),
                ("len_max", C.c_int32), ("n_islands", C.c_int32), ("island_len", C.c_int32), ("reserved", C.c_int32),
                ("chr_len", C.c_int64 * EA_SYNTH_MAX_CHR)]


class Shard(C.Structure):
    """ea_shard: rows [row_lo, own_hi) packed (halo = [row_lo, own_lo)), keys [key_lo, key_hi) owned."""
    _fields_ = [("row_lo", C.c_int64), ("own_lo", C.c_int64), ("own_hi", C.c_int64), ("key_lo", C.c_int64), ("key_hi", C.c_int64)]


_I32P, _U8P, _I64P, _F64P = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_double)
_H = C.c_void_p

# name -> (restype, argtypes); every symbol include/epialleler_b200.h declares
SIGNATURES = {
    "ea_version": (C.c_int, []),
    "ea_last_error": (C.c_char_p, []),
    "ea_device_count": (C.c_int, []),
    "ea_set_poll": (None, [C.c_void_p]),
    "ea_set_handle_poll": (None, [_H, C.c_void_p]),
    "ea_set_kernel_policy": (C.c_int, [C.c_int]),
    "ea_report_kernel": (C.c_char_p, [_H]),
    "ea_mhl_report": (C.c_int, [_H, C.c_char_p, C.c_int32, C.c_int32, C.c_double, C.POINTER(C.c_int64)]),
    "ea_mhl_report_fetch": (C.c_int, [_H, _I32P, _I32P, _I32P, _I32P, _I32P, _F64P, _F64P]),
    "ea_match_target": (C.c_int, [_H, C.c_int, _I32P, _I32P, _I32P, C.c_int64, C.c_int32, _I32P]),
    "ea_bed_report": (C.c_int, [_H, C.c_int, _I32P, _I32P, _I32P, C.c_int64, C.c_int32, C.POINTER(ReportParams), _U8P, _I32P, _U8P, _I64P]),
    "ea_pack": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, C.POINTER(C.c_void_p), _I64P, C.c_int64, C.c_int, C.POINTER(_H)]),
    "ea_pack_blob": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, _U8P, _I64P, C.c_int64, C.c_int, C.POINTER(_H)]),
    "ea_pack_blob_calls": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, _U8P, _I64P, C.c_int64, C.c_int, C.POINTER(_H)]),
    "ea_has_bytes": (C.c_int, [_H]),
    "ea_pack_h2d_bytes": (C.c_int64, [_H]),
    "ea_attach_device": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int, C.c_int, C.POINTER(_H)]),
    "ea_free": (None, [_H]),
    "ea_trim": (C.c_int, [C.c_int]),
    "ea_cx_report_text": (C.c_int, [_H, C.POINTER(C.c_char_p), C.c_int32, C.c_int, C.POINTER(C.c_int64)]),
    "ea_cx_report_text_fetch": (C.c_int, [_H, C.c_void_p]),
    "ea_base_freqs": (C.c_int, [_H, _U8P, _I32P, _I32P, C.c_int64, _F64P]),
    "ea_fisher_exact": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, _F64P]),
    "ea_host_alloc": (C.c_void_p, [C.c_size_t]),
    "ea_host_free": (None, [C.c_void_p]),
    "ea_nrow": (C.c_int64, [_H]),
    "ea_nbytes": (C.c_int64, [_H]),
    "ea_device": (C.c_int, [_H]),
    "ea_threshold_reads": (C.c_int, [_H, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_double, C.c_double, _U8P]),
    "ea_get_xm_beta": (C.c_int, [_H, C.c_char_p, C.c_char_p, _F64P]),
    "ea_cx_report": (C.c_int, [_H, _U8P, C.c_char_p, _I64P]),
    "ea_cx_report_fetch": (C.c_int, [_H] + [_I32P] * 6),
    "ea_cytosine_report": (C.c_int, [_H, C.POINTER(ReportParams), _U8P, _I64P]),
    "ea_cytosine_report_device": (C.c_int, [_H, C.POINTER(ReportParams), C.c_int, C.POINTER(Timing)]),
    "ea_stage_name": (C.c_char_p, [C.c_int]),
    "ea_flush_l2": (C.c_int, [_H, C.c_int64]),
    "ea_synth_device": (C.c_int, [C.POINTER(SynthCfg), C.c_int64, C.c_int64, C.c_int, C.POINTER(_H)]),
    "ea_synth_host": (C.c_int, [C.POINTER(SynthCfg), C.c_int64, C.c_int64, _I32P, _I32P, _I32P, _I64P, _U8P]),
    "ea_export_host": (C.c_int, [_H, _I32P, _I32P, _I32P, _I64P, _U8P]),
    "ea_pack_calls": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, C.POINTER(C.c_void_p), _I64P, C.c_int64, C.c_int, C.POINTER(_H)]),
    "ea_pack_calls_records": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(_H)]),
    "ea_ingest_begin": (C.c_int, [C.c_int, C.c_int64, C.POINTER(_H)]),
    "ea_ingest_append": (C.c_int64, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "ea_ingest_finish": (C.c_int, [_H, _I32P, _I32P, _I32P, _I32P, C.POINTER(_H)]),
    "ea_ingest_abort": (None, [_H]),
    "ea_plan_shards": (C.c_int, [_I32P, _I32P, _I32P, C.c_int64, _I64P, C.c_int64, C.c_int, C.POINTER(Shard)]),
    "ea_set_owned_range": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64]),
    "ea_pack_multi": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, _U8P, _I64P, C.c_int64, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(_H)]),
    "ea_pack_multi_records": (C.c_int, [_I32P, _I32P, _I32P, _I32P, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int), C.c_int, C.POINTER(_H)]),
    "ea_multi_ndev": (C.c_int, [_H]),
    "ea_multi_handle": (_H, [_H, C.c_int]),
    "ea_multi_shard": (C.c_int, [_H, C.c_int, C.POINTER(Shard)]),
    "ea_multi_cytosine_report": (C.c_int, [_H, C.POINTER(ReportParams), _U8P, _I64P]),
    "ea_multi_report_fetch": (C.c_int, [_H] + [_I32P] * 6),
    "ea_multi_free": (None, [_H]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C epialleler_b200/csrc). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library ever diverge
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise EaError(rc, lib().ea_last_error().decode())
