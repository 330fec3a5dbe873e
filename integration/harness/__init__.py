"""ctypes front-end of integration/harness/shim_harness.cpp: the reader's `std::vector<std::string>` (one pageable heap
string per template) and the exact calls the R shim's handle_of() makes on it.  Used by bench.py's end-to-end leg and by
the packer tests; not part of the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libshim_harness.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        L = C.CDLL(SO)
        L.harness_strings_new.restype = C.c_void_p
        L.harness_strings_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.harness_strings_free.argtypes = [C.c_void_p]
        L.harness_strings_count.restype = C.c_int64
        L.harness_strings_count.argtypes = [C.c_void_p]
        L.harness_pack.restype = C.c_int
        L.harness_pack.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.harness_pack_records.restype = C.c_int
        L.harness_pack_records.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
        L.harness_pack_multi_records.restype = C.c_int
        L.harness_pack_multi_records.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


class ReaderStrings:
    """The templates of a host BamData as the reference's reader holds them (heap std::strings) + the four columns as
    plain pageable int vectors (R's integer vectors)."""

    def __init__(self, bam):
        blob, off = bam.blob()
        blob = np.ascontiguousarray(blob if blob.shape[0] else np.zeros(1, np.uint8))
        self.n = int(off.shape[0] - 1)
        self._v = lib().harness_strings_new(blob.ctypes.data, off.ctypes.data, self.n)
        self.cols = [np.array(getattr(bam, k), dtype=np.int32, copy=True) for k in ("rname", "strand", "start", "templid")]
        self.nrow = int(self.cols[0].shape[0])

    def pack(self, device=0, calls_only=True, records=False):
        """handle_of() of the shim -> epialleler_b200.api.DeviceBam.  records=True: the table-free call
        (ea_pack_calls_records on the vector's std::string objects); else pointer/length tables + ea_pack_calls / ea_pack."""
        from epialleler_b200 import _native as N
        from epialleler_b200.api import DeviceBam
        h = C.c_void_p()
        if records and calls_only:
            rc = lib().harness_pack_records(self._v, *[c.ctypes.data for c in self.cols], self.nrow, device, C.byref(h))
            if rc == -1000:
                raise RuntimeError("std::string layout not understood by the probe")
            N.check(rc)
        else:
            N.check(lib().harness_pack(self._v, *[c.ctypes.data for c in self.cols], self.nrow, device, 1 if calls_only else 0, C.byref(h)))
        return DeviceBam(h, device)

    def pack_multi(self, devices, rname_levels=()):
        """the shim's multi-GPU pack (ea_pack_multi_records) -> epialleler_b200.api.MultiBam"""
        from epialleler_b200 import _native as N
        from epialleler_b200.api import MultiBam
        dv = (C.c_int * len(devices))(*devices)
        m = C.c_void_p()
        rc = lib().harness_pack_multi_records(self._v, *[c.ctypes.data for c in self.cols], self.nrow, dv, len(devices), C.byref(m))
        if rc == -1000:
            raise RuntimeError("std::string layout not understood by the probe")
        N.check(rc)
        return MultiBam.from_handle(m, self.nrow, list(devices), rname_levels)

    def close(self):
        if self._v and _lib is not None:
            _lib.harness_strings_free(self._v)
            self._v = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
