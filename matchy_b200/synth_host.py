"""Host-only binding of the synthetic workload generators (csrc/synth.cpp + csrc/mxy_builder.cpp as libmatchy_synth.so).

`bench.py --impl reference` and the cpu_baseline leg build their database and log sample through THIS module, so that the
CPU arm maps no product library: libmatchy_b200.so (CUDA kernels, C ABI) is never loaded by it.  Same code, same bytes as
matchy_b200.synth (which goes through the product library and adds the device generator)."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmatchy_synth.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python __graft_entry__.py` (build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.mgen_db.restype = C.c_void_p
        L.mgen_db.argtypes = [C.c_int, C.c_double]
        L.mgen_log.argtypes = [C.c_int, C.c_double, C.c_uint64, C.c_void_p, C.c_size_t, C.c_int]
        L.mxyb_build.argtypes = [C.c_void_p]
        L.mxyb_bytes.restype = C.c_void_p
        L.mxyb_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.mxyb_free.argtypes = [C.c_void_p]
        L.mxyb_error.restype = C.c_char_p
        L.mxyb_error.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def build_db(config: int, scale: float = 1.0) -> bytes:
    L = lib()
    h = L.mgen_db(config, float(scale))
    if not h:
        raise ValueError("bad config/scale")
    try:
        if L.mxyb_build(h) != 0:
            raise ValueError(L.mxyb_error(h).decode())
        n = C.c_size_t()
        p = L.mxyb_bytes(h, C.byref(n))
        return C.string_at(p, n.value)
    finally:
        L.mxyb_free(h)


def gen_log(config: int, nbytes: int, scale: float = 1.0, offset: int = 0, threads: int = 0, out=None) -> np.ndarray:
    if out is None:
        out = np.empty(nbytes, dtype=np.uint8)
    rc = lib().mgen_log(config, float(scale), int(offset), C.c_void_p(out.ctypes.data), int(nbytes), int(threads))
    if rc != 0:
        raise ValueError("mgen_log failed: %d" % rc)
    return out
