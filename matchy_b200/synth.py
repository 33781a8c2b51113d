"""Synthetic databases and logs for the BASELINE.json configs (bench + tests), via the mgen_* C ABI (csrc/synth.cpp)."""
import ctypes as C

import numpy as np

from . import _native as N

BLOCK = 65536


def build_db(config: int, scale: float = 1.0) -> bytes:
    """The config's `.mxy` database, written by the in-tree writer (reference format)."""
    L = N.lib()
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
    """nbytes of the config's log stream starting at byte `offset` (a multiple of 64 KiB); newline-terminated."""
    L = N.lib()
    if out is None:
        out = np.empty(nbytes, dtype=np.uint8)
    rc = L.mgen_log(config, float(scale), int(offset), C.c_void_p(out.ctypes.data), int(nbytes), int(threads))
    if rc != 0:
        raise ValueError("mgen_log failed: %d" % rc)
    return out


def gen_log_device(device: int, config: int, dev_ptr: int, nbytes: int, scale: float = 1.0, offset: int = 0) -> None:
    """Fill device memory [dev_ptr, dev_ptr + nbytes) on GPU `device` with the same bytes gen_log() would produce (offset and
    nbytes multiples of 64 KiB).  Raises without a CUDA device: the generator runs as a kernel (csrc/synth_device.cu)."""
    rc = N.lib().mgen_log_device(int(device), int(config), float(scale), int(offset), C.c_void_p(dev_ptr), int(nbytes))
    if rc != 0:
        raise ValueError("mgen_log_device failed: %d" % rc)
