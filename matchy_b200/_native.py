"""ctypes loader for libmatchy_b200.so (CUDA kernels + C ABI, see include/matchy_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback: if the shared
object is missing, or there is no CUDA device when an engine is created, this raises.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libmatchy_b200.so")
PSL_PATH = os.path.join(PKG_DIR, "data", "public_suffix_list.dat")


class MgpuMatch(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("len", C.c_uint32), ("item_type", C.c_uint8), ("kind", C.c_uint8),
                ("prefix_len", C.c_uint8), ("reserved", C.c_uint8), ("n_ids", C.c_uint32), ("ids_index", C.c_uint32),
                ("data_offset", C.c_uint32), ("pad", C.c_uint32)]


class MgpuIdPair(C.Structure):
    _fields_ = [("pattern_id", C.c_uint32), ("data_offset", C.c_uint32)]


class MgpuCounters(C.Structure):
    _fields_ = [("lines", C.c_uint64), ("bytes", C.c_uint64), ("candidates", C.c_uint64), ("matches", C.c_uint64),
                ("by_type", C.c_uint64 * 12)]


class MgpuTiming(C.Structure):
    _fields_ = [("kernel_ms", C.c_float * 5), ("launches", C.c_uint32 * 5), ("total_ms", C.c_float), ("chunks", C.c_uint32),
                ("scan_ms", C.c_float), ("aux_launches", C.c_uint32)]


class MgpuDbInfo(C.Structure):
    _fields_ = [("node_count", C.c_uint32), ("record_bits", C.c_uint32), ("ip_version", C.c_uint32), ("match_mode", C.c_uint32),
                ("has_ip", C.c_uint32), ("has_literal", C.c_uint32), ("has_glob", C.c_uint32),
                ("literal_count", C.c_uint32), ("glob_count", C.c_uint32), ("ac_node_count", C.c_uint32),
                ("tree_bytes", C.c_uint64), ("literal_bytes", C.c_uint64), ("paraglob_bytes", C.c_uint64), ("file_bytes", C.c_uint64)]


# every symbol include/matchy_b200.h declares: name -> (restype, argtypes)
_VP, _SZ, _U8P = C.c_void_p, C.c_size_t, C.c_void_p
SYMBOLS = {
    "mgpu_create": (_VP, [C.c_int, _SZ]),
    "mgpu_destroy": (None, [_VP]),
    "mgpu_last_error": (C.c_char_p, []),
    "mgpu_set_psl": (C.c_int, [_VP, _U8P, _SZ]),
    "mgpu_db_upload": (C.c_int, [_VP, _U8P, _SZ]),
    "mgpu_db_info_get": (C.c_int, [_VP, C.POINTER(MgpuDbInfo)]),
    "mgpu_default_flags": (C.c_uint32, [_VP]),
    "mgpu_scan": (C.c_int, [_VP, _U8P, _SZ, C.c_uint64, C.c_uint32]),
    "mgpu_scan_device": (C.c_int, [_VP, _U8P, _SZ, C.c_uint64, C.c_uint32]),
    "mgpu_results": (C.c_int, [_VP, C.POINTER(C.POINTER(MgpuMatch)), C.POINTER(_SZ), C.POINTER(C.POINTER(MgpuIdPair)), C.POINTER(_SZ)]),
    "mgpu_counters_get": (C.c_int, [_VP, C.POINTER(MgpuCounters)]),
    "mgpu_timing_get": (C.c_int, [_VP, C.POINTER(MgpuTiming)]),
    "mgpu_set_keep_results": (None, [_VP, C.c_int]),
    "mgpu_set_ac_mode": (None, [_VP, C.c_int]),
    "mgpu_set_option": (C.c_int, [_VP, C.c_char_p, C.c_uint64]),
    "mgpu_debug_get": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "mgpu_extract": (C.c_int64, [_VP, _U8P, _SZ, C.c_uint32, C.POINTER(C.c_uint64), _SZ]),
    "mgpu_lookup_string": (C.c_int, [_VP, _U8P, _SZ, C.POINTER(MgpuIdPair), _SZ]),
    "mgpu_lookup_ip": (C.c_int, [_VP, _U8P, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]),
    "mgpu_dev_alloc": (_VP, [_VP, _SZ]),
    "mgpu_dev_free": (None, [_VP, _VP]),
    "mgpu_dev_upload": (C.c_int, [_VP, _VP, _VP, _SZ]),
    "mgpu_dev_download": (C.c_int, [_VP, _VP, _VP, _SZ]),
    "mgpu_host_alloc_pinned": (_VP, [_SZ]),
    "mgpu_host_free_pinned": (None, [_VP]),
    "mgpu_flush_l2": (C.c_int, [_VP]),
    "mxyr_open": (_VP, [_U8P, _SZ]),
    "mxyr_close": (None, [_VP]),
    "mxyr_data_json": (_SZ, [_VP, C.c_uint32, C.POINTER(C.c_char_p)]),
    "mxyr_ndjson": (_SZ, [_VP, C.POINTER(MgpuMatch), _SZ, C.POINTER(MgpuIdPair), _U8P, C.c_uint64, C.c_char_p, C.POINTER(_VP)]),
    "mxyr_ndjson_sequential": (_SZ, [_VP, C.POINTER(MgpuMatch), _SZ, C.POINTER(MgpuIdPair), _U8P, C.c_uint64, C.c_char_p, C.c_char_p, C.POINTER(_VP)]),
    "mxyb_new": (_VP, [C.c_int]),
    "mxyb_free": (None, [_VP]),
    "mxyb_error": (C.c_char_p, [_VP]),
    "mxyb_data_begin": (None, [_VP]),
    "mxyb_data_str": (None, [_VP, C.c_char_p, C.c_char_p, _SZ]),
    "mxyb_data_i32": (None, [_VP, C.c_char_p, C.c_int32]),
    "mxyb_data_u16": (None, [_VP, C.c_char_p, C.c_uint16]),
    "mxyb_data_u32": (None, [_VP, C.c_char_p, C.c_uint32]),
    "mxyb_data_u64": (None, [_VP, C.c_char_p, C.c_uint64]),
    "mxyb_data_f64": (None, [_VP, C.c_char_p, C.c_double]),
    "mxyb_data_bool": (None, [_VP, C.c_char_p, C.c_int]),
    "mxyb_data_commit": (C.c_uint32, [_VP]),
    "mxyb_add": (C.c_int, [_VP, C.c_int, C.c_char_p, _SZ, C.c_uint32]),
    "mxyb_set_epoch": (None, [_VP, C.c_uint64]),
    "mxyb_set_type": (None, [_VP, C.c_char_p]),
    "mxyb_set_description": (None, [_VP, C.c_char_p, C.c_char_p]),
    "mxyb_build": (C.c_int, [_VP]),
    "mxyb_bytes": (_VP, [_VP, C.POINTER(_SZ)]),
    "mxyb_counts": (None, [_VP, C.POINTER(C.c_uint64)]),
    "mxyb_save": (C.c_int, [_VP, C.c_char_p]),
    "mxyb_xxh64": (C.c_uint64, [C.c_char_p, _SZ]),
    "mgen_db": (_VP, [C.c_int, C.c_double]),
    "mgen_log": (C.c_int, [C.c_int, C.c_double, C.c_uint64, _VP, _SZ, C.c_int]),
    "mgen_log_device": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_uint64, _VP, _SZ]),
}

_lib = None


def lib():
    """The loaded library with typed entry points.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "matchy_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().mgpu_last_error().decode("utf-8", "replace")


def as_ptr(data):
    """bytes / bytearray / memoryview / numpy uint8 array -> (void pointer, length, keepalive)."""
    if isinstance(data, bytes):
        return C.cast(C.c_char_p(data), C.c_void_p), len(data), data
    if isinstance(data, (bytearray, memoryview)):
        mv = memoryview(data)
        if mv.readonly:
            b = bytes(mv)
            return C.cast(C.c_char_p(b), C.c_void_p), len(b), b
        arr = (C.c_char * len(mv)).from_buffer(mv) if len(mv) else (C.c_char * 1)()
        return C.cast(arr, C.c_void_p), len(mv), arr
    import numpy as np
    a = np.ascontiguousarray(data, dtype=np.uint8)
    return C.c_void_p(a.ctypes.data), int(a.size), a
