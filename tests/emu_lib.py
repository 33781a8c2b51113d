"""ctypes binding of tests/host_emulation/emu.cpp (host emulation of the device logic).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "host_emulation", "emu.cpp"),
       os.path.join(ROOT, "matchy_b200", "csrc", "mxy_reader.cpp"),
       os.path.join(ROOT, "matchy_b200", "csrc", "mxy_builder.cpp")]
DEPS = SRC + [os.path.join(ROOT, "matchy_b200", "csrc", f) for f in ("device_fns.cuh", "tokenize.cuh", "crypto_addr.cuh", "db_prepare.h", "mxy_reader.h", "mxy_builder.h", "host_sort.h", "synth_gen.h")]
SO = os.path.join(ROOT, "tests", "host_emulation", "libemu.so")
PSL_PATH = os.path.join(ROOT, "matchy_b200", "data", "public_suffix_list.dat")


class Match(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("len", C.c_uint32), ("item_type", C.c_uint8), ("kind", C.c_uint8),
                ("prefix_len", C.c_uint8), ("reserved", C.c_uint8), ("n_ids", C.c_uint32), ("ids_index", C.c_uint32),
                ("data_offset", C.c_uint32), ("pad", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("lines", C.c_uint64), ("bytes", C.c_uint64), ("candidates", C.c_uint64), ("matches", C.c_uint64), ("by_type", C.c_uint64 * 12)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(p) for p in DEPS):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", SO] + SRC)
        L = C.CDLL(SO)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_char_p, C.c_size_t]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_error.restype = C.c_char_p
        L.emu_error.argtypes = [C.c_void_p]
        L.emu_db_upload.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.emu_set_anchored.argtypes = [C.c_void_p, C.c_int]
        L.emu_is_anchored_exact.argtypes = [C.c_void_p]
        L.emu_set_generic.argtypes = [C.c_void_p, C.c_int]
        L.emu_is_fast.argtypes = [C.c_void_p]
        L.emu_filter_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.emu_default_flags.restype = C.c_uint32
        L.emu_default_flags.argtypes = [C.c_void_p]
        L.emu_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int]
        L.emu_results.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Match)), C.POINTER(C.c_size_t), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_size_t)]
        L.emu_counters_get.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.emu_tokens.restype = C.c_int64
        L.emu_tokens.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t]
        L.emu_sort_records.restype = C.c_double
        L.emu_sort_records.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint, C.c_int]
        L.emu_records_sorted.restype = C.c_int
        L.emu_records_sorted.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
        L.emu_repack_ids.restype = C.c_size_t
        L.emu_repack_ids.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint]
        _lib = L
    return _lib


class Emu:
    def __init__(self, mxy: bytes = None):
        self.L = lib()
        psl = open(PSL_PATH, "rb").read()
        self.h = self.L.emu_create(psl, len(psl))
        assert self.h
        if mxy is not None:
            rc = self.L.emu_db_upload(self.h, mxy, len(mxy))
            if rc != 0:
                raise RuntimeError("emu upload: " + self.L.emu_error(self.h).decode())

    def __del__(self):
        try:
            if self.h:
                self.L.emu_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def default_flags(self):
        return self.L.emu_default_flags(self.h)

    def set_anchored(self, on: bool):
        """Choose between the anchored literal search and the Aho-Corasick formulation (both must give the oracle's answer)."""
        self.L.emu_set_anchored(self.h, 1 if on else 0)

    def set_generic(self, on: bool):
        """True: never use the constant-time string filters (generic lithash + acglob path)."""
        self.L.emu_set_generic(self.h, 1 if on else 0)

    def is_fast(self):
        """Does the uploaded database qualify for the fast string path?"""
        return bool(self.L.emu_is_fast(self.h))

    def filter_stats(self):
        """(tokens passing the literal filter, tokens passing the glob filter, tokens tested) of the last scan."""
        out = (C.c_uint64 * 3)()
        self.L.emu_filter_stats(self.h, out)
        return tuple(int(x) for x in out)

    def anchored_exact(self):
        return bool(self.L.emu_is_anchored_exact(self.h))

    def scan(self, data: bytes, flags=None, base=0, chunk_bytes=0, nwarps=1, misalign=0, lookups=True):
        if flags is None:
            flags = self.default_flags()
        buf = (C.c_char * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
        rc = self.L.emu_scan(self.h, C.cast(buf, C.c_void_p), len(data), base, flags, chunk_bytes, nwarps, misalign, 1 if lookups else 0)
        if rc != 0:
            raise RuntimeError("emu scan rc=%d" % rc)
        recs = C.POINTER(Match)()
        ids = C.POINTER(C.c_uint32)()
        nr, ni = C.c_size_t(), C.c_size_t()
        self.L.emu_results(self.h, C.byref(recs), C.byref(nr), C.byref(ids), C.byref(ni))
        out = []
        for k in range(nr.value):
            r = recs[k]
            pairs = tuple((int(ids[2 * (r.ids_index + j)]), int(ids[2 * (r.ids_index + j) + 1])) for j in range(r.n_ids))
            out.append((int(r.offset), int(r.len), int(r.item_type), int(r.kind), int(r.prefix_len), int(r.data_offset), pairs))
        out.sort()
        c = Counters()
        self.L.emu_counters_get(self.h, C.byref(c))
        cnt = [int(c.lines), int(c.bytes), int(c.candidates), int(c.matches)] + [int(x) for x in c.by_type]
        return out, cnt

    def tokens(self, data: bytes, flags=31, chunk_bytes=0, nwarps=1, misalign=0):
        self.scan(data, flags=flags, chunk_bytes=chunk_bytes, nwarps=nwarps, misalign=misalign, lookups=False)
        cap = max(16, 2 * len(data))
        out = (C.c_uint64 * (3 * cap))()
        n = self.L.emu_tokens(self.h, out, cap)
        return [(int(out[3 * k]), int(out[3 * k + 1]), int(out[3 * k + 2])) for k in range(n)]
