"""ctypes binding of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under matchy_b200/ does.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SRC = os.path.join(ROOT, "oracle", "oracle.cpp")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
PSL_PATH = os.path.join(ROOT, "matchy_b200", "data", "public_suffix_list.dat")

X_DOMAINS, X_EMAILS, X_IPV4, X_IPV6, X_HASHES = 1, 2, 4, 8, 16
X_DEFAULT = X_DOMAINS | X_EMAILS | X_IPV4 | X_IPV6 | X_HASHES
TYPE_NAMES = ["Domain", "Email", "IPv4", "IPv6", "MD5", "SHA1", "SHA256", "SHA384", "SHA512", "Bitcoin", "Ethereum", "Monero"]


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(ORACLE_SRC):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", ORACLE_SO, ORACLE_SRC])
    return ORACLE_SO


class MatchRec(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("len", C.c_uint32), ("item_type", C.c_uint8), ("kind", C.c_uint8),
                ("prefix_len", C.c_uint8), ("reserved", C.c_uint8), ("n_ids", C.c_uint32), ("ids_index", C.c_uint32),
                ("data_offset", C.c_uint32), ("pad", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int]
        L.orc_error.restype = C.c_char_p
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_default_flags.restype = C.c_uint32
        L.orc_default_flags.argtypes = [C.c_void_p]
        L.orc_has.argtypes = [C.c_void_p, C.c_int]
        L.orc_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_size_t]
        L.orc_scan_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_scan_mt_keep.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_set_cache.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_n_matches.restype = C.c_size_t
        L.orc_n_matches.argtypes = [C.c_void_p]
        L.orc_matches.restype = C.POINTER(MatchRec)
        L.orc_matches.argtypes = [C.c_void_p]
        L.orc_n_ids.restype = C.c_size_t
        L.orc_n_ids.argtypes = [C.c_void_p]
        L.orc_ids.restype = C.POINTER(C.c_uint32)
        L.orc_ids.argtypes = [C.c_void_p]
        L.orc_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_digest.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p]
        L.orc_cryptoaddr.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        L.orc_extract.restype = C.c_size_t
        L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint64), C.c_size_t]
        L.orc_ndjson.restype = C.c_void_p
        L.orc_ndjson.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.POINTER(C.c_size_t)]
        L.orc_lookup_ip4.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
        L.orc_lookup_ip6.argtypes = [C.c_void_p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
        L.orc_parse_ipv6.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint16)]
        L.orc_lookup_string.restype = C.c_size_t
        L.orc_lookup_string.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t]
        L.orc_data_json.restype = C.c_void_p
        L.orc_data_json.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_size_t)]
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_ipv6_display.restype = C.c_char_p
        L.orc_ipv6_display.argtypes = [C.POINTER(C.c_uint16)]
        _lib = L
    return _lib


def _buf(data):
    """bytes / bytearray / numpy array → (pointer, length, keepalive)."""
    if isinstance(data, (bytes, bytearray)):
        b = (C.c_char * len(data)).from_buffer_copy(data) if len(data) else (C.c_char * 1)()
        return C.cast(b, C.c_void_p), len(data), b
    import numpy as np
    a = np.ascontiguousarray(data, dtype=np.uint8)
    return C.c_void_p(a.ctypes.data), a.size, a


class Oracle:
    """One opened .mxy database + extractor, as `matchy match` would have them."""

    def __init__(self, mxy: bytes):
        self.L = lib()
        self._db = bytes(mxy)
        self.h = self.L.orc_open(self._db, len(self._db), PSL_PATH.encode(), 1)
        err = self.L.orc_error(self.h).decode()
        if err:
            raise RuntimeError("oracle: " + err)

    def close(self):
        if self.h:
            self.L.orc_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_flags(self):
        return self.L.orc_default_flags(self.h)

    def set_cache(self, capacity: int):
        """Per-thread LRU query cache of scan_mt / scan_mt_keep (`matchy match --cache-size`, default 10000; 0 = off)."""
        self.L.orc_set_cache(self.h, int(capacity))

    def extract(self, data, flags=X_DEFAULT):
        p, n, keep = _buf(data)
        cap = max(16, n)
        out = (C.c_uint64 * (3 * cap))()
        cnt = self.L.orc_extract(self.h, p, n, flags, out, cap)
        assert cnt <= cap
        return [(int(out[3 * k]), int(out[3 * k + 1]), int(out[3 * k + 2])) for k in range(cnt)]

    def extract_strings(self, data: bytes, flags=X_DEFAULT):
        return [(TYPE_NAMES[t], data[s:e]) for t, s, e in self.extract(data, flags)]

    def scan(self, data, flags=None, base=0, chunk_size=0):
        """Returns (records, counters); records = sorted list of
        (offset, len, item_type, kind, prefix_len, data_offset, ((pattern_id, data_offset), ...))."""
        if flags is None:
            flags = self.default_flags()
        p, n, keep = _buf(data)
        rc = self.L.orc_scan(self.h, p, n, base, flags, chunk_size)
        if rc != 0:
            raise RuntimeError("oracle scan failed rc=%d" % rc)
        return self._collect(), self.counters()

    def scan_mt(self, data, flags=None, threads=0):
        """Multi-threaded scan of newline-aligned shards (cpu_baseline leg); counters only."""
        if flags is None:
            flags = self.default_flags()
        p, n, keep = _buf(data)
        out = (C.c_uint64 * 16)()
        rc = self.L.orc_scan_mt(self.h, p, n, flags, threads, out)
        if rc != 0:
            raise RuntimeError("oracle scan_mt failed rc=%d" % rc)
        return [int(x) for x in out]

    REC_DTYPE = [("offset", "<u8"), ("len", "<u4"), ("item_type", "u1"), ("kind", "u1"), ("prefix_len", "u1"), ("reserved", "u1"),
                 ("n_ids", "<u4"), ("ids_index", "<u4"), ("data_offset", "<u4"), ("pad", "<u4")]
    ID_DTYPE = [("pattern_id", "<u4"), ("data_offset", "<u4")]

    def scan_mt_keep(self, data, flags=None, threads=0, base=0):
        """Multi-threaded scan that keeps the records: (recs, ids, counters) with recs / ids as numpy structured arrays in
        exactly the layout and order mgpu_results() returns (sorted by (offset, item_type, len), ids re-packed in record
        order), so a multi-GiB record-exact comparison is two array compares."""
        import numpy as np
        if flags is None:
            flags = self.default_flags()
        p, n, keep = _buf(data)
        out = (C.c_uint64 * 16)()
        rc = self.L.orc_scan_mt_keep(self.h, p, n, base, flags, threads, out)
        if rc != 0:
            raise RuntimeError("oracle scan_mt_keep failed rc=%d" % rc)
        nr, ni = self.L.orc_n_matches(self.h), self.L.orc_n_ids(self.h)
        recs = (np.frombuffer((C.c_uint8 * (nr * 32)).from_address(C.addressof(self.L.orc_matches(self.h).contents)), dtype=np.dtype(self.REC_DTYPE)).copy()
                if nr else np.zeros(0, np.dtype(self.REC_DTYPE)))
        ids = (np.frombuffer((C.c_uint8 * (ni * 8)).from_address(C.addressof(self.L.orc_ids(self.h).contents)), dtype=np.dtype(self.ID_DTYPE)).copy()
               if ni else np.zeros(0, np.dtype(self.ID_DTYPE)))
        return recs, ids, [int(x) for x in out]

    def _collect(self):
        n = self.L.orc_n_matches(self.h)
        recs = self.L.orc_matches(self.h)
        ids = self.L.orc_ids(self.h)
        out = []
        for k in range(n):
            r = recs[k]
            pairs = tuple((int(ids[2 * (r.ids_index + j)]), int(ids[2 * (r.ids_index + j) + 1])) for j in range(r.n_ids))
            out.append((int(r.offset), int(r.len), int(r.item_type), int(r.kind), int(r.prefix_len), int(r.data_offset), pairs))
        out.sort()
        return out

    def counters(self):
        c = (C.c_uint64 * 16)()
        self.L.orc_counters(self.h, c)
        return [int(x) for x in c]

    def ndjson(self, data, source="", base=0):
        p, n, keep = _buf(data)
        ln = C.c_size_t()
        ptr = self.L.orc_ndjson(self.h, p, base, source.encode(), C.byref(ln))
        return C.string_at(ptr, ln.value)

    def lookup_ip4(self, addr: int):
        off, pl = C.c_uint32(), C.c_uint8()
        rc = self.L.orc_lookup_ip4(self.h, addr, C.byref(off), C.byref(pl))
        return (rc, off.value, pl.value)

    def lookup_ip6(self, segs):
        a = (C.c_uint16 * 8)(*segs)
        off, pl = C.c_uint32(), C.c_uint8()
        rc = self.L.orc_lookup_ip6(self.h, a, C.byref(off), C.byref(pl))
        return (rc, off.value, pl.value)

    def lookup_string(self, q: bytes):
        out = (C.c_uint32 * 512)()
        n = self.L.orc_lookup_string(self.h, q, len(q), out, 256)
        return [(int(out[2 * k]), int(out[2 * k + 1])) for k in range(min(n, 256))]

    def data_json(self, off: int):
        ln = C.c_size_t()
        ptr = self.L.orc_data_json(self.h, off, C.byref(ln))
        return C.string_at(ptr, ln.value).decode()


def parse_ipv6(s: bytes):
    a = (C.c_uint16 * 8)()
    ok = lib().orc_parse_ipv6(s, len(s), a)
    return list(a) if ok else None


def ipv6_display(segs):
    a = (C.c_uint16 * 8)(*segs)
    return lib().orc_ipv6_display(a).decode()


def xxh64(b: bytes):
    return lib().orc_xxh64(b, len(b))


def digest(what: str, msg: bytes) -> bytes:
    """SHA-256 / Keccak-256 as the oracle's crypto-address validators compute them (known-answer tests)."""
    out = C.create_string_buffer(32)
    lib().orc_digest({"sha256": 0, "keccak256": 1}[what], msg, len(msg), out)
    return out.raw


def cryptoaddr_valid(kind: str, text: bytes) -> bool:
    return bool(lib().orc_cryptoaddr({"bitcoin_base58": 0, "bitcoin_bech32": 1, "ethereum": 2, "monero": 3}[kind], text, len(text)))
