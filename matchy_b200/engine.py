"""Thin object wrapper over the mgpu_* C ABI: one Engine == one mgpu_ctx == one GPU."""
import ctypes as C

import numpy as np

from . import _native as N

X_DOMAINS, X_EMAILS, X_IPV4, X_IPV6, X_HASHES = 1, 2, 4, 8, 16
X_BITCOIN, X_ETHEREUM, X_MONERO = 32, 64, 128
X_CRYPTO = X_BITCOIN | X_ETHEREUM | X_MONERO
X_SUPPORTED = 0xFF
ITEM_TYPE_NAMES = ["Domain", "Email", "IPv4", "IPv6", "MD5", "SHA1", "SHA256", "SHA384", "SHA512", "Bitcoin", "Ethereum", "Monero"]
KIND_IP, KIND_PATTERN = 1, 2
NO_DATA = 0xFFFFFFFF
KERNEL_NAMES = ["tokenize", "token", "iptrie", "lithash", "strings"]  # MGPU_K_* order (include/matchy_b200.h)


class EngineError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise EngineError("%s failed (%d): %s" % (what, rc, N.last_error()))


class Engine:
    def __init__(self, device=0, chunk_bytes=0, psl_path=None, fused=None):
        """fused: None = the library's default first stage (tokenize_kernel + token_kernel unless MATCHY_B200_FUSED is set);
        True / False = the single-pass scan_kernel / the kernel pair for this engine (the choice is fixed at creation)."""
        import os
        self.L = N.lib()
        saved = {k: os.environ.get(k) for k in ("MATCHY_B200_FUSED", "MATCHY_B200_UNFUSED")}
        try:
            if fused is not None:
                os.environ.pop("MATCHY_B200_FUSED", None); os.environ.pop("MATCHY_B200_UNFUSED", None)
                os.environ["MATCHY_B200_FUSED" if fused else "MATCHY_B200_UNFUSED"] = "1"
            self.h = self.L.mgpu_create(int(device), int(chunk_bytes))
        finally:
            if fused is not None:
                for k, v in saved.items():
                    os.environ.pop(k, None)
                    if v is not None:
                        os.environ[k] = v
        if not self.h:
            raise EngineError("mgpu_create(device=%d) failed: %s (matchy_b200 has no CPU fallback)" % (device, N.last_error()))
        self.device = device
        with open(psl_path or N.PSL_PATH, "rb") as f:
            psl = f.read()
        _check(self.L.mgpu_set_psl(self.h, psl, len(psl)), "mgpu_set_psl")
        self._db_bytes = None

    def close(self):
        if getattr(self, "h", None):
            self.L.mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- database
    def upload(self, mxy: bytes):
        self._db_bytes = mxy
        p, n, keep = N.as_ptr(mxy)
        _check(self.L.mgpu_db_upload(self.h, p, n), "mgpu_db_upload")

    def db_info(self):
        info = N.MgpuDbInfo()
        _check(self.L.mgpu_db_info_get(self.h, C.byref(info)), "mgpu_db_info_get")
        return {k: getattr(info, k) for k, _ in N.MgpuDbInfo._fields_}

    def default_flags(self):
        return self.L.mgpu_default_flags(self.h)

    # -- scanning
    def scan(self, data, flags=None, base=0):
        """Scan a host buffer (bytes / numpy uint8).  Returns (records, ids) as numpy structured arrays (views of the engine's
        result buffers, valid until the next scan; see results())."""
        if flags is None:
            flags = self.default_flags()
        p, n, keep = N.as_ptr(data)
        _check(self.L.mgpu_scan(self.h, p, n, int(base), int(flags)), "mgpu_scan")
        return self.results()

    def scan_device(self, dev_ptr, nbytes, flags=None, base=0):
        if flags is None:
            flags = self.default_flags()
        _check(self.L.mgpu_scan_device(self.h, C.c_void_p(dev_ptr), int(nbytes), int(base), int(flags)), "mgpu_scan_device")
        return self.results()

    REC_DTYPE = np.dtype([("offset", "<u8"), ("len", "<u4"), ("item_type", "u1"), ("kind", "u1"), ("prefix_len", "u1"), ("reserved", "u1"),
                          ("n_ids", "<u4"), ("ids_index", "<u4"), ("data_offset", "<u4"), ("pad", "<u4")])
    ID_DTYPE = np.dtype([("pattern_id", "<u4"), ("data_offset", "<u4")])

    def results(self):
        recs = C.POINTER(N.MgpuMatch)()
        ids = C.POINTER(N.MgpuIdPair)()
        nr, ni = C.c_size_t(), C.c_size_t()
        _check(self.L.mgpu_results(self.h, C.byref(recs), C.byref(nr), C.byref(ids), C.byref(ni)), "mgpu_results")
        # zero-copy views of the engine's pinned result buffers: like the C ABI's, they stay valid until the next scan on
        # this engine (take .copy() to keep them longer)
        r = (np.frombuffer((C.c_uint8 * (nr.value * 32)).from_address(C.addressof(recs.contents)), dtype=self.REC_DTYPE)
             if nr.value else np.zeros(0, self.REC_DTYPE))
        i = (np.frombuffer((C.c_uint8 * (ni.value * 8)).from_address(C.addressof(ids.contents)), dtype=self.ID_DTYPE)
             if ni.value else np.zeros(0, self.ID_DTYPE))
        return r, i

    def records_as_tuples(self):
        """[(offset, len, item_type, kind, prefix_len, data_offset, ((pattern_id, data_offset), ...)), ...] sorted."""
        r, i = self.results()
        out = []
        for k in range(len(r)):
            x = r[k]
            a, n = int(x["ids_index"]), int(x["n_ids"])
            pairs = tuple((int(i[j]["pattern_id"]), int(i[j]["data_offset"])) for j in range(a, a + n))
            out.append((int(x["offset"]), int(x["len"]), int(x["item_type"]), int(x["kind"]), int(x["prefix_len"]), int(x["data_offset"]), pairs))
        out.sort()
        return out

    def counters(self):
        c = N.MgpuCounters()
        _check(self.L.mgpu_counters_get(self.h, C.byref(c)), "mgpu_counters_get")
        return {"lines": int(c.lines), "bytes": int(c.bytes), "candidates": int(c.candidates), "matches": int(c.matches),
                "by_type": [int(v) for v in c.by_type]}

    def counters_list(self):
        c = self.counters()
        return [c["lines"], c["bytes"], c["candidates"], c["matches"]] + c["by_type"]

    def timing(self):
        t = N.MgpuTiming()
        _check(self.L.mgpu_timing_get(self.h, C.byref(t)), "mgpu_timing_get")
        return {"kernel_ms": {KERNEL_NAMES[k]: float(t.kernel_ms[k]) for k in range(5)},
                "launches": {KERNEL_NAMES[k]: int(t.launches[k]) for k in range(5)},
                "total_ms": float(t.total_ms), "chunks": int(t.chunks), "scan_ms": float(t.scan_ms), "aux_launches": int(t.aux_launches)}

    def set_keep_results(self, keep: bool):
        self.L.mgpu_set_keep_results(self.h, 1 if keep else 0)

    def set_ac_mode(self, mode: int):
        """0: automatic; 1: force the reference's Aho-Corasick walk instead of anchored walks (same results)."""
        self.L.mgpu_set_ac_mode(self.h, int(mode))

    def set_option(self, key: str, value: int):
        """Test / debug switches of the engine (include/matchy_b200.h: mgpu_set_option)."""
        _check(self.L.mgpu_set_option(self.h, key.encode(), int(value)), "mgpu_set_option(%s)" % key)

    def debug_counters(self):
        """Sums of the token-list audit since set_option("verify_tokens", 1)."""
        out = (C.c_uint64 * 64)()
        _check(self.L.mgpu_debug_get(self.h, out), "mgpu_debug_get")
        self.debug_events = [(int(out[16 + 2 * k]) >> 32, int(out[16 + 2 * k]) & 0xFFFFFFFF, int(out[17 + 2 * k]) >> 32, int(out[17 + 2 * k]) & 0xFFFFFFFF)
                             for k in range(min(21, int(out[13])))]  # (block, base, warp, active mask) of the first partial-warp events
        self.host_us = dict(zip(("scan_call", "sort", "id_repack", "launch_and_gather"), (int(out[60 + k]) for k in range(4))))  # last resident scan, host side
        self.arrived_sorted = bool(out[59])  # the last scan's records came off the device already in (offset, item_type, len) order
        names = ("poisoned", "padding", "ipv4", "ipv6", "lookup_hits", "unknown", "slots", "_7", "trie_thread_hits", "trie_warp_hits", "trie_block_hits", "trie_partial_warps", "trie_n_mismatch")
        return dict(zip(names, (int(x) for x in out)))

    def extract_array(self, data, flags=X_SUPPORTED):
        """uint64 array [n, 3] of (item_type, start, end) rows sorted by (start, item_type)."""
        p, n, keep = N.as_ptr(data)
        cap = 1 << 16
        while True:
            out = (C.c_uint64 * (3 * cap))()
            cnt = self.L.mgpu_extract(self.h, p, n, int(flags), out, cap)
            if cnt < 0:
                _check(int(cnt), "mgpu_extract")
            if cnt <= cap:
                return np.frombuffer(out, dtype=np.uint64, count=3 * cnt).reshape(-1, 3).copy()
            cap = int(cnt)

    def extract(self, data, flags=X_SUPPORTED):
        """[(item_type, start, end), ...] sorted by (start, item_type)."""
        return [(int(t), int(s), int(e)) for t, s, e in self.extract_array(data, flags)]

    def lookup_string(self, q: bytes):
        out = (N.MgpuIdPair * 4096)()
        n = self.L.mgpu_lookup_string(self.h, q, len(q), out, 4096)
        if n < 0:
            _check(n, "mgpu_lookup_string")
        return [(int(out[k].pattern_id), int(out[k].data_offset)) for k in range(min(n, 4096))]

    def lookup_ip(self, packed: bytes):
        """packed = 4 (IPv4) or 16 (IPv6) network-order bytes -> (found, data_offset, prefix_len)."""
        is6 = len(packed) == 16
        buf = packed + b"\0" * (16 - len(packed))
        off, pl = C.c_uint32(), C.c_uint8()
        rc = self.L.mgpu_lookup_ip(self.h, buf, 1 if is6 else 0, C.byref(off), C.byref(pl))
        if rc < 0:
            _check(rc, "mgpu_lookup_ip")
        return bool(rc), off.value, pl.value

    # -- HBM-resident inputs
    def dev_alloc(self, nbytes):
        p = self.L.mgpu_dev_alloc(self.h, int(nbytes))
        if not p:
            raise EngineError("mgpu_dev_alloc(%d) failed: %s" % (nbytes, N.last_error()))
        return p

    def dev_free(self, p):
        self.L.mgpu_dev_free(self.h, C.c_void_p(p))

    def dev_upload(self, dst, data, dst_offset=0):
        p, n, keep = N.as_ptr(data)
        _check(self.L.mgpu_dev_upload(self.h, C.c_void_p(dst + dst_offset), p, n), "mgpu_dev_upload")

    def flush_l2(self):
        _check(self.L.mgpu_flush_l2(self.h), "mgpu_flush_l2")


class RecordFormatter:
    """Host-side rendering of matched records (mxyr_*): data decode + `matchy match` NDJSON."""

    def __init__(self, mxy: bytes):
        self.L = N.lib()
        self._keep = mxy
        p, n, _ = N.as_ptr(mxy)
        self.h = self.L.mxyr_open(p, n)
        if not self.h:
            raise EngineError("not a valid .mxy database")

    def __del__(self):
        try:
            if self.h:
                self.L.mxyr_close(self.h)
                self.h = None
        except Exception:
            pass

    def data_json(self, data_offset: int) -> str:
        out = C.c_char_p()
        n = self.L.mxyr_data_json(self.h, int(data_offset), C.byref(out))
        return C.string_at(out, n).decode("utf-8")

    def ndjson(self, recs: np.ndarray, ids: np.ndarray, log, base=0, source="", timestamp=None) -> bytes:
        """`matchy match` lines for these records.  timestamp=None: parallel mode ("0.000", raw matched_text); a string:
        sequential / follow mode (that timestamp, canonical address text — bin/match_processor/sequential.rs:205-390)."""
        if len(recs) == 0:
            return b""
        p, n, keep = N.as_ptr(log)
        recs = np.ascontiguousarray(recs)
        ids = np.ascontiguousarray(ids) if len(ids) else np.zeros(1, Engine.ID_DTYPE)
        out = C.c_void_p()
        rp = C.cast(C.c_void_p(recs.ctypes.data), C.POINTER(N.MgpuMatch))
        ip = C.cast(C.c_void_p(ids.ctypes.data), C.POINTER(N.MgpuIdPair))
        if timestamp is None:
            ln = self.L.mxyr_ndjson(self.h, rp, len(recs), ip, p, int(base), source.encode(), C.byref(out))
        else:
            ln = self.L.mxyr_ndjson_sequential(self.h, rp, len(recs), ip, p, int(base), source.encode(), str(timestamp).encode(), C.byref(out))
        return C.string_at(out, ln)
