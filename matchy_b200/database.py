"""Mirror of the reference's unified `Database` (crates/matchy/src/database.rs): open an `.mxy` file, upload its
sections unchanged into HBM, answer `lookup*` through the device tables.  One Database == one Engine (one GPU)."""
import ipaddress
import json
import mmap
import os

from . import builder as B
from . import engine as E


class DatabaseError(Exception):
    pass


class QueryResult:
    """`QueryResult::{Ip{data,prefix_len} | Pattern{pattern_ids,data} | NotFound}` (database.rs)."""
    __slots__ = ("kind", "data", "prefix_len", "pattern_ids")

    def __init__(self, kind, data=None, prefix_len=None, pattern_ids=None):
        self.kind, self.data, self.prefix_len, self.pattern_ids = kind, data, prefix_len, pattern_ids

    @staticmethod
    def not_found():
        return QueryResult("NotFound")

    def is_not_found(self):
        return self.kind == "NotFound"

    def __eq__(self, o):
        return isinstance(o, QueryResult) and (self.kind, self.data, self.prefix_len, self.pattern_ids) == (o.kind, o.data, o.prefix_len, o.pattern_ids)

    def __repr__(self):
        if self.kind == "Ip":
            return "QueryResult.Ip(prefix_len=%d, data=%r)" % (self.prefix_len, self.data)
        if self.kind == "Pattern":
            return "QueryResult.Pattern(pattern_ids=%r, data=%r)" % (self.pattern_ids, self.data)
        return "QueryResult.NotFound"


class DatabaseOpener:
    """`Database::from(path)[.cache_capacity(n) | .no_cache()].open()` (database.rs:306-308, 433, 586-618).
    The query cache of the reference is result-neutral and has no device counterpart; the knobs are accepted and ignored."""

    def __init__(self, path):
        self._path, self._device, self._chunk = path, 0, 0

    def cache_capacity(self, _n): return self
    def no_cache(self): return self
    def device(self, d): self._device = int(d); return self
    def chunk_bytes(self, n): self._chunk = int(n); return self

    def open(self):
        path = os.fspath(self._path)
        low = path.lower()
        if low.endswith(".csv"):
            return Database.from_bytes(B.build_from_csv(path), self._device, self._chunk)
        if low.endswith(".json"):
            return Database.from_bytes(B.build_from_json(path), self._device, self._chunk)
        try:
            with open(path, "rb") as f:
                if os.fstat(f.fileno()).st_size == 0:
                    raise DatabaseError("empty database file: " + path)
                mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        except OSError as e:
            raise DatabaseError(str(e))
        return Database(bytes(mm), self._device, self._chunk)


class Database:
    def __init__(self, mxy: bytes, device=0, chunk_bytes=0):
        self._bytes = mxy
        self.engine = E.Engine(device, chunk_bytes)
        try:
            self.engine.upload(mxy)
        except E.EngineError as e:
            raise DatabaseError(str(e))
        self._fmt = E.RecordFormatter(mxy)
        self._info = self.engine.db_info()

    # constructors --------------------------------------------------------------------------------------
    @staticmethod
    def from_(path):
        return DatabaseOpener(path)

    @staticmethod
    def open(path, device=0):
        return DatabaseOpener(path).device(device).open()

    @staticmethod
    def from_bytes(mxy: bytes, device=0, chunk_bytes=0):
        return Database(mxy, device, chunk_bytes)

    def close(self):
        self.engine.close()

    # capabilities (database.rs:1081-1098) ----------------------------------------------------------------
    def has_ip_data(self): return bool(self._info["has_ip"])
    def has_literal_data(self): return bool(self._info["has_literal"])
    def has_glob_data(self): return bool(self._info["has_glob"])
    def has_string_data(self): return self.has_literal_data() or self.has_glob_data()
    def mode(self): return "CaseInsensitive" if self._info["match_mode"] == 1 else "CaseSensitive"
    def info(self): return dict(self._info)
    def raw_bytes(self): return self._bytes

    # lookups ---------------------------------------------------------------------------------------------
    def decode(self, data_offset):
        return json.loads(self._fmt.data_json(data_offset))

    def lookup_ip(self, addr):
        """`lookup_ip` (database.rs:837-855): None when the DB has no IP tree, else Ip / NotFound."""
        if not self.has_ip_data():
            return None
        a = ipaddress.ip_address(addr) if not isinstance(addr, (ipaddress.IPv4Address, ipaddress.IPv6Address)) else addr
        found, off, pl = self.engine.lookup_ip(a.packed)
        if not found:
            return QueryResult.not_found()
        return QueryResult("Ip", data=self.decode(off), prefix_len=pl)

    def lookup_string(self, s):
        """`lookup_string_uncached` (database.rs:911-981): literal id first, then ascending glob ids."""
        if not self.has_string_data():
            return None
        q = s.encode("utf-8") if isinstance(s, str) else bytes(s)
        pairs = self.engine.lookup_string(q)
        if not pairs:
            return QueryResult.not_found()
        return QueryResult("Pattern", pattern_ids=[p for p, _ in pairs],
                           data=[None if off == E.NO_DATA else self.decode(off) for _, off in pairs])

    def lookup(self, query):
        """`lookup` (database.rs:725-804): try `query.parse::<IpAddr>()` first, else the string path."""
        text = query.decode("utf-8") if isinstance(query, (bytes, bytearray)) else query
        try:
            a = _rust_parse_ip(text)
        except ValueError:
            a = None
        if a is not None:
            return self.lookup_ip(a)
        return self.lookup_string(text)


def _rust_parse_ip(text):
    """Rust `IpAddr::from_str`: strict dotted quad (no leading zeros) or RFC 4291 text; no zone ids, no whitespace."""
    if not text or text != text.strip() or "%" in text or "/" in text:
        raise ValueError(text)
    if ":" in text:
        return ipaddress.IPv6Address(text)
    parts = text.split(".")
    if len(parts) != 4 or any((not p.isdigit()) or len(p) > 3 or (len(p) > 1 and p[0] == "0") or int(p) > 255 or not p.isascii() for p in parts):
        raise ValueError(text)
    return ipaddress.IPv4Address(text)
