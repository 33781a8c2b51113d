"""`.mxy` database writer — mirrors the reference's ``DatabaseBuilder``
(crates/matchy-format/src/mmdb_builder.rs:186-760) and the CSV/JSON/text loaders of `matchy match`
(crates/matchy/src/bin/commands/match_cmd.rs:34-171).  The bytes are produced by the C++ writer in
csrc/mxy_builder.cpp through the mxyb_* C ABI."""
import csv
import json

from . import _native as N


class MatchMode:
    CaseSensitive = 0
    CaseInsensitive = 1


class DatabaseBuilder:
    """``DatabaseBuilder::new(mode)``; ``add_entry`` auto-detects IP/CIDR, glob or literal (literal:/glob:/ip: prefixes)."""

    AUTO, IP, LITERAL, GLOB = 0, 1, 2, 3

    def __init__(self, match_mode=MatchMode.CaseSensitive, build_epoch=None):
        self._L = N.lib()
        self._h = self._L.mxyb_new(1 if match_mode == MatchMode.CaseInsensitive else 0)
        if build_epoch is not None:
            self._L.mxyb_set_epoch(self._h, int(build_epoch))
        self._data_cache = {}

    def __del__(self):
        try:
            if self._h:
                self._L.mxyb_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- data values: python type -> MMDB type the way the reference's loaders pick them
    def encode_data(self, data):
        """Encode a flat metadata map, returning its data-section offset (deduplicated)."""
        key = tuple(sorted((k, type(v).__name__, v) for k, v in (data or {}).items()))
        off = self._data_cache.get(key)
        if off is not None:
            return off
        L, h = self._L, self._h
        L.mxyb_data_begin(h)
        for k, v in (data or {}).items():
            kb = k.encode()
            if isinstance(v, bool):
                L.mxyb_data_bool(h, kb, int(v))
            elif isinstance(v, int):
                if -(2 ** 63) <= v < 2 ** 63:
                    L.mxyb_data_i32(h, kb, ((v + 2 ** 31) % 2 ** 32) - 2 ** 31)  # `i as i32` (match_cmd.rs:83-84)
                else:
                    L.mxyb_data_u64(h, kb, v)
            elif isinstance(v, float):
                L.mxyb_data_f64(h, kb, v)
            elif isinstance(v, U16):
                L.mxyb_data_u16(h, kb, v.v)
            elif isinstance(v, U32):
                L.mxyb_data_u32(h, kb, v.v)
            elif isinstance(v, U64):
                L.mxyb_data_u64(h, kb, v.v)
            else:
                vb = str(v).encode()
                L.mxyb_data_str(h, kb, vb, len(vb))
        off = L.mxyb_data_commit(h)
        self._data_cache[key] = off
        return off

    def _add(self, kind, key, data):
        kb = key.encode() if isinstance(key, str) else bytes(key)
        off = data if isinstance(data, int) and not isinstance(data, bool) else self.encode_data(data)
        if self._L.mxyb_add(self._h, kind, kb, len(kb), off) != 0:
            raise ValueError(self._L.mxyb_error(self._h).decode() or "invalid entry: %r" % key)

    def add_entry(self, key, data=None):
        self._add(self.AUTO, key, data)

    def add_ip(self, key, data=None):
        self._add(self.IP, key, data)

    def add_literal(self, key, data=None):
        self._add(self.LITERAL, key, data)

    def add_glob(self, key, data=None):
        self._add(self.GLOB, key, data)

    def set_database_type(self, t):
        self._L.mxyb_set_type(self._h, t.encode())

    def set_description(self, lang, text):
        self._L.mxyb_set_description(self._h, lang.encode(), text.encode())

    def stats(self):
        import ctypes as C
        out = (C.c_uint64 * 3)()
        self._L.mxyb_counts(self._h, out)
        return {"ip_entries": out[0], "literal_entries": out[1], "glob_entries": out[2]}

    def build(self) -> bytes:
        import ctypes as C
        if self._L.mxyb_build(self._h) != 0:
            raise ValueError(self._L.mxyb_error(self._h).decode())
        n = C.c_size_t()
        p = self._L.mxyb_bytes(self._h, C.byref(n))
        return C.string_at(p, n.value)


class U16:
    def __init__(self, v): self.v = int(v)


class U32:
    def __init__(self, v): self.v = int(v)


class U64:
    def __init__(self, v): self.v = int(v)


def _csv_value(value):
    """match_cmd.rs:82-93: i64 -> Int32, u64 -> Uint64, f64 -> Double, true/false -> Bool, else String."""
    try:
        if value.strip() == value and value and (value.lstrip("+-").isdigit()):
            i = int(value)
            if -(2 ** 63) <= i < 2 ** 63:
                return i
            if 0 <= i < 2 ** 64:
                return U64(i)
    except ValueError:
        pass
    try:
        if value.strip() == value and value.lower() not in ("nan", "inf", "+inf", "-inf", "infinity", "+infinity", "-infinity") or \
                value in ("NaN", "inf", "-inf", "+inf", "infinity", "-infinity", "+infinity"):
            return float(value)
    except ValueError:
        pass
    if value in ("true", "false"):
        return value == "true"
    return value


def add_csv_file(b, path):
    """One CSV input of `matchy build -f csv` / `matchy match DB.csv` (build_cmd.rs:168-229): the `entry` / `key` column
    is the indicator, every other non-empty cell a metadata value typed by _csv_value.  Returns the number of entries."""
    n = 0
    with open(path, newline="") as f:
        rd = csv.reader(f)
        headers = next(rd)
        try:
            col = next(i for i, h in enumerate(headers) if h in ("entry", "key"))
        except StopIteration:
            raise ValueError("CSV must have an 'entry' or 'key' column. Found headers: " + ", ".join(headers))
        for row in rd:
            if not row:
                continue
            data = {}
            for i, name in enumerate(headers):
                if i != col and i < len(row) and row[i] != "":
                    data[name] = _csv_value(row[i])
            b.add_entry(row[col], data)
            n += 1
    return n


def add_json_file(b, path):
    """`[{"key": ..., "data": {...}}, ...]` (build_cmd.rs:231-278)."""
    n = 0
    with open(path) as f:
        for k, item in enumerate(json.load(f)):
            if not isinstance(item, dict) or not isinstance(item.get("key"), str):
                raise ValueError("Missing 'key' field at index %d" % k)
            data = item.get("data") or {}
            flat = {k2: v for k2, v in data.items() if not isinstance(v, (dict, list))}
            if len(flat) != len(data):
                raise ValueError("nested metadata values are not supported by this writer yet")
            b.add_entry(item["key"], flat)
            n += 1
    return n


def add_text_file(b, path):
    """One indicator per line, blank lines and `#` comments skipped (build_cmd.rs:98-166)."""
    n = 0
    with open(path) as f:
        for line in f:
            e = line.strip()
            if e and not e.startswith("#"):
                b.add_entry(e, {})
                n += 1
    return n


def build_from_csv(path, build_epoch=None) -> bytes:
    b = DatabaseBuilder(MatchMode.CaseSensitive, build_epoch)
    add_csv_file(b, path)
    return b.build()


def build_from_json(path, build_epoch=None) -> bytes:
    b = DatabaseBuilder(MatchMode.CaseSensitive, build_epoch)
    add_json_file(b, path)
    return b.build()


def build_from_text(path, build_epoch=None) -> bytes:
    b = DatabaseBuilder(MatchMode.CaseSensitive, build_epoch)
    add_text_file(b, path)
    return b.build()
