"""The reference's PUBLIC C ABI (include/matchy/matchy.h: matchy_open / matchy_query / matchy_extractor_* / matchy_builder_*)
served by libmatchy_b200.so.  CPU part: the header is plain C, every declared symbol is exported, the builder half works
without a GPU and the device half refuses to.  GPU part: tests/capi/capi_check.c — a C program written against the
header alone, following the reference's own C smoke test (crates/matchy/tests/test_c_api.c) — runs end to end."""
import ctypes as C
import os
import re
import subprocess

import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "matchy", "matchy.h")
SRC = os.path.join(ROOT, "tests", "capi", "capi_check.c")


def _declared():
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(matchy_[a-z0-9_]+)\s*\(", text)))


def _compile(tmp_path):
    from matchy_b200 import _native as N
    exe = str(tmp_path / "capi_check")
    libdir = os.path.dirname(N.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, SRC,
                           "-L", libdir, "-lmatchy_b200", "-Wl,-rpath," + libdir])
    return exe


def test_public_abi_symbols_exported(built):
    from matchy_b200 import _native as N
    names = _declared()
    assert len(names) == 40, names
    lib = C.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_header_is_plain_c_and_program_links(built, tmp_path):
    _compile(tmp_path)
    # and the header also parses as C++ (namespace matchy + extern "C")
    cc = tmp_path / "hdr.cpp"
    cc.write_text('#include "matchy/matchy.h"\nint main() { return matchy::matchy_version() == nullptr; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(cc)])


def test_struct_layouts_match_the_reference_abi(built, tmp_path):
    """Sizes / offsets a cbindgen build of the reference produces on x86-64 (matchy.h:361-556)."""
    c = tmp_path / "lay.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "matchy/matchy.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", '
                 "sizeof(matchy_result_t), offsetof(matchy_result_t, _data_cache), sizeof(matchy_open_options_t), offsetof(matchy_open_options_t, reload_callback), "
                 "sizeof(matchy_entry_data_t), offsetof(matchy_entry_data_t, value), offsetof(matchy_entry_data_t, data_size), sizeof(matchy_match_t), "
                 "sizeof(matchy_matches_t), sizeof(matchy_stats_t)); return 0; }\n")
    exe = str(tmp_path / "lay")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(c)])
    assert subprocess.check_output([exe]).split() == [b"24", b"8", b"24", b"8", b"32", b"8", b"24", b"32", b"24", b"56"]


def _lib():
    from matchy_b200 import _native as N
    L = C.CDLL(N.LIB_PATH)
    L.matchy_builder_new.restype = C.c_void_p
    L.matchy_builder_add.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.matchy_builder_build.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    L.matchy_builder_free.argtypes = [C.c_void_p]
    L.matchy_builder_set_case_insensitive.argtypes = [C.c_void_p, C.c_bool]
    L.matchy_open_buffer.restype = C.c_void_p
    L.matchy_open_buffer.argtypes = [C.c_char_p, C.c_size_t]
    L.matchy_open.restype = C.c_void_p
    L.matchy_open.argtypes = [C.c_char_p]
    L.matchy_extractor_create.restype = C.c_void_p
    L.matchy_item_type_name.restype = C.c_char_p
    L.matchy_item_type_name.argtypes = [C.c_uint8]
    L.matchy_version.restype = C.c_char_p
    return L


def _build(L, entries, ci=False):
    b = L.matchy_builder_new()
    assert b
    if ci:
        assert L.matchy_builder_set_case_insensitive(b, True) == 0
    for k, j, rc in entries:
        assert L.matchy_builder_add(b, k, j) == rc, (k, j)
    buf, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    assert L.matchy_builder_build(b, C.byref(buf), C.byref(n)) == 0
    data = bytes(bytearray(buf[: n.value]))
    C.CDLL(None).free(buf)
    L.matchy_builder_free(b)
    return data


def test_builder_half_on_cpu_json_typing(built):
    """matchy_builder_add: serde number typing (matchy-data-format/src/lib.rs:120-163), {"value": x} wrapping, error codes;
    the produced file is read back by the oracle."""
    L = _lib()
    db = _build(L, [
        (b"10.1.0.0/16", b'{"u16":65535,"u32":65536,"u64":4294967296,"i32":-2147483648,"dbl":-2147483649,"f":2.5,"s":"a\\u00e9\\ud83d\\ude00\\n","t":true,"arr":[1,[2,{"k":"v"}]],"dup":1,"dup":2}', 0),
        (b"evil.com", b"42", 0),
        (b"*.evil.com", b'"x"', 0),
        (b"bad1", b"null", -2), (b"bad2", b'{"a":null}', -2), (b"bad3", b"{", -2), (b"bad4", b'{"a":1}x', -2), (b"bad5", b"[1,]", -2),
        (b"bad6", b"01", -2), (b"bad7", b'"\\ud800"', -2), (b"bad8", b"1e999", -2),
        (None, b"{}", -5), (b"k", None, -5),
    ])
    o = O.Oracle(db)
    found, off, plen = o.lookup_ip4((10 << 24) | (1 << 16) | (2 << 8) | 3)
    assert found and plen == 16
    import json
    v = json.loads(o.data_json(off))
    assert v == {"u16": 65535, "u32": 65536, "u64": 4294967296, "i32": -2147483648, "dbl": -2147483649.0, "f": 2.5, "s": "aé\U0001F600\n",
                 "t": True, "arr": [1, [2, {"k": "v"}]], "dup": 2}
    assert json.loads(o.data_json(o.lookup_string(b"evil.com")[0][1])) == {"value": 42}
    assert json.loads(o.data_json(o.lookup_string(b"a.evil.com")[0][1])) == {"value": "x"}
    assert o.lookup_string(b"bad1") == []
    # typed widths survive: 65535 is stored as uint16 (type 5), 65536 as uint32 (6), 2^32 as uint64 (9), negatives as int32 (8)
    assert L.matchy_item_type_name(11) == b"Monero" and L.matchy_item_type_name(12) == b"Unknown"
    assert L.matchy_version() == b"1.2.2"


def test_device_half_refuses_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib()
    db = _build(L, [(b"evil.com", b"{}", 0)])
    assert L.matchy_open_buffer(db, len(db)) is None
    assert L.matchy_open(b"/nonexistent") is None
    assert L.matchy_extractor_create(255) is None


@pytest.mark.gpu
def test_c_program_against_public_header(built, tmp_path):
    exe = _compile(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "capi.mxy")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "capi ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_query_and_extract_agree_with_oracle(built, small_dbs):
    """matchy_query / matchy_extractor_extract_chunk on config 1's database and log vs the oracle's lookups / extraction."""
    L = _lib()

    class Result(C.Structure):
        _fields_ = [("found", C.c_bool), ("prefix_len", C.c_uint8), ("cache", C.c_void_p), ("db", C.c_void_p)]

    class Match(C.Structure):
        _fields_ = [("item_type", C.c_uint8), ("value", C.c_char_p), ("start", C.c_size_t), ("end", C.c_size_t)]

    class Matches(C.Structure):
        _fields_ = [("items", C.POINTER(Match)), ("count", C.c_size_t), ("internal", C.c_void_p)]

    L.matchy_query.restype = Result
    L.matchy_query.argtypes = [C.c_void_p, C.c_char_p]
    L.matchy_free_result.argtypes = [C.POINTER(Result)]
    L.matchy_result_to_json.restype = C.c_void_p
    L.matchy_result_to_json.argtypes = [C.POINTER(Result)]
    L.matchy_free_string.argtypes = [C.c_void_p]
    L.matchy_close.argtypes = [C.c_void_p]
    L.matchy_extractor_extract_chunk.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(Matches)]
    L.matchy_matches_free.argtypes = [C.POINTER(Matches)]
    L.matchy_extractor_free.argtypes = [C.c_void_p]
    db, log = small_dbs[1]
    orc = O.Oracle(db)
    chunk = log[: 256 * 1024]
    chunk = chunk[: chunk.rfind(b"\n") + 1]
    ex = L.matchy_extractor_create(0x1F)  # everything but the crypto-address extractors
    assert ex
    ms = Matches()
    assert L.matchy_extractor_extract_chunk(ex, chunk, len(chunk), C.byref(ms)) == 0
    got = [(ms.items[k].item_type, ms.items[k].start, ms.items[k].end) for k in range(ms.count)]
    values = {chunk[ms.items[k].start:ms.items[k].end] for k in range(ms.count)}
    L.matchy_matches_free(C.byref(ms))
    L.matchy_extractor_free(ex)
    want = orc.extract(chunk, 0x1F)
    assert sorted(got) == sorted(want) and len(got) > 1000
    order = {3: 0, 2: 1, 1: 2, 0: 3, 4: 4, 5: 4, 6: 4, 7: 4, 8: 4}
    assert got == sorted(got, key=lambda t: (order[t[0]], t[1]))  # extract_from_chunk's grouping
    h = L.matchy_open_buffer(db, len(db))
    assert h
    import json
    n_found = 0
    for q in sorted(values)[:400] + [b"45.0.0.1", b"2001:db8::1", b"nope.invalid"]:
        r = L.matchy_query(h, q)
        try:
            text = q.decode()
            import ipaddress
            try:
                a = ipaddress.ip_address(text)
                if a.version == 4:
                    f, off, pl = orc.lookup_ip4(int(a))
                else:
                    f, off, pl = orc.lookup_ip6([int.from_bytes(a.packed[2 * k:2 * k + 2], "big") for k in range(8)])
                exp = (True, pl, off) if f else (False, 0, None)
            except ValueError:
                pairs = orc.lookup_string(q)
                exp = (True, 0, pairs[0][1]) if pairs and pairs[0][1] != 0xFFFFFFFF else (False, 0, None)
        except UnicodeDecodeError:
            exp = (False, 0, None)
        assert (r.found, r.prefix_len) == exp[:2], q
        if r.found:
            n_found += 1
            p = L.matchy_result_to_json(C.byref(r))
            assert json.loads(C.string_at(p)) == json.loads(orc.data_json(exp[2]))
            L.matchy_free_string(p)
        L.matchy_free_result(C.byref(r))
    L.matchy_close(h)


def test_data_codec_round_trips_of_the_reference(built):
    """The reference's data-format tests (matchy-data-format/src/lib.rs:1054-1290: map, array, complex nested structure, string
    size classes, deduplication) as writer -> reader round trips: our encoder (behind matchy_builder_add) against the oracle's
    decoder, which follows DataDecoder (lib.rs:635-1044)."""
    import json
    L = _lib()
    nested = {"threat_level": "high", "category": "malware", "confidence": 0.98, "first_seen": 1704067200,
              "indicators": {"ip_count": 42, "domain_count": 15}, "tags": ["botnet", "c2"], "active": True}
    sizes = {"short": "x" * 28, "medium": "x" * 100, "long": "x" * 1000, "huge": "y" * 70000}  # size classes <29, <285, <65821, beyond
    arr = {"value": ["tag1", "tag2", 123, False]}
    entries = [
        (b"10.0.0.1", json.dumps(nested).encode(), 0),
        (b"10.0.0.2", json.dumps({"country": "US", "asn": 13335, "score": 0.95}).encode(), 0),
        (b"10.0.0.3", json.dumps(arr["value"]).encode(), 0),          # bare array -> {"value": [...]}
        (b"10.0.0.4", json.dumps(sizes).encode(), 0),
        (b"10.0.0.5", json.dumps(nested).encode(), 0),                # same value again: deduplicated
        (b"same.example.com", json.dumps(nested).encode(), 0),
        (b"10.0.0.6", json.dumps({"a": {"b": {"c": {"d": [1, [2, [3, {"e": "deep"}]]]}}}}).encode(), 0),
        (b"10.0.0.7", b'{"u16max":65535,"u32min":65536,"u32max":4294967295,"u64min":4294967296,"u64max":18446744073709551615,"i32min":-2147483648,"m1":-1}', 0),
    ]
    db = _build(L, entries)
    o = O.Oracle(db)

    def data_of(last_octet):
        f, off, pl = o.lookup_ip4((10 << 24) | last_octet)
        assert f and pl == 32
        return off, json.loads(o.data_json(off))

    off1, v1 = data_of(1)
    assert v1 == nested
    assert data_of(2)[1] == {"country": "US", "asn": 13335, "score": 0.95}
    assert data_of(3)[1] == arr
    assert data_of(4)[1] == sizes
    off5, v5 = data_of(5)
    assert off5 == off1 and v5 == nested
    assert o.lookup_string(b"same.example.com")[0][1] == off1
    assert data_of(6)[1] == {"a": {"b": {"c": {"d": [1, [2, [3, {"e": "deep"}]]]}}}}
    assert data_of(7)[1] == {"u16max": 65535, "u32min": 65536, "u32max": 4294967295, "u64min": 4294967296, "u64max": 18446744073709551615,
                             "i32min": -2147483648, "m1": -1}


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_c_api", "test_c_api_extensions"])
def test_reference_c_test_programs_pass(built, name, tmp_path):
    """The reference's OWN C test programs (crates/matchy/tests/test_c_api.c, test_c_api_extensions.c), compiled unchanged
    against the reference's own header by __graft_entry__.build_reference_c_tests() and linked against libmatchy_b200.so,
    run to completion on the GPU.  (The sources stay in /root/reference; on a box without them the prebuilt binaries run.)"""
    exe = os.path.join(ROOT, "tests", "capi", "_ref", name)
    if not os.path.exists(exe):
        pytest.skip("reference C tests were not built (no /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
