"""CPU suite: the C-ABI library loads and exports everything include/matchy_b200.h declares; host-side pieces
(writer, section locator, NDJSON renderer, generators, API mirror) behave; no compute calls need a GPU."""
import ctypes as C
import json
import os
import re

import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from matchy_b200 import _native as N
    hdr = open(os.path.join(ROOT, "include", "matchy_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:mgpu|mxyr|mxyb|mgen)_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    lib = C.CDLL(N.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(N.SYMBOLS), declared ^ set(N.SYMBOLS)


def test_no_cpu_fallback(built):
    """Without a CUDA device engine creation fails loudly (this container has none)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from matchy_b200 import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)


def test_product_does_not_import_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "matchy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "host_emulation" not in text.replace("tests/host_emulation can", ""), f


def test_writer_layout_and_metadata(built):
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1700000000)
    b.add_entry("10.0.0.0/8", {"threat_level": "low", "score": 1})
    b.add_entry("evil.com", {"threat_level": "high", "score": 5})
    b.add_entry("*.evil.com", {"threat_level": "high", "score": 5})  # same data -> deduplicated offset
    db = b.build()
    assert db[-2000:].count(b"\xab\xcd\xefMaxMind.com") == 1
    assert b"MMDB_PATTERN\0\0\0\0" in db and b"MMDB_LITERAL\0\0\0\0" in db and b"PARAGLOB" in db and b"LHSH" in db and b"ACLH" in db
    o = O.Oracle(db)
    lit, glob = o.lookup_string(b"evil.com"), o.lookup_string(b"x.evil.com")
    assert lit[0][1] == glob[0][1]
    assert json.loads(o.data_json(lit[0][1])) == {"score": 5, "threat_level": "high"}
    assert b.stats() == {"ip_entries": 1, "literal_entries": 1, "glob_entries": 1}


def test_entry_type_detection(built):
    """mmdb_builder.rs:392-429."""
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1)
    for key in ("1.2.3.4", "10.0.0.0/8", "2001:db8::/32", "ip:9.9.9.9"):
        b.add_entry(key, {})
    for key in ("*.example.com", "file[0-9].txt", "glob:no-wildcards.com"):
        b.add_entry(key, {})
    for key in ("evil.com", "literal:*.not-a-glob.com", "[unclosed", "1.2.3.4/33"):
        b.add_entry(key, {})
    assert b.stats() == {"ip_entries": 4, "literal_entries": 4, "glob_entries": 3}
    with pytest.raises(ValueError):
        b.add_entry("glob:[unclosed", {})
    with pytest.raises(ValueError):
        b.add_entry("ip:not-an-ip", {})


def test_24bit_record_overflow_is_refused(built):
    """The reference silently truncates records that do not fit (matchy-ip-trie/src/lib.rs:467-475); the writer refuses."""
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1)
    big = "x" * 4000
    for i in range(4300):
        b.add_entry("10.%d.%d.0/24" % (i // 256, i % 256), {"blob": big + str(i)})
    with pytest.raises(ValueError):
        b.build()


def test_ndjson_renderer_matches_oracle(small_dbs):
    """mxyr_ndjson (product, host) == oracle NDJSON for the same records."""
    import numpy as np
    from matchy_b200 import Engine, RecordFormatter
    for cfg in (1, 3, 5):
        db, log = small_dbs[cfg]
        log = log[:300000] + b"src=2001:d00::1 ::ffff:1.2.3.4 x\n"
        orc = O.Oracle(db)
        recs, _ = orc.scan(log, chunk_size=128 * 1024)
        r = np.zeros(len(recs), Engine.REC_DTYPE)
        ids = []
        for k, (off, ln, ty, kind, pl, doff, pairs) in enumerate(recs):
            r[k] = (off, ln, ty, kind, pl, 0, len(pairs), len(ids), doff, 0)
            ids.extend(pairs)
        i = np.array(ids, Engine.ID_DTYPE) if ids else np.zeros(0, Engine.ID_DTYPE)
        got = RecordFormatter(db).ndjson(r, i, log, 0, 'we"ird\\src.log')
        orc.scan(log, chunk_size=128 * 1024)
        assert sorted(got.splitlines()) == sorted(orc.ndjson(log, 'we"ird\\src.log').splitlines())
        for line in got.splitlines()[:50]:
            obj = json.loads(line)
            assert list(obj) == sorted(obj) and obj["timestamp"] == "0.000"


def test_generators_are_deterministic_and_block_addressable(built):
    from matchy_b200 import synth
    a = synth.gen_log(2, 4 * 65536, 0.01)
    b = synth.gen_log(2, 2 * 65536, 0.01, offset=2 * 65536)
    assert bytes(a[2 * 65536:]) == bytes(b)
    assert bytes(synth.gen_log(2, 4 * 65536, 0.01, threads=3)) == bytes(a)
    assert a[-1] == 10 and a[65535] == 10
    c = synth.gen_log(3, 100000, 0.01)
    assert c[-1] == 10
    assert synth.build_db(1, 0.01) == synth.build_db(1, 0.01)


def test_file_reader_next_batch(tmp_path):
    """FileReader::next_batch (processing/mod.rs:206-251): every batch ends at a newline except the tail."""
    from matchy_b200.processing import FileReader
    p = tmp_path / "x.log"
    lines = [b"line %d 1.2.3.%d\n" % (i, i % 250) for i in range(1000)]
    p.write_bytes(b"".join(lines) + b"tail-without-newline")
    fr = FileReader(p, chunk_size=1000)
    got = [b.data for b in fr.batches()]
    assert b"".join(got) == p.read_bytes()
    assert all(g.endswith(b"\n") for g in got[:-1]) and got[-1] == b"tail-without-newline"
    assert max(len(g) for g in got) <= 1000 + 30


def test_csv_loader_value_typing(tmp_path, built):
    """match_cmd.rs:77-97: i64 -> Int32, f64 -> Double, true/false -> Bool, else String; empty cells skipped."""
    from matchy_b200 import builder
    p = tmp_path / "db.csv"
    p.write_text("entry,score,ratio,active,note\n1.2.3.4,7,0.5,true,hello\nevil.com,,,false,\n*.bad.org,-3,x,TRUE,y z\n")
    db = builder.build_from_csv(str(p), build_epoch=1)
    o = O.Oracle(db)
    rc, off, pl = o.lookup_ip4(0x01020304)
    assert json.loads(o.data_json(off)) == {"active": True, "note": "hello", "ratio": 0.5, "score": 7}
    assert json.loads(o.data_json(o.lookup_string(b"evil.com")[0][1])) == {"active": False}
    assert json.loads(o.data_json(o.lookup_string(b"a.bad.org")[0][1])) == {"active": "TRUE", "note": "y z", "ratio": "x", "score": -3}


def test_oracle_multithreaded_counters_equal_single_thread(small_dbs):
    db, log = small_dbs[5]
    o = O.Oracle(db)
    _, cnt = o.scan(log, chunk_size=128 * 1024)
    assert o.scan_mt(log, threads=4) == cnt


def test_build_subcommand(built, tmp_path, capsys):
    """`python -m matchy_b200 build` == the `matchy build` contract (bin/commands/build_cmd.rs): text / csv / json inputs, several
    files, metadata options, read-only output; the file is read back through the oracle."""
    import stat
    from matchy_b200.__main__ import main
    txt = tmp_path / "a.txt"
    txt.write_text("# comment\n\n10.0.0.0/8\n  evil.com  \n*.bad.org\nliteral:*.star.com\n")
    txt2 = tmp_path / "b.txt"
    txt2.write_text("2001:db8::/32\nother.net\n")
    out = tmp_path / "t.mxy"
    assert main(["build", str(txt), str(txt2), "-o", str(out), "-t", "MyCompany-Intel", "-d", "test db", "--desc-lang", "de", "-v"]) == 0
    msg = capsys.readouterr().out
    assert "Total entries:   6" in msg and "IP entries:      2" in msg and "Literal entries: 3" in msg and "Glob entries:    1" in msg
    assert stat.S_IMODE(os.stat(out).st_mode) == 0o444
    db = out.read_bytes()
    assert b"MyCompany-Intel" in db[-600:] and b"test db" in db[-600:] and b"de" in db[-600:]
    o = O.Oracle(db)
    assert o.lookup_ip4(10 << 24 | 5)[0] and o.lookup_ip6([0x2001, 0xdb8, 1, 0, 0, 0, 0, 1])[0]
    assert o.lookup_string(b"evil.com") and o.lookup_string(b"x.bad.org") and o.lookup_string(b"*.star.com") and not o.lookup_string(b"a.star.com")
    # rebuilding over the read-only file works (fs::write would fail on 0444 only for non-owners; root owns it here)
    assert main(["build", str(txt), "-o", str(out)]) == 0
    assert "Database built: " in capsys.readouterr().out
    csvf = tmp_path / "c.csv"
    csvf.write_text("entry,threat_level,score,ratio,flag,note\n1.2.3.4,high,7,0.5,true,\nEvil.COM,low,-3,,false,\"a,b\"\n")
    assert main(["build", str(csvf), "-o", str(out), "-f", "csv", "-i"]) == 0
    o = O.Oracle(out.read_bytes())
    f, off, pl = o.lookup_ip4(0x01020304)
    assert f and pl == 32 and json.loads(o.data_json(off)) == {"threat_level": "high", "score": 7, "ratio": 0.5, "flag": True}
    hit = o.lookup_string(b"evil.com")  # case-insensitive database
    assert hit and json.loads(o.data_json(hit[0][1])) == {"threat_level": "low", "score": -3, "flag": False, "note": "a,b"}
    js = tmp_path / "d.json"
    js.write_text(json.dumps([{"key": "9.9.9.9", "data": {"k": "v", "n": 5}}, {"key": "*.x.io"}]))
    assert main(["build", str(js), "-o", str(out), "-f", "json"]) == 0
    o = O.Oracle(out.read_bytes())
    f, off, pl = o.lookup_ip4(0x09090909)
    assert f and json.loads(o.data_json(off)) == {"k": "v", "n": 5}
    assert o.lookup_string(b"a.x.io")
    capsys.readouterr()
    assert main(["build", str(js), "-o", str(out), "-f", "xml"]) == 1
    assert main(["build", str(js), "-o", str(out), "-f", "misp"]) == 1
    assert main(["build", str(txt), "-o", str(out), "-t", "threatdb"]) == 1
    bad = tmp_path / "bad.csv"
    bad.write_text("a,b\n1,2\n")
    assert main(["build", str(bad), "-o", str(out), "-f", "csv"]) == 1
    assert "must have an 'entry' or 'key' column" in capsys.readouterr().err


def test_build_from_misp_events(built, tmp_path, capsys):
    """`matchy build -f misp` (misp_importer.rs): event + attribute + object metadata, composite values split, ignored types,
    non-event files skipped, malformed events refused."""
    from matchy_b200.__main__ import main
    from matchy_b200.misp_importer import domain_of_url
    # the reference's own unit vectors (misp_importer.rs:1158-1177)
    assert domain_of_url("http://example.com/path") == "example.com"
    assert domain_of_url("https://test.org:8080/") == "test.org"
    assert domain_of_url("example.net") == "example.net"
    assert domain_of_url("http://evil.com?param=value") == "evil.com"
    ev = {"Event": {"uuid": "u-1", "info": "Test Event", "threat_level_id": "2", "analysis": 1, "date": "2024-05-06", "Orgc": {"name": "CIRCL"},
                    "Tag": [{"name": "tlp:green"}, {"name": "apt"}],
                    "Attribute": [{"type": "ip-src", "value": "192.168.1.1", "category": "Network activity", "to_ids": True, "comment": "", "Tag": [{"name": "c2"}]},
                                  {"type": "ip-dst|port", "value": "10.1.2.3|443"},
                                  {"type": "domain|ip", "value": "bad.example|172.16.0.9", "comment": "pair"},
                                  {"type": "url", "value": "https://dl.evil.test:8080/a.exe"},
                                  {"type": "filename|sha256", "value": "x.exe|" + "ab" * 32},
                                  {"type": "ip-src/netmask", "value": "203.0.113.0/24"},
                                  {"type": "comment", "value": "never.added.example"},
                                  {"type": "port", "value": 8080},
                                  {"type": "md5", "value": None},
                                  {"type": "brand-new-type", "value": "odd-value"},
                                  {"type": "brand-new-type", "value": "z" * 1000}],
                    "Object": [{"name": "file", "comment": "dropper", "Attribute": [{"type": "sha1", "value": "cd" * 20, "object_relation": "sha1"}]}]}}
    f1 = tmp_path / "event1.json"
    f1.write_text(json.dumps(ev))
    (tmp_path / "manifest.json").write_text(json.dumps({"u-1": {"info": "x"}}))
    (tmp_path / "notes.json").write_text("[1, 2, 3]")
    out = tmp_path / "m.mxy"
    assert main(["build", str(f1), str(tmp_path / "manifest.json"), str(tmp_path / "notes.json"), "-o", str(out), "-f", "misp", "-v"]) == 0
    cap = capsys.readouterr()
    assert "Skipped 2 non-MISP file(s)" in cap.err and "manifest.json: metadata file" in cap.err and "notes.json: not a MISP event" in cap.err
    assert "IP entries:      4" in cap.out and "Glob entries:    0" in cap.out
    db = out.read_bytes()
    assert b"MISP-ThreatIntel" in db[-600:]
    o = O.Oracle(db)
    base = {"event_info": "Test Event", "event_uuid": "u-1", "threat_level": "Medium", "analysis": "Ongoing", "event_date": "2024-05-06", "org_name": "CIRCL"}
    f, off, pl = o.lookup_ip4(0xC0A80101)
    assert f and pl == 32 and json.loads(o.data_json(off)) == dict(base, type="ip-src", category="Network activity", to_ids=True, tags="tlp:green,apt,c2")
    f, off, pl = o.lookup_ip4(0x0A010203)
    assert f and json.loads(o.data_json(off)) == dict(base, type="ip-dst|port", tags="tlp:green,apt")
    assert o.lookup_ip4(0xAC100009)[0] and o.lookup_ip4(0xCB007142)[2] == 24
    hit = o.lookup_string(b"bad.example")
    assert hit and json.loads(o.data_json(hit[0][1])) == dict(base, type="domain|ip", comment="pair", tags="tlp:green,apt")
    for lit in (b"dl.evil.test", b"https://dl.evil.test:8080/a.exe", b"x.exe", b"ab" * 32, b"odd-value"):
        assert o.lookup_string(lit), lit
    hit = o.lookup_string(b"cd" * 20)
    assert hit and json.loads(o.data_json(hit[0][1])) == dict(base, type="sha1", object_type="file", object_comment="dropper", tags="tlp:green,apt")
    for lit in (b"never.added.example", b"8080", b"z" * 1000):
        assert not o.lookup_string(lit), lit
    # something that claims to be an event but is not one is an error, not a skip
    f2 = tmp_path / "broken.json"
    f2.write_text('{"Event": {"threat_level_id": 999, "Attribute": []}}')
    assert main(["build", str(f1), str(f2), "-o", str(out), "-f", "misp"]) == 1
    assert "Failed to parse MISP JSON in broken.json" in capsys.readouterr().err
    f3 = tmp_path / "badip.json"
    f3.write_text(json.dumps({"Event": {"Attribute": [{"type": "ip-src", "value": "not-an-ip"}]}}))
    assert main(["build", str(f3), "-o", str(out), "-f", "misp"]) == 1


def _patch_metadata_uint(db: bytes, key: bytes, value: int) -> bytes:
    """The database with metadata field `key` re-encoded as an 8-byte MMDB uint64 holding `value`."""
    mk = db.rfind(b"\xab\xcd\xefMaxMind.com")
    at = db.index(key, mk) + len(key)
    ctrl = db[at]
    assert ctrl >> 5 in (5, 6) or (ctrl >> 5 == 0 and db[at + 1] == 2), "expected a uint16/32/64 value"
    size = ctrl & 31
    old_len = 1 + size if ctrl >> 5 else 2 + size
    return db[:at] + bytes([8, 2]) + value.to_bytes(8, "big") + db[at + old_len:]  # extended type 9 (uint64), 8 bytes


def test_validate_rejects_wrapping_section_offsets(built, tmp_path):
    """A crafted 64-bit section offset near 2^64 must not wrap past the bounds checks (mxy_reader.cpp:331)."""
    from matchy_b200 import DatabaseBuilder, _native as N
    b = DatabaseBuilder(build_epoch=1)
    b.add_glob("*.evil.com", {"k": "v"})
    b.add_entry("bad.example.com", {"k": "w"})
    b.add_entry("10.0.0.0/8", {"k": "x"})
    db = b.build()
    L = C.CDLL(N.LIB_PATH)
    L.matchy_validate.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p)]
    good = tmp_path / "good.mxy"
    good.write_bytes(db)
    assert L.matchy_validate(str(good).encode(), 0, None) == 0
    for key, value in ((b"pattern_section_offset", 0xFFFFFFFFFFFFFFF9), (b"literal_section_offset", 0xFFFFFFFFFFFFFFE1),
                       (b"pattern_section_offset", len(db) - 4), (b"node_count", 0x3000000000000001), (b"node_count", 0xFFFFFFFF)):
        bad = tmp_path / "bad.mxy"
        bad.write_bytes(_patch_metadata_uint(db, key, value))
        msg = C.c_char_p()
        assert L.matchy_validate(str(bad).encode(), 0, C.byref(msg)) < 0, (key, hex(value))
        assert N.lib().mxyr_open(bad.read_bytes(), bad.stat().st_size) in (None, 0)


def test_validate_rejects_corrupt_tree_records(built, tmp_path):
    """matchy_validate runs the device-side preparation too (db_prepare.h): a tree record that is neither a node, the empty
    marker nor a pointer into the data section is refused, and so is the upload of such a file (validation.rs:256 walks the
    tree the same way)."""
    from matchy_b200 import DatabaseBuilder, _native as N
    b = DatabaseBuilder(build_epoch=1)
    for k in range(50):
        b.add_entry("10.%d.0.0/16" % k, {"n": k})
    db = bytearray(b.build())
    L = C.CDLL(N.LIB_PATH)
    L.matchy_validate.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p)]
    path = tmp_path / "t.mxy"
    path.write_bytes(db)
    assert L.matchy_validate(str(path).encode(), 0, None) == 0
    at = db.rfind(b"node_count") + len(b"node_count")
    assert db[at] == 0xC4  # uint32, 4 bytes
    nodes = int.from_bytes(db[at + 1:at + 5], "big")
    for bad_record in (b"\xff\xff\xf0", (nodes + 5).to_bytes(3, "big")):  # far past the data section; inside the 16-byte separator
        d = bytearray(db)
        d[3:6] = bad_record  # right record of node 0
        path.write_bytes(d)
        msg = C.c_char_p()
        assert L.matchy_validate(str(path).encode(), 0, C.byref(msg)) < 0
        assert b"tree" in (msg.value or b"").lower()


def test_worker_stats_count_crypto_types():
    from matchy_b200.processing import WorkerStats
    s = WorkerStats()
    s.add_counters({"lines": 3, "bytes": 100, "candidates": 78, "matches": 2, "by_type": list(range(1, 13))})
    assert (s.bitcoin_count, s.ethereum_count, s.monero_count) == (10, 11, 12)
    assert s.as_vector()[-3:] == [10, 11, 12] and s.domain_count == 1 and s.sha512_count == 9


def test_sequential_order_is_line_then_extractor_group_then_position():
    """match_modes.sequential_order: the order `matchy match --threads 1` prints matches in (sequential.rs:205-390 over
    extract_from_line, lib.rs:1472-1521: domains, IPv4, e-mails, IPv6, hashes, crypto)."""
    import numpy as np
    from matchy_b200 import match_modes as M
    from matchy_b200.engine import Engine
    data = b"a 1.2.3.4 evil.com\nx@evil.org 2001:db8::1 evil.net\n\n5d41402abc4b2a76b9719d911017c592 9.9.9.9\n"
    rows = [(36, 8, 0), (2, 7, 2), (10, 8, 0), (19, 10, 1), (21, 8, 0), (30, 11, 3), (52, 32, 4), (85, 7, 2)]  # (offset, len, item type), shuffled
    recs = np.zeros(len(rows), Engine.REC_DTYPE)
    for k, (o, l, t) in enumerate(rows):
        recs[k]["offset"], recs[k]["len"], recs[k]["item_type"] = o + 1000, l, t
    got = [tuple(int(x) for x in (recs[i]["offset"] - 1000, recs[i]["item_type"])) for i in M.sequential_order(recs, data, base=1000)]
    # line 0: domain evil.com, then IPv4; line 1: domains (evil.org inside the e-mail at 21, evil.net at 36), e-mail, IPv6; line 3: IPv4 before the hash
    assert got == [(10, 0), (2, 2), (21, 0), (36, 0), (19, 1), (30, 3), (85, 2), (52, 4)]


def test_follow_files_batches_and_rotation(tmp_path):
    """match_modes.follow_files: existing content first, appended bytes next, a truncated file starts over (follow.rs:209-262)."""
    import threading
    import time
    from matchy_b200 import match_modes as M
    p = tmp_path / "a.log"
    p.write_bytes(b"one\ntwo\n")
    seen, out = [], __import__("io").BytesIO()

    def scan_batch(data, source):
        seen.append((bytes(data), source))
        return b"%d\n" % len(data)

    def writer():
        time.sleep(0.4)
        with open(p, "ab") as f:
            f.write(b"three\n")
        time.sleep(0.4)
        p.write_bytes(b"new\n")  # rotation: the file is shorter than what was read
    th = threading.Thread(target=writer)
    th.start()
    total = M.follow_files([str(p)], scan_batch, out, poll=0.05, idle_exit=1.0)
    th.join()
    assert [d for d, _ in seen] == [b"one\ntwo\n", b"three\n", b"new\n"] and total == 18
    assert out.getvalue() == b"8\n6\n4\n"
    with pytest.raises(ValueError):
        M.follow_files(["-"], scan_batch, out)


def _repr_c_offsets(fields):
    """Offsets of a #[repr(C)] struct from (name, size, align) in declaration order."""
    off, out = 0, {}
    for name, size, align in fields:
        off = (off + align - 1) // align * align
        out[name] = off
        off += size
    return out, off


def test_mxy_section_headers_follow_the_reference_struct_layouts(built):
    """the byte layout of what mxy_builder.cpp writes, pinned against the reference's #[repr(C)] declarations
    field by field — ParaglobHeader (matchy-paraglob/src/offset_format.rs:73-178, 112 bytes, asserted there at :464),
    PatternEntry 16 / SingleWildcard 8 / GlobSegmentIndex 8 / GlobSegmentHeader 12 (:470-475), LiteralHashHeader + HashEntry +
    PatternMapping (matchy-literal-hash/src/lib.rs:80-136)."""
    import struct
    from matchy_b200 import DatabaseBuilder
    u32 = lambda n: (n, 4, 4)
    pg_fields = [("magic", 8, 1)] + [u32(n) for n in (
        "version", "match_mode", "ac_node_count", "ac_nodes_offset", "ac_edges_size", "ac_patterns_size", "pattern_count", "patterns_offset",
        "pattern_strings_offset", "pattern_strings_size", "meta_word_mapping_count", "meta_word_mappings_offset", "pattern_refs_size",
        "wildcard_count", "total_buffer_size")] + [("endianness", 1, 1), ("reserved", 3, 1)] + [u32(n) for n in (
        "data_section_offset", "data_section_size", "mapping_table_offset", "mapping_count", "data_flags", "reserved_v2",
        "ac_literal_map_offset", "ac_literal_map_count", "glob_segments_offset", "glob_segments_size")]
    po, psize = _repr_c_offsets(pg_fields)
    assert psize == 112
    lh_fields = [("magic", 4, 1)] + [u32(n) for n in ("version", "entry_count", "table_size", "strings_offset", "strings_size", "num_shards", "shard_bits")]
    lo, lsize = _repr_c_offsets(lh_fields)
    assert lsize == 32
    he, hesize = _repr_c_offsets([("hash", 8, 8), ("string_offset", 4, 4), ("pattern_id", 4, 4)])
    assert hesize == 16 and he["string_offset"] == 8

    b = DatabaseBuilder(build_epoch=1700000000)
    globs = ["*.evil-%d.com" % i for i in range(5)] + ["mal-*", "*[0-9].*.bad-attack.org"]
    lits = ["lit-%d.example.com" % i for i in range(40)]
    for g in globs:
        b.add_glob(g, {"k": 1})
    b.add_glob("*", {"k": 2})  # a pure wildcard (SingleWildcard entry)
    for s in lits:
        b.add_literal(s, {"k": 3})
    db = b.build()

    # ---- PARAGLOB ----
    p0 = db.index(b"PARAGLOB")
    f = lambda name: struct.unpack_from("<I", db, p0 + po[name])[0]
    assert f("version") == 5 and f("match_mode") == 0 and db[p0 + po["endianness"]] == 1
    assert struct.unpack_from("<I", db, p0 - 4)[0] == f("total_buffer_size")  # [total][paraglob_size][paraglob]... (mmdb_builder.rs:529-550)
    assert f("pattern_count") == len(globs) + 1 and f("wildcard_count") == 1
    assert f("ac_nodes_offset") % 64 == 0 and f("ac_nodes_offset") >= 112 and f("ac_node_count") > 0
    assert f("patterns_offset") % 8 == 0 and f("patterns_offset") >= f("ac_nodes_offset") + f("ac_edges_size")
    assert f("pattern_strings_offset") == f("patterns_offset") + 16 * f("pattern_count")  # PatternEntry = 16 bytes
    for i, g in enumerate(globs + ["*"]):  # PatternEntry {pattern_id u32, pattern_type u8, pad[3], string_offset u32, string_len u32}
        pid, ptype, soff, slen = struct.unpack_from("<IB3xII", db, p0 + f("patterns_offset") + 16 * i)
        assert pid == i and slen == len(g) and db[p0 + soff:p0 + soff + slen + 1] == g.encode() + b"\0"
    assert f("data_section_offset") == 0 and f("mapping_count") == 0
    ao = f("ac_literal_map_offset")
    assert db[p0 + ao:p0 + ao + 4] == b"ACLH" and f("ac_literal_map_count") > 0
    go, gsz = f("glob_segments_offset"), f("glob_segments_size")
    assert go % 8 == 0 and go + gsz == f("total_buffer_size")
    first_hdr = None
    for i in range(f("pattern_count")):  # GlobSegmentIndex {first_segment_offset u32, segment_count u16, reserved u16} = 8 bytes
        first, cnt = struct.unpack_from("<IH", db, p0 + go + 8 * i)
        if first_hdr is None:
            first_hdr = first
        assert first >= go + 8 * f("pattern_count") and (first - first_hdr) % 12 == 0  # GlobSegmentHeader = 12 bytes
        assert cnt >= 1
    assert first_hdr == go + 8 * f("pattern_count")

    # ---- LHSH ----  header 32 B, shard offset table (num_shards + 1) x u32, slot table, string pool, [count][PatternMapping x count]
    # (matchy-literal-hash/src/lib.rs:236-330)
    l0 = db.index(b"LHSH")
    g = lambda name: struct.unpack_from("<I", db, l0 + lo[name])[0]
    assert g("version") == 1 and g("entry_count") == len(lits)
    assert g("num_shards") == 1 << g("shard_bits") and g("table_size") >= g("entry_count")
    table0 = 32 + 4 * (g("num_shards") + 1)
    assert g("strings_offset") == table0 + 16 * g("table_size")
    shard_off = struct.unpack_from("<%dI" % (g("num_shards") + 1), db, l0 + 32)
    assert shard_off[0] == 0 and shard_off[-1] == g("table_size") and list(shard_off) == sorted(shard_off)
    seen = set()
    for s_ in range(g("table_size")):  # HashEntry {hash u64, string_offset u32, pattern_id u32}; empty = string_offset 0xFFFFFFFF
        h, so, pid = struct.unpack_from("<QII", db, l0 + table0 + 16 * s_)
        if so != 0xFFFFFFFF:
            assert pid < len(lits) and so < g("strings_size")
            seen.add(pid)
    assert seen == set(range(len(lits)))
    m0 = l0 + g("strings_offset") + g("strings_size")
    assert struct.unpack_from("<I", db, m0)[0] == len(lits)  # PatternMapping {pattern_id u32, data_offset u32} = 8 bytes each
    assert sorted(struct.unpack_from("<II", db, m0 + 4 + 8 * k)[0] for k in range(len(lits))) == list(range(len(lits)))
