"""CPU suite: the device logic (tokenizer bit arithmetic, validation, trie / hash / AC+glob walkers), compiled for the
host by tests/host_emulation/emu.cpp, against the oracle.  This is how kernel logic is debugged without a GPU; the
real kernels are checked by test_gpu_parity.py.  Nothing here is a product path."""
import random

import pytest

import emu_lib as E
import oracle_lib as O
from test_gpu_parity import CRYPTO_FRAGS, FRAGS, _fuzz_text, _monero_like_word


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_configs(small_dbs, cfg):
    db, log = small_dbs[cfg]
    log = log[:400000]
    orc, emu = O.Oracle(db), E.Emu(db)
    want = orc.scan(log, chunk_size=128 * 1024)
    for chunk, nwarps, mis in ((0, 1, 0), (70000, 3, 5), (0, 7, 15), (150000, 64, 9)):
        assert emu.scan(log, chunk_bytes=chunk, nwarps=nwarps, misalign=mis) == want, (chunk, nwarps, mis)
    assert emu.is_fast()  # every BASELINE config qualifies for the constant-time string filters
    lit, glob, tested = emu.filter_stats()
    if tested:
        assert lit + glob <= 0.05 * tested + len(want[0]) + 8, (lit, glob, tested)  # the filters must stay selective
    emu.set_generic(True)  # the generic lithash + acglob path must give the same answer
    assert emu.scan(log, chunk_bytes=150000, nwarps=5) == want
    if cfg in (2, 5):
        assert emu.anchored_exact()
        emu.set_anchored(False)  # the Aho-Corasick formulation must give the same answer as the anchored walks
        assert emu.scan(log) == want


def test_extractor_fuzz(small_dbs):
    db, _ = small_dbs[1]
    orc, emu = O.Oracle(db), E.Emu(db)
    rng = random.Random(7)
    for it in range(600):
        data = _fuzz_text(rng, rng.randint(0, 60) if it % 10 else rng.randint(200, 900))
        flags = rng.choice([31, 31, 31, 1, 2, 4, 8, 16, 5, 10, 21])
        want = sorted((s, t, e) for t, s, e in orc.extract(data, flags))
        got = sorted((s, t, e) for t, s, e in emu.tokens(data, flags, nwarps=rng.choice([1, 1, 2, 3, 5]), misalign=rng.choice([0, 0, 3, 15])))
        assert got == want, (it, flags, data[:160])


def test_scan_fuzz_with_hits(small_dbs):
    db, log = small_dbs[5]
    orc, emu = O.Oracle(db), E.Emu(db)
    rng = random.Random(99)
    words = [w for w in log.replace(b"=", b" ").replace(b"\"", b" ").split() if b"." in w or len(w) in (32, 40, 64)]
    for it in range(40):
        parts = []
        for _ in range(rng.randint(5, 300)):
            parts.append(rng.choice(words) if rng.random() < 0.5 else rng.choice(FRAGS))
            parts.append(rng.choice([b" ", b"\n", b" ", b"=", b","]))
        data = b"".join(parts)
        assert emu.scan(data, nwarps=rng.choice([1, 2, 4])) == orc.scan(data), it


def test_long_tokens_cross_tiles(small_dbs):
    """Words longer than a 32-byte slice, a 1 KiB tile and a whole warp range keep exact extents."""
    db, _ = small_dbs[5]
    orc, emu = O.Oracle(db), E.Emu(db)
    for n in (31, 32, 33, 63, 64, 65, 1000, 1023, 1024, 1025, 3000, 5000):
        for data in (b"a" * n + b".com x", b"x " + b"b" * n + b".evil.org\n", b"0" * n + b" ", b"q=" + b"1" * 7 + b"." + b"c" * n + b".net",
                     b"u" * n + b"@mail.example.com ", b"\xc3\xa9" * (n // 2) + b".fr "):
            want = sorted((s, t, e) for t, s, e in orc.extract(data, 31))
            for nwarps in (1, 2, 5):
                got = sorted((s, t, e) for t, s, e in emu.tokens(data, 31, nwarps=nwarps))
                assert got == want, (n, data[:40], nwarps)


def test_glob_backtracking_budget(built):
    """The 100 000-step budget of match_segments_impl is spent exactly like the reference spends it."""
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1)
    b.add_glob("*aaa*aaa*aaa*aaa*aaa*bbb", {"x": 1})
    b.add_glob("*.example.*", {"x": 2})
    db = b.build()
    orc, emu = O.Oracle(db), E.Emu(db)
    for text in (b"a" * 60 + b".example.com", b"a" * 200 + b".example.com", b"a" * 60 + b"bbb.example.com", b"aaa" * 5 + b"bbb"):
        data = b"host=" + text + b" \n"
        assert emu.scan(data) == orc.scan(data), text


def test_case_insensitive_database(built):
    from matchy_b200 import DatabaseBuilder, MatchMode
    b = DatabaseBuilder(MatchMode.CaseInsensitive, build_epoch=1)
    b.add_entry("Evil.Example.COM", {"x": 1})
    b.add_entry("*.BadSite.org", {"x": 2})
    b.add_entry("5D41402ABC4B2A76B9719D911017C592", {"x": 3})
    db = b.build()
    orc, emu = O.Oracle(db), E.Emu(db)
    data = b"a Evil.example.com b evil.EXAMPLE.com c www.badsite.ORG d 5d41402abc4b2a76b9719d911017c592 e WWW.BADSITE.org\n"
    want = orc.scan(data)
    assert len(want[0]) >= 4
    assert emu.scan(data) == want


def test_literal_search_formulations_agree(built):
    """Anchored walks (device default) vs Aho-Corasick walk on overlapping / nested / repeated literals; a database with a
    2-byte literal must switch the anchored mode off."""
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1)
    for g in ("*abcab*", "*bcabc*", "*cab*", "*.evil.com", "*evil.com*", "*l.c*", "*aaa*", "*aaaa*", "x*aaa*y", "*abc*abc*", "abcabc"):
        b.add_glob(g, {"g": g})
    db = b.build()
    orc, emu = O.Oracle(db), E.Emu(db)
    assert emu.anchored_exact()
    texts = [b"abcabcab", b"xabcabcy", b"aaaaaaa", b"xaaay", b"www.evil.com", b"evil.com.evil.com", b"cabcabcabc.abcab", b"l.c", b"ab", b"abc"]
    data = b" ".join(texts) + b"\n" + b"".join(b"h=" + t + b".example.com\n" for t in texts)
    want = orc.scan(data)
    assert len(want[0]) >= 10
    assert emu.scan(data) == want
    emu.set_anchored(False)
    assert emu.scan(data) == want
    b.add_glob("zz", {"g": "short"})  # Literal-type pattern of 2 bytes goes into the automaton
    db2 = b.build()
    orc2, emu2 = O.Oracle(db2), E.Emu(db2)
    assert not emu2.anchored_exact()
    data2 = data + b"q=buzz.example.com zz.example.org\n"
    assert emu2.scan(data2) == orc2.scan(data2)


def test_fast_string_path_classification(built):
    """Suffix- / prefix-anchored globs of every key length, tokens shorter than the keys, literals of every length:
    the constant-time filters (string_filters) may never lose a match.  Unanchored and literal-type globs are covered by the
    per-position literal prefilter; a case-insensitive database must fall back to the generic path."""
    from matchy_b200 import DatabaseBuilder, MatchMode
    rng = random.Random(5)
    sfx = ["*m", "*om", "*.io", "*l.io", "*il.io", "*vil.io", "*evil.io", "*.evil.io", "*x.evil.io", "*.very-long-suffix.example.net", "*[0-9].bad.org",
           "a?c*zz.org", "*.??.uk"]
    pfx = ["a*", "ab*", "abc.*", "abcd*.z?", "abcde*[a-z]", "abcdef.*", "abcdefg*", "abcdefgh*", "abcdefghi.*", "prefix-longer-than-eight-*", "p.q*r?"]
    lits = ["a.io", "ab.io", "abc.com", "abcd.com", "abcde.io", "evil.io", "x.evil.io", "abcdefgh.com", "a-much-longer-literal.example.com",
            "5d41402abc4b2a76b9719d911017c592", "user@evil.io"]
    b = DatabaseBuilder(build_epoch=1)
    for g in sfx + pfx:
        b.add_glob(g, {"g": g})
    for l in lits:
        b.add_entry(l, {"l": l})
    db = b.build()
    orc, emu = O.Oracle(db), E.Emu(db)
    assert emu.is_fast()
    toks = [b"a.io", b"ab.io", b"abc.com", b"abcd.zz", b"abcde.az", b"abcdef.com", b"abcdefg.com", b"abcdefgh.com", b"abcdefghi.com", b"evil.io", b"x.evil.io",
            b"xx.evil.io", b"l.io", b"il.io", b"vil.io", b"m.com", b"a.om", b"prefix-longer-than-eight-x.com", b"prefix-longer-than-eight.com",
            b"q.very-long-suffix.example.net", b"very-long-suffix.example.net", b"a7.bad.org", b"ax.bad.org", b"abc.qzz.org", b"a.cc.uk", b"p.qr.rs",
            b"a-much-longer-literal.example.com", b"x-much-longer-literal.example.com", b"user@evil.io", b"u@x.evil.io", b"5d41402abc4b2a76b9719d911017c592",
            b"5d41402abc4b2a76b9719d911017c593", b"abc.io", b"b.io", b"a.b.c.d.e.io", b"zz.org", b"ac.zz.org", b"abc.zz.org"]
    data = b"".join(b"k=" + t + b" \n" for t in toks)
    for _ in range(300):
        data += b"h=" + rng.choice(toks)[:rng.randint(1, 40)] + rng.choice([b"", b".io", b".com", b"m", b".evil.io", b".uk"]) + b" "
    want = orc.scan(data)
    assert len(want[0]) >= 30
    assert emu.scan(data) == want
    emu.set_generic(True)
    assert emu.scan(data) == want
    # fall-backs
    for extra, mode in (("*mid*", None), ("glob:plain.example.com", None), (None, MatchMode.CaseInsensitive)):
        b2 = DatabaseBuilder(mode, build_epoch=1) if mode is not None else DatabaseBuilder(build_epoch=1)
        for g in sfx[:6] + pfx[:4]:
            b2.add_glob(g, {"g": g})
        for l in lits:
            b2.add_entry(l, {"l": l})
        if extra and extra.startswith("glob:"):
            b2.add_entry(extra, {"g": extra})
        elif extra:
            b2.add_glob(extra, {"g": extra})
        db2 = b2.build()
        orc2, emu2 = O.Oracle(db2), E.Emu(db2)
        # unanchored / literal-typed globs stay on the fast path (generic_literal_scan); case-insensitive databases do not
        assert emu2.is_fast() == (mode is None), (extra, mode)
        d2 = data + b"q=a.mid.com z=xplain.example.com.evil.io Q=ABC.COM mid.org amid.st.org xmi.d.com plain.example.co\n"
        want2 = orc2.scan(d2)
        assert emu2.scan(d2) == want2, (extra, mode)
        if mode is None:
            lit, glob, tested = emu2.filter_stats()
            assert glob < tested  # the scan must not flag everything
            emu2.set_generic(True)
            assert emu2.scan(d2) == want2


def test_tld_fast_front_end(small_dbs):
    """Last-label decisions: accepted TLDs, multi-label-only suffixes, labels of 7 / 8 / 9 bytes, non-ASCII TLDs, numeric last labels."""
    db, _ = small_dbs[1]
    orc, emu = O.Oracle(db), E.Emu(db)
    words = [b"a.com", b"a.co.uk", b"a.b.kawasaki.jp", b"city.kawasaki.jp", b"x.ck", b"www.ck", b"a.b.ck", b"a.website", b"a.websitex", b"a.shopping", b"a.shoppin",
             b"a.photography", b"a.international", b"a.xn--p1ai", "a.\u4e2d\u56fd".encode(), "a.\u0440\u0444".encode(), b"a.\xff\xfe", b"a.b.1", b"1.2.3.4", b"a.html",
             b"x.y.compute.amazonaws.com", b"a.s3.amazonaws.com", b"a.blogspot.com", b"abcdefg.h", b"a.bcdefgh", b"a.museum", b"a.co", b"co.uk", b"uk",
             b"a.nom.br", b"a.b.nom.br", b"a.appspot.com", b"a.Com", b"a.COM", b"a.c-m", b"a.travelersinsurance", b"a.x.travelersinsurance"]
    data = b" ".join(words) + b"\n" + b"".join(b"<" + w + b"> " for w in words)
    want = sorted((s, t, e) for t, s, e in orc.extract(data, 31))
    assert len(want) >= 20
    assert sorted((s, t, e) for t, s, e in emu.tokens(data, 31)) == want


def test_crypto_address_extraction(small_dbs):
    """Host emulation of the crypto-address path (long-word queue of the tokenizer + crypto_addr.cuh) against the oracle."""
    db, log = small_dbs[1]
    orc, emu = O.Oracle(db), E.Emu(db)
    rng = random.Random(17)
    monero_ok = _monero_like_word()
    seen = set()
    for it in range(150):
        parts = []
        for _ in range(rng.randint(1, 50)):
            f = rng.choice(CRYPTO_FRAGS + [monero_ok]) if rng.random() < 0.6 else rng.choice(FRAGS)
            parts.append(f)
            parts.append(rng.choice([b" ", b" ", b"\n", b"/", b"=", b":", b"", b".", b",", b"\"", b"-"]))
        data = b"".join(parts)
        want = sorted((s, t, e) for t, s, e in orc.extract(data, 0xFF))
        got = sorted((s, t, e) for t, s, e in emu.tokens(data, 0xFF, nwarps=rng.choice([1, 2, 3]), misalign=rng.choice([0, 7])))
        assert got == want, (it, data[:200])
        seen |= {t for _, t, _ in want}
    assert {9, 10, 11} <= seen
    data = log[:300000] + b"pay 1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa or 0x5aeda56215b167893e80b4fe645ba6d5bab767de now\n"
    flags = orc.default_flags() | 0xE0
    assert emu.scan(data, flags=flags, nwarps=3) == orc.scan(data, flags=flags)


def test_ipv6_mask_parser_equals_reference_parser():
    """parse_ipv6_run_masks (device_fns.cuh, straight-line mask arithmetic) == the oracle's restatement of Rust's
    Ipv6Addr::from_str on every hex/colon run that holds a "::" (the only runs the extractor ever parses, lib.rs:1044-1116)."""
    import ctypes as C
    import random
    L = E.lib()
    L.emu_parse_ipv6_masks.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32)]
    rng = random.Random(20261018)
    fixed = [b"2001:db8::1", b"1::2", b"1:2:3:4:5:6::7", b"1:2:3:4:5:6:7::8", b"1::2::3", b"1:::2", b":1::2", b"1::2:", b"12345::1", b"1::12345",
             b"a:b:c:d:e:f::1", b"FFFF::ffff", b"0::0", b"1:2::3:4:5:6:7", b"1:2::3:4:5:6:7:8", b"::1:2:3", b"1:2:3::", b"abcd:ef01::2345:6789",
             b"1::", b"::", b"1:2:3:4:5:6:7:8::9", b"0000::0000:0000", b"1::2:3:4:5:6:7"]
    cases = list(fixed)
    alphabet = b"0123456789abcdefABCDEF"
    for _ in range(20000):
        # fields of 0..5 digits joined by ':' (empty fields make "::" / ":::"), plus some fully random strings
        if rng.random() < 0.85:
            k = rng.randint(1, 10)
            fields = [bytes(rng.choice(alphabet) for _ in range(rng.choice((0, 1, 1, 2, 3, 4, 4, 4, 5)))) for _ in range(k)]
            s = b":".join(fields)
        else:
            s = bytes(rng.choice(alphabet + b"::::") for _ in range(rng.randint(2, 44)))
        cases.append(s)
    checked = 0
    for s in cases:
        if b"::" not in s or len(s) < 2:
            continue
        want = O.parse_ipv6(s) if len(s) <= 39 else None
        w = (C.c_uint32 * 4)()
        got = L.emu_parse_ipv6_masks(s + b"\0" * 16, len(s), w)
        if want is None:
            assert got == 0, s
        else:
            assert got == 1, s
            assert [(w[k // 2] >> (0 if k & 1 else 16)) & 0xFFFF for k in range(8)] == want, s
        checked += 1
    assert checked > 5000


def test_byte_category_planes_reproduce_the_class_bits():
    """scan_kernel classifies a byte through a 4-bit category (four bit planes, one table word) instead of the tokenizer's
    eight class bits: the planes must give back exactly class_bits() for every byte value."""
    import emu_lib
    L = emu_lib.lib()
    L.emu_class_bits_via_planes.restype = L.emu_class_bits.restype = __import__("ctypes").c_uint32
    for b in range(256):
        assert L.emu_class_bits_via_planes(b) == L.emu_class_bits(b), b


def test_hexdot_ipv4_parser_equals_general_parser():
    """scan_kernel parses numeric-queue words (hex digits and dots only) with parse_ipv4_hexdot: same verdict and address as
    parse_ipv4_words on valid addresses, every near miss, and random words over the alphabet."""
    import ctypes as C
    import random
    import emu_lib
    L = emu_lib.lib()
    L.emu_parse_ipv4_both.restype = C.c_uint32
    L.emu_parse_ipv4_both.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32)]
    rng = random.Random(7)
    words = [b"1.2.3.4", b"255.255.255.255", b"256.1.1.1", b"01.2.3.4", b"1.02.3.4", b"0.0.0.0", b"1.2.3", b"1.2.3.4.5", b"1..2.3", b".1.2.3", b"1.2.3.",
             b"a.b.c.d", b"1.2.3.4a", b"cafe.de", b"1.2.3.f", b"999.999.999.999", b"1.2.3.1000", b"10.20.30.40", b"192.168.001.1", b"192.168.1.001",
             b"25.5.2.55", b"1.1.1.1", b"12.34.56.78", b"249.250.251.252", b"199.200.201.202", b"100.099.1.1", b"0.00.0.0", b"1234.1.1.1"]
    for _ in range(60000):
        kind = rng.random()
        if kind < 0.5:
            w = ".".join(str(rng.choice([rng.randrange(0, 300), rng.randrange(0, 30), rng.randrange(250, 260)])) for _ in range(4)).encode()
        elif kind < 0.75:
            w = ".".join(("%0*d" % (rng.randint(1, 4), rng.randrange(0, 400))) for _ in range(rng.randint(3, 5))).encode()
        else:
            w = bytes(rng.choice(b"0123456789.abcdefABCDEF..00") for _ in range(rng.randint(1, 18)))
        words.append(w)
    out = (C.c_uint32 * 2)()
    for w in words:
        for tail in (b"\0" * 24, b"9.9.9.9 zzzzzzzzzzzzzzzzzzzz", b"\xff" * 24):  # whatever follows the word must not matter
            r = L.emu_parse_ipv4_both(w + tail, len(w), out)
            assert r in (0, 3), (w, r)
            assert r == 0 or out[0] == out[1], w


def test_result_sort_equals_a_plain_sort():
    """host_sort.h (offset-range partition + per-bucket sorts on a thread pool) against numpy's lexsort: spread-out offsets,
    clustered offsets, ties on offset decided by (item_type, len), sizes around the thresholds."""
    import numpy as np
    import emu_lib
    L = emu_lib.lib()
    dt = np.dtype([("offset", "<u8"), ("len", "<u4"), ("item_type", "u1"), ("kind", "u1"), ("prefix_len", "u1"), ("reserved", "u1"),
                   ("n_ids", "<u4"), ("ids_index", "<u4"), ("data_offset", "<u4"), ("pad", "<u4")])
    rng = np.random.default_rng(7)
    for n, threads, shape in ((0, 4, "spread"), (1, 4, "spread"), (4095, 4, "spread"), (4096, 4, "spread"), (60000, 8, "spread"), (300000, 16, "spread"),
                              (100000, 8, "cluster"), (50000, 3, "ties"), (50000, 0, "spread"), (70000, 8, "one")):
        r = np.zeros(n, dtype=dt)
        if shape == "spread":
            r["offset"] = rng.integers(5_000_000_000, 15_000_000_000, n)
        elif shape == "cluster":
            r["offset"] = np.where(rng.random(n) < 0.9, rng.integers(10**9, 10**9 + 4000, n), rng.integers(0, 2**40, n))
        elif shape == "ties":
            r["offset"] = rng.integers(0, 3000, n)
        else:
            r["offset"] = 12345
        r["item_type"] = rng.integers(0, 12, n)
        r["len"] = rng.integers(1, 300, n)
        r["ids_index"] = np.arange(n)  # payload: every record must survive
        want = r[np.lexsort((r["len"], r["item_type"], r["offset"]))]
        key = lambda a: (a["offset"].tolist(), a["item_type"].tolist(), a["len"].tolist())
        lo, hi = (int(r["offset"].min()), int(r["offset"].max())) if n else (0, 0)
        for rng_lo, rng_hi in ((lo, hi), (0, 2**41), (lo + 1, hi)):  # exact range, generous range, WRONG range (falls back to a plain sort)
            got = r.copy()
            L.emu_sort_records(got.ctypes.data, n, rng_lo, rng_hi, threads, 1)
            assert key(got) == key(want), (n, threads, shape, rng_lo)
            assert sorted(got["ids_index"].tolist()) == list(range(n)), (n, threads, shape)
    # the order check that lets the host skip its sort when the device has sorted: true on sorted input, false for ONE swapped
    # neighbour pair anywhere (slice edges of the pool included), ties in offset decided by item_type, then len
    for n, threads in ((0, 4), (1, 4), (2, 0), (70000, 4), (200000, 16), (200000, 0)):
        r = np.zeros(n, dtype=dt)
        r["offset"] = np.sort(rng.integers(0, 2**38, n)) if n else 0
        r["item_type"] = 3
        r["len"] = 7
        assert L.emu_records_sorted(r.ctypes.data, n, threads) == 1
        if n >= 2:
            for at in {1, n // 2, n - 1} | ({n * k // threads for k in range(1, threads)} if threads else set()):
                b = r.copy()
                b[[at - 1, at]] = b[[at, at - 1]]
                if b["offset"][at - 1] != b["offset"][at]:
                    assert L.emu_records_sorted(b.ctypes.data, n, threads) == 0, (n, threads, at)
            t = r.copy()
            t["offset"][:] = 5
            t["item_type"][n // 2:] = 4
            assert L.emu_records_sorted(t.ctypes.data, n, threads) == 1
            if n >= 4:  # (the last two records share offset and item_type: len decides)
                t["len"][n - 1] = 6
                assert L.emu_records_sorted(t.ctypes.data, n, threads) == 0
                t["len"][n - 1] = 7
            t["item_type"][0] = 5
            assert L.emu_records_sorted(t.ctypes.data, n, threads) == 0
    # id pairs gathered in record order
    idt = np.dtype([("pattern_id", "<u4"), ("data_offset", "<u4")])
    for n, threads, p_pattern in ((10, 4, 0.5), (5000, 4, 0.5), (40000, 8, 0.5), (40000, 0, 0.5), (120000, 8, 0.02), (50000, 16, 0.0), (50000, 5, 1.0)):
        r = np.zeros(n, dtype=dt)
        r["kind"] = np.where(rng.random(n) < p_pattern, 2, 1)  # 1 = IP record (no ids), 2 = pattern record
        r["n_ids"] = np.where(r["kind"] == 2, rng.integers(1, 5, n), 0)
        total = int(r["n_ids"].sum())
        ids = np.zeros(total + 7, dtype=idt)
        ids["pattern_id"] = rng.permutation(total + 7)
        starts = np.concatenate(([0], np.cumsum(r["n_ids"])[:-1]))
        order = rng.permutation(n)  # the device's arrival order: record k's pairs sit at a random place
        where = np.zeros(n, dtype=np.int64)
        where[order] = np.concatenate(([0], np.cumsum(r["n_ids"][order])[:-1]))
        r["ids_index"] = where
        want_ids = np.concatenate([ids["pattern_id"][where[k]:where[k] + r["n_ids"][k]] for k in range(n)]) if total else np.zeros(0)
        out = np.zeros(total, dtype=idt)
        got_r = r.copy()
        assert L.emu_repack_ids(got_r.ctypes.data, n, ids.ctypes.data, out.ctypes.data, threads) == total
        assert out["pattern_id"].tolist() == want_ids.tolist(), (n, threads)
        pat = got_r["kind"] == 2
        assert got_r["ids_index"][pat].tolist() == starts[pat].tolist(), (n, threads)
