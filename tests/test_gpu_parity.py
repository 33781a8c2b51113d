"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.  Bit-exact."""
import random

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines(small_dbs):
    from matchy_b200 import Engine
    out = {}
    for cfg, (db, log) in small_dbs.items():
        e = Engine(0, chunk_bytes=8 << 20)
        e.upload(db)
        out[cfg] = (e, O.Oracle(db), log)
    return out


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_config_parity(engines, cfg):
    eng, orc, log = engines[cfg]
    eng.scan(log)
    want, wcnt = orc.scan(log, chunk_size=128 * 1024)
    assert eng.counters_list() == wcnt
    assert eng.records_as_tuples() == want
    assert len(want) > 0


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_device_result_sort(small_dbs, cfg):
    """The per-piece result sort on the device (cub radix sort of (offset, item_type) keys + gather, engine.cu: drain_records):
    forced for every piece ("device_sort_min" = 1; by default only pieces with >= 128 K records take it), over several pieces per
    scan, resident and from host memory — the records equal the oracle's AND arrive in order (the host's own sort does not run)."""
    import oracle_lib as O
    from matchy_b200 import Engine
    db, log = small_dbs[cfg]
    orc = O.Oracle(db)
    want, wcnt = orc.scan(log, chunk_size=128 * 1024)
    eng = Engine(0, chunk_bytes=192 << 10)  # ~6 pieces per MiB
    eng.upload(db)
    eng.set_option("device_sort_min", 1)
    eng.scan(log)  # host buffer: piece by piece, a piece that runs out of room is split and redone in place -> in order
    assert eng.counters_list() == wcnt and eng.records_as_tuples() == want
    eng.debug_counters()
    assert eng.arrived_sorted
    dev = eng.dev_alloc(len(log))
    eng.dev_upload(dev, log)
    # resident, pieces this small: the per-warp token reservations overflow the (chunk-sized) lists, the batch redoes those
    # pieces AFTER the others, so the records may arrive out of order — the host's sort is the safety net, the result the same
    eng.scan_device(dev, len(log), eng.default_flags())
    assert eng.counters_list() == wcnt and eng.records_as_tuples() == want
    eng.set_option("device_sort_min", 1 << 30)  # never: the records arrive in append order and the host sorts
    eng.scan_device(dev, len(log), eng.default_flags())
    assert eng.records_as_tuples() == want
    eng.dev_free(dev)
    eng.close()
    big = Engine(0, chunk_bytes=16 << 20)  # resident, one roomy piece: sorted on the device, nothing redone
    big.upload(db)
    big.set_option("device_sort_min", 1)
    dev = big.dev_alloc(len(log))
    big.dev_upload(dev, log)
    big.scan_device(dev, len(log), big.default_flags())
    assert big.counters_list() == wcnt and big.records_as_tuples() == want
    big.debug_counters()
    assert big.arrived_sorted
    big.dev_free(dev)
    big.close()


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_ndjson_parity(engines, small_dbs, cfg):
    """sorted NDJSON of the GPU path == sorted NDJSON of the oracle (the `matchy match` output contract)."""
    from matchy_b200 import RecordFormatter
    eng, orc, log = engines[cfg]
    recs, ids = eng.scan(log)
    fmt = RecordFormatter(small_dbs[cfg][0])
    got = sorted(fmt.ndjson(recs, ids, log, 0, "test.log").splitlines())
    orc.scan(log, chunk_size=128 * 1024)
    want = sorted(orc.ndjson(log, "test.log").splitlines())
    assert got == want


def test_chunking_is_invisible(engines, small_dbs):
    """Results do not depend on how the engine cuts the buffer (newline-aligned pieces, FileReader::next_batch)."""
    from matchy_b200 import Engine
    db, log = small_dbs[5]
    ref = None
    for chunk in (64 << 10, 200 << 10, 8 << 20):
        e = Engine(0, chunk_bytes=chunk)
        e.upload(db)
        e.scan(log, base=1 << 33)
        cur = (e.records_as_tuples(), e.counters_list())
        if ref is None:
            ref = cur
        assert cur == ref
        e.close()
    orc = O.Oracle(db)
    want, wcnt = orc.scan(log, base=1 << 33, chunk_size=128 * 1024)
    assert ref == (want, wcnt)


def test_device_resident_scan_matches_host_scan(engines):
    eng, orc, log = engines[2]
    eng.scan(log)
    host = (eng.records_as_tuples(), eng.counters_list())
    p = eng.dev_alloc(len(log))
    try:
        eng.dev_upload(p, log)
        eng.scan_device(p, len(log))
        assert (eng.records_as_tuples(), eng.counters_list()) == host
    finally:
        eng.dev_free(p)


FRAGS = [b"1.2.3.4", b"10.0.0.1", b"256.1.1.1", b"1.2.3", b"1.2.3.4.5", b"01.2.3.4", b"192.168.001.1", b"evil.com", b"www.example.co.uk",
         b"a.b", b".com", b"com.", b"-a.com", b"a-.com", b"a..com", b"EVIL.COM", b"Evil.com", b"x_y.com", b"user@example.com", b"user@.com",
         b"a@b", b"@", b"@@", b"u.s.e.r@sub.evil.org", b"a..b@c.com", b"123@456.com", b"u+tag@ex-ample.net", b"::", b":::", b"::1",
         b"2001:db8::1", b"fe80::1ff:fe23:4567:890a", b"2001:db8:0:0:0:0:2:1", b"2001:db8::85a3::7334", b"1:2:3:4:5:6:7::", b"12345::1:2:3",
         b"::ffff:1.2.3.4", b"a::b:c:d:e:f:1", b"FE90::1:2:3:4", b"5d41402abc4b2a76b9719d911017c592", b"5D41402ABC4B2A76B9719D911017C59",
         b"da39a3ee5e6b4b0d3255bfef95601890afd80709", b"e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
         b"0123456789abcdef" * 6, b"0123456789abcdef" * 8, b"0123456789abcdef" * 9, b"g" * 32, b"\xc3\xa9xample.com", b"\xff\xfe.com",
         b"caf\xc3\xa9.fr", b"\xe2\x82\xac.org", b"\xc3.com", b" ", b"\t", b"\n", b"\r\n", b"/", b",", b";", b":", b"(", b")", b"[", b"]",
         b"{", b"}", b"<", b">", b"\"", b"'", b"=", b"&", b"?", b"-", b"_", b".", b"%", b"#", b"|", b"\\", b"*", b"!", b"~", b"http://",
         b"key=", b"a" * 70 + b".com", b"x" * 1500 + b".org", b"b." * 600 + b"com", b"9" * 40, b"1." * 20, b"deadbeef"]
# crypto-address material (matchy-extractor/src/lib.rs:3240-3626 and near misses of every rule)
CRYPTO_FRAGS = [b"1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa", b"3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64", b"bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq",
                b"bc1p0xlxvlhemja6c4dqv22uapctqupfhlxm9h8z3k2e72q4k9hcz7vqzk5jj0", b"1A1zP1eP5QGefi2DMPTfTL5SLmv7Divf00", b"1A1zP1eP",
                b"1111111111111111111114oLvT2", b"bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdQ", b"bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdx",
                b"tb1qw508d6qejxtdg4y5r3zarvary0c5xw7kxpjzsx", b"BC1QAR0SRRR7XFKVY5L643LYDNW9RE59GTZZWF5MDQ", b"bc1" + b"q" * 30,
                b"0x5aeda56215b167893e80b4fe645ba6d5bab767de", b"0x5aAeb6053F3E94C9b9A09f33669435E7Ef1BeAed", b"0x5aAeb6053f3e94c9b9a09f33669435e7ef1beaed",
                b"0x5AEDA56215B167893E80B4FE645BA6D5BAB767DE", b"0x5aeda56215b167893e80b4fe645ba6d5bab7", b"0x5aeda56215b167893e80b4fe645ba6d5bab767dg",
                b"0xfB6916095ca1df60bB79Ce92cE3Ea74c37c5d359", b"0xfB6916095ca1df60bB79Ce92cE3Ea74c37c5d35B", b"0X5aeda56215b167893e80b4fe645ba6d5bab767de",
                b"44AFFq5kSiGBoZ4NMDwYtN18obc8AemS33DBLWs3H7otXft3XjrpDtQGv7SqSsaBYBb98uNbr2VBBEt7f2wfn3RVGQBEP3A",
                b"4" + b"A" * 94, b"8" + b"z" * 100, b"3" * 30, b"1" * 40, b"3" + b"1" * 25, b"1" + b"2" * 61, b"1" + b"2" * 62, b"13" * 13,
                b"0123456789abcdef0123456789abcdef", b"3f" * 16, b"1a" * 20, b"x" * 26, b"y" * 25, b"4" * 89, b"4" * 90, b"8" * 110, b"8" * 111]


def _fuzz_text(rng, n):
    parts = []
    for _ in range(n):
        f = rng.choice(FRAGS)
        if rng.random() < 0.1:
            f = bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 6)))
        parts.append(f)
        if rng.random() < 0.5:
            parts.append(rng.choice([b" ", b" ", b"\n", b"/", b"=", b":", b"@", b"", b"."]))
    return b"".join(parts)


def test_extractor_fuzz(engines):
    """Adversarial token soup: the device extractor returns exactly the oracle's (type, start, end) set."""
    eng, orc, _ = engines[1]
    rng = random.Random(1234)
    for it in range(400):
        data = _fuzz_text(rng, rng.randint(0, 60) if it % 8 else rng.randint(300, 3000))
        flags = rng.choice([31, 31, 31, 1, 2, 4, 8, 16, 5, 10, 21])
        want = sorted((s, t, e) for t, s, e in orc.extract(data, flags))
        got = sorted((s, t, e) for t, s, e in eng.extract(data, flags))
        assert got == want, (it, flags, data[:200])


def test_scan_fuzz_with_hits(engines, small_dbs):
    """Token soup salted with real indicators of the mixed database: same records, same counters."""
    from matchy_b200 import synth  # noqa: F401
    eng, orc, log = engines[5]
    rng = random.Random(99)
    words = [w for w in log.replace(b"=", b" ").replace(b"\"", b" ").split() if b"." in w or len(w) in (32, 40, 64)]
    for it in range(60):
        parts = []
        for _ in range(rng.randint(5, 400)):
            parts.append(rng.choice(words) if rng.random() < 0.5 else rng.choice(FRAGS))
            parts.append(rng.choice([b" ", b"\n", b" ", b"=", b","]))
        data = b"".join(parts)
        eng.scan(data)
        want, wcnt = orc.scan(data)
        assert eng.counters_list() == wcnt, it
        assert eng.records_as_tuples() == want, it


def test_dense_hits_drain_the_deferred_filter_queue(engines, small_dbs):
    """48 MiB in which a third of the tokens are indicators of config 2's database or near misses that share their key bytes:
    every warp's stage-2 queue (token kernel, defer_push) fills and drains many times, the record staging of the exact and
    IP kernels sees dense output.  Records and counters must equal the oracle's."""
    import numpy as np
    eng, orc, log = engines[2]
    want0, _ = orc.scan(log, chunk_size=128 * 1024)
    hits = sorted({log[r[0]:r[0] + r[1]] for r in want0})
    assert len(hits) > 20
    rng = random.Random(4242)
    plain = [w for w in log.replace(b"=", b" ").replace(b"\"", b" ").split() if b"." in w][:4000]
    pool = []
    for h in hits:
        pool.append(h)
        pool.append(bytes([h[0] ^ 1]) + h[1:])                  # same tail, different head
        pool.append(h[:-1] + bytes([h[-1] ^ 1]))                # same head, different tail
        pool.append(b"x" + h)                                   # suffix-anchored globs still match, literals do not
        pool.append(h[: len(h) // 2] + b"0" + h[len(h) // 2:])  # same 8-byte head and tail, different middle
    lines = []
    for _ in range(3000):
        toks = [rng.choice(pool) if rng.random() < 0.35 else rng.choice(plain) for _ in range(rng.randint(3, 9))]
        lines.append(b" ".join(toks) + b"\n")
    block = b"".join(lines)
    data = (block * (48 * 1024 * 1024 // len(block) + 1))[: 48 * 1024 * 1024]
    data = data[: data.rfind(b"\n") + 1]
    big = Engine_for(eng)
    big.scan(np.frombuffer(data, dtype=np.uint8))
    want, wcnt = orc.scan(data, chunk_size=4 << 20)
    assert big.counters_list() == wcnt
    assert big.records_as_tuples() == want
    assert len(want) > 200000


def Engine_for(eng):
    """An engine with 64 MiB pieces holding the same database as `eng` (the module's engines use 8 MiB pieces: too little per warp)."""
    from matchy_b200 import Engine
    e = Engine(0, chunk_bytes=64 << 20)
    e.upload(eng._db_bytes)
    return e


@pytest.mark.parametrize("data", [b"", b"\n", b"x", b"1.2.3.4", b"evil.com", b"no newline at the end 8.8.8.8", b"\n\n\n", b" " * 5000,
                                  b"a" * 5000, b"a." * 5000 + b"com\n", b"@" * 3000, b":" * 3000, b"." * 3000])
def test_edge_inputs(engines, data):
    eng, orc, _ = engines[5]
    eng.scan(data)
    want, wcnt = orc.scan(data)
    assert eng.counters_list() == wcnt
    assert eng.records_as_tuples() == want


def test_single_lookups(engines, small_dbs):
    """Database.lookup / lookup_ip through the device tables == the oracle's single-query helpers."""
    eng, orc, log = engines[5]
    toks = [w for w in log.replace(b"=", b" ").replace(b"\"", b" ").split() if b"." in w][:300]
    for w in toks + [b"", b"nothing.example.invalid", b"a" * 300]:
        assert eng.lookup_string(w) == orc.lookup_string(w), w
    rng = random.Random(5)
    for _ in range(300):
        a = rng.getrandbits(32)
        rc, off, pl = orc.lookup_ip4(a)
        found, goff, gpl = eng.lookup_ip(a.to_bytes(4, "big"))
        assert (found, goff if found else 0, gpl if found else 0) == (rc == 1, off if rc == 1 else 0, pl if rc == 1 else 0)


def test_full_size_properties(built):
    """At a size the oracle cannot cover quickly: splitting the log must not change the summed counters, and
    scanning the same buffer twice is idempotent (size-independent properties)."""
    from matchy_b200 import Engine, synth
    db = synth.build_db(5, 0.02)
    log = synth.gen_log(5, 256 << 20, 0.02)
    e = Engine(0, chunk_bytes=64 << 20)
    e.upload(db)
    e.set_keep_results(True)
    e.scan(log)
    whole = e.counters_list()
    recs_whole = e.results()[0].copy()
    e.scan(log)
    assert e.counters_list() == whole
    assert np.array_equal(e.results()[0], recs_whole)
    half = (len(log) // 2) // 65536 * 65536  # block edges are line edges
    acc = [0] * 16
    n = 0
    for a, b in ((0, half), (half, len(log))):
        e.scan(log[a:b], base=a)
        acc = [x + y for x, y in zip(acc, e.counters_list())]
        n += len(e.results()[0])
    assert acc == whole
    assert n == len(recs_whole)
    # a 16 MiB slice against the oracle
    orc = O.Oracle(db)
    sl = log[:16 << 20].tobytes()
    e.scan(sl)
    want, wcnt = orc.scan(sl, chunk_size=128 * 1024)
    assert e.counters_list() == wcnt
    assert e.records_as_tuples() == want


def test_literal_search_formulations_agree_on_device(engines, small_dbs):
    """The anchored-walk fast path and the Aho-Corasick walk give identical records on the device."""
    from matchy_b200 import DatabaseBuilder, Engine
    for cfg in (2, 5):
        eng, orc, log = engines[cfg]
        eng.scan(log)
        a = (eng.records_as_tuples(), eng.counters_list())
        for mode in (1, 2):  # generic kernels: Aho-Corasick walk / anchored walks; mode 0 = constant-time filters + exact kernel
            eng.set_ac_mode(mode)
            eng.scan(log)
            b = (eng.records_as_tuples(), eng.counters_list())
            eng.set_ac_mode(0)
            assert a == b, mode
    bld = DatabaseBuilder(build_epoch=1)
    for g in ("*abcab*", "*bcabc*", "*cab*", "*.evil.com", "*evil.com*", "*aaa*", "*aaaa*", "*abc*abc*", "abcabc", "zz"):
        bld.add_glob(g, {"g": g})
    db = bld.build()
    e = Engine(0)
    e.upload(db)
    orc = O.Oracle(db)
    data = b"".join(b"h=" + t + b".example.com\n" for t in [b"abcabcab", b"xabcabcy", b"aaaaaaa", b"www.evil.com", b"buzz", b"cabcabcabc.abcab"])
    e.scan(data)
    want, wcnt = orc.scan(data)
    assert e.records_as_tuples() == want and e.counters_list() == wcnt and len(want) >= 5


def test_fast_string_path_on_device(built):
    """Anchored globs of every key length + literals through the constant-time filters, against the oracle; then the
    same database with an unanchored glob added (generic kernels)."""
    from matchy_b200 import DatabaseBuilder, Engine
    rng = random.Random(5)
    sfx = ["*m", "*om", "*.io", "*l.io", "*il.io", "*vil.io", "*evil.io", "*.evil.io", "*x.evil.io", "*.very-long-suffix.example.net", "*[0-9].bad.org",
           "a?c*zz.org", "*.??.uk"]
    pfx = ["a*", "ab*", "abc.*", "abcd*.z?", "abcde*[a-z]", "abcdef.*", "abcdefg*", "abcdefgh*", "abcdefghi.*", "prefix-longer-than-eight-*", "p.q*r?"]
    lits = ["a.io", "ab.io", "abc.com", "abcd.com", "abcde.io", "evil.io", "x.evil.io", "abcdefgh.com", "a-much-longer-literal.example.com",
            "5d41402abc4b2a76b9719d911017c592", "user@evil.io"]
    toks = [b"a.io", b"ab.io", b"abc.com", b"abcd.zz", b"abcde.az", b"abcdef.com", b"abcdefg.com", b"abcdefgh.com", b"abcdefghi.com", b"evil.io", b"x.evil.io",
            b"xx.evil.io", b"l.io", b"il.io", b"vil.io", b"m.com", b"a.om", b"prefix-longer-than-eight-x.com", b"prefix-longer-than-eight.com",
            b"q.very-long-suffix.example.net", b"very-long-suffix.example.net", b"a7.bad.org", b"ax.bad.org", b"abc.qzz.org", b"a.cc.uk", b"p.qr.rs",
            b"a-much-longer-literal.example.com", b"x-much-longer-literal.example.com", b"user@evil.io", b"u@x.evil.io", b"5d41402abc4b2a76b9719d911017c592",
            b"5d41402abc4b2a76b9719d911017c593", b"abc.io", b"b.io", b"a.b.c.d.e.io", b"zz.org", b"ac.zz.org", b"abc.zz.org"]
    data = b"".join(b"k=" + t + b" \n" for t in toks)
    for _ in range(3000):
        data += b"h=" + rng.choice(toks)[:rng.randint(1, 40)] + rng.choice([b"", b".io", b".com", b"m", b".evil.io", b".uk"]) + rng.choice([b" ", b"\n"])
    for extra in (None, "*mid*"):
        b = DatabaseBuilder(build_epoch=1)
        for g in sfx + pfx + ([extra] if extra else []):
            b.add_glob(g, {"g": g})
        for l in lits:
            b.add_entry(l, {"l": l})
        db = b.build()
        e = Engine(0, chunk_bytes=1 << 20)
        e.upload(db)
        orc = O.Oracle(db)
        want, wcnt = orc.scan(data)
        assert len(want) >= 30
        e.scan(data)
        assert e.counters_list() == wcnt
        assert e.records_as_tuples() == want
        e.close()


def test_crypto_address_extraction_on_device(engines):
    """All extractors on (Extractor::new()): Bitcoin / Ethereum / Monero tokens agree with the oracle on the reference's own
    vectors, near misses of every rule and random token soup; and a scan with the crypto extractors on keeps the match set."""
    eng, orc, log = engines[1]
    rng = random.Random(17)
    monero_ok = _monero_like_word()
    seen = set()
    for it in range(120):
        parts = []
        for _ in range(rng.randint(1, 50)):
            f = rng.choice(CRYPTO_FRAGS + [monero_ok]) if rng.random() < 0.6 else rng.choice(FRAGS)
            parts.append(f)
            parts.append(rng.choice([b" ", b" ", b"\n", b"/", b"=", b":", b"", b".", b",", b"\"", b"-"]))
        data = b"".join(parts)
        want = sorted((s, t, e) for t, s, e in orc.extract(data, 0xFF))
        got = sorted((s, t, e) for t, s, e in eng.extract(data, 0xFF))
        assert got == want, (it, data[:200])
        seen |= {t for _, t, _ in want}
    assert {9, 10, 11} <= seen  # every crypto type was exercised
    data = log[:2_000_000] + b"pay 1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa or 0x5aeda56215b167893e80b4fe645ba6d5bab767de now\n"
    eng.scan(data, flags=eng.default_flags() | 0xE0)
    got, cnt = eng.records_as_tuples(), eng.counters_list()
    want, wcnt = orc.scan(data, flags=orc.default_flags() | 0xE0, chunk_size=128 * 1024)
    assert cnt == wcnt and got == want
    assert cnt[4 + 9] >= 1 and cnt[4 + 10] >= 1


def _monero_like_word():
    """A word that satisfies the reference's Monero rule (plain base58 of payload + Keccak-256[..4], '4'/'8' first, 90..110 chars)."""
    A = b"123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"

    def b58(raw):
        n, out = int.from_bytes(raw, "big"), bytearray()
        while n:
            n, r = divmod(n, 58)
            out.append(A[r])
        return bytes(reversed(out))
    for first in range(1, 256):
        p = bytes([first]) + bytes(range(1, 65))
        s = b58(p + O.digest("keccak256", p)[:4])
        if 90 <= len(s) <= 110 and s[:1] in (b"4", b"8"):
            return s
    raise AssertionError


def test_match_cli_output_contract(small_dbs, tmp_path):
    """`python -m matchy_b200 match` prints exactly the oracle's sorted `matchy match` NDJSON for the same files."""
    import subprocess
    import sys
    db, log = small_dbs[5]
    import gzip
    dbp, l1, l2 = tmp_path / "t.mxy", tmp_path / "a.log", tmp_path / "b.log.GZ"  # .gz in any case is inflated (file_reader.rs:60-75)
    dbp.write_bytes(db)
    cut = log.rfind(b"\n", 0, len(log) // 2) + 1
    l1.write_bytes(log[:cut])
    l2.write_bytes(gzip.compress(log[cut:]) + gzip.compress(b"second member: 45.0.0.1 is ignored like flate2's GzDecoder does\n"))
    r = subprocess.run([sys.executable, "-m", "matchy_b200", "match", str(dbp), str(l1), str(l2), "--extractors=-crypto", "--stats"],
                       capture_output=True, cwd=str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    assert r.returncode == 0, r.stderr[-500:]
    orc = O.Oracle(db)
    want = []
    for path, part in ((l1, log[:cut]), (l2, log[cut:])):
        orc.scan(part, chunk_size=128 * 1024)
        want += orc.ndjson(part, str(path)).splitlines()
    assert sorted(r.stdout.splitlines()) == sorted(want) and len(want) > 0
    assert b'"matches"' in r.stderr


def _expected_extract_output(orc, data: bytes, flags, fmt="json", unique=False):
    """`matchy extract` restated around the ORACLE's extractor, one line at a time (bin/commands/extract_cmd.rs:200-289,
    LineScanner bin/cli_utils.rs:9-95): trim ASCII whitespace, skip empty lines, extract_from_line order."""
    ws = b" \t\n\x0c\r"
    names = ["domain", "email", "ipv4", "ipv6", "md5", "sha1", "sha256", "sha384", "sha512", "bitcoin", "ethereum", "monero"]
    group = {0: 0, 2: 1, 1: 2, 3: 3, 4: 4, 5: 4, 6: 4, 7: 4, 8: 4, 9: 5, 10: 6, 11: 7}
    out, seen, lines = [], set(), 0
    if fmt == "csv":
        out.append(b"type,value\n")
    for raw in data.split(b"\n"):
        line = raw.strip(ws)
        if not line:
            continue
        lines += 1
        items = sorted(orc.extract(line, flags), key=lambda t: (group[t[0]], t[1]))
        for t, s, e in items:
            text = line[s:e]
            if unique:
                if text in seen:
                    continue
                seen.add(text)
            if fmt == "json":
                out.append(b'{"type":"%s","value":"%s"}\n' % (names[t].encode(), text.replace(b"\\", b"\\\\").replace(b'"', b'\\"')))
            elif fmt == "csv":
                out.append(b'%s,"%s"\n' % (names[t].encode(), text.replace(b'"', b'""')))
            else:
                out.append(text + b"\n")
    return b"".join(out), lines


@pytest.mark.gpu
def test_extract_subcommand_matches_per_line_oracle(built, small_dbs, tmp_path, capfdbinary):
    """`python -m matchy_b200 extract` == the reference's `matchy extract` contract, checked against the oracle run line by line."""
    from matchy_b200.__main__ import main
    db, log = small_dbs[5]
    orc = O.Oracle(db)
    tricky = (b"\x0cevil.com after a form feed\n"            # FF is trimmed by LineScanner but is no token boundary
              b"  \t padded.example.org \r\n"
              b"\n   \n\x0c\n"                                # empty after trimming: not lines at all
              b"mid\x0cdle.com x\x0c\n"                       # FF inside the line stays; trailing FF is trimmed
              b"dup.example.com dup.example.com 10.1.2.3 10.1.2.3\n"
              b'quote "a\\b.example.net" user@mail.example.com 2001:db8::7 5d41402abc4b2a76b9719d911017c592\n'
              b"1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa 0x5aAeb6053F3E94C9b9A09f33669435E7Ef1BeAed\n"
              b"last line without newline evil.org")
    data = tricky[: tricky.rfind(b"\n") + 1] + log[: 96 * 1024][: log[: 96 * 1024].rfind(b"\n") + 1] + tricky[tricky.rfind(b"\n") + 1:]
    path = tmp_path / "in.log"
    path.write_bytes(data)
    for args, flags, fmt, uniq in ((["--format", "json"], 0xFF, "json", False), (["--format", "csv", "-u"], 0xFF, "csv", True),
                                   (["--format", "text", "--types", "ip"], 0xF0 | 0x0C, "text", False),
                                   (["--types", "domain,email"], 0xF0 | 0x03, "json", False)):
        capfdbinary.readouterr()
        assert main(["extract", str(path)] + args + ["--stats"]) == 0
        got = capfdbinary.readouterr()
        want, lines = _expected_extract_output(orc, data, flags, fmt, uniq)
        assert got.out == want, args
        assert want.count(b"\n") > 100
        assert ("Lines processed: %s" % format(lines, ",")).encode() in got.err
    small = tmp_path / "small.log"
    small.write_bytes(b"\x0cevil.com after a form feed\n  \t padded.example.org 10.1.2.3\n")
    capfdbinary.readouterr()
    assert main(["extract", str(small), "--format", "text", "--show-candidates"]) == 0
    got = capfdbinary.readouterr()
    assert got.out == b"evil.com\npadded.example.org\n10.1.2.3\n"
    assert b"[CANDIDATE] Domain at 0-8: evil.com" in got.err and b"[CANDIDATE] IPv4 at 19-27: 10.1.2.3" in got.err  # spans index the trimmed line
    assert main(["extract", str(path), "--format", "xml"]) == 1
    assert main(["extract", str(path), "--types", "bogus"]) == 1


def test_paraglob_integration_vectors_on_device(built):
    """The reference's paraglob integration pattern sets (matchy-paraglob/tests/integration_tests.rs:11-258, ported in
    test_oracle_kats.py): the device's single-string lookup returns the oracle's (pattern id, data offset) pairs for every text."""
    from matchy_b200 import Engine
    from test_oracle_kats import _globs
    sets = [(["*.txt", "test*", "*file*"], 0, ["document.txt", "test_case", "myfile.dat", "nomatch"]),
            (["hello", "world", "test"], 0, ["hello", "world", "hello world", "nomatch"]),
            (["*test*", "*test*", "hello", "hello"], 0, ["test123", "hello"]),
            (["*.txt", "*file*", "test*"], 0, ["testfile.txt"]),
            (["Test*", "HELLO"], 0, ["Test123", "test123", "HELLO", "hello"]),
            (["Test*", "HELLO"], 1, ["Test123", "test123", "HELLO", "hello"]),
            (["*test*", "test*", "*test"], 0, ["test", "testing", "mytest", "mytesting"]),
            (["*.rs", "*.toml", "Cargo.*", "src/*", "*.md"], 0, ["main.rs", "Cargo.toml", "src/lib.rs", "README.md", "test.py"]),
            (["pattern_%d_*" % i for i in range(1000)], 0, ["pattern_500_test", "pattern_999_data", "nomatch"]),
            (["hello", "*.txt", "test_*"], 0, ["hello.txt", "test_file.txt"]),
            (["*", "?", "**"], 0, ["test", "a"])]
    eng = Engine(0, chunk_bytes=1 << 20)
    for patterns, mode, texts in sets:
        orc = _globs(patterns, mode)
        eng.upload(orc._db)
        for t in texts:
            assert eng.lookup_string(t.encode()) == orc.lookup_string(t.encode()), (patterns[:3], t)
    eng.close()


def test_device_generator_equals_host_generator(built):
    """mgen_log_device (csrc/synth_device.cu) writes the bytes mgen_log (csrc/synth.cpp) writes: same integer code
    (csrc/synth_gen.h) on both sides, every config / line family, several offsets, full and 1 % database scale."""
    import ctypes as C
    from matchy_b200 import Engine, synth, _native as N
    eng = Engine(0, chunk_bytes=1 << 20)
    nbytes = 48 * 65536
    dev = eng.dev_alloc(nbytes)
    got = np.empty(nbytes, dtype=np.uint8)
    try:
        for cfg in (1, 2, 3, 4, 5):
            for scale in (1.0, 0.01):
                for off in (0, 5 * 65536, 123457 * 65536):
                    synth.gen_log_device(0, cfg, dev, nbytes, scale, offset=off)
                    assert N.lib().mgpu_dev_download(C.c_void_p(eng.h), C.c_void_p(got.ctypes.data), C.c_void_p(dev), nbytes) == 0
                    want = synth.gen_log(cfg, nbytes, scale, offset=off)
                    assert np.array_equal(got, want), (cfg, scale, off)
    finally:
        eng.dev_free(dev)
    eng.close()


def _canon(line):
    """A `matchy match` JSON line without its timestamp, IPv6 matched_text in canonical form (what sequential mode prints)."""
    import json
    d = json.loads(line)
    d.pop("timestamp")
    if ":" in d["matched_text"]:
        d["matched_text"] = O.ipv6_display(O.parse_ipv6(d["matched_text"].encode()))
    return json.dumps(d, sort_keys=True)


def test_sequential_mode_and_follow_mode_cli(small_dbs, tmp_path):
    """`match --threads 1`: the oracle's match set, in line order, with wall-clock timestamps and canonical address text;
    `match --follow`: the same lines for content that arrives while the command runs."""
    import json
    import subprocess
    import sys
    import time
    db, log = small_dbs[5]
    log = log[: log.rfind(b"\n", 0, 400_000) + 1] + b"v6 2001:0DB8:0000:0000:0000:0000:0000:0001 and 2001:db8::0:1 end\n"
    dbp, lp = tmp_path / "t.mxy", tmp_path / "a.log"
    dbp.write_bytes(db)
    lp.write_bytes(log)
    root = str(__import__("pathlib").Path(__file__).resolve().parents[1])
    orc = O.Oracle(db)
    orc.scan(log, flags=orc.default_flags() | 0xE0, chunk_size=128 * 1024)
    want = sorted(_canon(l) for l in orc.ndjson(log, str(lp)).splitlines())
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "matchy_b200", "match", str(dbp), str(lp), "--threads", "1"], capture_output=True, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    lines = r.stdout.splitlines()
    assert sorted(_canon(l) for l in lines) == want and len(want) > 50
    stamps = [float(json.loads(l)["timestamp"]) for l in lines]
    assert all(t0 - 1 <= s <= time.time() + 1 for s in stamps)
    # line order: the position of each match's line in the log never decreases
    starts, pos = [0] + [i + 1 for i, b in enumerate(log) if b == 10], 0
    import bisect
    prev = -1
    for l in lines:
        text = json.loads(l)["matched_text"].encode()
        at = log.find(text, starts[max(prev, 0)]) if ":" not in text.decode() else -1
        if at >= 0:
            ln = bisect.bisect_right(starts, at) - 1
            assert ln >= prev
            prev = ln
    # follow mode: half of the log is there at the start, the rest arrives later
    cut = log.rfind(b"\n", 0, len(log) // 2) + 1
    lp.write_bytes(log[:cut])
    proc = subprocess.Popen([sys.executable, "-m", "matchy_b200", "match", str(dbp), str(lp), "--follow", "--follow-idle-exit", "12"],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=root)
    time.sleep(6)  # (the existing half is processed before the rest is appended, or — slow engine start-up — together with it: both are valid;
    #                 the idle limit is longer than this wait, so the command cannot leave before the second half arrives)
    with open(lp, "ab") as f:
        f.write(log[cut:])
    out, err = proc.communicate(timeout=120)
    assert proc.returncode == 0, err[-500:]
    assert sorted(_canon(l) for l in out.splitlines()) == want


def test_watching_database_reloads_into_the_same_engine(built, tmp_path):
    """WatchingDatabase (watching_database.rs:103-191): a changed database file is uploaded into the running engine, the
    generation counter moves, scans see the new indicators; a corrupt file is refused and the old database stays."""
    from matchy_b200 import DatabaseBuilder
    from matchy_b200.match_modes import WatchingDatabase
    from matchy_b200.database import DatabaseError

    def make(entries):
        b = DatabaseBuilder(build_epoch=1)
        for e in entries:
            b.add_entry(e, {"e": e})
        return b.build()
    p = tmp_path / "w.mxy"
    p.write_bytes(make(["evil.com", "10.0.0.0/8"]))
    seen = []
    db = WatchingDatabase.from_(str(p)).no_thread().on_reload(lambda gen, path: seen.append(gen)).open()
    data = b"a evil.com b bad.org c 10.1.2.3 d 192.168.1.1\n"
    recs, _ = db.engine.scan(data)
    assert sorted(bytes(data[int(r["offset"]):int(r["offset"]) + int(r["len"])]) for r in recs) == [b"10.1.2.3", b"evil.com"]
    assert db.generation() == 1 and not db.check()
    os_stat = p.stat()
    p.write_bytes(make(["bad.org", "192.168.0.0/16"]))
    __import__("os").utime(p, ns=(os_stat.st_atime_ns, os_stat.st_mtime_ns + 10_000_000))
    assert db.check() and db.generation() == 2 and seen == [2]
    recs, _ = db.engine.scan(data)
    assert sorted(bytes(data[int(r["offset"]):int(r["offset"]) + int(r["len"])]) for r in recs) == [b"192.168.1.1", b"bad.org"]
    assert db.lookup("bad.org").kind == "Pattern" and db.lookup("evil.com").is_not_found()
    p.write_bytes(b"not a database" * 100)
    with pytest.raises(DatabaseError):
        db.check()
    assert db.generation() == 2
    recs, _ = db.engine.scan(data)
    assert len(recs) == 2  # the old tables are still there
    db.close()

