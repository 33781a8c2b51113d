"""Pins the CPU oracle against the reference's own known-answer tests (ported; file:line cited per block).

The reference cannot be compiled here (no Rust toolchain), so these vectors are what anchors parity:
  * extractor KATs          crates/matchy-extractor/src/lib.rs:1922-3235
  * longest-prefix tests    crates/matchy/tests/test_ip_longest_prefix_match.rs, test_ip_exact_match.rs
  * literal + glob tests    crates/matchy/tests/test_literal_hash.rs, matchy-paraglob/src/glob.rs:465-706,
                            matchy-paraglob/tests/integration_tests.rs
  * worker smoke            crates/matchy/src/processing/mod.rs:614-702
  * XXH64 vectors           public xxHash specification (cross-checked with the python `xxhash` module)
"""
import pytest

import oracle_lib as O


@pytest.fixture(scope="module")
def orc(built):
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1)
    b.add_entry("1.2.3.4", {"k": "v"})
    return O.Oracle(b.build())


def ex(orc, line, flags=O.X_DEFAULT):
    return orc.extract_strings(line, flags)


def only(items, kind):
    return [v for k, v in items if k == kind]


def hashes(items):
    return [(k, v) for k, v in items if k in ("MD5", "SHA1", "SHA256", "SHA384", "SHA512")]


# ---- domains (lib.rs:2093-2180, 2222-2365, 2877-2921) ----------------------------------------------------
def test_domain_basic(orc):
    assert ex(orc, b"Visit example.com for more info") == [("Domain", b"example.com")]
    assert only(ex(orc, b"Check google.com and github.com"), "Domain") == [b"google.com", b"github.com"]
    assert only(ex(orc, b"Visit api.example.com today"), "Domain") == [b"api.example.com"]
    assert only(ex(orc, b"Go to https://www.example.com/path"), "Domain") == [b"www.example.com"]


def test_domain_log_line(orc):
    got = only(ex(orc, b"2024-01-15 10:32:45 GET /api evil.example.com 192.168.1.1 - malware.badsite.org"), "Domain")
    assert b"evil.example.com" in got and b"malware.badsite.org" in got


def test_unicode_domains(orc):
    assert only(ex(orc, "Visit münchen.de for info".encode()), "Domain") == ["münchen.de".encode()]
    got = only(ex(orc, "Check café.fr and example.com".encode()), "Domain")
    assert b"example.com" in got and "café.fr".encode() in got


def test_binary_junk_and_invalid_utf8(orc):
    assert b"evil.com" in only(ex(orc, b"Log: \xff\xfe evil.com \x80"), "Domain")
    assert only(ex(orc, b"Visit \xff\xc0.com"), "Domain") == []


def test_false_positives_rejected(orc):
    assert not any(d.endswith(b".com") for d in only(ex(orc, b"This is blah.community stuff"), "Domain"))
    assert only(ex(orc, b"Request: host=api.example.com method=GET path=/test"), "Domain") == [b"api.example.com"]
    assert only(ex(orc, b"Invalid domain: Kagi%20Assistant.app"), "Domain") == []
    assert only(ex(orc, b"Visit .app or .com for info"), "Domain") == []


def test_domain_quirks_from_survey(orc):
    # trailing dot, leading dot, underscore neighbours, '&' not a boundary, case-sensitive PSL
    assert only(ex(orc, b"evil.com. "), "Domain") == []
    assert only(ex(orc, b" .evil.com "), "Domain") == []
    assert only(ex(orc, b"_evil.com evil.com_"), "Domain") == []
    assert only(ex(orc, b"?host=evil.com&x"), "Domain") == []
    assert only(ex(orc, b"host=evil.com "), "Domain") == [b"evil.com"]
    assert only(ex(orc, b"EVIL.COM Evil.com"), "Domain") == [b"Evil.com"]
    assert only(ex(orc, b"www.example.co.uk"), "Domain") == [b"www.example.co.uk"]


# ---- IPv4 (lib.rs:2182-2220, 2368-2384, 2629-2684) -------------------------------------------------------
def test_ipv4(orc):
    assert only(ex(orc, b"Server at 192.168.1.1 responded"), "IPv4") == [b"192.168.1.1"]
    assert only(ex(orc, b"Traffic from 10.0.0.5 to 172.16.0.10"), "IPv4") == [b"10.0.0.5", b"172.16.0.10"]
    assert only(ex(orc, b"Not IPs: 256.1.1.1 1.2.3.999 1.2.3"), "IPv4") == []
    assert only(ex(orc, b"Invalid IP: 2025.36.0.72591908"), "IPv4") == []
    assert only(ex(orc, b"Invalid IP: 460.1.1.2"), "IPv4") == []
    assert only(ex(orc, b"Invalid IP: 26.0..26.0"), "IPv4") == []
    assert only(ex(orc, b"1.2.3.4. 1.2.3.4-x v1.2.3.4 192.168.01.1"), "IPv4") == []


def test_mixed(orc):
    got = ex(orc, b"Request from 10.1.2.3 to api.example.com at 192.168.1.100")
    assert only(got, "IPv4") == [b"10.1.2.3", b"192.168.1.100"] and only(got, "Domain") == [b"api.example.com"]
    got = ex(orc, b"2024-01-15 user@example.com from 10.1.2.3 accessed api.test.com")
    assert only(got, "Email") == [b"user@example.com"] and only(got, "IPv4") == [b"10.1.2.3"]
    assert sorted(only(got, "Domain")) == [b"api.test.com", b"example.com"]


# ---- e-mail (lib.rs:2416-2472, 2686-2798) ------------------------------------------------------------------
def test_email(orc):
    assert only(ex(orc, b"Contact user@example.com for info"), "Email") == [b"user@example.com"]
    assert only(ex(orc, b"Email alice@test.com or bob@company.org"), "Email") == [b"alice@test.com", b"bob@company.org"]
    assert only(ex(orc, b"Send to user+tag@example.com"), "Email") == [b"user+tag@example.com"]
    assert only(ex(orc, b"Invalid email: s...@example.com"), "Email") == []
    assert only(ex(orc, b"Invalid email: .@example.com"), "Email") == []
    assert only(ex(orc, b"Valid email: 34480FE2-5610-4973-AA09-3ABB60D38D55@example.com"), "Email") == [b"34480FE2-5610-4973-AA09-3ABB60D38D55@example.com"]
    assert only(ex(orc, b"Invalid email: user@192.168.1.222"), "Email") == []
    assert only(ex(orc, b"Invalid email: test@Uv3.peer"), "Email") == []
    assert only(ex(orc, b"odd but accepted: user@.com"), "Email") == [b"user@.com"]


# ---- IPv6 (lib.rs:2517-2624, 2801-2874, 2924-2946) ---------------------------------------------------------
def test_ipv6(orc):
    assert only(ex(orc, b"Server at 2001:db8:85a3::8a2e:370:7334 responded"), "IPv6") == [b"2001:db8:85a3::8a2e:370:7334"]
    assert only(ex(orc, b"Connecting to 2001:db8::1"), "IPv6") == [b"2001:db8::1"]
    assert only(ex(orc, b"Address 2001:0db8::1 connects to 2606:2800:220:1::248"), "IPv6") == [b"2001:0db8::1", b"2606:2800:220:1::248"]
    got = ex(orc, b"IPv4: 192.168.1.1 IPv6: 2001:db8::1")
    assert only(got, "IPv4") == [b"192.168.1.1"] and only(got, "IPv6") == [b"2001:db8::1"]
    for line in (b"Tiny IPv6: e::f", b"Tiny IPv6: ce::A", b"Tiny IPv6: e::add", b"Invalid IPv6: FEC0050519FB::c", b"Invalid IPv6: 7::31BD71E4",
                 b"Link-local address: fe80::1 and fe80::dead:beef"):
        assert only(ex(orc, line), "IPv6") == [], line


def test_ipv6_text_semantics():
    assert O.parse_ipv6(b"2001:0db8::1") == [0x2001, 0xdb8, 0, 0, 0, 0, 0, 1]
    assert O.ipv6_display([0x2001, 0xdb8, 0, 0, 0, 0, 0, 1]) == "2001:db8::1"
    assert O.ipv6_display([0x2001, 0xdb8, 0x85a3, 0, 0x8a2e, 0x370, 0x7334, 1]) == "2001:db8:85a3:0:8a2e:370:7334:1"
    assert O.ipv6_display([0, 0, 0, 0, 0, 0xffff, 0x0102, 0x0304]) == "::ffff:1.2.3.4"
    assert O.ipv6_display([1, 0, 0, 2, 0, 0, 0, 3]) == "1:0:0:2::3"
    assert O.parse_ipv6(b"1:2:3:4:5:6:7:8") == [1, 2, 3, 4, 5, 6, 7, 8]
    assert O.parse_ipv6(b"1:2:3:4:5:6:7::") == [1, 2, 3, 4, 5, 6, 7, 0]
    for bad in (b"1::2::3", b"12345::1", b"1:::2", b":1::2", b"1:2:3:4:5:6:7:8:9", b"1:2:3:4:5:6:7:8::"):
        assert O.parse_ipv6(bad) is None, bad


# ---- hashes (lib.rs:2038-2059, 2949-3235) ------------------------------------------------------------------
def test_hashes(orc):
    assert hashes(ex(orc, b"File hash: 5d41402abc4b2a76b9719d911017c592 uploaded")) == [("MD5", b"5d41402abc4b2a76b9719d911017c592")]
    assert hashes(ex(orc, b"SHA1: 2fd4e1c67a2d28fced849ee1bb76e7391b93eb12 verified")) == [("SHA1", b"2fd4e1c67a2d28fced849ee1bb76e7391b93eb12")]
    assert hashes(ex(orc, b"SHA256: 2c26b46b68ffc68ff99b453c1d30413413422d706483bfa0f98a5e886266e7ae detected"))[0][0] == "SHA256"
    assert hashes(ex(orc, b"SHA384: cb00753f45a35e8bb5a03d699ac65007272c32ab0eded1631a8b605a43ff5bed8086072ba1e7cc2358baeca134c825a7 verified"))[0][0] == "SHA384"
    sha512 = b"cf83e1357eefb8bdf1542850d66d8007d620e4050b5715dc83f4a921d36ce9ce47d0d13c5d85f2b0ff8318d2877eec2f63b931bd47417a81a538327af927da3e"
    assert hashes(ex(orc, b"SHA512: " + sha512 + b" found")) == [("SHA512", sha512)]
    assert [k for k, _ in hashes(ex(orc, b"MD5: 5d41402abc4b2a76b9719d911017c592 SHA1: 2fd4e1c67a2d28fced849ee1bb76e7391b93eb12"))] == ["MD5", "SHA1"]
    assert hashes(ex(orc, b"Hash: 5D41402ABC4B2A76B9719D911017C592 found")) == [("MD5", b"5D41402ABC4B2A76B9719D911017C592")]
    assert hashes(ex(orc, b"Hash: 5d41402AbC4b2A76b9719D911017c592 mixed")) == [("MD5", b"5d41402AbC4b2A76b9719D911017c592")]
    assert hashes(ex(orc, b"Hash: 5d41402abc4b2a76b9719d91101 invalid")) == []
    assert hashes(ex(orc, b"Hash: 5d41402abc4b2a76b9719d911017c5gz invalid")) == []
    assert hashes(ex(orc, b"Hash: [5d41402abc4b2a76b9719d911017c592] in brackets")) == [("MD5", b"5d41402abc4b2a76b9719d911017c592")]
    got = ex(orc, b"2024-01-15 malware.exe MD5=5d41402abc4b2a76b9719d911017c592 detected from 192.168.1.100")
    assert len(hashes(got)) == 1 and only(got, "IPv4") == [b"192.168.1.100"]
    assert [k for k, _ in hashes(ex(orc, b"Line1: 5d41402abc4b2a76b9719d911017c592\nLine2: 2fd4e1c67a2d28fced849ee1bb76e7391b93eb12\n"))] == ["MD5", "SHA1"]
    assert hashes(ex(orc, b"Hash: 5d41402abc4b2a76b9719d911017c592 should not extract", O.X_DEFAULT & ~O.X_HASHES)) == []
    assert hashes(ex(orc, b"UUID: 550e8400-e29b-41d4-a716-446655440000 not a hash")) == []


def test_chunk_order_is_by_type(orc):
    """extract_from_chunk order: IPv6, IPv4, e-mail, domain, hashes (lib.rs:449-472)."""
    line = b"Check example.com user@test.com 192.168.1.1 2001:db8::1 5d41402abc4b2a76b9719d911017c592"
    assert [k for k, _ in ex(orc, line)] == ["IPv6", "IPv4", "Email", "Domain", "Domain", "MD5"]


def test_xxh64_vectors():
    assert O.xxh64(b"") == 0xef46db3751d8e999
    assert O.xxh64(b"a") == 0xd24ec4f1a98c6e5b
    assert O.xxh64(b"evil.com") == 0x4a37aa533dbb4ae5
    assert O.xxh64(b"5d41402abc4b2a76b9719d911017c592") == 0xca5c562425799b4d
    try:
        import xxhash
    except ImportError:
        return
    for n in (0, 1, 3, 4, 7, 8, 15, 31, 32, 33, 63, 64, 100, 257):
        b = bytes((i * 7 + 3) & 255 for i in range(n))
        assert O.xxh64(b) == xxhash.xxh64(b, seed=0).intdigest()


# ---- longest-prefix match (crates/matchy/tests/test_ip_longest_prefix_match.rs:13-320, test_ip_exact_match.rs) ----
def _db(entries, mode=0):
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(mode, build_epoch=1)
    for k, v in entries:
        b.add_entry(k, v)
    return O.Oracle(b.build())


def _ip4(s):
    a, b, c, d = (int(x) for x in s.split("."))
    return (a << 24) | (b << 16) | (c << 8) | d


def test_lpm_specific_first_and_last():
    for order in ([("192.0.2.1/32", {"type": "specific"}), ("192.0.2.0/24", {"type": "general"})],
                  [("192.0.2.0/24", {"type": "general"}), ("192.0.2.1/32", {"type": "specific"})]):
        o = _db(order)
        rc, off, pl = o.lookup_ip4(_ip4("192.0.2.1"))
        assert rc == 1 and pl == 32 and '"specific"' in o.data_json(off)
        rc, off, pl = o.lookup_ip4(_ip4("192.0.2.2"))
        assert rc == 1 and '"general"' in o.data_json(off)
        assert o.lookup_ip4(_ip4("192.0.3.1"))[0] == 0


def test_lpm_three_levels_any_order():
    import itertools
    ents = [("10.0.0.0/8", {"level": "8"}), ("10.1.0.0/16", {"level": "16"}), ("10.1.1.0/24", {"level": "24"}), ("10.1.1.1/32", {"level": "32"})]
    for perm in itertools.permutations(ents):
        o = _db(list(perm))
        for ip, lvl in (("10.1.1.1", "32"), ("10.1.1.2", "24"), ("10.1.2.1", "16"), ("10.2.0.1", "8")):
            rc, off, pl = o.lookup_ip4(_ip4(ip))
            assert rc == 1 and ('"%s"' % lvl) in o.data_json(off), (ip, perm)
        assert o.lookup_ip4(_ip4("11.0.0.1"))[0] == 0


def test_lpm_ipv6_and_v4_in_v6_tree():
    o = _db([("2001:db8::/32", {"n": "wide"}), ("2001:db8::1/128", {"n": "host"}), ("192.0.2.0/24", {"n": "v4"})])
    rc, off, pl = o.lookup_ip6([0x2001, 0xdb8, 0, 0, 0, 0, 0, 1])
    assert rc == 1 and pl == 128 and '"host"' in o.data_json(off)
    rc, off, pl = o.lookup_ip6([0x2001, 0xdb8, 0, 0, 0, 0, 0, 2])
    assert rc == 1 and '"wide"' in o.data_json(off)
    rc, off, pl = o.lookup_ip4(_ip4("192.0.2.77"))
    assert rc == 1 and pl == 24 and '"v4"' in o.data_json(off)
    assert o.lookup_ip4(_ip4("192.0.3.77"))[0] == 0


def test_exact_ip_entries():
    o = _db([("1.2.3.4", {"a": 1}), ("5.6.7.8", {"a": 2})])
    assert o.lookup_ip4(_ip4("1.2.3.4"))[0] == 1 and o.lookup_ip4(_ip4("1.2.3.5"))[0] == 0
    assert o.lookup_ip4(_ip4("5.6.7.8"))[2] == 32


# ---- literals + globs (crates/matchy/tests/test_literal_hash.rs:6-266, glob.rs:465-706, integration_tests.rs) ----
def test_literal_and_glob_both_match():
    o = _db([("test.com", {"t": "lit"}), ("*.com", {"t": "glob"}), ("evil.example.com", {"t": "lit2"})])
    r = o.lookup_string(b"test.com")
    assert len(r) == 2  # literal id first (literal id space), then the glob
    assert '"lit"' in o.data_json(r[0][1]) and '"glob"' in o.data_json(r[1][1])
    assert len(o.lookup_string(b"other.com")) == 1
    assert o.lookup_string(b"other.org") == []
    assert len(o.lookup_string(b"evil.example.com")) == 2


GLOB_CASES = [  # (pattern, text, matches) — glob.rs:465-706 cases, restricted to what Paraglob::find_all can reach:
    # literals shorter than 3 bytes never enter the automaton (paraglob_offset.rs:553-555), and a `glob` entry without
    # wildcards is a substring match (:1153-1156).  Literal bytes of the class patterns add up to a multiple of 4 so that
    # the class items stay 4-byte aligned (see test_glob_quirks for the misaligned case).
    ("hello", "hello", True), ("hello", "say hello world", True), ("*.txt", "file.txt", True), ("*.txt", "file.log", False),
    ("test_*", "test_file", True), ("*test*", "my_test_file", True), ("file?.txt", "file1.txt", True), ("file?.txt", "file12.txt", False),
    ("file[0-9].txt", "file5.txt", True), ("file[0-9].txt", "filea.txt", False), ("file[!0-9].txt", "filea.txt", True),
    ("file[!0-9].txt", "file5.txt", False), ("*.evil.com", "www.evil.com", True), ("*.evil.com", "evil.com", False),
    ("abc*def*ghi", "abcXXdefYYghi", True), ("abc*def*ghi", "abcXXdefYY", False), ("a*b*c", "aXXbYYc", False),
    ("*", "anything", True), ("abc*", "abc", True), ("[abc]atch", "batch", True), ("[abc]atch", "datch", False),
    ("[a-cx-z]1.ex", "y1.ex", True), ("[a-cx-z]1.ex", "m1.ex", False), ("caf?.fr", "café.fr", True), ("*é*", "café.fr", False),
    ("*éé*", "caféé.fr", True),
]


@pytest.mark.parametrize("pattern,text,want", GLOB_CASES)
def test_glob_cases(pattern, text, want):
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1)
    b.add_glob(pattern, {"p": pattern})
    o = O.Oracle(b.build())
    assert (len(o.lookup_string(text.encode())) == 1) == want


def test_glob_quirks():
    # a glob whose literals are all shorter than 3 bytes is unreachable; `glob:` entry without
    # wildcards behaves as a substring match (quirk 12)
    o = _db([("a*", {"x": 1}), ("glob:test.com", {"x": 2})])
    assert o.lookup_string(b"abc") == []
    assert len(o.lookup_string(b"mytest.com.evil.org")) == 1
    # char-class items that land on a misaligned address make zerocopy's Ref::from_prefix fail; the error leaves the
    # match through `?`, so the pattern can never match (paraglob_offset.rs:1572-1577).  5 literal bytes => misaligned.
    assert _db([("[abc]at.io", {"x": 1})]).lookup_string(b"bat.io") == []
    assert len(_db([("[abc]atx.info", {"x": 1})]).lookup_string(b"batx.info")) == 1
    # case-insensitive database
    o = _db([("*.Evil.COM", {"x": 1}), ("Bad.Example", {"x": 2})], mode=1)
    assert len(o.lookup_string(b"WWW.EVIL.com")) == 1 and len(o.lookup_string(b"bad.EXAMPLE")) == 1


def test_worker_smoke():
    """processing/mod.rs:614-702: the worker finds 1.2.3.4, and evil.com + 8.8.8.8."""
    o = _db([("1.2.3.4", {"threat": "x"}), ("evil.com", {"threat": "y"}), ("8.8.8.8", {"threat": "z"})])
    recs, cnt = o.scan(b"Connection from 1.2.3.4\n")
    assert len(recs) == 1 and recs[0][0] == 16 and recs[0][1] == 7
    data = b"Visit evil.com or connect to 8.8.8.8\n"
    recs, cnt = o.scan(data)
    assert sorted(data[r[0]:r[0] + r[1]] for r in recs) == [b"8.8.8.8", b"evil.com"]
    assert cnt[0] == 1 and cnt[3] == 2
    nd = o.ndjson(data, "in.log").decode().splitlines()
    assert nd[0] == '{"cidr":"8.8.8.8/32","data":{"threat":"z"},"match_type":"ip","matched_text":"8.8.8.8","prefix_len":32,"source":"in.log","timestamp":"0.000"}'
    assert nd[1] == '{"data":[{"threat":"y"}],"match_type":"pattern","matched_text":"evil.com","pattern_count":1,"source":"in.log","timestamp":"0.000"}'


def test_next_batch_chunking_is_result_neutral(small_dbs):
    db, log = small_dbs[5]
    o = O.Oracle(db)
    whole = o.scan(log[:300000] + b"tail without newline 1.2.3.4")
    for cs in (1000, 4096, 128 * 1024):
        assert o.scan(log[:300000] + b"tail without newline 1.2.3.4", chunk_size=cs) == whole


# ---------------------------------------------------------------------------------------------------------
# Crypto-address extractors (matchy-extractor/src/lib.rs:1269-1409, 1799-1920; KATs :3240-3626)
# ---------------------------------------------------------------------------------------------------------
X_ALL = 0xFF  # every extractor, like Extractor::new()


def _crypto(o, data):
    return [(t, s) for t, s in o.extract_strings(data, X_ALL) if t in ("Bitcoin", "Ethereum", "Monero")]


def test_crypto_digests_known_answers():
    import hashlib
    for msg in (b"", b"abc", b"a" * 55, b"a" * 56, b"a" * 64, b"The quick brown fox jumps over the lazy dog", bytes(range(200))):
        assert O.digest("sha256", msg) == hashlib.sha256(msg).digest(), msg
    # Keccak-256 (the pre-standard padding tiny-keccak's Keccak::v256 uses — not hashlib's sha3_256)
    assert O.digest("keccak256", b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert O.digest("keccak256", b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert O.digest("keccak256", b"a" * 135) != O.digest("keccak256", b"a" * 136)  # one block vs two
    assert O.digest("keccak256", b"a" * 200) != hashlib.sha3_256(b"a" * 200).digest()


def test_bitcoin_kats(small_dbs):
    o = O.Oracle(small_dbs[1][0])
    # lib.rs:3240-3298: legacy, P2SH, bech32
    assert _crypto(o, b"Send to 1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa for payment") == [("Bitcoin", b"1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa")]
    assert _crypto(o, b"Payment to 3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64 confirmed") == [("Bitcoin", b"3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64")]
    assert _crypto(o, b"Withdraw to bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq") == [("Bitcoin", b"bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq")]
    # :3300-3341: bad checksum, too short
    assert _crypto(o, b"Fake address 1A1zP1eP5QGefi2DMPTfTL5SLmv7Divf00 is invalid") == []
    assert _crypto(o, b"Short address 1A1zP1eP is invalid") == []
    # :3608-3626: chunk mode, one per line
    chunk = b"Line1: 1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa\nLine2: 3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64\n"
    assert [s for _, s in _crypto(o, chunk)] == [b"1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa", b"3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64"]
    # bech32 details of bech32 0.11 `decode`: Bech32m (BIP-350 taproot vector) accepted too; mixed case, wrong hrp, bad checksum rejected
    assert O.cryptoaddr_valid("bitcoin_bech32", b"bc1p0xlxvlhemja6c4dqv22uapctqupfhlxm9h8z3k2e72q4k9hcz7vqzk5jj0")
    assert not O.cryptoaddr_valid("bitcoin_bech32", b"bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdQ")
    assert not O.cryptoaddr_valid("bitcoin_bech32", b"bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdx")
    assert not O.cryptoaddr_valid("bitcoin_bech32", b"tb1qw508d6qejxtdg4y5r3zarvary0c5xw7kxpjzsx")
    assert _crypto(o, b"x bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq.") == []  # '.' is not a boundary: the word is longer
    # Base58Check: leading '1's are zero bytes; '0', 'O', 'I', 'l' are not in the alphabet
    assert O.cryptoaddr_valid("bitcoin_base58", b"1111111111111111111114oLvT2")
    assert not O.cryptoaddr_valid("bitcoin_base58", b"1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNO")


def test_ethereum_kats(small_dbs):
    o = O.Oracle(small_dbs[1][0])
    # lib.rs:3343-3447
    assert _crypto(o, b"Send to 0x5aeda56215b167893e80b4fe645ba6d5bab767de") == [("Ethereum", b"0x5aeda56215b167893e80b4fe645ba6d5bab767de")]
    assert _crypto(o, b"Send to 0x5aAeb6053F3E94C9b9A09f33669435E7Ef1BeAed") == [("Ethereum", b"0x5aAeb6053F3E94C9b9A09f33669435E7Ef1BeAed")]
    assert _crypto(o, b"Bad address 0x5aAeb6053f3e94c9b9a09f33669435e7ef1beaed") == []       # mixed case, wrong checksum
    assert _crypto(o, b"Short address 0x5aeda56215b167893e80b4fe645ba6d5bab7") == []
    assert _crypto(o, b"Invalid 0x5aeda56215b167893e80b4fe645ba6d5bab767dg") == []
    # :3588-3606 in a log line; all upper case needs no checksum; boundaries on both sides are required
    line = b"2025-01-15 10:32:45 Transaction to=0x5aeda56215b167893e80b4fe645ba6d5bab767de value=1000000000000000000"
    assert _crypto(o, line) == [("Ethereum", b"0x5aeda56215b167893e80b4fe645ba6d5bab767de")]
    assert _crypto(o, b"a 0x5AEDA56215B167893E80B4FE645BA6D5BAB767DE b") == [("Ethereum", b"0x5AEDA56215B167893E80B4FE645BA6D5BAB767DE")]
    assert _crypto(o, b"a x0x5aeda56215b167893e80b4fe645ba6d5bab767de b") == []
    assert _crypto(o, b"a 0x5aeda56215b167893e80b4fe645ba6d5bab767de0 b") == []
    # other EIP-55 test vectors (EIP-55 text)
    for a in (b"0xfB6916095ca1df60bB79Ce92cE3Ea74c37c5d359", b"0xdbF03B407c01E7cD3CBea99509d93f8DDDC8C6FB", b"0xD1220A0cf47c7B9Be7A2E6BA89F429762e7b9aDb"):
        assert O.cryptoaddr_valid("ethereum", a)
        assert not O.cryptoaddr_valid("ethereum", a[:-1] + (b"B" if a[-1:] == b"b" else b"b"))


def test_monero_kats(small_dbs):
    o = O.Oracle(small_dbs[1][0])
    # lib.rs:3449-3518.  The reference decodes the WHOLE string with bs58 (not Monero's block-wise base58), so the real
    # address of :3455 does not pass its Keccak check — the reference's own test tolerates that outcome; restated as is.
    addr = b"44AFFq5kSiGBoZ4NMDwYtN18obc8AemS33DBLWs3H7otXft3XjrpDtQGv7SqSsaBYBb98uNbr2VBBEt7f2wfn3RVGQBEP3A"
    assert _crypto(o, b"Donate to " + addr) == []
    assert _crypto(o, b"Fake 1AdUndXHHZ6cfufTMvppY6JwXNouMBzSkbLYfpAV5Usx3skxNgYeYTRj5UzqtReoS44qo9mtmXCqY45DJ852K5Jv2684Rge") == []
    assert _crypto(o, b"Short 4AdUndXHHZ6cfufTMvppY6JwXNouMBzSkbLYfpAV5Usx") == []
    # a string built to satisfy the reference's rule (payload ‖ Keccak256(payload)[..4], plain base58) IS extracted
    good = _monero_like(b"\x12" + bytes(range(64)))
    assert 90 <= len(good) <= 110 and good[:1] in (b"4", b"8")
    assert _crypto(o, b"to " + good + b" ok") == [("Monero", good)]
    assert _crypto(o, b"to " + good[:-1] + (b"2" if good[-1:] != b"2" else b"3") + b" ok") == []


ALPHABET58 = b"123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"


def _b58(raw: bytes) -> bytes:
    n = int.from_bytes(raw, "big")
    out = bytearray()
    while n:
        n, r = divmod(n, 58)
        out.append(ALPHABET58[r])
    out.extend(ALPHABET58[0:1] * (len(raw) - len(raw.lstrip(b"\0"))))
    return bytes(reversed(out))


def _monero_like(payload: bytes) -> bytes:
    """Search the payload's last byte until plain base58 of payload ‖ keccak[..4] starts with '4' or '8' and is 90..110 long."""
    for tail in range(256):
        for first in range(1, 256):
            p = bytes([first]) + payload[1:-1] + bytes([tail])
            s = _b58(p + O.digest("keccak256", p)[:4])
            if 90 <= len(s) <= 110 and s[:1] in (b"4", b"8"):
                return s
    raise AssertionError("no monero-like string found")


def test_crypto_mixed_and_disabled(small_dbs):
    o = O.Oracle(small_dbs[1][0])
    # lib.rs:3520-3557
    line = b"Transaction from 192.168.1.1 to bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq via example.com"
    got = o.extract_strings(line, X_ALL)
    assert ("IPv4", b"192.168.1.1") in got and ("Bitcoin", b"bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq") in got and ("Domain", b"example.com") in got
    # :3559-3586 disabled extractors find nothing
    line = b"BTC: 1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa ETH: 0x5aeda56215b167893e80b4fe645ba6d5bab767de"
    assert _crypto(o, line) == [("Bitcoin", b"1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa"), ("Ethereum", b"0x5aeda56215b167893e80b4fe645ba6d5bab767de")]
    assert [x for x in o.extract_strings(line, O.X_DEFAULT) if x[0] in ("Bitcoin", "Ethereum", "Monero")] == []


def test_tree_record_readers():
    """SearchTree::read_24bit_record / read_28bit_record with the reference's own vectors (matchy-format/src/mmdb/tree.rs:323-376),
    a 32-bit one of the same shape, and calculate_data_offset (:378-398: record - node_count - 16); the oracle's reader and
    the device code's reader (tree_record in device_fns.cuh, through the host emulation) both have to return them."""
    import ctypes as C
    import emu_lib as E
    L, M = O.lib(), E.lib()
    L.orc_tree_record.restype = C.c_int64
    L.orc_tree_record.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint32, C.c_int]
    M.emu_tree_record.restype = C.c_uint32
    M.emu_tree_record.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    d24 = bytes([0, 0, 1, 0, 0, 2]) + bytes(994)
    d28 = bytes([0, 0, 1, 0x12, 0, 0, 2]) + bytes(993)
    d32 = bytes([0x01, 0x02, 0x03, 0x04, 0xA0, 0xB0, 0xC0, 0xD0]) + bytes(992)
    cases = [(d24, 24, 60, 0, 0, 1), (d24, 24, 60, 0, 1, 2), (d28, 28, 70, 0, 0, 0x1000001), (d28, 28, 70, 0, 1, 0x2000002),
             (d32, 32, 80, 0, 0, 0x01020304), (d32, 32, 80, 0, 1, 0xA0B0C0D0), (d24, 24, 60, 1, 0, 0)]
    for data, bits, tree_size, node, side, want in cases:
        assert L.orc_tree_record(data, tree_size, 10, bits, node, side) == want
        assert M.emu_tree_record(data, 10, bits, node, side) == want
    assert L.orc_tree_record(d24, 60, 10, 24, 10, 0) == -1  # node index out of range is an error in the reference (tree.rs:133-138)
    # calculate_data_offset: a record above node_count + 16 points (record - node_count - 16) bytes into the data section
    from matchy_b200 import DatabaseBuilder
    b = DatabaseBuilder(build_epoch=1)
    b.add_entry("1.2.3.4", {"a": "first"})
    b.add_entry("5.6.7.8", {"a": "second value"})
    o = O.Oracle(b.build())
    f1, off1, _ = o.lookup_ip4(0x01020304)
    f2, off2, _ = o.lookup_ip4(0x05060708)
    assert f1 and f2 and off1 == 0 and off2 > off1
    import json
    assert json.loads(o.data_json(off1)) == {"a": "first"} and json.loads(o.data_json(off2)) == {"a": "second value"}


def _globs(patterns, mode=0):
    from matchy_b200 import DatabaseBuilder, MatchMode
    b = DatabaseBuilder(MatchMode.CaseInsensitive if mode else MatchMode.CaseSensitive, build_epoch=1)
    for p in patterns:
        b.add_glob(p, {"p": p})
    return O.Oracle(b.build())


def test_paraglob_integration_vectors():
    """matchy-paraglob/tests/integration_tests.rs:11-258, pattern sets and expectations as written there (find_all over a
    paraglob; entries without wildcards are literal-typed patterns = substring matches, quirk 12)."""
    n = lambda o, text: len(o.lookup_string(text.encode()))
    ids = lambda o, text: {p for p, _ in o.lookup_string(text.encode())}
    o = _globs(["*.txt", "test*", "*file*"])                                  # test_basic_wildcards
    assert n(o, "document.txt") and n(o, "test_case") and n(o, "myfile.dat") and not n(o, "nomatch")
    o = _globs(["hello", "world", "test"])                                    # test_exact_string_matching
    assert n(o, "hello") == 1 and n(o, "world") == 1 and n(o, "hello world") == 2 and n(o, "nomatch") == 0
    o = _globs(["*test*", "*test*", "hello", "hello"])                        # test_duplicate_pattern_deduplication
    assert n(o, "test123") == 1 and n(o, "hello") == 1
    o = _globs(["*.txt", "*file*", "test*"])                                  # test_multiple_patterns_matching_same_text
    assert ids(o, "testfile.txt") == {0, 1, 2}
    o = _globs(["Test*", "HELLO"])                                            # test_case_sensitivity
    assert n(o, "Test123") and not n(o, "test123") and n(o, "HELLO") and not n(o, "hello")
    o = _globs(["Test*", "HELLO"], mode=1)                                    # test_case_insensitivity
    assert n(o, "Test123") and n(o, "test123") and n(o, "HELLO") and n(o, "hello")
    o = _globs(["test"])                                                      # test_empty_string_queries
    assert n(o, "") == 0 and n(o, "test") == 1
    o = _globs(["exact_match", "another_literal", "third"])                   # test_pure_literal_patterns
    assert n(o, "exact_match") == 1 and n(o, "another_literal") == 1 and n(o, "nomatch") == 0
    o = _globs(["*test*", "test*", "*test"])                                  # test_overlapping_literal_patterns
    assert (n(o, "test"), n(o, "testing"), n(o, "mytest"), n(o, "mytesting")) == (3, 2, 2, 1)
    o = _globs(["*.rs", "*.toml", "Cargo.*", "src/*", "*.md"])                # test_real_world_file_patterns
    assert n(o, "main.rs") and n(o, "Cargo.toml") >= 2 and n(o, "src/lib.rs") >= 2 and n(o, "README.md") and not n(o, "test.py")
    o = _globs(["pattern_%d_*" % i for i in range(1000)])                     # test_large_pattern_set
    assert 500 in ids(o, "pattern_500_test") and 999 in ids(o, "pattern_999_data") and not n(o, "nomatch")
    o = _globs(["hello", "*.txt", "test_*"])                                  # test_combined_literal_and_glob_patterns
    assert ids(o, "hello.txt") == {0, 1} and ids(o, "test_file.txt") == {1, 2}
    o = _globs(["*", "?", "**"])                                              # test_pure_wildcard_patterns
    assert n(o, "test") >= 2 and n(o, "a") >= 3
