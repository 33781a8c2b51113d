// synth_gen.h — the deterministic log / indicator generators of the five BASELINE.json configs, as ONE piece of integer
// code that compiles for the host (synth.cpp: databases, host-side logs, tests) and for the device (synth_device.cu:
// config 5's 200 GB are generated in HBM).  The reference ships no fixtures for these shapes;
// the shapes borrow from its own bench generators (crates/matchy/src/bin/commands/bench/pattern.rs:46-110,
// bench/literal.rs:51) and its log example (examples/generate_logs.rs:47-85).
//
// Everything is counter-based: indicator i of a family is a pure function of i, so the log generator plants hits without
// holding the database, and log block b (64 KiB) is a pure function of (config, counts, b), so any byte range can be
// regenerated on any rank, on either side of the PCIe bus, with identical bytes (tests/test_abi_and_host.py,
// tests/test_gpu_parity.py::test_device_generator_equals_host_generator).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SG_HD __host__ __device__
#else
#define SG_HD
#endif

namespace sgen {

static const uint64_t SEED0 = 0x6d61746368790001ULL;
static const size_t BLOCK = 65536;

SG_HD inline uint64_t mix(uint64_t x) {  // splitmix64 finalizer
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
struct Rng {
  uint64_t s;
  SG_HD explicit Rng(uint64_t seed) : s(seed) {}
  SG_HD uint64_t next() { s += 0x9E3779B97F4A7C15ULL; return mix(s); }
  SG_HD uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
  SG_HD bool chance(uint32_t per_million) { return below(1000000) < per_million; }
};

// word tables as functions of the index (string literals are addressable on both sides)
#define SG_TABLE(name, mask, ...)                                   \
  SG_HD inline const char* name(uint32_t i) {                       \
    const char* const t[] = {__VA_ARGS__};                          \
    return t[i & (mask)];                                           \
  }
SG_TABLE(WORDS, 31, "alpha", "bravo", "cobalt", "delta", "ember", "falcon", "garnet", "harbor", "indigo", "jasper", "kernel", "lumen", "meteor", "nimbus", "onyx",
         "prism", "quartz", "raven", "sierra", "tundra", "umbra", "vector", "willow", "xenon", "yonder", "zephyr", "anchor", "beacon", "cipher", "drift", "echo", "flint")
SG_TABLE(DWS, 15, "cdn", "api", "mail", "login", "update", "static", "track", "files", "portal", "secure", "img", "ads", "sync", "auth", "push", "edge")
SG_TABLE(TLDS, 7, "com", "net", "org", "info", "biz", "io", "ru", "cn")
SG_TABLE(CATS, 15, "mal", "phish", "c2", "spam", "bot", "scan", "tor", "proxy", "miner", "drop", "exfil", "rat", "worm", "adware", "fraud", "apt")
SG_TABLE(ZONES, 15, "example", "contoso", "fabrikam", "northwind", "tailspin", "wingtip", "adatum", "litware", "proseware", "fourth", "lucerne", "trey", "woodgrove",
         "humongous", "margie", "blueyonder")
SG_TABLE(LEVELS, 3, "low", "medium", "high", "critical")
SG_TABLE(PATHS, 7, "index.html", "api/v1/items", "static/app.js", "images/logo.png", "login", "search", "assets/site.css", "download/file.zip")
SG_TABLE(UAS, 3, "Mozilla/5.0 (X11; Linux x86_64) AppleWebKit/537.36 (KHTML, like Gecko) Chrome/120.0 Safari/537.36",
         "Mozilla/5.0 (Windows NT 10.0; Win64; x64; rv:121.0) Gecko/20100101 Firefox/121.0", "curl/8.4.0",
         "Mozilla/5.0 (Macintosh; Intel Mac OS X 10_15_7) AppleWebKit/605.1.15 (KHTML, like Gecko) Version/17.1 Safari/605.1.15")
#undef SG_TABLE
SG_HD inline const char* SOURCES(uint32_t i) { const char* const t[] = {"feed-a", "feed-b", "osint", "internal", "partner"}; return t[i % 5]; }
SG_HD inline const char* MONTHS(uint32_t i) {
  const char* const t[] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
  return t[i % 12];
}

struct Out {  // bounded appender
  uint8_t* p; size_t n, cap;
  SG_HD void ch(char c) { if (n < cap) p[n] = (uint8_t)c; n++; }
  SG_HD void str(const char* s) { while (*s) ch(*s++); }
  SG_HD void num(uint64_t v) { char b[24]; int k = 0; do { b[k++] = char('0' + v % 10); v /= 10; } while (v); while (k) ch(b[--k]); }
  SG_HD void num2(uint32_t v) { ch(char('0' + v / 10 % 10)); ch(char('0' + v % 10)); }
  SG_HD void hexnum(uint32_t v) {  // printf("%x")
    char b[8]; int k = 0;
    do { b[k++] = "0123456789abcdef"[v & 15]; v >>= 4; } while (v);
    while (k) ch(b[--k]);
  }
  SG_HD void hex(uint64_t seed, int digits) {
    uint64_t x = 0;
    for (int i = 0; i < digits; i++) { if ((i & 15) == 0) x = mix(seed + (uint64_t)(i >> 4)); ch("0123456789abcdef"[(x >> ((i & 15) * 4)) & 15]); }
  }
  SG_HD void ip4(uint32_t a) { num(a >> 24); ch('.'); num((a >> 16) & 255); ch('.'); num((a >> 8) & 255); ch('.'); num(a & 255); }
};

// ---- indicator families (pure functions of the index) -----------------------------------------------------
SG_HD inline void lit_domain(Out& o, uint64_t i) {  // "{cat}-{svc}-{i}.{zone}.{tld}"
  const uint64_t h = mix(i * 2654435761ULL + 17);
  o.str(CATS((uint32_t)(h & 15))); o.ch('-'); o.str(DWS((uint32_t)((h >> 4) & 15))); o.ch('-'); o.num(i); o.ch('.');
  o.str(ZONES((uint32_t)((h >> 8) & 15))); o.ch('.'); o.str(TLDS((uint32_t)((h >> 12) & 7)));
}
SG_HD inline int glob_shape(uint64_t i) { const uint32_t r = (uint32_t)(i % 100); return r < 60 ? 0 : r < 85 ? 1 : r < 99 ? 2 : 3; }
SG_HD inline void glob_pattern(Out& o, uint64_t i) {
  const uint64_t h = mix(i * 0x9E3779B1ULL + 5);
  const char *w = WORDS((uint32_t)(h & 31)), *d = DWS((uint32_t)((h >> 5) & 15)), *t = TLDS((uint32_t)((h >> 9) & 7));
  switch (glob_shape(i)) {
    case 0: o.str("*."); o.str(w); o.ch('-'); o.str(d); o.ch('-'); o.num(i); o.ch('.'); o.str(t); break;
    case 1: o.str("*.evil-"); o.num(i); o.ch('.'); o.str(t); break;
    case 2: o.str(w); o.ch('-'); o.str(d); o.ch('-'); o.num(i); o.str(".*"); break;
    default: o.str("*[0-9].*."); o.str(w); o.str("-attack-"); o.num(i); o.ch('.'); o.str(t); break;
  }
}
SG_HD inline void glob_hit(Out& o, uint64_t i, uint64_t r) {  // a domain that pattern i matches
  const uint64_t h = mix(i * 0x9E3779B1ULL + 5);
  const char *w = WORDS((uint32_t)(h & 31)), *d = DWS((uint32_t)((h >> 5) & 15)), *t = TLDS((uint32_t)((h >> 9) & 7));
  switch (glob_shape(i)) {
    case 0: o.str(WORDS((uint32_t)(r & 31))); o.num((uint32_t)(r >> 8) % 100); o.ch('.'); o.str(w); o.ch('-'); o.str(d); o.ch('-'); o.num(i); o.ch('.'); o.str(t); break;
    case 1: o.str("www"); o.num((uint32_t)(r & 7)); o.str(".evil-"); o.num(i); o.ch('.'); o.str(t); break;
    case 2: o.str(w); o.ch('-'); o.str(d); o.ch('-'); o.num(i); o.ch('.'); o.str(ZONES((uint32_t)(r & 15))); o.ch('.'); o.str(TLDS((uint32_t)((r >> 4) & 7))); break;
    default: o.str("node"); o.num((uint32_t)(r % 10)); o.ch('.'); o.str(DWS((uint32_t)((r >> 4) & 15))); o.ch('.'); o.str(w); o.str("-attack-"); o.num(i); o.ch('.'); o.str(t); break;
  }
}
SG_HD inline int hash_digits(uint64_t i) { const uint32_t r = (uint32_t)(i % 10); return r < 4 ? 32 : r < 6 ? 40 : 64; }  // 40% MD5, 20% SHA1, 40% SHA256
SG_HD inline uint64_t hash_seed(uint64_t i) { return mix(i + 0xABCDEF12345ULL) | 1; }
SG_HD inline void hash_text(Out& o, uint64_t i) { o.hex(hash_seed(i), hash_digits(i)); }

// IPv4 prefixes: clustered under 4096 /16 parents so that the tree stays below 2^24 records (SURVEY §7 hard parts)
struct Prefix4 { uint32_t addr, plen; };
SG_HD inline Prefix4 ip4_prefix(uint64_t i, int cfg) {
  const uint64_t h = mix(i * 0xD6E8FEB86659FD93ULL + 3);
  uint32_t parent = (uint32_t)(mix((h & 4095) + 99) >> 32) & 0xFFFF0000u;
  if ((parent >> 24) == 0 || (parent >> 24) >= 224 || (parent >> 24) == 127 || (parent >> 24) == 10) parent = (parent & 0x00FF0000u) | 0x2D000000u;
  uint32_t plen = 16 + (uint32_t)((h >> 12) % 17);  // /16../32
  if (cfg == 1 && (i & 3) != 3) plen = 32;          // cfg 1: three host addresses for every CIDR
  const uint32_t low = (uint32_t)(h >> 20) & 0xFFFFu;
  uint32_t addr = parent | low;
  if (plen < 32) addr &= ~((1u << (32 - plen)) - 1u);
  return Prefix4{addr, plen};
}
struct Prefix6 { uint64_t hi; uint32_t plen; };  // the low 64 address bits are zero
SG_HD inline Prefix6 ip6_prefix(uint64_t i) {  // /32../64 under 256 /32 parents 2001:0dXX::/32 … (documentation-style space)
  const uint64_t h = mix(i * 0xA0761D6478BD642FULL + 11);
  const uint32_t plen = 32 + (uint32_t)((h >> 8) % 33);
  uint64_t hi = ((uint64_t)0x20010d00u | (h & 255)) << 32 | (uint32_t)(h >> 24);
  if (plen < 64) hi &= ~((1ULL << (64 - plen)) - 1);
  return Prefix6{hi, plen};
}

struct Counts { uint64_t ip4, ip6, lit, glob, hash; int cfg; };
SG_HD inline uint64_t scaled(uint64_t v, double scale) { const uint64_t r = (uint64_t)((double)v * scale); return r < 8 ? (v < 8 ? v : 8) : r; }
SG_HD inline Counts counts_for(int cfg, double scale) {
  switch (cfg) {
    case 1: return Counts{scaled(4000, scale), scaled(200, scale), scaled(3000, scale), scaled(800, scale), scaled(2000, scale), 1};
    case 2: return Counts{0, 0, scaled(1000000, scale), scaled(100000, scale), 0, 2};
    case 3: return Counts{scaled(850000, scale), scaled(150000, scale), 0, 0, 0, 3};
    case 4: return Counts{0, 0, 0, 0, scaled(5000000, scale), 4};
    default: return Counts{scaled(850000, scale), scaled(150000, scale), scaled(1000000, scale), scaled(100000, scale), scaled(2900000, scale), 5};
  }
}

// ---- log lines -----------------------------------------------------------------------------------------------
SG_HD inline void timestamp(Out& o, Rng& g) {  // 2025-03-14T09:26:53Z
  o.str("2025-"); o.num2(1 + g.below(12)); o.ch('-'); o.num2(1 + g.below(28)); o.ch('T'); o.num2(g.below(24)); o.ch(':');
  o.num2(g.below(60)); o.ch(':'); o.num2(g.below(60)); o.ch('Z');
}
SG_HD inline uint32_t random_public_ip(Rng& g) {
  uint32_t a = (uint32_t)g.next();
  const uint32_t top = a >> 24;
  if (top == 0 || top >= 224 || top == 127 || top == 10) a = (a & 0x00FFFFFFu) | 0x53000000u;
  return a;
}
SG_HD inline void benign_domain(Out& o, Rng& g, bool mixed_case) {  // depth 2-5, PSL-valid, not in any database
  const uint32_t depth = 2 + g.below(4);
  const size_t start = o.n;
  for (uint32_t k = 0; k + 2 < depth; k++) { o.str(DWS(g.below(16))); o.num(g.below(100)); o.ch('.'); }
  o.str(WORDS(g.below(32))); o.ch('-'); o.str(ZONES(g.below(16))); o.ch('.'); o.str(TLDS(g.below(8)));
  if (mixed_case && start < o.cap) { uint8_t& c = o.p[start]; if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32); }
}
SG_HD inline void hit_ip4(Out& o, Rng& g, const Counts& c) {  // an address inside a random database prefix
  const Prefix4 k = ip4_prefix(g.next() % c.ip4, c.cfg);
  uint32_t a = k.addr;
  if (k.plen < 32) a |= (uint32_t)g.next() & ((1u << (32 - k.plen)) - 1u);
  o.ip4(a);
}
// compressed form with "::" (the extractor only anchors on "::"): the four groups of the high 64 bits, "::", the last group | 1
SG_HD inline void ip6_text(Out& o, uint64_t hi, uint32_t last_group) {
  for (int k = 0; k < 4; k++) { o.hexnum((uint32_t)(hi >> (48 - 16 * k)) & 0xFFFFu); if (k < 3) o.ch(':'); }
  o.str("::"); o.hexnum((last_group & 0xFFFFu) | 1u);
}

SG_HD inline void line_nginx(Out& o, Rng& g, const Counts& c) {
  const bool hit = g.chance(1000);  // 0.1 % of lines
  const uint32_t what = hit ? g.below(4) : 99;
  if (what == 0 && c.ip4) hit_ip4(o, g, c); else o.ip4(random_public_ip(g));
  o.str(" - - ["); o.num2(1 + g.below(28)); o.ch('/'); o.str(MONTHS(g.below(12))); o.str("/2025:"); o.num2(g.below(24)); o.ch(':');
  o.num2(g.below(60)); o.ch(':'); o.num2(g.below(60)); o.str(" +0000] \"GET /"); o.str(PATHS(g.below(8)));
  if (what == 3 && c.hash) { o.str("?h="); hash_text(o, g.next() % c.hash); }
  o.str(" HTTP/1.1\" "); o.num(g.below(10) ? 200 : 404); o.ch(' '); o.num(200 + g.below(50000)); o.str(" \"http://");
  if (what == 1 && c.lit) lit_domain(o, g.next() % c.lit);
  else if (what == 2 && c.glob) { const uint64_t r = g.next(); const uint64_t i = g.next() % c.glob; glob_hit(o, i, r); }  // (argument order of the first generator)
  else benign_domain(o, g, false);
  o.ch('/'); o.str(PATHS(g.below(8))); o.str("\" \""); o.str(UAS(g.below(4))); o.str("\"\n");
}
SG_HD inline void line_dns(Out& o, Rng& g, const Counts& c) {
  const bool hit = g.chance(5000);  // 0.5 %
  timestamp(o, g); o.str(" dns01 client="); o.ip4(random_public_ip(g)); o.str(" query=");
  if (hit && (g.below(2) ? c.lit != 0 : c.glob == 0) && c.lit) lit_domain(o, g.next() % c.lit);
  else if (hit && c.glob) { const uint64_t r = g.next(); const uint64_t i = g.next() % c.glob; glob_hit(o, i, r); }
  else benign_domain(o, g, g.chance(20000));
  o.str(" type="); o.str(g.below(4) ? "A" : "AAAA"); o.str(" rcode=NOERROR upstream="); benign_domain(o, g, g.chance(20000));
  if (g.below(2)) { o.str(" referer=https://"); benign_domain(o, g, false); o.ch('/'); o.str(PATHS(g.below(8))); }
  o.str(" latency="); o.num(g.below(400)); o.str("ms\n");
}
SG_HD inline void line_fw(Out& o, Rng& g, const Counts& c) {
  const bool hit = g.chance(20000);  // 2 %
  const bool v6 = g.chance(100000);  // 10 %
  timestamp(o, g); o.str(" fw01 "); o.str(g.below(8) ? "ACCEPT" : "DROP"); o.str(g.below(3) ? " TCP" : " UDP"); o.str(" src=");
  if (v6) {
    if (hit && c.ip6) { const Prefix6 k = ip6_prefix(g.next() % c.ip6); (void)g.next(); ip6_text(o, k.hi, 0); }
    else { const uint64_t x = g.next(), y = g.next(); ip6_text(o, ((uint64_t)(0x2a000000u | (uint32_t)(x & 0xFFFFFF)) << 32) | (y >> 48), 0); }
    o.str(" dst="); { const uint64_t y = g.next(); ip6_text(o, ((uint64_t)0x2a001450u << 32) | (y >> 48), 0); }
  } else {
    if (hit && c.ip4) hit_ip4(o, g, c); else o.ip4(random_public_ip(g));
    o.str(" dst="); o.ip4(random_public_ip(g));
  }
  o.str(" sport="); o.num(1024 + g.below(60000)); o.str(" dport="); o.num(g.below(4) ? 443 : 1 + g.below(65000));
  o.str(" bytes="); o.num(g.below(1000000)); o.str(" iface=eth"); o.num(g.below(4)); o.str(" rule="); o.num(g.below(500)); o.ch('\n');
}
SG_HD inline void line_edr(Out& o, Rng& g, const Counts& c) {
  const bool hit = g.chance(2000);  // 0.2 %, drawn from <= 5000 distinct indicators
  timestamp(o, g); o.str(" host=ws-"); o.num(g.below(5000)); o.str(" pid="); o.num(g.below(65536)); o.str(" ppid="); o.num(g.below(65536));
  o.str(" user=u"); o.num(g.below(2000)); o.str(" image=C:\\Program Files\\"); o.str(WORDS(g.below(32))); o.ch('\\'); o.str(DWS(g.below(16)));
  o.str(".exe md5=");
  const uint64_t pick = c.hash ? (uint64_t)g.below(5000) * (c.hash / 5000 ? c.hash / 5000 : 1) % c.hash : 0;
  const int hd = c.hash ? hash_digits(pick) : 0;
  if (hit && hd == 32) hash_text(o, pick); else o.hex(g.next(), 32);
  if (g.below(4) == 0) { o.str(" sha1="); if (hit && hd == 40) hash_text(o, pick); else o.hex(g.next(), 40); }
  o.str(" sha256=");
  if (hit && hd == 64) hash_text(o, pick); else o.hex(g.next(), 64);
  o.str(" cmdline=\""); o.str(DWS(g.below(16))); o.str(".exe --config C:\\ProgramData\\"); o.str(ZONES(g.below(16))); o.str("\\settings.json --threads ");
  o.num(1 + g.below(16)); o.str("\" parent=explorer.exe integrity="); o.str(LEVELS(g.below(4))); o.str(" session="); o.num(g.below(10)); o.ch('\n');
}

// One 64 KiB block of the config's stream: whole lines, then a filler line up to the block edge ('#', blanks, newline: no
// token, no anchor).  `line` is scratch for one line (LINE_CAP bytes).
static const size_t LINE_CAP = 1024;
SG_HD inline void gen_block(int cfg, const Counts& c, uint64_t block, uint8_t* out, uint8_t* line) {
  Rng g(mix(SEED0 + (uint64_t)cfg) ^ mix(block * 0x2545F4914F6CDD1DULL + 1));
  size_t n = 0;
  const int family = cfg == 5 ? 1 + (int)(block & 3) : cfg;
  for (;;) {
    Out o{line, 0, LINE_CAP};
    switch (family) {
      case 1: line_nginx(o, g, c); break;
      case 2: line_dns(o, g, c); break;
      case 3: line_fw(o, g, c); break;
      default: line_edr(o, g, c); break;
    }
    if (o.n > LINE_CAP || n + o.n + 2 > BLOCK) break;
    for (size_t k = 0; k < o.n; k++) out[n + k] = line[k];
    n += o.n;
  }
  if (n < BLOCK) {
    out[n++] = '#';
    while (n + 1 < BLOCK) out[n++] = ' ';
    out[BLOCK - 1] = '\n';
  }
}

}  // namespace sgen
