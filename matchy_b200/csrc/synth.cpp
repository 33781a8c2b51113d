// synth.cpp — deterministic synthetic databases and logs for the five BASELINE.json configs (bench + tests).
//
// The reference ships no fixtures for these shapes; the generators borrow the pattern shapes of the
// reference's own bench generators (crates/matchy/src/bin/commands/bench/pattern.rs:46-110, bench/literal.rs:51) and
// its log example (examples/generate_logs.rs:47-85).  Everything is counter-based: indicator i of a family is a pure
// function of i, so the log generator can plant hits without holding the database, and log block b (64 KiB) is a pure
// function of (config, b), so any byte range can be regenerated on any rank.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "mxy_builder.h"

struct mxyb_builder;
extern "C" mxyb_builder* mxyb_new(int);
extern "C" void mxyb_free(mxyb_builder*);
mxy::DatabaseBuilder& mxyb_inner(mxyb_builder*);  // defined in mxy_builder.cpp

namespace {

using mxy::DataValue;
using mxy::IpKey;
typedef unsigned __int128 u128;

const uint64_t SEED0 = 0x6d61746368790001ULL;
const size_t BLOCK = 65536;

inline uint64_t mix(uint64_t x) {  // splitmix64 finalizer
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { s += 0x9E3779B97F4A7C15ULL; return mix(s); }
  uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
  bool chance(uint32_t per_million) { return below(1000000) < per_million; }
};

const char* WORDS[32] = {"alpha", "bravo", "cobalt", "delta", "ember", "falcon", "garnet", "harbor", "indigo", "jasper", "kernel",
                         "lumen", "meteor", "nimbus", "onyx", "prism", "quartz", "raven", "sierra", "tundra", "umbra", "vector",
                         "willow", "xenon", "yonder", "zephyr", "anchor", "beacon", "cipher", "drift", "echo", "flint"};
const char* DWS[16] = {"cdn", "api", "mail", "login", "update", "static", "track", "files", "portal", "secure", "img", "ads", "sync", "auth", "push", "edge"};
const char* TLDS[8] = {"com", "net", "org", "info", "biz", "io", "ru", "cn"};
const char* CATS[16] = {"mal", "phish", "c2", "spam", "bot", "scan", "tor", "proxy", "miner", "drop", "exfil", "rat", "worm", "adware", "fraud", "apt"};
const char* ZONES[16] = {"example", "contoso", "fabrikam", "northwind", "tailspin", "wingtip", "adatum", "litware", "proseware", "fourth", "lucerne",
                         "trey", "woodgrove", "humongous", "margie", "blueyonder"};
const char* LEVELS[4] = {"low", "medium", "high", "critical"};
const char* SOURCES[5] = {"feed-a", "feed-b", "osint", "internal", "partner"};
const char* MONTHS[12] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
const char* UAS[4] = {"Mozilla/5.0 (X11; Linux x86_64) AppleWebKit/537.36 (KHTML, like Gecko) Chrome/120.0 Safari/537.36",
                      "Mozilla/5.0 (Windows NT 10.0; Win64; x64; rv:121.0) Gecko/20100101 Firefox/121.0",
                      "curl/8.4.0", "Mozilla/5.0 (Macintosh; Intel Mac OS X 10_15_7) AppleWebKit/605.1.15 (KHTML, like Gecko) Version/17.1 Safari/605.1.15"};
const char* PATHS[8] = {"index.html", "api/v1/items", "static/app.js", "images/logo.png", "login", "search", "assets/site.css", "download/file.zip"};

struct Out {  // bounded appender
  uint8_t* p; size_t n, cap;
  void ch(char c) { if (n < cap) p[n] = (uint8_t)c; n++; }
  void str(const char* s) { while (*s) ch(*s++); }
  void str(const std::string& s) { for (char c : s) ch(c); }
  void num(uint64_t v) { char b[24]; int k = 0; do { b[k++] = char('0' + v % 10); v /= 10; } while (v); while (k) ch(b[--k]); }
  void num2(uint32_t v) { ch(char('0' + v / 10 % 10)); ch(char('0' + v % 10)); }
  void hex(uint64_t seed, int digits) {
    static const char* H = "0123456789abcdef";
    uint64_t x = 0;
    for (int i = 0; i < digits; i++) { if ((i & 15) == 0) x = mix(seed + (uint64_t)(i >> 4)); ch(H[(x >> ((i & 15) * 4)) & 15]); }
  }
  void ip4(uint32_t a) { num(a >> 24); ch('.'); num((a >> 16) & 255); ch('.'); num((a >> 8) & 255); ch('.'); num(a & 255); }
};

// ---- indicator families (pure functions of the index) -----------------------------------------------------
std::string lit_domain(uint64_t i) {  // "{cat}-{svc}-{i}.{zone}.{tld}"
  uint64_t h = mix(i * 2654435761ULL + 17);
  char b[96];
  snprintf(b, sizeof b, "%s-%s-%llu.%s.%s", CATS[h & 15], DWS[(h >> 4) & 15], (unsigned long long)i, ZONES[(h >> 8) & 15], TLDS[(h >> 12) & 7]);
  return b;
}
int glob_shape(uint64_t i) { uint32_t r = (uint32_t)(i % 100); return r < 60 ? 0 : r < 85 ? 1 : r < 99 ? 2 : 3; }
std::string glob_pattern(uint64_t i) {
  uint64_t h = mix(i * 0x9E3779B1ULL + 5);
  char b[128];
  const char *w = WORDS[h & 31], *d = DWS[(h >> 5) & 15], *t = TLDS[(h >> 9) & 7];
  switch (glob_shape(i)) {
    case 0: snprintf(b, sizeof b, "*.%s-%s-%llu.%s", w, d, (unsigned long long)i, t); break;
    case 1: snprintf(b, sizeof b, "*.evil-%llu.%s", (unsigned long long)i, t); break;
    case 2: snprintf(b, sizeof b, "%s-%s-%llu.*", w, d, (unsigned long long)i); break;
    default: snprintf(b, sizeof b, "*[0-9].*.%s-attack-%llu.%s", w, (unsigned long long)i, t); break;
  }
  return b;
}
std::string glob_hit(uint64_t i, uint64_t r) {  // a domain that pattern i matches
  uint64_t h = mix(i * 0x9E3779B1ULL + 5);
  char b[160];
  const char *w = WORDS[h & 31], *d = DWS[(h >> 5) & 15], *t = TLDS[(h >> 9) & 7];
  switch (glob_shape(i)) {
    case 0: snprintf(b, sizeof b, "%s%u.%s-%s-%llu.%s", WORDS[r & 31], (unsigned)(r >> 8) % 100, w, d, (unsigned long long)i, t); break;
    case 1: snprintf(b, sizeof b, "www%u.evil-%llu.%s", (unsigned)(r & 7), (unsigned long long)i, t); break;
    case 2: snprintf(b, sizeof b, "%s-%s-%llu.%s.%s", w, d, (unsigned long long)i, ZONES[r & 15], TLDS[(r >> 4) & 7]); break;
    default: snprintf(b, sizeof b, "node%u.%s.%s-attack-%llu.%s", (unsigned)(r % 10), DWS[(r >> 4) & 15], w, (unsigned long long)i, t); break;
  }
  return b;
}
int hash_digits(uint64_t i) { uint32_t r = (uint32_t)(i % 10); return r < 4 ? 32 : r < 6 ? 40 : 64; }  // 40% MD5, 20% SHA1, 40% SHA256
uint64_t hash_seed(uint64_t i) { return mix(i + 0xABCDEF12345ULL) | 1; }
std::string hash_text(uint64_t i) {
  std::string s((size_t)hash_digits(i), '0');
  Out o{(uint8_t*)&s[0], 0, s.size()};
  o.hex(hash_seed(i), hash_digits(i));
  return s;
}
// IPv4 prefixes: clustered under 4096 /16 parents so that the tree stays below 2^24 records (SURVEY §7 hard parts)
IpKey ip4_prefix(uint64_t i, int cfg) {
  uint64_t h = mix(i * 0xD6E8FEB86659FD93ULL + 3);
  uint32_t parent = (uint32_t)(mix((h & 4095) + 99) >> 32) & 0xFFFF0000u;
  if ((parent >> 24) == 0 || (parent >> 24) >= 224 || (parent >> 24) == 127 || (parent >> 24) == 10) parent = (parent & 0x00FF0000u) | 0x2D000000u;
  uint32_t plen = 16 + (uint32_t)((h >> 12) % 17);  // /16../32
  if (cfg == 1 && (i & 3) != 3) plen = 32;          // cfg 1: three host addresses for every CIDR
  uint32_t low = (uint32_t)(h >> 20) & 0xFFFFu;
  uint32_t addr = parent | low;
  if (plen < 32) addr &= ~((1u << (32 - plen)) - 1u);
  IpKey k; k.bits = addr; k.v6 = false; k.prefix = (uint8_t)plen;
  return k;
}
IpKey ip6_prefix(uint64_t i) {  // /32../64 under 256 /32 parents 2001:0dXX::/32 … (documentation-style space)
  uint64_t h = mix(i * 0xA0761D6478BD642FULL + 11);
  uint32_t plen = 32 + (uint32_t)((h >> 8) % 33);
  uint64_t hi = ((uint64_t)0x20010d00u | (h & 255)) << 32 | (uint32_t)(h >> 24);
  if (plen < 64) hi &= ~((1ULL << (64 - plen)) - 1);
  IpKey k; k.bits = (u128)hi << 64; k.v6 = true; k.prefix = (uint8_t)plen;
  return k;
}

struct Counts { uint64_t ip4, ip6, lit, glob, hash; int cfg; };
Counts counts_for(int cfg, double scale) {
  auto sc = [&](uint64_t v) { uint64_t r = (uint64_t)((double)v * scale); return r < 8 ? std::min<uint64_t>(v, 8) : r; };
  switch (cfg) {
    case 1: return {sc(4000), sc(200), sc(3000), sc(800), sc(2000), 1};
    case 2: return {0, 0, sc(1000000), sc(100000), 0, 2};
    case 3: return {sc(850000), sc(150000), 0, 0, 0, 3};
    case 4: return {0, 0, 0, 0, sc(5000000), 4};
    default: return {sc(850000), sc(150000), sc(1000000), sc(100000), sc(2900000), 5};
  }
}

DataValue meta_row(uint64_t i) {  // ~20 distinct rows: dedup of the data section is exercised
  uint32_t r = (uint32_t)(mix(i + 77) % 20);
  DataValue m = DataValue::Map();
  m.map["threat_level"] = DataValue::String(LEVELS[r & 3]);
  m.map["category"] = DataValue::String(CATS[(r * 7) & 15]);
  m.map["source"] = DataValue::String(SOURCES[r % 5]);
  return m;
}

// ---- log lines -----------------------------------------------------------------------------------------------
void timestamp(Out& o, Rng& g) {  // 2025-03-14T09:26:53Z
  o.str("2025-"); o.num2(1 + g.below(12)); o.ch('-'); o.num2(1 + g.below(28)); o.ch('T'); o.num2(g.below(24)); o.ch(':');
  o.num2(g.below(60)); o.ch(':'); o.num2(g.below(60)); o.ch('Z');
}
uint32_t random_public_ip(Rng& g) {
  uint32_t a = (uint32_t)g.next();
  uint32_t top = a >> 24;
  if (top == 0 || top >= 224 || top == 127 || top == 10) a = (a & 0x00FFFFFFu) | 0x53000000u;
  return a;
}
void benign_domain(Out& o, Rng& g, bool mixed_case) {  // depth 2-5, PSL-valid, not in any database
  uint32_t depth = 2 + g.below(4);
  size_t start = o.n;
  for (uint32_t k = 0; k + 2 < depth; k++) { o.str(DWS[g.below(16)]); o.num(g.below(100)); o.ch('.'); }
  o.str(WORDS[g.below(32)]); o.ch('-'); o.str(ZONES[g.below(16)]); o.ch('.'); o.str(TLDS[g.below(8)]);
  if (mixed_case && start < o.cap) { uint8_t& c = o.p[start]; if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32); }
}
void hit_ip4(Out& o, Rng& g, const Counts& c) {  // an address inside a random database prefix
  IpKey k = ip4_prefix(g.next() % c.ip4, c.cfg);
  uint32_t a = (uint32_t)k.bits;
  if (k.prefix < 32) a |= (uint32_t)g.next() & ((1u << (32 - k.prefix)) - 1u);
  o.ip4(a);
}
void ip6_text(Out& o, u128 bits) {  // compressed form with "::" (the extractor only anchors on "::")
  uint16_t s[8];
  for (int k = 0; k < 8; k++) s[k] = (uint16_t)(bits >> (112 - 16 * k));
  char b[8];
  for (int k = 0; k < 4; k++) { snprintf(b, sizeof b, "%x", s[k]); o.str(b); if (k < 3) o.ch(':'); }
  o.str("::"); snprintf(b, sizeof b, "%x", (unsigned)s[7] | 1u); o.str(b);
}

void line_nginx(Out& o, Rng& g, const Counts& c, int cfg) {
  bool hit = g.chance(1000);  // 0.1 % of lines
  uint32_t what = hit ? g.below(4) : 99;
  if (what == 0 && c.ip4) hit_ip4(o, g, c); else o.ip4(random_public_ip(g));
  o.str(" - - ["); o.num2(1 + g.below(28)); o.ch('/'); o.str(MONTHS[g.below(12)]); o.str("/2025:"); o.num2(g.below(24)); o.ch(':');
  o.num2(g.below(60)); o.ch(':'); o.num2(g.below(60)); o.str(" +0000] \"GET /"); o.str(PATHS[g.below(8)]);
  if (what == 3 && c.hash) { o.str("?h="); o.str(hash_text(g.next() % c.hash)); }
  o.str(" HTTP/1.1\" "); o.num(g.below(10) ? 200 : 404); o.ch(' '); o.num(200 + g.below(50000)); o.str(" \"http://");
  if (what == 1 && c.lit) o.str(lit_domain(g.next() % c.lit));
  else if (what == 2 && c.glob) o.str(glob_hit(g.next() % c.glob, g.next()));
  else benign_domain(o, g, false);
  o.ch('/'); o.str(PATHS[g.below(8)]); o.str("\" \""); o.str(UAS[g.below(4)]); o.str("\"\n");
  (void)cfg;
}
void line_dns(Out& o, Rng& g, const Counts& c) {
  bool hit = g.chance(5000);  // 0.5 %
  timestamp(o, g); o.str(" dns01 client="); o.ip4(random_public_ip(g)); o.str(" query=");
  if (hit && (g.below(2) ? c.lit != 0 : c.glob == 0) && c.lit) o.str(lit_domain(g.next() % c.lit));
  else if (hit && c.glob) o.str(glob_hit(g.next() % c.glob, g.next()));
  else benign_domain(o, g, g.chance(20000));
  o.str(" type="); o.str(g.below(4) ? "A" : "AAAA"); o.str(" rcode=NOERROR upstream="); benign_domain(o, g, g.chance(20000));
  if (g.below(2)) { o.str(" referer=https://"); benign_domain(o, g, false); o.ch('/'); o.str(PATHS[g.below(8)]); }
  o.str(" latency="); o.num(g.below(400)); o.str("ms\n");
}
void line_fw(Out& o, Rng& g, const Counts& c) {
  bool hit = g.chance(20000);  // 2 %
  bool v6 = g.chance(100000);  // 10 %
  timestamp(o, g); o.str(" fw01 "); o.str(g.below(8) ? "ACCEPT" : "DROP"); o.str(g.below(3) ? " TCP" : " UDP"); o.str(" src=");
  if (v6) {
    if (hit && c.ip6) { IpKey k = ip6_prefix(g.next() % c.ip6); ip6_text(o, k.bits | ((u128)(g.next() & 0xFFFF) << 16)); }
    else ip6_text(o, ((u128)(0x2a000000u | (uint32_t)(g.next() & 0xFFFFFF)) << 96) | ((u128)g.next() << 16));
    o.str(" dst="); ip6_text(o, ((u128)0x2a001450u << 96) | ((u128)g.next() << 16));
  } else {
    if (hit && c.ip4) hit_ip4(o, g, c); else o.ip4(random_public_ip(g));
    o.str(" dst="); o.ip4(random_public_ip(g));
  }
  o.str(" sport="); o.num(1024 + g.below(60000)); o.str(" dport="); o.num(g.below(4) ? 443 : 1 + g.below(65000));
  o.str(" bytes="); o.num(g.below(1000000)); o.str(" iface=eth"); o.num(g.below(4)); o.str(" rule="); o.num(g.below(500)); o.ch('\n');
}
void line_edr(Out& o, Rng& g, const Counts& c) {
  bool hit = g.chance(2000);  // 0.2 %, drawn from <= 5000 distinct indicators
  timestamp(o, g); o.str(" host=ws-"); o.num(g.below(5000)); o.str(" pid="); o.num(g.below(65536)); o.str(" ppid="); o.num(g.below(65536));
  o.str(" user=u"); o.num(g.below(2000)); o.str(" image=C:\\Program Files\\"); o.str(WORDS[g.below(32)]); o.ch('\\'); o.str(DWS[g.below(16)]);
  o.str(".exe md5=");
  uint64_t pick = c.hash ? (uint64_t)g.below(5000) * (c.hash / 5000 ? c.hash / 5000 : 1) % c.hash : 0;
  int hd = c.hash ? hash_digits(pick) : 0;
  if (hit && hd == 32) o.str(hash_text(pick)); else o.hex(g.next(), 32);
  if (g.below(4) == 0) { o.str(" sha1="); if (hit && hd == 40) o.str(hash_text(pick)); else o.hex(g.next(), 40); }
  o.str(" sha256=");
  if (hit && hd == 64) o.str(hash_text(pick)); else o.hex(g.next(), 64);
  o.str(" cmdline=\""); o.str(DWS[g.below(16)]); o.str(".exe --config C:\\ProgramData\\"); o.str(ZONES[g.below(16)]); o.str("\\settings.json --threads ");
  o.num(1 + g.below(16)); o.str("\" parent=explorer.exe integrity="); o.str(LEVELS[g.below(4)]); o.str(" session="); o.num(g.below(10)); o.ch('\n');
}

void gen_block(int cfg, const Counts& c, uint64_t block, uint8_t* out) {
  Rng g(mix(SEED0 + (uint64_t)cfg) ^ mix(block * 0x2545F4914F6CDD1DULL + 1));
  size_t n = 0;
  uint8_t tmp[1024];
  for (;;) {
    Out o{tmp, 0, sizeof tmp};
    int family = cfg;
    if (cfg == 5) family = 1 + (int)(block & 3);
    switch (family) {
      case 1: line_nginx(o, g, c, cfg); break;
      case 2: line_dns(o, g, c); break;
      case 3: line_fw(o, g, c); break;
      default: line_edr(o, g, c); break;
    }
    if (o.n > sizeof tmp || n + o.n + 2 > BLOCK) break;
    memcpy(out + n, tmp, o.n);
    n += o.n;
  }
  // filler line up to the block edge: '#', blanks, newline (no token, no anchor)
  if (n < BLOCK) {
    out[n++] = '#';
    while (n + 1 < BLOCK) out[n++] = ' ';
    out[BLOCK - 1] = '\n';
  }
}

}  // namespace

extern "C" {

mxyb_builder* mgen_db(int cfg, double scale) {
  if (cfg < 1 || cfg > 5 || !(scale > 0)) return nullptr;
  Counts c = counts_for(cfg, scale);
  mxyb_builder* hb = mxyb_new(0);
  mxy::DatabaseBuilder& b = mxyb_inner(hb);
  b.set_build_epoch(1760000000ULL + (uint64_t)cfg);
  std::vector<uint32_t> offs(20);
  for (uint32_t r = 0; r < 20; r++) {
    // encode each distinct row once; entries reuse the offsets (what encode_and_deduplicate_data achieves)
    uint64_t i = 0;
    while ((uint32_t)(mix(i + 77) % 20) != r) i++;
    offs[r] = b.encode_data(meta_row(i));
  }
  auto off_of = [&](uint64_t i) { return offs[(uint32_t)(mix(i + 77) % 20)]; };
  // duplicate prefixes are skipped: with two data values for one prefix the reference's unstable sort decides which one
  // survives (mmdb_builder.rs:485-487), so a duplicate-free list is the only input with one well-defined tree
  std::unordered_set<uint64_t> seen4;
  for (uint64_t i = 0; i < c.ip4; i++) {
    IpKey k = ip4_prefix(i, cfg);
    if (!seen4.insert(((uint64_t)(uint32_t)k.bits << 8) | k.prefix).second) continue;
    b.add_ip_raw(k, off_of(i));
  }
  std::unordered_set<uint64_t> seen6;
  for (uint64_t i = 0; i < c.ip6; i++) {
    IpKey k = ip6_prefix(i);
    if (!seen6.insert((uint64_t)(k.bits >> 64) ^ ((uint64_t)k.prefix << 56)).second) continue;
    b.add_ip_raw(k, off_of(i + 1000003));
  }
  for (uint64_t i = 0; i < c.lit; i++) b.add_literal(lit_domain(i), off_of(i + 2000003));
  for (uint64_t i = 0; i < c.glob; i++) b.add_glob(glob_pattern(i), off_of(i + 3000017));
  for (uint64_t i = 0; i < c.hash; i++) b.add_literal(hash_text(i), off_of(i + 4000037));
  return hb;
}

int mgen_log(int cfg, double scale, uint64_t offset, uint8_t* out, size_t len, int threads) {
  if (cfg < 1 || cfg > 5 || !(scale > 0) || offset % BLOCK != 0) return -3;
  Counts c = counts_for(cfg, scale);
  size_t nblocks = (len + BLOCK - 1) / BLOCK;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(nblocks, 1));
  uint64_t b0 = offset / BLOCK;
  auto work = [&](int t) {
    std::vector<uint8_t> tmp(BLOCK);
    for (size_t k = (size_t)t; k < nblocks; k += (size_t)threads) {
      size_t o = k * BLOCK;
      if (o + BLOCK <= len) gen_block(cfg, c, b0 + k, out + o);
      else {
        gen_block(cfg, c, b0 + k, tmp.data());
        size_t m = len - o;
        memcpy(out + o, tmp.data(), m);
        size_t e = m;  // cut at the last complete line, blank the rest
        while (e > 0 && out[o + e - 1] != '\n') e--;
        for (size_t j = e; j < m; j++) out[o + j] = ' ';
        out[o + m - 1] = '\n';
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  return 0;
}

}  // extern "C"
