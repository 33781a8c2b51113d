// mxy_builder.cpp — `.mxy` writer.  See mxy_builder.h for the reference files this restates.
#include "mxy_builder.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <ctime>
#include <deque>
#include <unordered_set>

namespace mxy {

// ---------------------------------------------------------------------------------------------
// XXH64 (public specification; xxhash-rust 0.8.15 `xxh64`, seed 0 — matchy-literal-hash/src/lib.rs:666-671)
// ---------------------------------------------------------------------------------------------
static const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL,
                      P4 = 9650029242287828579ULL, P5 = 2870177450012600261ULL;
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xxround(uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl64(acc, 31); return acc * P1; }
static inline uint64_t xxmerge(uint64_t acc, uint64_t v) { v = xxround(0, v); acc ^= v; return acc * P1 + P4; }

uint64_t xxh64(const uint8_t* p, size_t len, uint64_t seed) {
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t* lim = end - 32;
    do {
      v1 = xxround(v1, rd64(p)); v2 = xxround(v2, rd64(p + 8));
      v3 = xxround(v3, rd64(p + 16)); v4 = xxround(v4, rd64(p + 24));
      p += 32;
    } while (p <= lim);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xxmerge(h, v1); h = xxmerge(h, v2); h = xxmerge(h, v3); h = xxmerge(h, v4);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) { h ^= xxround(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
  if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
  while (p < end) { h ^= (*p) * P5; h = rotl64(h, 11) * P1; p++; }
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

static void put32(std::vector<uint8_t>& b, uint32_t v) { for (int k = 0; k < 4; k++) b.push_back(uint8_t(v >> (8 * k))); }
static void put16(std::vector<uint8_t>& b, uint16_t v) { b.push_back(uint8_t(v)); b.push_back(uint8_t(v >> 8)); }
static void put64(std::vector<uint8_t>& b, uint64_t v) { for (int k = 0; k < 8; k++) b.push_back(uint8_t(v >> (8 * k))); }
static void set32(std::vector<uint8_t>& b, size_t off, uint32_t v) { for (int k = 0; k < 4; k++) b[off + k] = uint8_t(v >> (8 * k)); }

// ---------------------------------------------------------------------------------------------
// Rust std::net text parsers (core::net::parser)
// ---------------------------------------------------------------------------------------------
static bool read_v4(const char* s, size_t n, size_t& pos, uint32_t& out) {
  size_t p = pos; uint32_t addr = 0;
  for (int k = 0; k < 4; k++) {
    if (k > 0) { if (p >= n || s[p] != '.') return false; p++; }
    size_t st = p; uint32_t v = 0;
    while (p < n && s[p] >= '0' && s[p] <= '9') {
      v = v * 10 + uint32_t(s[p] - '0'); p++;
      if (p - st > 3) return false;
    }
    size_t d = p - st;
    if (d == 0 || v > 255 || (d > 1 && s[st] == '0')) return false;
    addr = (addr << 8) | v;
  }
  pos = p; out = addr;
  return true;
}

bool parse_ipv4_text(const char* s, size_t n, uint32_t& out) {
  size_t pos = 0;
  return read_v4(s, n, pos, out) && pos == n;
}

static size_t v6_groups(const char* s, size_t n, size_t& pos, uint16_t* groups, size_t limit, bool& v4) {
  v4 = false;
  for (size_t i = 0; i < limit; i++) {
    if (i + 1 < limit) {  // embedded IPv4 needs two groups
      size_t p = pos; bool ok = true;
      if (i > 0) { if (p < n && s[p] == ':') p++; else ok = false; }
      uint32_t a;
      if (ok && read_v4(s, n, p, a)) {
        groups[i] = uint16_t(a >> 16); groups[i + 1] = uint16_t(a);
        pos = p; v4 = true;
        return i + 2;
      }
    }
    size_t p = pos;
    if (i > 0) { if (p < n && s[p] == ':') p++; else return i; }
    uint32_t v = 0; int digits = 0;
    while (p < n) {
      char c = s[p]; int d;
      if (c >= '0' && c <= '9') d = c - '0';
      else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
      else break;
      v = v * 16 + uint32_t(d); digits++; p++;
      if (digits > 4) return i;
    }
    if (digits == 0) return i;
    groups[i] = uint16_t(v);
    pos = p;
  }
  return limit;
}

bool parse_ipv6_text(const char* s, size_t n, uint16_t out[8]) {
  size_t pos = 0; bool v4 = false;
  uint16_t head[8] = {0};
  size_t hs = v6_groups(s, n, pos, head, 8, v4);
  if (hs == 8) { if (pos != n) return false; memcpy(out, head, 16); return true; }
  if (v4) return false;
  if (!(pos + 1 < n && s[pos] == ':' && s[pos + 1] == ':')) return false;
  pos += 2;
  uint16_t tail[7] = {0};
  size_t limit = 8 - (hs + 1);
  size_t ts = limit ? v6_groups(s, n, pos, tail, limit, v4) : 0;
  if (pos != n) return false;
  for (size_t k = 0; k < ts; k++) head[8 - ts + k] = tail[k];
  memcpy(out, head, 16);
  return true;
}

// ---------------------------------------------------------------------------------------------
// DataEncoder — matchy-data-format/src/lib.rs:257-623
// ---------------------------------------------------------------------------------------------
static void enc_size(uint8_t type_id, size_t size, std::vector<uint8_t>& b) {  // :605-622
  uint8_t tb = uint8_t(type_id << 5);
  if (size < 29) b.push_back(tb | uint8_t(size));
  else if (size < 29 + 256) { b.push_back(tb | 29); b.push_back(uint8_t(size - 29)); }
  else if (size < 29 + 256 + 65536) { b.push_back(tb | 30); size_t a = size - 29 - 256; b.push_back(uint8_t(a >> 8)); b.push_back(uint8_t(a)); }
  else { b.push_back(tb | 31); size_t a = size - 29 - 256 - 65536; b.push_back(uint8_t(a >> 16)); b.push_back(uint8_t(a >> 8)); b.push_back(uint8_t(a)); }
}
static void enc_pointer(uint32_t off, std::vector<uint8_t>& b) {  // :374-428
  if (off < 2048) { b.push_back(0x20 | uint8_t((off >> 8) & 7)); b.push_back(uint8_t(off)); }
  else if (off < 2048 + 524288) { uint32_t a = off - 2048; b.push_back(0x20 | (1 << 3) | uint8_t((a >> 16) & 7)); b.push_back(uint8_t(a >> 8)); b.push_back(uint8_t(a)); }
  else if (off < 2048 + 524288 + 134217728) { uint32_t a = off - 526336; b.push_back(0x20 | (2 << 3) | uint8_t((a >> 24) & 7)); b.push_back(uint8_t(a >> 16)); b.push_back(uint8_t(a >> 8)); b.push_back(uint8_t(a)); }
  else { b.push_back(0x20 | (3 << 3)); for (int k = 3; k >= 0; k--) b.push_back(uint8_t(off >> (8 * k))); }
}
static void enc_string(const std::string& s, std::vector<uint8_t>& b) { enc_size(2, s.size(), b); b.insert(b.end(), s.begin(), s.end()); }
static void enc_array_head(size_t size, std::vector<uint8_t>& b) {  // :526-548 (size bytes, then the extended type byte)
  if (size < 29) b.push_back(uint8_t(size));
  else if (size < 29 + 256) { b.push_back(29); b.push_back(uint8_t(size - 29)); }
  else if (size < 29 + 256 + 65536) { b.push_back(30); size_t a = size - 29 - 256; b.push_back(uint8_t(a >> 8)); b.push_back(uint8_t(a)); }
  else { b.push_back(31); size_t a = size - 29 - 256 - 65536; b.push_back(uint8_t(a >> 16)); b.push_back(uint8_t(a >> 8)); b.push_back(uint8_t(a)); }
  b.push_back(0x04);
}
static void enc_scalar(const DataValue& v, std::vector<uint8_t>& b) {
  switch (v.type) {
    case DataValue::DOUBLE: { b.push_back(0x68); uint64_t bits; memcpy(&bits, &v.dbl, 8); for (int k = 7; k >= 0; k--) b.push_back(uint8_t(bits >> (8 * k))); break; }
    case DataValue::BYTES: enc_size(4, v.str.size(), b); b.insert(b.end(), v.str.begin(), v.str.end()); break;
    case DataValue::UINT16: b.push_back(0xA2); b.push_back(uint8_t(v.u >> 8)); b.push_back(uint8_t(v.u)); break;
    case DataValue::UINT32: b.push_back(0xC4); for (int k = 3; k >= 0; k--) b.push_back(uint8_t(v.u >> (8 * k))); break;
    case DataValue::INT32: { b.push_back(0x04); b.push_back(0x01); uint32_t x = (uint32_t)v.i32; for (int k = 3; k >= 0; k--) b.push_back(uint8_t(x >> (8 * k))); break; }
    case DataValue::UINT64: b.push_back(0x08); b.push_back(0x02); for (int k = 7; k >= 0; k--) b.push_back(uint8_t(v.u >> (8 * k))); break;
    case DataValue::UINT128: b.push_back(0x10); b.push_back(0x03); for (int k = 15; k >= 0; k--) b.push_back(uint8_t(v.big >> (8 * k))); break;
    case DataValue::BOOL: b.push_back(v.u ? 0x01 : 0x00); b.push_back(0x07); break;
    case DataValue::FLOAT: { b.push_back(0x04); b.push_back(0x08); uint32_t bits; memcpy(&bits, &v.flt, 4); for (int k = 3; k >= 0; k--) b.push_back(uint8_t(bits >> (8 * k))); break; }
    default: break;
  }
}

void DataEncoder::encode_plain(const DataValue& v, std::vector<uint8_t>& b) {  // encode_to_buffer :355-371
  switch (v.type) {
    case DataValue::STRING: enc_string(v.str, b); break;
    case DataValue::MAP:
      enc_size(7, v.map.size(), b);
      for (auto& kv : v.map) { enc_string(kv.first, b); encode_plain(kv.second, b); }
      break;
    case DataValue::ARRAY:
      enc_array_head(v.arr.size(), b);
      for (auto& e : v.arr) encode_plain(e, b);
      break;
    default: enc_scalar(v, b);
  }
}

void DataEncoder::intern_string(const std::string& s) {
  auto it = strings_.find(s);
  if (it != strings_.end()) enc_pointer(it->second, buf_);
  else { uint32_t off = (uint32_t)buf_.size(); enc_string(s, buf_); strings_[s] = off; }
}

void DataEncoder::encode_interned(const DataValue& v) {  // :333-352, 462-486, 526-554
  switch (v.type) {
    case DataValue::STRING: intern_string(v.str); break;
    case DataValue::MAP:
      enc_size(7, v.map.size(), buf_);
      for (auto& kv : v.map) { intern_string(kv.first); encode_interned(kv.second); }
      break;
    case DataValue::ARRAY:
      enc_array_head(v.arr.size(), buf_);
      for (auto& e : v.arr) encode_interned(e);
      break;
    default: enc_scalar(v, buf_);
  }
}

uint32_t DataEncoder::encode(const DataValue& v) {  // :294-318
  std::vector<uint8_t> tmp;
  encode_plain(v, tmp);
  std::string key((const char*)tmp.data(), tmp.size());
  auto it = dedup_.find(key);
  if (it != dedup_.end()) return it->second;
  uint32_t off = (uint32_t)buf_.size();
  encode_interned(v);
  dedup_[key] = off;
  return off;
}

// ---------------------------------------------------------------------------------------------
// glob syntax — matchy-paraglob/src/glob.rs:307-451, paraglob_offset.rs:93-159
// ---------------------------------------------------------------------------------------------
static bool next_char(const std::string& s, size_t& pos, uint32_t& cp, size_t& len) {  // str::chars() on valid UTF-8
  if (pos >= s.size()) return false;
  uint8_t c = (uint8_t)s[pos];
  if (c < 0x80) { cp = c; len = 1; }
  else if (c < 0xE0) { cp = ((c & 0x1F) << 6) | (uint8_t(s[pos + 1]) & 0x3F); len = 2; }
  else if (c < 0xF0) { cp = ((c & 0x0F) << 12) | ((uint8_t(s[pos + 1]) & 0x3F) << 6) | (uint8_t(s[pos + 2]) & 0x3F); len = 3; }
  else { cp = ((c & 0x07) << 18) | ((uint8_t(s[pos + 1]) & 0x3F) << 12) | ((uint8_t(s[pos + 2]) & 0x3F) << 6) | (uint8_t(s[pos + 3]) & 0x3F); len = 4; }
  pos += len;
  return true;
}

bool glob_is_glob(const std::string& p) {
  bool esc = false; size_t pos = 0; uint32_t c; size_t l;
  while (next_char(p, pos, c, l)) {
    if (esc) { esc = false; continue; }
    if (c == '\\') esc = true;
    else if (c == '*' || c == '?' || c == '[') return true;
  }
  return false;
}

std::vector<std::string> glob_extract_literals(const std::string& p) {
  std::vector<std::string> lits; std::string cur;
  size_t pos = 0; uint32_t c; size_t l; bool esc = false;
  while (true) {
    size_t at = pos;
    if (!next_char(p, pos, c, l)) break;
    if (esc) { cur.append(p, at, l); esc = false; continue; }
    if (c == '\\') esc = true;
    else if (c == '*' || c == '?') { if (!cur.empty()) { lits.push_back(cur); cur.clear(); } }
    else if (c == '[') {
      if (!cur.empty()) { lits.push_back(cur); cur.clear(); }
      int depth = 1;
      while (next_char(p, pos, c, l)) {
        if (c == '\\') next_char(p, pos, c, l);
        else if (c == '[') depth++;
        else if (c == ']') { if (--depth == 0) break; }
      }
    } else cur.append(p, at, l);
  }
  if (!cur.empty()) lits.push_back(cur);
  return lits;
}

bool parse_glob(const std::string& p, std::vector<GlobSeg>& out, std::string& err) {
  std::vector<GlobSeg> segs; std::string lit;
  auto flush = [&]() { if (!lit.empty()) { GlobSeg g; g.type = GlobSeg::LITERAL; g.lit = lit; segs.push_back(g); lit.clear(); } };
  size_t pos = 0; uint32_t c; size_t l;
  while (true) {
    size_t at = pos;
    if (!next_char(p, pos, c, l)) break;
    if (c == '*') { flush(); GlobSeg g; g.type = GlobSeg::STAR; segs.push_back(g); }
    else if (c == '?') { flush(); GlobSeg g; g.type = GlobSeg::QUESTION; segs.push_back(g); }
    else if (c == '[') {
      flush();
      GlobSeg g; g.type = GlobSeg::CLASS;
      {  // negation
        size_t pp = pos; uint32_t nc; size_t nl;
        if (next_char(p, pp, nc, nl) && (nc == '!' || nc == '^')) { g.negated = true; pos = pp; }
      }
      bool have_prev = false, expect_end = false; uint32_t prev = 0;
      for (;;) {
        uint32_t cc; size_t cl;
        if (!next_char(p, pos, cc, cl)) { err = "Unclosed character class"; return false; }
        if (cc == ']' && (!g.items.empty() || have_prev)) {
          if (have_prev) g.items.push_back({false, prev, 0});
          break;
        }
        bool peek_some = pos < p.size();
        bool peek_close = peek_some && p[pos] == ']';
        if (cc == '-' && have_prev && peek_some && !peek_close) expect_end = true;
        else if (expect_end) {
          if (prev > cc) { err = "Invalid character range"; return false; }
          g.items.push_back({true, prev, cc});
          have_prev = false; expect_end = false;
        } else {
          if (have_prev) g.items.push_back({false, prev, 0});
          prev = cc; have_prev = true;
        }
      }
      if (g.items.empty()) { err = "Empty character class"; return false; }
      segs.push_back(g);
    } else if (c == '\\') {
      size_t a2 = pos;
      if (!next_char(p, pos, c, l)) { err = "Trailing backslash in pattern"; return false; }
      lit.append(p, a2, l);
    } else lit.append(p, at, l);
  }
  flush();
  // optimize_segments: merge consecutive literals
  out.clear();
  std::string buf;
  for (auto& s : segs) {
    if (s.type == GlobSeg::LITERAL) buf += s.lit;
    else {
      if (!buf.empty()) { GlobSeg g; g.type = GlobSeg::LITERAL; g.lit = buf; out.push_back(g); buf.clear(); }
      out.push_back(s);
    }
  }
  if (!buf.empty()) { GlobSeg g; g.type = GlobSeg::LITERAL; g.lit = buf; out.push_back(g); }
  return true;
}

// ---------------------------------------------------------------------------------------------
// entry typing — mmdb_builder.rs:338-429
// ---------------------------------------------------------------------------------------------
static bool parse_ip_addr(const std::string& s, IpKey& k) {
  uint32_t v4; uint16_t v6[8];
  if (parse_ipv4_text(s.data(), s.size(), v4)) { k.bits = v4; k.v6 = false; return true; }
  if (parse_ipv6_text(s.data(), s.size(), v6)) {
    u128 b = 0; for (int i = 0; i < 8; i++) b = (b << 16) | v6[i];
    k.bits = b; k.v6 = true; return true;
  }
  return false;
}

bool DatabaseBuilder::parse_ip_entry(const std::string& key, IpKey& out) {
  if (parse_ip_addr(key, out)) { out.prefix = out.v6 ? 128 : 32; return true; }
  size_t slash = key.find('/');
  if (slash != std::string::npos) {
    std::string a = key.substr(0, slash), pstr = key.substr(slash + 1);
    // str::parse::<u8>: optional '+', then 1+ decimal digits, value <= 255
    size_t i = 0; if (!pstr.empty() && pstr[0] == '+') i = 1;
    bool ok = i < pstr.size(); uint32_t pv = 0;
    for (; ok && i < pstr.size(); i++) { if (pstr[i] < '0' || pstr[i] > '9') ok = false; else { pv = pv * 10 + uint32_t(pstr[i] - '0'); if (pv > 255) ok = false; } }
    if (ok && parse_ip_addr(a, out)) {
      uint32_t maxp = out.v6 ? 128 : 32;
      if (pv <= maxp) { out.prefix = (uint8_t)pv; return true; }
    }
  }
  return false;
}

bool DatabaseBuilder::add_ip(const std::string& s, uint32_t data_offset) {
  IpKey k;
  if (!parse_ip_entry(s, k)) { error_ = "Invalid IP address or CIDR: " + s; return false; }
  ips_.push_back({k, data_offset});
  return true;
}

bool DatabaseBuilder::add_entry_at(const std::string& key, uint32_t data_offset) {
  std::string err;
  std::vector<GlobSeg> segs;
  if (key.compare(0, 8, "literal:") == 0) { add_literal(key.substr(8), data_offset); return true; }
  if (key.compare(0, 5, "glob:") == 0) {
    std::string g = key.substr(5);
    if (!parse_glob(g, segs, err)) { error_ = "Invalid glob pattern syntax: " + err; return false; }
    add_glob(g, data_offset);
    return true;
  }
  if (key.compare(0, 3, "ip:") == 0) return add_ip(key.substr(3), data_offset);
  IpKey k;
  if (parse_ip_entry(key, k)) { ips_.push_back({k, data_offset}); return true; }
  if (key.find('*') != std::string::npos || key.find('?') != std::string::npos || key.find('[') != std::string::npos) {
    if (parse_glob(key, segs, err)) { add_glob(key, data_offset); return true; }
  }
  add_literal(key, data_offset);
  return true;
}

bool DatabaseBuilder::add_entry(const std::string& key, const DataValue& data_map) {
  // the reference types the key first (errors leave the data section untouched), then encodes the data
  std::string err; std::vector<GlobSeg> segs; IpKey k;
  if (key.compare(0, 5, "glob:") == 0 && !parse_glob(key.substr(5), segs, err)) { error_ = "Invalid glob pattern syntax: " + err; return false; }
  if (key.compare(0, 3, "ip:") == 0 && !parse_ip_entry(key.substr(3), k)) { error_ = "Invalid IP address or CIDR: " + key.substr(3); return false; }
  return add_entry_at(key, data_.encode(data_map));
}

// ---------------------------------------------------------------------------------------------
// IP tree — matchy-ip-trie/src/lib.rs:142-546
// ---------------------------------------------------------------------------------------------
namespace {
struct Ptr { uint8_t kind = 0; uint8_t plen = 0; uint32_t v = 0; };  // 0 Empty, 1 Node(v), 2 Data(v, plen)
struct TNode { Ptr l, r; };
struct Tree {
  std::vector<TNode> nodes;
  bool v6;
  explicit Tree(bool is6) : v6(is6) { nodes.emplace_back(); }
  uint32_t alloc() { nodes.emplace_back(); return (uint32_t)nodes.size() - 1; }
  void backfill(uint32_t id, uint32_t off, uint8_t plen) {  // :333-380 (explicit stack instead of recursion)
    std::vector<uint32_t> st{id};
    while (!st.empty()) {
      uint32_t n = st.back(); st.pop_back();
      for (int side = 0; side < 2; side++) {
        Ptr& p = side ? nodes[n].r : nodes[n].l;
        if (p.kind == 0) { p.kind = 2; p.v = off; p.plen = plen; }
        else if (p.kind == 2) { if (plen > p.plen) { p.v = off; p.plen = plen; } }
        else st.push_back(p.v);
      }
    }
  }
  void insert_bits(u128 bits, uint8_t plen, uint32_t off) {  // :185-310
    uint32_t node = 0;
    for (unsigned depth = 0; depth < plen; depth++) {
      int bit = int((bits >> (127 - depth)) & 1);
      Ptr child = bit ? nodes[node].r : nodes[node].l;
      auto set = [&](Ptr np) { if (bit) nodes[node].r = np; else nodes[node].l = np; };
      if (depth + 1 == plen) {
        if (child.kind == 0) set(Ptr{2, plen, off});
        else if (child.kind == 2) { if (plen >= child.plen) set(Ptr{2, plen, off}); }
        else backfill(child.v, off, plen);
        return;
      }
      if (child.kind == 0) { uint32_t id = alloc(); set(Ptr{1, 0, id}); node = id; }
      else if (child.kind == 1) node = child.v;
      else {
        uint32_t id = alloc();
        nodes[id].l = child; nodes[id].r = child;
        set(Ptr{1, 0, id});
        node = id;
      }
    }
  }
  void insert(const IpKey& k, uint32_t off) {  // :142-182
    if (!k.v6) {
      if (v6) insert_bits(k.bits, uint8_t(96 + k.prefix), off);
      else insert_bits(k.bits << 96, k.prefix, off);
    } else insert_bits(k.bits, k.prefix, off);
  }
  void serialize(int record_bits, std::vector<uint8_t>& out) const {  // :385-546
    uint32_t nc = (uint32_t)nodes.size();
    size_t nb = record_bits == 24 ? 6 : record_bits == 28 ? 7 : 8;
    out.assign((size_t)nc * nb, 0);
    auto val = [&](const Ptr& p) -> uint32_t { return p.kind == 0 ? nc : p.kind == 1 ? p.v : nc + 16 + p.v; };
    for (size_t i = 0; i < nodes.size(); i++) {
      uint32_t l = val(nodes[i].l), r = val(nodes[i].r);
      uint8_t* t = out.data() + i * nb;
      if (record_bits == 24) {
        t[0] = uint8_t(l >> 16); t[1] = uint8_t(l >> 8); t[2] = uint8_t(l);
        t[3] = uint8_t(r >> 16); t[4] = uint8_t(r >> 8); t[5] = uint8_t(r);
      } else if (record_bits == 28) {
        t[0] = uint8_t(l >> 16); t[1] = uint8_t(l >> 8); t[2] = uint8_t(l);
        t[3] = uint8_t((((l >> 24) & 15) << 4) | ((r >> 24) & 15));
        t[4] = uint8_t(r >> 16); t[5] = uint8_t(r >> 8); t[6] = uint8_t(r);
      } else {
        t[0] = uint8_t(l >> 24); t[1] = uint8_t(l >> 16); t[2] = uint8_t(l >> 8); t[3] = uint8_t(l);
        t[4] = uint8_t(r >> 24); t[5] = uint8_t(r >> 16); t[6] = uint8_t(r >> 8); t[7] = uint8_t(r);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Aho-Corasick — matchy-ac/src/lib.rs:201-516
// ---------------------------------------------------------------------------------------------
struct AcState {
  std::vector<std::pair<uint8_t, uint32_t>> tr;  // kept sorted by byte
  uint32_t failure = 0;
  std::vector<uint32_t> outputs;
  uint32_t get(uint8_t ch) const {
    for (auto& e : tr) { if (e.first == ch) return e.second; if (e.first > ch) break; }
    return 0xFFFFFFFFu;
  }
  void put(uint8_t ch, uint32_t to) {
    auto it = tr.begin();
    while (it != tr.end() && it->first < ch) ++it;
    tr.insert(it, {ch, to});
  }
};

static std::string ascii_lower(const std::string& s) {  // ASCII only; non-ASCII + CaseInsensitive is out of scope
  std::string r = s;
  for (auto& c : r) if (c >= 'A' && c <= 'Z') c = char(c + 32);
  return r;
}

static void build_ac(const std::vector<std::string>& lits, MatchMode mode, std::vector<uint8_t>& buf, uint32_t& node_count) {
  std::vector<AcState> st(1);
  for (size_t id = 0; id < lits.size(); id++) {  // add_pattern :201-235
    std::string p = mode == MatchMode::CaseInsensitive ? ascii_lower(lits[id]) : lits[id];
    uint32_t cur = 0;
    for (unsigned char ch : p) {
      uint32_t nx = st[cur].get(ch);
      if (nx == 0xFFFFFFFFu) { nx = (uint32_t)st.size(); st.emplace_back(); st[cur].put(ch, nx); }
      cur = nx;
    }
    st[cur].outputs.push_back((uint32_t)id);
  }
  {  // build_failure_links :237-301
    std::deque<uint32_t> q;
    for (auto& e : st[0].tr) { st[e.second].failure = 0; q.push_back(e.second); }
    while (!q.empty()) {
      uint32_t s = q.front(); q.pop_front();
      for (size_t ei = 0; ei < st[s].tr.size(); ei++) {
        uint8_t ch = st[s].tr[ei].first; uint32_t nx = st[s].tr[ei].second;
        q.push_back(nx);
        uint32_t fail = st[s].failure; bool found = false;
        while (fail != 0) {
          uint32_t t = st[fail].get(ch);
          if (t != 0xFFFFFFFFu) { st[nx].failure = t; found = true; break; }
          fail = st[fail].failure;
        }
        if (!found) {
          uint32_t t = st[0].get(ch);
          st[nx].failure = (t != 0xFFFFFFFFu && t != nx) ? t : 0;
        }
        uint32_t suf = st[nx].failure;
        while (suf != 0) {
          if (!st[suf].outputs.empty()) { std::vector<uint32_t> so = st[suf].outputs; st[nx].outputs.insert(st[nx].outputs.end(), so.begin(), so.end()); }
          suf = st[suf].failure;
        }
      }
    }
  }
  // serialize :304-516
  size_t n = st.size();
  node_count = (uint32_t)n;
  size_t nodes_size = n * 20, sparse_edges = 0, dense_count = 0, total_patterns = 0;
  for (auto& s : st) {
    size_t k = s.tr.size();
    if (k >= 2 && k <= 8) sparse_edges += k; else if (k >= 9) dense_count++;
    total_patterns += s.outputs.size();
  }
  size_t edges_start = nodes_size, edges_size = sparse_edges * 8;
  size_t unaligned_dense = edges_start + edges_size;
  size_t dense_pad = dense_count ? (64 - unaligned_dense % 64) % 64 : 0;
  size_t dense_start = unaligned_dense + dense_pad;
  size_t patterns_start = dense_start + dense_count * 1024;
  size_t total = patterns_start + total_patterns * 4;
  buf.assign(total, 0);
  size_t eo = edges_start, dn = dense_start, po = patterns_start;
  auto w32 = [&](size_t off, uint32_t v) { memcpy(buf.data() + off, &v, 4); };
  for (size_t i = 0; i < n; i++) {
    const AcState& s = st[i];
    size_t k = s.tr.size();
    uint8_t kind = k == 0 ? 0 : k == 1 ? 1 : k <= 8 ? 2 : 3;
    uint32_t edges_offset = 0, one_target = 0; uint8_t one_char = 0;
    if (kind == 1) { one_char = s.tr[0].first; one_target = s.tr[0].second * 20; edges_offset = one_target; }
    else if (kind == 2) {
      edges_offset = (uint32_t)eo;
      for (auto& e : s.tr) { buf[eo] = e.first; w32(eo + 4, e.second * 20); eo += 8; }
    } else if (kind == 3) {
      edges_offset = (uint32_t)dn;
      for (auto& e : s.tr) w32(dn + (size_t)e.first * 4, e.second * 20);
      dn += 1024;
    }
    uint32_t patterns_offset = s.outputs.empty() ? 0 : (uint32_t)po;
    for (uint32_t pid : s.outputs) { w32(po, pid); po += 4; }
    uint8_t* nd = buf.data() + i * 20;
    nd[0] = kind; nd[1] = one_char;
    nd[2] = kind == 1 ? 0 : (uint8_t)std::min<size_t>(k, 255);
    nd[3] = (uint8_t)std::min<size_t>(s.outputs.size(), 255);
    w32(i * 20 + 4, one_target);
    w32(i * 20 + 8, s.failure * 20);
    w32(i * 20 + 12, edges_offset);
    w32(i * 20 + 16, patterns_offset);
  }
}

// rustc-hash 2.1.1 FxHasher on a u32 (matchy-paraglob/src/literal_hash.rs:95-99).  The crate is absent from
// the reference tree; this is a best-effort restatement used ONLY to place ACLH entries the way the Rust
// builder would.  Our own readers never hash literal ids (they scan all slots into a dense index), so a
// mismatch here can only affect whether the Rust reader finds entries in files written by this builder.
static uint64_t fx_hash_u32(uint32_t id) {
  uint64_t h = (uint64_t)id * 0xf1357aea2e62a9c5ULL;
  return (h << 26) | (h >> 38);
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// Paraglob section — paraglob_offset.rs:277-298, 368-522, 524-888 ; literal_hash.rs:121-200
// ---------------------------------------------------------------------------------------------
bool DatabaseBuilder::build_paraglob(std::vector<uint8_t>& section) {
  struct Pat { std::string s; int kind; std::vector<std::string> lits; };  // kind 0 Literal, 1 Glob, 2 PureWildcard
  std::vector<Pat> pats;
  std::unordered_map<std::string, uint32_t> seen;
  std::vector<uint32_t> mapping;  // one data offset per glob ENTRY (duplicates included, mmdb_builder.rs:521-544)
  for (auto& g : globs_) {
    mapping.push_back(g.data_offset);
    if (seen.count(g.s)) continue;
    if (g.s.empty()) { error_ = "Empty pattern"; return false; }
    Pat p; p.s = g.s;
    if (glob_is_glob(g.s)) { p.lits = glob_extract_literals(g.s); p.kind = p.lits.empty() ? 2 : 1; }
    else p.kind = 0;
    seen[g.s] = (uint32_t)pats.size();
    pats.push_back(std::move(p));
  }
  // AC literals (dedup, first-seen order) and literal → pattern ids
  std::vector<std::string> ac_lits;
  std::unordered_map<std::string, uint32_t> lit_id;
  std::vector<std::vector<uint32_t>> lit_pats;
  auto add_lit = [&](const std::string& l, uint32_t pid) {
    auto it = lit_id.find(l);
    uint32_t id;
    if (it == lit_id.end()) { id = (uint32_t)ac_lits.size(); lit_id[l] = id; ac_lits.push_back(l); lit_pats.emplace_back(); }
    else id = it->second;
    lit_pats[id].push_back(pid);
  };
  for (size_t i = 0; i < pats.size(); i++) {
    if (pats[i].kind == 0) add_lit(pats[i].s, (uint32_t)i);
    else if (pats[i].kind == 1) for (auto& l : pats[i].lits) { if (l.size() < 3) continue; add_lit(l, (uint32_t)i); }
  }
  std::vector<uint8_t> ac; uint32_t ac_nodes = 0;
  if (!ac_lits.empty()) build_ac(ac_lits, mode_, ac, ac_nodes);

  // glob segments
  std::vector<uint8_t> seg_index, seg_headers, seg_strings, seg_classes;
  struct H { uint8_t type, flags; uint32_t len, off; };
  std::vector<H> hdrs;
  std::vector<std::pair<uint32_t, uint16_t>> index;  // (first header idx, count)
  for (auto& p : pats) {
    std::vector<GlobSeg> segs; std::string err;
    if (!parse_glob(p.s, segs, err)) { error_ = "Invalid glob pattern: " + err; return false; }
    index.push_back({(uint32_t)hdrs.size(), (uint16_t)segs.size()});
    for (auto& s : segs) {
      if (s.type == GlobSeg::LITERAL) { hdrs.push_back({0, 0, (uint32_t)s.lit.size(), (uint32_t)seg_strings.size()}); seg_strings.insert(seg_strings.end(), s.lit.begin(), s.lit.end()); }
      else if (s.type == GlobSeg::STAR) hdrs.push_back({1, 0, 0, 0});
      else if (s.type == GlobSeg::QUESTION) hdrs.push_back({2, 0, 0, 0});
      else {
        hdrs.push_back({3, uint8_t(s.negated ? 1 : 0), (uint32_t)(s.items.size() * 12), (uint32_t)seg_classes.size()});
        for (auto& it : s.items) { seg_classes.push_back(it.range ? 1 : 0); seg_classes.push_back(0); seg_classes.push_back(0); seg_classes.push_back(0); put32(seg_classes, it.a); put32(seg_classes, it.range ? it.b : 0); }
      }
    }
  }

  // ACLH
  std::vector<uint8_t> aclh;
  if (!ac_lits.empty()) {
    size_t nlit = ac_lits.size();
    size_t table_size = std::max<size_t>((nlit * 5 + 3) / 4, 16);
    std::vector<uint8_t> lists; std::vector<uint32_t> list_off(nlit);
    for (size_t i = 0; i < nlit; i++) { list_off[i] = (uint32_t)lists.size(); for (uint32_t pid : lit_pats[i]) put32(lists, pid); }
    std::vector<uint32_t> slots(table_size, 0xFFFFFFFFu);
    for (size_t i = 0; i < nlit; i++) {
      size_t slot = (size_t)(fx_hash_u32((uint32_t)i) % table_size);
      while (slots[slot] != 0xFFFFFFFFu) slot = (slot + 1) % table_size;
      slots[slot] = (uint32_t)i;
    }
    aclh.insert(aclh.end(), {'A', 'C', 'L', 'H'});
    put32(aclh, 1); put32(aclh, (uint32_t)nlit); put32(aclh, (uint32_t)table_size);
    put32(aclh, (uint32_t)(24 + table_size * 16)); put32(aclh, (uint32_t)lists.size());
    for (size_t s = 0; s < table_size; s++) {
      uint32_t id = slots[s];
      if (id == 0xFFFFFFFFu) { put32(aclh, id); put32(aclh, 0); put32(aclh, 0); put32(aclh, 0); }
      else { put32(aclh, id); put32(aclh, list_off[id]); put32(aclh, (uint32_t)lit_pats[id].size()); put32(aclh, 0); }
    }
    aclh.insert(aclh.end(), lists.begin(), lists.end());
  }

  // layout (build_internal_v3)
  auto align = [](size_t x, size_t a) { return x + (a - x % a) % a; };
  size_t ac_start = align(112, 64);
  size_t patterns_start = align(ac_start + ac.size(), 8);
  size_t strings_start = patterns_start + pats.size() * 16;
  std::vector<uint8_t> strs; std::vector<uint32_t> str_off;
  for (auto& p : pats) { str_off.push_back((uint32_t)strs.size()); strs.insert(strs.end(), p.s.begin(), p.s.end()); strs.push_back(0); }
  size_t wild_start = align(strings_start + strs.size(), 8);
  std::vector<uint32_t> wild;
  for (size_t i = 0; i < pats.size(); i++) if (pats[i].kind == 2) wild.push_back((uint32_t)i);
  size_t data_start = wild_start + wild.size() * 8;  // no inline data: DatabaseBuilder calls add_pattern without data
  size_t mappings_start = align(data_start, 4);
  size_t aclh_start = mappings_start;
  size_t glob_start = align(aclh_start + aclh.size(), 8);
  size_t index_size = index.size() * 8, headers_size = hdrs.size() * 12;
  size_t glob_index_end = glob_start + index_size;
  size_t glob_size = index_size + headers_size + seg_strings.size() + seg_classes.size();
  size_t total = glob_start + glob_size;

  std::vector<uint8_t> b(total, 0);
  memcpy(b.data(), "PARAGLOB", 8);
  set32(b, 8, 5); set32(b, 12, mode_ == MatchMode::CaseInsensitive ? 1 : 0);
  set32(b, 16, ac_nodes); set32(b, 20, (uint32_t)ac_start); set32(b, 24, (uint32_t)ac.size()); set32(b, 28, 0);
  set32(b, 32, (uint32_t)pats.size()); set32(b, 36, (uint32_t)patterns_start);
  set32(b, 40, (uint32_t)strings_start); set32(b, 44, (uint32_t)strs.size());
  set32(b, 60, (uint32_t)wild.size()); set32(b, 64, (uint32_t)total);
  b[68] = 0x01;
  set32(b, 96, (uint32_t)aclh_start); set32(b, 100, (uint32_t)ac_lits.size());
  set32(b, 104, (uint32_t)glob_start); set32(b, 108, (uint32_t)glob_size);
  if (!ac.empty()) memcpy(b.data() + ac_start, ac.data(), ac.size());
  for (size_t i = 0; i < pats.size(); i++) {
    size_t eo = patterns_start + i * 16;
    set32(b, eo, (uint32_t)i); b[eo + 4] = pats[i].kind == 0 ? 0 : 1;
    set32(b, eo + 8, (uint32_t)(strings_start + str_off[i])); set32(b, eo + 12, (uint32_t)pats[i].s.size());
  }
  memcpy(b.data() + strings_start, strs.data(), strs.size());
  for (size_t i = 0; i < wild.size(); i++) { set32(b, wild_start + i * 8, wild[i]); set32(b, wild_start + i * 8 + 4, (uint32_t)(strings_start + str_off[wild[i]])); }
  if (!aclh.empty()) memcpy(b.data() + aclh_start, aclh.data(), aclh.size());
  size_t hdr_base = glob_index_end, str_base = hdr_base + headers_size, cls_base = str_base + seg_strings.size();
  for (size_t i = 0; i < index.size(); i++) {
    size_t o = glob_start + i * 8;
    set32(b, o, (uint32_t)(hdr_base + (size_t)index[i].first * 12));
    b[o + 4] = uint8_t(index[i].second); b[o + 5] = uint8_t(index[i].second >> 8);
  }
  for (size_t i = 0; i < hdrs.size(); i++) {
    size_t o = hdr_base + i * 12;
    b[o] = hdrs[i].type; b[o + 1] = hdrs[i].flags;
    set32(b, o + 4, hdrs[i].len);
    uint32_t off = 0;
    if (hdrs[i].len > 0) off = (uint32_t)((hdrs[i].type == 0 ? str_base : cls_base) + hdrs[i].off);
    set32(b, o + 8, off);
  }
  if (!seg_strings.empty()) memcpy(b.data() + str_base, seg_strings.data(), seg_strings.size());
  if (!seg_classes.empty()) memcpy(b.data() + cls_base, seg_classes.data(), seg_classes.size());

  // pattern section: [total][paraglob_size][paraglob][count][data offsets]  (mmdb_builder.rs:529-550)
  section.clear();
  put32(section, 0); put32(section, (uint32_t)b.size());
  section.insert(section.end(), b.begin(), b.end());
  put32(section, (uint32_t)mapping.size());
  for (uint32_t off : mapping) put32(section, off);
  set32(section, 0, (uint32_t)section.size());
  return true;
}

// ---------------------------------------------------------------------------------------------
// Literal hash — matchy-literal-hash/src/lib.rs:160-354, 589-663
// ---------------------------------------------------------------------------------------------
bool DatabaseBuilder::build_literal_hash(std::vector<uint8_t>& out) {
  size_t n = literals_.size();
  uint32_t shard_bits = n < 10000 ? 4 : n < 100000 ? 5 : 6;
  size_t num_shards = (size_t)1 << shard_bits;
  struct E { uint32_t idx; uint64_t hash; };
  std::vector<std::vector<E>> buckets(num_shards);
  std::vector<std::string> norm(n);
  for (size_t i = 0; i < n; i++) {
    norm[i] = mode_ == MatchMode::CaseInsensitive ? ascii_lower(literals_[i].s) : literals_[i].s;
    uint64_t h = xxh64((const uint8_t*)norm[i].data(), norm[i].size(), 0);
    buckets[(size_t)(h % num_shards)].push_back({(uint32_t)i, h});
  }
  struct Slot { uint64_t hash; uint32_t soff, pid; };
  std::vector<Slot> table; std::vector<uint8_t> pool;
  std::vector<uint32_t> shard_offsets(num_shards + 1, 0);
  for (size_t s = 0; s < num_shards; s++) {
    shard_offsets[s] = (uint32_t)table.size();
    auto& es = buckets[s];
    if (es.empty()) continue;
    size_t needed = (size_t)std::ceil((double)es.size() / 0.60);
    size_t cap = 16; while (cap < needed) cap <<= 1;
    size_t base = table.size();
    table.resize(base + cap, Slot{0, 0xFFFFFFFFu, 0});
    // every entry's string goes into the pool; a repeated hash keeps only the LAST entry in the table
    // (FxHashMap::insert overwrites, :636-639)
    std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> last;
    std::vector<uint64_t> order;
    for (auto& e : es) {
      uint32_t so = (uint32_t)pool.size();
      put16(pool, (uint16_t)norm[e.idx].size());
      pool.insert(pool.end(), norm[e.idx].begin(), norm[e.idx].end());
      pool.push_back(0);
      if (!last.count(e.hash)) order.push_back(e.hash);
      last[e.hash] = {so, e.idx};
    }
    for (uint64_t h : order) {
      size_t pos = (size_t)h & (cap - 1);
      while (table[base + pos].soff != 0xFFFFFFFFu) pos = (pos + 1) & (cap - 1);
      table[base + pos] = Slot{h, last[h].first, last[h].second};
    }
  }
  shard_offsets[num_shards] = (uint32_t)table.size();
  size_t entry_count = 0;
  for (auto& t : table) if (t.soff != 0xFFFFFFFFu) entry_count++;
  size_t strings_offset = 32 + (num_shards + 1) * 4 + table.size() * 16;
  out.clear();
  out.insert(out.end(), {'L', 'H', 'S', 'H'});
  put32(out, 1); put32(out, (uint32_t)entry_count); put32(out, (uint32_t)table.size());
  put32(out, (uint32_t)strings_offset); put32(out, (uint32_t)pool.size());
  put32(out, (uint32_t)num_shards); put32(out, shard_bits);
  for (uint32_t o : shard_offsets) put32(out, o);
  for (auto& t : table) { put64(out, t.hash); put32(out, t.soff); put32(out, t.pid); }
  out.insert(out.end(), pool.begin(), pool.end());
  put32(out, (uint32_t)n);
  for (size_t i = 0; i < n; i++) { put32(out, (uint32_t)i); put32(out, literals_[i].data_offset); }
  return true;
}

// ---------------------------------------------------------------------------------------------
// build — mmdb_builder.rs:432-760
// ---------------------------------------------------------------------------------------------
bool DatabaseBuilder::build(std::vector<uint8_t>& db) {
  std::vector<uint8_t> tree_bytes; uint32_t node_count; int record_bits = 24; int ip_version = 4;
  if (!ips_.empty()) {
    bool needs_v6 = false;
    for (auto& e : ips_) needs_v6 |= e.k.v6;
    size_t est = ips_.size();
    record_bits = est > 200000000 ? 32 : est > 15000000 ? 28 : 24;
    std::vector<IpEntry> sorted = ips_;
    std::stable_sort(sorted.begin(), sorted.end(), [](const IpEntry& a, const IpEntry& b) {
      if (a.k.prefix != b.k.prefix) return a.k.prefix > b.k.prefix;   // more specific first
      if (a.k.v6 != b.k.v6) return !a.k.v6;                           // IpAddr: V4 < V6
      return a.k.bits < b.k.bits;
    });
    Tree t(needs_v6);
    t.nodes.reserve(est + est / 2);
    for (auto& e : sorted) t.insert(e.k, e.data_offset);
    node_count = (uint32_t)t.nodes.size();
    uint64_t max_rec = (uint64_t)node_count + 16 + data_.bytes().size();
    if (record_bits < 32 && max_rec >= (1ull << record_bits)) {
      // the reference would silently truncate records here (matchy-ip-trie/src/lib.rs:467-475) and write a corrupt tree
      error_ = "IP tree records overflow the " + std::to_string(record_bits) + "-bit record size chosen from the entry count";
      return false;
    }
    t.serialize(record_bits, tree_bytes);
    ip_version = needs_v6 ? 6 : 4;
  } else {
    Tree t(false);
    node_count = 1;
    t.serialize(24, tree_bytes);
  }
  std::vector<uint8_t> glob_section, literal_section;
  bool has_globs = !globs_.empty(), has_literals = !literals_.empty();
  if (has_globs && !build_paraglob(glob_section)) return false;
  if (has_literals && !build_literal_hash(literal_section)) return false;

  db.clear();
  db.insert(db.end(), tree_bytes.begin(), tree_bytes.end());
  db.insert(db.end(), 16, 0);
  db.insert(db.end(), data_.bytes().begin(), data_.bytes().end());
  size_t pad = 0;
  if (has_globs) { pad = (4 - (db.size() + 16) % 4) % 4; db.insert(db.end(), pad, 0); }
  size_t tree_sep = tree_bytes.size() + 16, data_size = data_.bytes().size();
  uint32_t pattern_offset = has_globs ? (uint32_t)(tree_sep + data_size + pad + 16) : 0;
  uint32_t literal_offset = 0;
  if (has_literals) literal_offset = has_globs ? (uint32_t)(tree_sep + data_size + pad + 16 + glob_section.size() + 16) : (uint32_t)(tree_sep + data_size + 16);

  DataValue meta = DataValue::Map();
  meta.map["binary_format_major_version"] = DataValue::Uint16(2);
  meta.map["binary_format_minor_version"] = DataValue::Uint16(0);
  meta.map["build_epoch"] = DataValue::Uint64(have_epoch_ ? build_epoch_ : (uint64_t)time(nullptr));
  std::string db_type = have_type_ ? database_type_
                        : (has_globs || has_literals) ? (!ips_.empty() ? "Paraglob-Combined-IP-Pattern" : "Paraglob-Pattern")
                                                      : "Paraglob-IP";
  meta.map["database_type"] = DataValue::String(db_type);
  DataValue desc = DataValue::Map();
  if (description_.empty()) desc.map["en"] = DataValue::String("Paraglob unified database with IP and pattern matching");
  else for (auto& kv : description_) desc.map[kv.first] = DataValue::String(kv.second);
  meta.map["description"] = desc;
  DataValue langs = DataValue::Array(); langs.arr.push_back(DataValue::String("en"));
  meta.map["languages"] = langs;
  meta.map["ip_version"] = DataValue::Uint16((uint16_t)ip_version);
  meta.map["node_count"] = DataValue::Uint32(node_count);
  meta.map["record_size"] = DataValue::Uint16((uint16_t)record_bits);
  meta.map["ip_entry_count"] = DataValue::Uint32((uint32_t)ips_.size());
  meta.map["literal_entry_count"] = DataValue::Uint32((uint32_t)literals_.size());
  meta.map["glob_entry_count"] = DataValue::Uint32((uint32_t)globs_.size());
  meta.map["match_mode"] = DataValue::Uint16(mode_ == MatchMode::CaseInsensitive ? 1 : 0);
  meta.map["pattern_section_offset"] = DataValue::Uint32(pattern_offset);
  meta.map["literal_section_offset"] = DataValue::Uint32(literal_offset);
  DataEncoder menc;
  menc.encode(meta);

  if (has_globs) { const char m[] = "MMDB_PATTERN\0\0\0"; db.insert(db.end(), m, m + 16); db.insert(db.end(), glob_section.begin(), glob_section.end()); }
  if (has_literals) { const char m[] = "MMDB_LITERAL\0\0\0"; db.insert(db.end(), m, m + 16); db.insert(db.end(), literal_section.begin(), literal_section.end()); }
  const uint8_t marker[14] = {0xAB, 0xCD, 0xEF, 'M', 'a', 'x', 'M', 'i', 'n', 'd', '.', 'c', 'o', 'm'};
  db.insert(db.end(), marker, marker + 14);
  db.insert(db.end(), menc.bytes().begin(), menc.bytes().end());
  return true;
}

}  // namespace mxy

// =============================================================================================
// C ABI (declared in include/matchy_b200.h)
// =============================================================================================
using namespace mxy;

struct mxyb_builder {
  DatabaseBuilder b;
  DataValue cur = DataValue::Map();
  std::vector<uint8_t> out;
  std::string err;
};

mxy::DatabaseBuilder& mxyb_inner(mxyb_builder* h) { return h->b; }  // for the synthetic generators (synth.cpp)

extern "C" {

mxyb_builder* mxyb_new(int case_insensitive) {
  auto* h = new mxyb_builder();
  h->b.set_mode(case_insensitive ? MatchMode::CaseInsensitive : MatchMode::CaseSensitive);
  return h;
}
void mxyb_free(mxyb_builder* h) { delete h; }
const char* mxyb_error(mxyb_builder* h) { h->err = h->b.error(); return h->err.c_str(); }

// flat data map under construction (CSV-row shaped metadata)
void mxyb_data_begin(mxyb_builder* h) { h->cur = DataValue::Map(); }
void mxyb_data_str(mxyb_builder* h, const char* k, const char* v, size_t vlen) { h->cur.map[k] = DataValue::String(std::string(v, vlen)); }
void mxyb_data_i32(mxyb_builder* h, const char* k, int32_t v) { h->cur.map[k] = DataValue::Int32(v); }
void mxyb_data_u16(mxyb_builder* h, const char* k, uint16_t v) { h->cur.map[k] = DataValue::Uint16(v); }
void mxyb_data_u32(mxyb_builder* h, const char* k, uint32_t v) { h->cur.map[k] = DataValue::Uint32(v); }
void mxyb_data_u64(mxyb_builder* h, const char* k, uint64_t v) { h->cur.map[k] = DataValue::Uint64(v); }
void mxyb_data_f64(mxyb_builder* h, const char* k, double v) { h->cur.map[k] = DataValue::Double(v); }
void mxyb_data_bool(mxyb_builder* h, const char* k, int v) { h->cur.map[k] = DataValue::Bool(v != 0); }
uint32_t mxyb_data_commit(mxyb_builder* h) { return h->b.encode_data(h->cur); }

// kind: 0 auto-detect (with literal:/glob:/ip: prefixes), 1 ip, 2 literal, 3 glob.  0 on success, -1 on error.
int mxyb_add(mxyb_builder* h, int kind, const char* key, size_t klen, uint32_t data_offset) {
  std::string k(key, klen);
  switch (kind) {
    case 0: return h->b.add_entry_at(k, data_offset) ? 0 : -1;
    case 1: return h->b.add_ip(k, data_offset) ? 0 : -1;
    case 2: h->b.add_literal(k, data_offset); return 0;
    case 3: {
      std::vector<GlobSeg> segs; std::string err;
      if (!parse_glob(k, segs, err)) return -1;
      h->b.add_glob(k, data_offset); return 0;
    }
  }
  return -1;
}
void mxyb_set_epoch(mxyb_builder* h, uint64_t e) { h->b.set_build_epoch(e); }
void mxyb_set_type(mxyb_builder* h, const char* t) { h->b.set_database_type(t); }
void mxyb_set_description(mxyb_builder* h, const char* lang, const char* text) { h->b.set_description(lang, text); }
int mxyb_build(mxyb_builder* h) { return h->b.build(h->out) ? 0 : -1; }
const uint8_t* mxyb_bytes(mxyb_builder* h, size_t* len) { *len = h->out.size(); return h->out.data(); }
void mxyb_counts(mxyb_builder* h, uint64_t* out3) { out3[0] = h->b.ip_count(); out3[1] = h->b.literal_count(); out3[2] = h->b.glob_count(); }
int mxyb_save(mxyb_builder* h, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  size_t w = fwrite(h->out.data(), 1, h->out.size(), f);
  fclose(f);
  return w == h->out.size() ? 0 : -1;
}
uint64_t mxyb_xxh64(const uint8_t* p, size_t n) { return xxh64(p, n, 0); }

}  // extern "C"
