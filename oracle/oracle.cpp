// oracle.cpp — CPU restatement of matchy's log-scan hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for the CUDA engine in matchy_b200/.  It is NOT part of the
// product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it.  The product path never links or calls anything in oracle/.
//
// It follows the reference (matchylabs/matchy v1.2.2, Rust) function by function; every function
// cites the reference file:line it restates (paths relative to the reference root).  The Rust
// reference cannot be compiled in this environment (no cargo/rustc), so the oracle is pinned by
// the reference's own known-answer tests, ported into tests/test_oracle_kats.py
// (extractor KATs matchy-extractor/src/lib.rs:1922-3235, LPM tests
// crates/matchy/tests/test_ip_longest_prefix_match.rs, literal/glob tests
// crates/matchy/tests/test_literal_hash.rs, glob KATs matchy-paraglob/src/glob.rs:465-706,
// XXH64 vectors cross-checked against the python `xxhash` module).
//
// Deliberately sequential and literal: the per-anchor loops, `last_end` bookkeeping and chunk
// splitting are kept exactly as the reference has them, so that the (differently structured)
// GPU tokenizer is checked against the reference's formulation, not against its own.
//
// Results-neutral things NOT restated: the thread-local LRU query cache (database.rs:725-804),
// sampled timers and DatabaseStats counters.

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <list>
#include <string>
#include <unordered_map>
#include <thread>
#include <unordered_set>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;

// ===========================================================================================
// MMDB data-section decoder — crates/matchy-data-format/src/lib.rs:635-1047
// ===========================================================================================
struct Val {
  enum T { PTR, STR, DBL, BYTES, U16, U32, MAP, I32, U64, U128, ARR, BOOL, FLT } t = U16;
  std::string s;                 // STR / BYTES
  double d = 0;                  // DBL
  float f = 0;                   // FLT
  uint64_t u = 0;                // U16/U32/U64/PTR/BOOL
  u128 big = 0;                  // U128
  int32_t i = 0;                 // I32
  std::map<std::string, Val> m;  // MAP (HashMap in the reference; order is irrelevant, JSON sorts)
  std::vector<Val> a;            // ARR
};

struct Decoder {
  const uint8_t* buf;
  size_t len;
  bool ok = true;

  Decoder(const uint8_t* b, size_t l) : buf(b), len(l) {}

  // lib.rs:981-1013
  size_t decode_size(size_t& cur, uint8_t size_bits) {
    if (size_bits <= 28) return size_bits;
    if (size_bits == 29) {
      if (cur >= len) { ok = false; return 0; }
      size_t s = buf[cur]; cur += 1; return 29 + s;
    }
    if (size_bits == 30) {
      if (cur + 2 > len) { ok = false; return 0; }
      size_t s = (size_t(buf[cur]) << 8) | buf[cur + 1]; cur += 2; return 29 + 256 + s;
    }
    if (cur + 3 > len) { ok = false; return 0; }
    size_t s = (size_t(buf[cur]) << 16) | (size_t(buf[cur + 1]) << 8) | buf[cur + 2];
    cur += 3;
    return 29 + 256 + 65536 + s;
  }

  uint64_t read_be(size_t& cur, size_t n) {
    uint64_t v = 0;
    for (size_t k = 0; k < n; k++) v = (v << 8) | buf[cur + k];
    cur += n;
    return v;
  }

  // lib.rs:665-687
  Val decode_at(size_t& cur) {
    Val v;
    if (cur >= len) { ok = false; return v; }
    uint8_t ctrl = buf[cur++];
    uint8_t type_id = ctrl >> 5, payload = ctrl & 0x1F;
    switch (type_id) {
      case 0: return decode_extended(cur, payload);
      case 1: {  // pointer, lib.rs:721-771
        uint8_t size_bits = (payload >> 3) & 3;
        uint32_t low3 = payload & 7, off = 0;
        if (size_bits == 0) {
          if (cur >= len) { ok = false; return v; }
          off = (low3 << 8) | buf[cur]; cur += 1;
        } else if (size_bits == 1) {
          if (cur + 1 >= len) { ok = false; return v; }
          off = 2048 + ((low3 << 16) | (uint32_t(buf[cur]) << 8) | buf[cur + 1]); cur += 2;
        } else if (size_bits == 2) {
          if (cur + 2 >= len) { ok = false; return v; }
          off = 526336 + ((low3 << 24) | (uint32_t(buf[cur]) << 16) | (uint32_t(buf[cur + 1]) << 8) | buf[cur + 2]);
          cur += 3;
        } else {
          if (cur + 3 >= len) { ok = false; return v; }
          off = (uint32_t)read_be(cur, 4);
        }
        v.t = Val::PTR; v.u = off; return v;
      }
      case 2: {  // string
        size_t n = decode_size(cur, payload);
        if (!ok || cur + n > len) { ok = false; return v; }
        v.t = Val::STR; v.s.assign((const char*)buf + cur, n); cur += n; return v;
      }
      case 3: {  // double
        if (cur + 8 > len) { ok = false; return v; }
        uint64_t bits = read_be(cur, 8);
        v.t = Val::DBL; memcpy(&v.d, &bits, 8); return v;
      }
      case 4: {  // bytes
        size_t n = decode_size(cur, payload);
        if (!ok || cur + n > len) { ok = false; return v; }
        v.t = Val::BYTES; v.s.assign((const char*)buf + cur, n); cur += n; return v;
      }
      case 5: {
        size_t n = decode_size(cur, payload);
        if (!ok || n > 2 || cur + n > len) { ok = false; return v; }
        v.t = Val::U16; v.u = read_be(cur, n); return v;
      }
      case 6: {
        size_t n = decode_size(cur, payload);
        if (!ok || n > 4 || cur + n > len) { ok = false; return v; }
        v.t = Val::U32; v.u = read_be(cur, n); return v;
      }
      default: {  // 7: map, lib.rs:854-878
        size_t count = decode_size(cur, payload);
        v.t = Val::MAP;
        for (size_t k = 0; ok && k < count; k++) {
          Val key = decode_at(cur);
          if (!ok) return v;
          std::string ks;
          if (key.t == Val::STR) ks = key.s;
          else if (key.t == Val::PTR) {
            Val kv = decode((uint32_t)key.u);
            if (!ok || kv.t != Val::STR) { ok = false; return v; }
            ks = kv.s;
          } else { ok = false; return v; }
          Val val = decode_at(cur);
          if (!ok) return v;
          v.m[ks] = val;
        }
        return v;
      }
    }
  }

  // lib.rs:689-719
  Val decode_extended(size_t& cur, uint8_t size_from_ctrl) {
    Val v;
    if (cur >= len) { ok = false; return v; }
    int type_id = 7 + buf[cur++];
    switch (type_id) {
      case 8: {  // int32, lib.rs:880-909
        size_t n = decode_size(cur, size_from_ctrl);
        if (!ok || n > 4 || cur + n > len) { ok = false; return v; }
        int32_t val = 0;
        if (n > 0) {
          if (buf[cur] & 0x80) val = -1;
          for (size_t k = 0; k < n; k++) val = (int32_t)(((uint32_t)val << 8) | buf[cur + k]);
        }
        cur += n;
        v.t = Val::I32; v.i = val; return v;
      }
      case 9: {
        size_t n = decode_size(cur, size_from_ctrl);
        if (!ok || n > 8 || cur + n > len) { ok = false; return v; }
        v.t = Val::U64; v.u = read_be(cur, n); return v;
      }
      case 10: {
        size_t n = decode_size(cur, size_from_ctrl);
        if (!ok || n > 16 || cur + n > len) { ok = false; return v; }
        u128 b = 0;
        for (size_t k = 0; k < n; k++) b = (b << 8) | buf[cur + k];
        cur += n;
        v.t = Val::U128; v.big = b; return v;
      }
      case 11: {
        size_t count = decode_size(cur, size_from_ctrl);
        v.t = Val::ARR;
        for (size_t k = 0; ok && k < count; k++) v.a.push_back(decode_at(cur));
        return v;
      }
      case 14: v.t = Val::BOOL; v.u = size_from_ctrl != 0; return v;
      case 15: {
        if (size_from_ctrl != 4 || cur + 4 > len) { ok = false; return v; }
        uint32_t bits = (uint32_t)read_be(cur, 4);
        v.t = Val::FLT; memcpy(&v.f, &bits, 4); return v;
      }
      default: ok = false; return v;
    }
  }

  // lib.rs:1016-1047
  Val resolve(Val v, int depth = 0) {
    if (depth > 64) { ok = false; return v; }
    if (v.t == Val::PTR) {
      size_t cur = (size_t)v.u;
      Val p = decode_at(cur);
      if (!ok) return p;
      return resolve(p, depth + 1);
    }
    if (v.t == Val::MAP) { for (auto& kv : v.m) kv.second = resolve(kv.second, depth + 1); }
    if (v.t == Val::ARR) { for (auto& e : v.a) e = resolve(e, depth + 1); }
    return v;
  }

  // lib.rs:654-663
  Val decode(uint32_t offset) {
    size_t cur = offset;
    Val v = decode_at(cur);
    if (!ok) return v;
    return resolve(v);
  }
};

// ===========================================================================================
// serde_json rendering — bin/cli_utils.rs:177-201 (data_value_to_json) + serde_json 1.0.145
// (no preserve_order ⇒ object keys sorted; compact separators)
// ===========================================================================================
static void json_escape(const std::string& s, std::string& out) {
  static const char* hex = "0123456789abcdef";
  out.push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\b': out += "\\b"; break;
      case '\f': out += "\\f"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      default:
        if (c < 0x20) { out += "\\u00"; out.push_back(hex[c >> 4]); out.push_back(hex[c & 15]); }
        else out.push_back((char)c);
    }
  }
  out.push_back('"');
}

static void json_double(double d, std::string& out) {
  if (!(d == d) || d > 1.79e308 || d < -1.79e308) { out += "null"; return; }
  char b[40];
  for (int prec = 1; prec <= 17; prec++) {
    snprintf(b, sizeof b, "%.*g", prec, d);
    if (strtod(b, nullptr) == d) break;
  }
  std::string s = b;
  // ryu prints integral floats as "1.0" and exponents as "1e21"/"1e-7"
  size_t e = s.find('e');
  if (e != std::string::npos) {
    std::string mant = s.substr(0, e), ex = s.substr(e + 1);
    bool neg = false; size_t k = 0;
    if (ex[0] == '-') { neg = true; k = 1; } else if (ex[0] == '+') k = 1;
    while (k + 1 < ex.size() && ex[k] == '0') k++;
    s = mant + "e" + (neg ? "-" : "") + ex.substr(k);
  } else if (s.find('.') == std::string::npos) s += ".0";
  out += s;
}

static void u128_dec(u128 v, std::string& out) {
  if (v == 0) { out += "0"; return; }
  char b[48]; int n = 0;
  while (v) { b[n++] = char('0' + (int)(v % 10)); v /= 10; }
  while (n) out.push_back(b[--n]);
}

static void json_val(const Val& v, std::string& out) {
  switch (v.t) {
    case Val::STR: json_escape(v.s, out); break;
    case Val::DBL: json_double(v.d, out); break;
    case Val::FLT: json_double((double)v.f, out); break;
    case Val::BYTES: {
      out.push_back('[');
      for (size_t k = 0; k < v.s.size(); k++) { if (k) out.push_back(','); out += std::to_string((unsigned)(uint8_t)v.s[k]); }
      out.push_back(']');
      break;
    }
    case Val::U16: case Val::U32: case Val::U64: out += std::to_string(v.u); break;
    case Val::U128: out.push_back('"'); u128_dec(v.big, out); out.push_back('"'); break;
    case Val::I32: out += std::to_string(v.i); break;
    case Val::BOOL: out += v.u ? "true" : "false"; break;
    case Val::MAP: {
      out.push_back('{');
      bool first = true;
      for (auto& kv : v.m) {  // std::map ⇒ byte-wise sorted keys, same as BTreeMap<String,_>
        if (!first) out.push_back(',');
        first = false;
        json_escape(kv.first, out); out.push_back(':'); json_val(kv.second, out);
      }
      out.push_back('}');
      break;
    }
    case Val::ARR: {
      out.push_back('[');
      for (size_t k = 0; k < v.a.size(); k++) { if (k) out.push_back(','); json_val(v.a[k], out); }
      out.push_back(']');
      break;
    }
    case Val::PTR: out += "\"<pointer>\""; break;
  }
}

// ===========================================================================================
// XXH64 (xxhash-rust 0.8.15 `xxh64`, seed 0) — matchy-literal-hash/src/lib.rs:43,666-671.
// Third-party, absent from the reference tree: restated from the public XXH64 specification.
// ===========================================================================================
static const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL,
                      P4 = 9650029242287828579ULL, P5 = 2870177450012600261ULL;
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
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

// ===========================================================================================
// Rust std semantics that the reference leans on
// ===========================================================================================
// str::from_utf8 (strict UTF-8: no overlongs, no surrogates, max U+10FFFF)
static bool valid_utf8(const uint8_t* s, size_t n) {
  size_t i = 0;
  while (i < n) {
    uint8_t c = s[i];
    if (c < 0x80) { i++; continue; }
    if (c >= 0xC2 && c <= 0xDF) {
      if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return false;
      i += 2;
    } else if (c >= 0xE0 && c <= 0xEF) {
      if (i + 2 >= n) return false;
      uint8_t c1 = s[i + 1], c2 = s[i + 2];
      if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80) return false;
      if (c == 0xE0 && c1 < 0xA0) return false;
      if (c == 0xED && c1 > 0x9F) return false;
      i += 3;
    } else if (c >= 0xF0 && c <= 0xF4) {
      if (i + 3 >= n) return false;
      uint8_t c1 = s[i + 1], c2 = s[i + 2], c3 = s[i + 3];
      if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80 || (c3 & 0xC0) != 0x80) return false;
      if (c == 0xF0 && c1 < 0x90) return false;
      if (c == 0xF4 && c1 > 0x8F) return false;
      i += 4;
    } else return false;
  }
  return true;
}

// core::net::parser — Ipv6Addr::from_str restricted to inputs without '.' (read_ipv6_addr,
// read_groups, read_number(16, Some(4), true)).  SURVEY §8 quirk 6.
static size_t v6_read_groups(const uint8_t* s, size_t n, size_t& pos, uint16_t* groups, size_t limit) {
  for (size_t i = 0; i < limit; i++) {
    size_t save = pos;
    if (i > 0) {
      if (pos < n && s[pos] == ':') pos++;
      else { pos = save; return i; }
    }
    // read_number: up to 4 hex digits; a 5th digit makes the whole group fail (atomically)
    uint32_t v = 0; int digits = 0; bool fail = false;
    while (pos < n) {
      uint8_t c = s[pos]; int d;
      if (c >= '0' && c <= '9') d = c - '0';
      else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
      else break;
      v = v * 16 + d; digits++; pos++;
      if (digits > 4) { fail = true; break; }
    }
    if (fail || digits == 0) { pos = save; return i; }
    groups[i] = (uint16_t)v;
  }
  return limit;
}

static bool parse_ipv6(const uint8_t* s, size_t n, uint16_t out[8]) {
  size_t pos = 0;
  uint16_t head[8] = {0};
  size_t head_size = v6_read_groups(s, n, pos, head, 8);
  if (head_size == 8) {
    if (pos != n) return false;
    memcpy(out, head, 16);
    return true;
  }
  if (!(pos + 1 < n && s[pos] == ':' && s[pos + 1] == ':')) return false;
  pos += 2;
  uint16_t tail[7] = {0};
  size_t limit = 8 - (head_size + 1);
  size_t tail_size = v6_read_groups(s, n, pos, tail, limit);
  if (pos != n) return false;
  for (size_t k = 0; k < tail_size; k++) head[8 - tail_size + k] = tail[k];
  memcpy(out, head, 16);
  return true;
}

// Ipv4Addr::from_str: exactly 4 decimal octets, 1-3 digits, no leading zeros, <=255
static bool parse_ipv4_strict(const uint8_t* s, size_t n, uint32_t& out) {
  size_t pos = 0; uint32_t addr = 0;
  for (int k = 0; k < 4; k++) {
    if (k > 0) { if (pos >= n || s[pos] != '.') return false; pos++; }
    size_t st = pos; uint32_t v = 0;
    while (pos < n && s[pos] >= '0' && s[pos] <= '9' && pos - st < 4) { v = v * 10 + (s[pos] - '0'); pos++; }
    size_t d = pos - st;
    if (d == 0 || d > 3 || v > 255 || (d > 1 && s[st] == '0')) return false;
    addr = (addr << 8) | v;
  }
  if (pos != n) return false;
  out = addr;
  return true;
}

// Ipv6Addr Display (RFC 5952; v4-mapped special form)
static std::string ipv6_display(const uint16_t seg[8]) {
  char b[64];
  bool mapped = seg[0] == 0 && seg[1] == 0 && seg[2] == 0 && seg[3] == 0 && seg[4] == 0 && seg[5] == 0xffff;
  if (mapped) {
    snprintf(b, sizeof b, "::ffff:%u.%u.%u.%u", seg[6] >> 8, seg[6] & 255, seg[7] >> 8, seg[7] & 255);
    return b;
  }
  int best_start = -1, best_len = 0, cs = -1, cl = 0;
  for (int i = 0; i < 8; i++) {
    if (seg[i] == 0) {
      if (cs < 0) { cs = i; cl = 0; }
      cl++;
      if (cl > best_len) { best_len = cl; best_start = cs; }
    } else cs = -1;
  }
  std::string out;
  auto fmt = [&](int a, int e) {
    for (int i = a; i < e; i++) {
      if (i > a) out.push_back(':');
      snprintf(b, sizeof b, "%x", seg[i]);
      out += b;
    }
  };
  if (best_len > 1) {
    fmt(0, best_start);
    out += "::";
    fmt(best_start + best_len, 8);
  } else fmt(0, 8);
  return out;
}

// ===========================================================================================
// Extractor — crates/matchy-extractor/src/lib.rs
// ===========================================================================================
enum ItemType : uint8_t {  // include/matchy/matchy.h:233-288
  T_DOMAIN = 0, T_EMAIL = 1, T_IPV4 = 2, T_IPV6 = 3, T_MD5 = 4, T_SHA1 = 5, T_SHA256 = 6, T_SHA384 = 7, T_SHA512 = 8,
  T_BITCOIN = 9, T_ETHEREUM = 10, T_MONERO = 11
};
enum ExtractFlags : uint32_t {  // matchy.h:188-228
  X_DOMAINS = 1, X_EMAILS = 2, X_IPV4 = 4, X_IPV6 = 8, X_HASHES = 16, X_BITCOIN = 32, X_ETHEREUM = 64, X_MONERO = 128
};


// ===========================================================================================
// Crypto-address validators — matchy-extractor/src/lib.rs:1799-1920.  Third-party crates absent from the tree, restated
// from their published algorithms: sha2 0.10 (FIPS 180-4 SHA-256), tiny-keccak 2.0 (Keccak-256, original 0x01 padding),
// bs58 0.5.1 (Bitcoin alphabet, decode_into), bech32 0.11.1 (`decode`: BIP-173 / BIP-350 checksum, either constant).
// Pinned by the reference's own address vectors (lib.rs:3240-3626) and by standard digests (tests/test_oracle_kats.py).
// ===========================================================================================
namespace cryptoaddr {

static inline uint32_t rotr32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
static void sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74,
      0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d,
      0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e,
      0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5,
      0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  std::vector<uint8_t> m(msg, msg + len);
  m.push_back(0x80);
  while (m.size() % 64 != 56) m.push_back(0);
  uint64_t bits = (uint64_t)len * 8;
  for (int k = 7; k >= 0; k--) m.push_back((uint8_t)(bits >> (8 * k)));
  for (size_t off = 0; off < m.size(); off += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)m[off + 4 * i] << 24) | ((uint32_t)m[off + 4 * i + 1] << 16) | ((uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25), ch = (e & f) ^ (~e & g), t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22), mj = (a & b) ^ (a & c) ^ (b & c), t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}

static inline uint64_t rotl64(uint64_t x, int r) { return r ? (x << r) | (x >> (64 - r)) : x; }
static void keccak_f1600(uint64_t st[25]) {
  static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
                                  0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                                  0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
                                  0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};  // [x + 5y]
  for (int round = 0; round < 24; round++) {
    uint64_t C[5], D[5], B[25];
    for (int x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) st[i] ^= D[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[round];
  }
}
// Keccak-256 as tiny-keccak's Keccak::v256: rate 136 bytes, domain byte 0x01 (NOT SHA3's 0x06)
static void keccak256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  uint64_t st[25] = {0};
  const size_t rate = 136;
  std::vector<uint8_t> m(msg, msg + len);
  m.push_back(0x01);
  while (m.size() % rate != 0) m.push_back(0);
  m.back() |= 0x80;
  for (size_t off = 0; off < m.size(); off += rate) {
    for (size_t i = 0; i < rate / 8; i++) {
      uint64_t v = 0;
      for (int k = 7; k >= 0; k--) v = (v << 8) | m[off + 8 * i + k];
      st[i] ^= v;
    }
    keccak_f1600(st);
  }
  for (int i = 0; i < 32; i++) out[i] = (uint8_t)(st[i / 8] >> (8 * (i % 8)));
}

// bs58 0.5.1 decode (Bitcoin alphabet): false on a character outside the alphabet
static bool bs58_decode(const uint8_t* s, size_t n, std::vector<uint8_t>& out) {
  static const char* A = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz";
  int8_t map[128];
  for (int i = 0; i < 128; i++) map[i] = -1;
  for (int i = 0; i < 58; i++) map[(int)A[i]] = (int8_t)i;
  std::vector<uint8_t> le;  // little-endian number
  for (size_t i = 0; i < n; i++) {
    if (s[i] > 127 || map[s[i]] < 0) return false;
    size_t val = (size_t)map[s[i]];
    for (auto& b : le) { val += (size_t)b * 58; b = (uint8_t)(val & 0xFF); val >>= 8; }
    while (val > 0) { le.push_back((uint8_t)(val & 0xFF)); val >>= 8; }
  }
  for (size_t i = 0; i < n && s[i] == '1'; i++) le.push_back(0);
  out.assign(le.rbegin(), le.rend());
  return true;
}
static bool bitcoin_base58(const uint8_t* s, size_t n) {  // validate_bitcoin_base58 :1799-1822
  std::vector<uint8_t> d;
  if (!bs58_decode(s, n, d) || d.size() < 5) return false;
  uint8_t h1[32], h2[32];
  sha256(d.data(), d.size() - 4, h1);
  sha256(h1, 32, h2);
  return memcmp(h2, d.data() + d.size() - 4, 4) == 0;
}
static bool monero(const uint8_t* s, size_t n) {  // validate_monero_address :1895-1920
  std::vector<uint8_t> d;
  if (!bs58_decode(s, n, d) || d.size() < 5) return false;
  uint8_t h[32];
  keccak256(d.data(), d.size() - 4, h);
  return memcmp(h, d.data() + d.size() - 4, 4) == 0;
}
// validate_bitcoin_bech32 :1825-1836 — bech32::decode(addr) succeeds and the human-readable part is "bc"
static bool bitcoin_bech32(const uint8_t* s, size_t n) {
  static const char* CH = "qpzry9x8gf2tvdw0s3jn54khce6mua7l";
  // check_characters: scanning from the end, everything after the LAST '1' must be a bech32 character (either case);
  // no mixed case anywhere; a separator must exist
  bool upper = false, lower = false;
  long sep = -1;
  for (long i = (long)n - 1; i >= 0; i--) {
    uint8_t ch = s[i];
    if (ch > 127) return false;  // (a non-ASCII char after the separator is invalid; before it, Hrp::parse rejects it)
    if (ch == '1' && sep < 0) sep = i;
    else if (sep < 0) {
      uint8_t lc = (ch >= 'A' && ch <= 'Z') ? (uint8_t)(ch + 32) : ch;
      if (!strchr(CH, lc) || lc == 0) return false;
    }
    if (ch >= 'A' && ch <= 'Z') upper = true; else if (ch >= 'a' && ch <= 'z') lower = true;
  }
  if (upper && lower) return false;
  if (sep < 0) return false;
  // Hrp::parse: 1..83 chars of ASCII 33..126; equality with "bc" is case-insensitive
  if (sep != 2) return false;
  if (!((s[0] == 'b' || s[0] == 'B') && (s[1] == 'c' || s[1] == 'C'))) return false;
  size_t dn = n - 3;
  if (dn < 6) return false;  // shorter than a checksum
  static const uint32_t GEN[5] = {0x3b6a57b2, 0x26508e6d, 0x1ea119fa, 0x3d4233dd, 0x2a1462b3};
  uint32_t chk = 1;
  auto step = [&](uint32_t v) {
    uint32_t b = chk >> 25;
    chk = ((chk & 0x1ffffff) << 5) ^ v;
    for (int i = 0; i < 5; i++) if ((b >> i) & 1) chk ^= GEN[i];
  };
  step('b' >> 5); step('c' >> 5); step(0); step('b' & 31); step('c' & 31);
  for (size_t i = 3; i < n; i++) {
    uint8_t lc = (s[i] >= 'A' && s[i] <= 'Z') ? (uint8_t)(s[i] + 32) : s[i];
    step((uint32_t)(strchr(CH, lc) - CH));
  }
  return chk == 1 || chk == 0x2bc830a3;  // Bech32 or Bech32m
}
// validate_ethereum_checksum :1841-1892 for "0x" + 40 hex digits
static bool ethereum(const uint8_t* s) {
  bool lower = false, upper = false;
  for (int i = 2; i < 42; i++) { if (s[i] >= 'a' && s[i] <= 'f') lower = true; if (s[i] >= 'A' && s[i] <= 'F') upper = true; }
  if (!(lower && upper)) return true;  // all one case (or digits only): nothing to verify
  uint8_t lc[40], h[32];
  for (int i = 0; i < 40; i++) lc[i] = (s[2 + i] >= 'A' && s[2 + i] <= 'F') ? (uint8_t)(s[2 + i] + 32) : s[2 + i];
  keccak256(lc, 40, h);
  for (int i = 0; i < 40; i++) {
    uint8_t c = s[2 + i];
    bool alpha = (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F');
    if (!alpha) continue;
    uint8_t nib = (i % 2 == 0) ? (h[i / 2] >> 4) : (h[i / 2] & 15);
    if ((c <= 'F') != (nib >= 8)) return false;
  }
  return true;
}

}  // namespace cryptoaddr

struct Item {
  uint8_t type;
  size_t start, end;
  uint32_t v4 = 0;
  uint16_t v6[8] = {0};
};

// lib.rs:1568-1593
static bool is_boundary(uint8_t b) {
  switch (b) {
    case ' ': case '\t': case '\n': case '\r': case '/': case ',': case ';': case ':': case '(': case ')':
    case '[': case ']': case '{': case '}': case '<': case '>': case '"': case '\'': case '@': case '=':
      return true;
    default: return false;
  }
}
static inline bool is_digit(uint8_t b) { return b >= '0' && b <= '9'; }
static inline bool is_alpha(uint8_t b) { return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z'); }
static inline bool is_hex(uint8_t b) { return is_digit(b) || (b >= 'a' && b <= 'f') || (b >= 'A' && b <= 'F'); }  // :1696-1717
// lib.rs:1597-1629 (DOMAIN_CHAR_LOOKUP incl. 0x80-0xFF)
static inline bool is_domain_char_fast(uint8_t b) { return is_digit(b) || is_alpha(b) || b == '-' || b == '.' || b >= 0x80; }
// lib.rs:1639-1641
static inline bool is_domain_char(uint8_t b) { return is_digit(b) || is_alpha(b) || b == '-' || b == '.'; }
// lib.rs:1644-1647
static inline bool is_email_local_char(uint8_t b) { return is_digit(b) || is_alpha(b) || b == '.' || b == '-' || b == '_' || b == '+'; }

struct Psl {
  std::unordered_set<std::string> set;  // lib.rs:1552-1563
  bool load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    std::string all; char b[65536]; size_t n;
    while ((n = fread(b, 1, sizeof b, f)) > 0) all.append(b, n);
    fclose(f);
    size_t i = 0;
    while (i <= all.size()) {
      size_t e = all.find('\n', i);
      if (e == std::string::npos) e = all.size();
      std::string line = all.substr(i, e - i);
      // str::trim (whitespace both ends)
      size_t a = 0, z = line.size();
      while (a < z && (line[a] == ' ' || line[a] == '\t' || line[a] == '\r' || line[a] == '\n' || line[a] == '\f' || line[a] == '\v')) a++;
      while (z > a && (line[z - 1] == ' ' || line[z - 1] == '\t' || line[z - 1] == '\r' || line[z - 1] == '\n' || line[z - 1] == '\f' || line[z - 1] == '\v')) z--;
      line = line.substr(a, z - a);
      if (!line.empty() && !(line.size() >= 2 && line[0] == '/' && line[1] == '/')) set.insert(line);
      if (e >= all.size()) break;
      i = e + 1;
    }
    return !set.empty();
  }
  // lib.rs:1671-1692  (shortest suffix first; returns position of the dot or -1)
  long find_valid_tld_suffix(const uint8_t* d, size_t n) const {
    for (size_t k = n; k-- > 0;) {
      if (d[k] == '.') {
        std::string suf((const char*)d + k + 1, n - k - 1);
        if (set.count(suf)) return (long)k;
      }
    }
    return -1;
  }
};

struct Extractor {
  uint32_t flags = 0xFF;
  size_t min_domain_labels = 2;         // lib.rs:47
  bool require_word_boundaries = true;  // lib.rs:48
  const Psl* psl = nullptr;

  // lib.rs:1742-1782
  static void find_word_boundaries(const uint8_t* c, size_t n, std::vector<size_t>& out) {
    if (n == 0) return;
    bool in_token = !is_boundary(c[0]);
    if (in_token) out.push_back(0);
    for (size_t i = 1; i < n; i++) {
      bool b = is_boundary(c[i]);
      if (in_token && b) { out.push_back(i); in_token = false; }
      else if (!in_token && !b) { out.push_back(i); in_token = true; }
    }
    if (in_token) out.push_back(n);
  }

  // lib.rs:1425-1456
  static bool is_ipv6_loopback_or_linklocal(const uint8_t* c, size_t n) {
    if (n == 3 && memcmp(c, "::1", 3) == 0) return true;
    if (n >= 4) {
      auto lc = [](uint8_t x) { return (x >= 'A' && x <= 'Z') ? uint8_t(x + 32) : x; };
      if (lc(c[0]) == 'f' && lc(c[1]) == 'e') {
        uint8_t t = lc(c[2]);
        if (t == '8' && lc(c[3]) == '0') return true;
        if (t == '8' || t == '9' || t == 'a' || t == 'b') return true;
      }
    }
    return false;
  }

  // lib.rs:1044-1116.  memmem::find_iter("::") = leftmost non-overlapping occurrences.
  void extract_ipv6_chunk(const uint8_t* c, size_t n, std::vector<Item>& out) const {
    size_t last_end = 0, scan = 0;
    while (scan + 1 < n) {
      const uint8_t* p = (const uint8_t*)memchr(c + scan, ':', n - scan);
      if (!p) break;
      size_t pos = p - c;
      if (pos + 1 >= n) break;
      if (c[pos + 1] != ':') { scan = pos + 1; continue; }
      scan = pos + 2;  // non-overlapping
      if (pos < last_end) continue;
      bool hex_before = pos > 0 && is_hex(c[pos - 1]);
      bool hex_after = pos + 2 < n && is_hex(c[pos + 2]);
      if (!hex_before && !hex_after) { last_end = pos + 2; continue; }
      size_t start = pos;
      while (start > 0) { uint8_t ch = c[start - 1]; if (!is_hex(ch) && ch != ':') break; start--; }
      size_t end = pos + 2;
      while (end < n) { uint8_t ch = c[end]; if (!is_hex(ch) && ch != ':') break; end++; }
      const uint8_t* cand = c + start; size_t cl = end - start;
      if (cl < 8) { last_end = end; continue; }
      if ((cand[0] == ':' && cand[1] == ':') || (cand[cl - 2] == ':' && cand[cl - 1] == ':')) { last_end = end; continue; }
      if (is_ipv6_loopback_or_linklocal(cand, cl)) { last_end = end; continue; }
      Item it; it.type = T_IPV6; it.start = start; it.end = end;
      if (parse_ipv6(cand, cl, it.v6)) { out.push_back(it); last_end = end; continue; }
      last_end = pos + 2;
    }
  }

  // lib.rs:813-869
  bool try_parse_ipv4(const uint8_t* c, size_t n, size_t start, uint32_t& ip, size_t& end) const {
    size_t pos = start;
    if (require_word_boundaries && start > 0 && !is_boundary(c[start - 1])) return false;
    uint32_t addr = 0;
    for (int k = 0; k < 4; k++) {
      uint32_t v = 0; int digits = 0; size_t ostart = pos;
      while (pos < n && is_digit(c[pos]) && digits < 3) { v = v * 10 + (c[pos] - '0'); pos++; digits++; }
      if (digits == 0) return false;
      if (v > 255) return false;
      if (digits > 1 && c[ostart] == '0') return false;
      addr = (addr << 8) | v;
      if (k < 3) { if (pos >= n || c[pos] != '.') return false; pos++; }
    }
    if (require_word_boundaries && pos < n && !is_boundary(c[pos])) return false;
    ip = addr; end = pos;
    return true;
  }

  // lib.rs:1120-1179
  void extract_ipv4_chunk(const uint8_t* c, size_t n, const std::vector<size_t>& dots, std::vector<Item>& out) const {
    size_t last_end = 0;
    for (size_t i = 0; i < dots.size(); i++) {
      size_t dot = dots[i];
      if (dot == 0 || dot + 6 > n) continue;
      if (!is_digit(c[dot - 1]) || !is_digit(c[dot + 1])) continue;
      size_t start = dot;
      while (start > 0 && (is_digit(c[start - 1]) || c[start - 1] == '.')) start--;
      if (start < last_end) continue;
      size_t end_search = std::min(start + 15, n);
      size_t cnt = 0;
      for (size_t j = i; j < dots.size() && dots[j] < end_search; j++) cnt++;
      if (cnt < 3) continue;
      uint32_t ip; size_t end;
      if (try_parse_ipv4(c, n, start, ip, end)) {
        Item it; it.type = T_IPV4; it.start = start; it.end = end; it.v4 = ip;
        out.push_back(it);
        last_end = end;
      }
    }
  }

  // lib.rs:891-950
  bool extract_email_at(const uint8_t* c, size_t n, size_t at, size_t& s, size_t& e) const {
    size_t start = at;
    while (start > 0 && is_email_local_char(c[start - 1])) start--;
    if (start == at) return false;
    if (require_word_boundaries && start > 0 && !is_boundary(c[start - 1])) return false;
    size_t end = at + 1;
    while (end < n && is_domain_char(c[end])) end++;
    if (end == at + 1) return false;
    if (require_word_boundaries && end < n && !is_boundary(c[end])) return false;
    const uint8_t* local = c + start; size_t ll = at - start;
    const uint8_t* dom = c + at + 1; size_t dl = end - at - 1;
    for (size_t k = 0; k + 1 < ll; k++) if (local[k] == '.' && local[k + 1] == '.') return false;
    bool has_letter = false;
    for (size_t k = 0; k < ll; k++) if (is_alpha(local[k])) { has_letter = true; break; }
    if (!has_letter) return false;
    if (!memchr(dom, '.', dl)) return false;
    if (psl->find_valid_tld_suffix(dom, dl) < 0) return false;
    s = start; e = end;
    return true;
  }

  // lib.rs:1182-1196
  void extract_emails_chunk(const uint8_t* c, size_t n, std::vector<Item>& out) const {
    for (size_t i = 0; i < n; i++) {
      if (c[i] != '@') continue;
      size_t s, e;
      if (extract_email_at(c, n, i, s, e) && valid_utf8(c + s, e - s)) {
        Item it; it.type = T_EMAIL; it.start = s; it.end = e;
        out.push_back(it);
      }
    }
  }

  // lib.rs:637-689
  bool is_valid_domain(const uint8_t* d, size_t n) const {
    size_t label_count = 0, label_start = 0;
    auto valid_label = [&](size_t a, size_t b) {
      if (a == b) return false;
      if (d[a] == '-' || d[b - 1] == '-') return false;
      return true;
    };
    for (size_t i = 0; i < n; i++) {
      if (d[i] == '.') {
        if (!valid_label(label_start, i)) return false;
        label_count++;
        label_start = i + 1;
      }
    }
    if (!valid_label(label_start, n)) return false;
    label_count++;
    return label_count >= min_domain_labels;
  }

  // lib.rs:537-628
  void extract_domains_chunk(const uint8_t* c, size_t n, const std::vector<size_t>& dots, std::vector<Item>& out) const {
    size_t last_domain_end = 0;
    for (size_t dot : dots) {
      if (dot < last_domain_end) continue;
      size_t start = dot;
      while (start > 0 && is_domain_char_fast(c[start - 1])) start--;
      size_t end = dot + 1;
      while (end < n && is_domain_char_fast(c[end])) end++;
      if (start >= dot || end <= dot + 1) continue;
      long tld = psl->find_valid_tld_suffix(c + start, end - start);
      if (tld < 0) continue;
      if (tld == 0) continue;
      if (require_word_boundaries) {
        if (start > 0 && !is_boundary(c[start - 1])) continue;
        if (end < n && !is_boundary(c[end])) continue;
      }
      if (is_valid_domain(c + start, end - start)) {
        if (!valid_utf8(c + start, end - start)) continue;
        Item it; it.type = T_DOMAIN; it.start = start; it.end = end;
        out.push_back(it);
        last_domain_end = end;
      }
    }
  }

  // lib.rs:1212-1250 ; HashType::from_len :160-169
  void extract_hashes_chunk(const uint8_t* c, const std::vector<size_t>& bounds, std::vector<Item>& out) const {
    for (size_t k = 0; k + 1 < bounds.size(); k += 2) {
      size_t s = bounds[k], e = bounds[k + 1], len = e - s;
      uint8_t t;
      switch (len) {
        case 32: t = T_MD5; break; case 40: t = T_SHA1; break; case 64: t = T_SHA256; break;
        case 96: t = T_SHA384; break; case 128: t = T_SHA512; break;
        default: continue;
      }
      bool all = true;
      for (size_t j = s; j < e; j++) if (!is_hex(c[j])) { all = false; break; }
      if (!all) continue;
      Item it; it.type = t; it.start = s; it.end = e;
      out.push_back(it);
    }
  }

  // lib.rs:1269-1319 — words of 26..62 bytes: "bc1…" -> bech32, '1…' / '3…' -> Base58Check
  void extract_bitcoin_chunk(const uint8_t* c, const std::vector<size_t>& bounds, std::vector<Item>& out) const {
    for (size_t k = 0; k + 1 < bounds.size(); k += 2) {
      size_t s = bounds[k], e = bounds[k + 1], len = e - s;
      if (len < 26 || len > 62) continue;
      bool ok = false;
      if (c[s] == 'b' && c[s + 1] == 'c' && c[s + 2] == '1') ok = valid_utf8(c + s, len) && cryptoaddr::bitcoin_bech32(c + s, len);
      else if (c[s] == '1' || c[s] == '3') ok = valid_utf8(c + s, len) && cryptoaddr::bitcoin_base58(c + s, len);
      if (ok) { Item it; it.type = T_BITCOIN; it.start = s; it.end = e; out.push_back(it); }
    }
  }
  // lib.rs:1328-1361 — every "0x" followed by 40 hex digits, boundary (or chunk edge) on both sides
  void extract_ethereum_chunk(const uint8_t* c, size_t n, std::vector<Item>& out) const {
    for (size_t s = 0; s + 1 < n;) {
      if (!(c[s] == '0' && c[s + 1] == 'x')) { s++; continue; }
      size_t at = s;
      s += 2;  // memmem find_iter: non-overlapping occurrences
      if (at + 42 > n) continue;
      if (require_word_boundaries && at > 0 && !is_boundary(c[at - 1])) continue;
      if (require_word_boundaries && at + 42 < n && !is_boundary(c[at + 42])) continue;
      bool hex = true;
      for (size_t j = at + 2; j < at + 42; j++) if (!is_hex(c[j])) { hex = false; break; }
      if (!hex) continue;
      if (cryptoaddr::ethereum(c + at)) { Item it; it.type = T_ETHEREUM; it.start = at; it.end = at + 42; out.push_back(it); }
    }
  }
  // lib.rs:1367-1409 — words of 90..110 bytes that start with '4' or '8'
  void extract_monero_chunk(const uint8_t* c, const std::vector<size_t>& bounds, std::vector<Item>& out) const {
    for (size_t k = 0; k + 1 < bounds.size(); k += 2) {
      size_t s = bounds[k], e = bounds[k + 1], len = e - s;
      if (len < 90 || len > 110) continue;
      if (c[s] != '4' && c[s] != '8') continue;
      if (valid_utf8(c + s, len) && cryptoaddr::monero(c + s, len)) { Item it; it.type = T_MONERO; it.start = s; it.end = e; out.push_back(it); }
    }
  }

  // lib.rs:409-488 — order: IPv6, IPv4, email, domain, hash, bitcoin, ethereum, monero
  void extract_from_chunk(const uint8_t* c, size_t n, std::vector<Item>& out) const {
    std::vector<size_t> bounds, dots;
    if (flags & (X_HASHES | X_BITCOIN | X_MONERO)) find_word_boundaries(c, n, bounds);
    // dots: the reference collects once if both ipv4+domains, else lazily; identical positions either way
    if (flags & (X_IPV4 | X_DOMAINS)) {
      // memchr_iter(b'.', chunk) (lib.rs:432-447)
      for (const uint8_t* q = c; (q = (const uint8_t*)memchr(q, '.', (size_t)(c + n - q))) != nullptr; q++) dots.push_back((size_t)(q - c));
    }
    if (flags & X_IPV6) extract_ipv6_chunk(c, n, out);
    if (flags & X_IPV4) extract_ipv4_chunk(c, n, dots, out);
    if (flags & X_EMAILS) extract_emails_chunk(c, n, out);
    if (flags & X_DOMAINS) extract_domains_chunk(c, n, dots, out);
    if (flags & X_HASHES) extract_hashes_chunk(c, bounds, out);
    if (flags & X_BITCOIN) extract_bitcoin_chunk(c, bounds, out);
    if (flags & X_ETHEREUM) extract_ethereum_chunk(c, n, out);
    if (flags & X_MONERO) extract_monero_chunk(c, bounds, out);
  }
};

// ===========================================================================================
// Database — crates/matchy/src/database.rs:649-713 (from_storage) and lookups
// ===========================================================================================
struct MatchRec {       // flat record, same fields as include/matchy_b200.h mgpu_match
  uint64_t offset;
  uint32_t len;
  uint8_t item_type, kind, prefix_len, reserved;
  uint32_t n_ids;
  uint32_t ids_index;
  uint32_t data_offset;
  uint32_t pad;
};
struct IdPair { uint32_t pattern_id, data_offset; };
static const uint32_t NO_DATA = 0xFFFFFFFFu;

struct Db {
  std::vector<uint8_t> owned;
  const uint8_t* data = nullptr;
  size_t len = 0;
  // header (mmdb/format.rs:35-87)
  uint32_t node_count = 0; int record_bits = 24; int ip_version = 4; size_t tree_size = 0;
  bool has_ip_header = false;
  int match_mode = 0;  // 0 CS, 1 CI
  // literal hash (matchy-literal-hash/src/lib.rs:382-457)
  bool has_literal = false;
  const uint8_t* lh = nullptr; size_t lh_len = 0;
  uint32_t lh_num_shards = 0, lh_strings_offset = 0, lh_strings_size = 0;
  size_t lh_table_start = 0, lh_mappings_start = 0;
  std::vector<uint32_t> lh_shard_offsets;
  bool lh_map_dense = false;  // mapping entry i has pattern_id i (checked at load) ⇒ O(1) index == linear scan result
  // paraglob (paraglob_offset.rs) + mappings (database.rs:1315-1394)
  bool has_glob = false;
  const uint8_t* pg = nullptr; size_t pg_len = 0;
  size_t pg_file_off = 0;             // file offset of the PARAGLOB buffer; with the page-aligned mmap of Database::open it
                                      // decides whether zerocopy's alignment checks pass (see aligned4 below)
  size_t map_off = 0, map_count = 0;  // absolute offset of u32 data offsets
  std::unordered_map<uint32_t, std::pair<uint32_t, uint32_t>> aclh;  // literal id → (abs offset in pg, count)
  Psl psl;
  std::string error;

  static size_t find_metadata_marker(const uint8_t* d, size_t n) {  // mmdb/format.rs:126-150
    static const uint8_t M[14] = {0xAB, 0xCD, 0xEF, 'M', 'a', 'x', 'M', 'i', 'n', 'd', '.', 'c', 'o', 'm'};
    if (n < 14) return (size_t)-1;
    size_t start = n > 128 * 1024 ? n - 128 * 1024 : 0, last = (size_t)-1;
    for (size_t i = start; i + 14 <= n; i++) if (memcmp(d + i, M, 14) == 0) last = i;
    return last;
  }

  static bool get_uint(const Val& map, const char* key, uint64_t& out) {
    auto it = map.m.find(key);
    if (it == map.m.end()) return false;
    const Val& v = it->second;
    if (v.t == Val::U16 || v.t == Val::U32 || v.t == Val::U64) { out = v.u; return true; }
    return false;
  }

  bool open(const uint8_t* d, size_t n) {
    data = d; len = n;
    size_t mk = find_metadata_marker(d, n);
    if (mk == (size_t)-1) { error = "no MMDB metadata marker"; return false; }
    Decoder dec(d + mk + 14, n - mk - 14);
    Val meta = dec.decode(0);
    if (!dec.ok || meta.t != Val::MAP) { error = "metadata decode failed"; return false; }
    uint64_t nc, rs, ipv;
    if (!get_uint(meta, "node_count", nc) || !get_uint(meta, "record_size", rs) || !get_uint(meta, "ip_version", ipv)) {
      error = "metadata missing fields"; return false;
    }
    node_count = (uint32_t)nc; record_bits = (int)rs; ip_version = (int)ipv;
    if (record_bits != 24 && record_bits != 28 && record_bits != 32) { error = "bad record size"; return false; }
    if (ip_version != 4 && ip_version != 6) { error = "bad ip version"; return false; }
    tree_size = (size_t)node_count * (record_bits == 24 ? 6 : record_bits == 28 ? 7 : 8);
    has_ip_header = true;
    {  // read_match_mode_from_metadata database.rs:1398-1415 (Uint16 only)
      auto it = meta.m.find("match_mode");
      if (it != meta.m.end() && it->second.t == Val::U16 && it->second.u == 1) match_mode = 1;
    }
    // find_pattern_section_fast database.rs:1218-1236 (Uint32 only)
    {
      auto it = meta.m.find("pattern_section_offset");
      if (it != meta.m.end() && it->second.t == Val::U32 && it->second.u != 0) {
        size_t off = (size_t)it->second.u;
        if (off + 8 > n) { error = "pattern section header truncated"; return false; }
        size_t pg_size = rd32(d + off + 4);
        size_t pstart = off + 8, pend = pstart + pg_size;
        if (pend > n) { error = "paraglob beyond file"; return false; }
        if (pend + 4 > n) { error = "mappings truncated"; return false; }
        size_t cnt = rd32(d + pend);
        if (pend + 4 + cnt * 4 > n) { error = "mappings out of bounds"; return false; }
        pg = d + pstart; pg_len = pg_size; pg_file_off = pstart; map_off = pend + 4; map_count = cnt; has_glob = true;
        if (pg_len < 112 || memcmp(pg, "PARAGLOB", 8) != 0) { error = "bad paraglob magic"; return false; }
        load_aclh();
      }
    }
    // find_literal_section_fast database.rs:1255-1278 ; :702-710
    {
      auto it = meta.m.find("literal_section_offset");
      if (it != meta.m.end() && it->second.t == Val::U32 && it->second.u != 0) {
        size_t off = (size_t)it->second.u;  // already points after the 16-byte marker
        if (off + 32 > n) { error = "literal section truncated"; return false; }
        lh = d + off; lh_len = n - off;
        if (memcmp(lh, "LHSH", 4) != 0 || rd32(lh + 4) != 1) { error = "bad LHSH header"; return false; }
        lh_strings_offset = rd32(lh + 16); lh_strings_size = rd32(lh + 20); lh_num_shards = rd32(lh + 24);
        for (uint32_t i = 0; i <= lh_num_shards; i++) {
          if (32 + (size_t)i * 4 + 4 > lh_len) { error = "shard table truncated"; return false; }
          lh_shard_offsets.push_back(rd32(lh + 32 + i * 4));
        }
        lh_table_start = 32 + ((size_t)lh_num_shards + 1) * 4;
        lh_mappings_start = (size_t)lh_strings_offset + lh_strings_size;
        has_literal = true;
        // does mapping entry i carry pattern_id i?  (mmdb_builder.rs:562-565 writes them that way)
        if (lh_mappings_start + 4 <= lh_len) {
          uint32_t cnt = rd32(lh + lh_mappings_start);
          bool dense = lh_mappings_start + 4 + (size_t)cnt * 8 <= lh_len;
          for (uint32_t i = 0; dense && i < cnt; i++) if (rd32(lh + lh_mappings_start + 4 + (size_t)i * 8) != i) dense = false;
          lh_map_dense = dense;
        }
      }
    }
    return true;
  }

  // ACLH: matchy-paraglob/src/literal_hash.rs:218-333.  The reader hashes the literal id with
  // rustc-hash 2.1.1 FxHasher (absent from the tree, numerically unpinned — SURVEY §8(c)); every entry of a
  // builder-produced table is reachable from its home slot, so "scan all slots into a map" returns
  // exactly what the probing lookup returns.
  void load_aclh() {
    uint32_t off = rd32(pg + 96), cnt = rd32(pg + 100);
    if (off == 0 || cnt == 0 || (size_t)off + 24 > pg_len) return;
    const uint8_t* h = pg + off;
    if (memcmp(h, "ACLH", 4) != 0) return;
    uint32_t table_size = rd32(h + 12), patterns_offset = rd32(h + 16);
    for (uint32_t s = 0; s < table_size; s++) {
      size_t eo = (size_t)off + 24 + (size_t)s * 16;
      if (eo + 16 > pg_len) break;
      uint32_t lit = rd32(pg + eo);
      if (lit == 0xFFFFFFFFu) continue;
      uint32_t po = rd32(pg + eo + 4), pc = rd32(pg + eo + 8);
      size_t abs = (size_t)off + patterns_offset + po;
      if (abs + (size_t)pc * 4 > pg_len) continue;  // read_pattern_list → None → empty
      if (!aclh.count(lit)) aclh[lit] = {(uint32_t)abs, pc};
    }
  }

  // ---- IP tree: matchy-format/src/mmdb/tree.rs:46-277 ----
  bool read_record(uint32_t node, int side, uint32_t& rec) const {
    if (node >= node_count) return false;
    if (record_bits == 24) {
      size_t o = (size_t)node * 6 + side * 3;
      if (o + 3 > tree_size) return false;
      rec = (uint32_t(data[o]) << 16) | (uint32_t(data[o + 1]) << 8) | data[o + 2];
    } else if (record_bits == 28) {
      size_t o = (size_t)node * 7;
      if (o + 7 > tree_size) return false;
      const uint8_t* b = data + o;
      if (side == 0) rec = (uint32_t((b[3] >> 4) & 15) << 24) | (uint32_t(b[0]) << 16) | (uint32_t(b[1]) << 8) | b[2];
      else rec = (uint32_t(b[3] & 15) << 24) | (uint32_t(b[4]) << 16) | (uint32_t(b[5]) << 8) | b[6];
    } else {
      size_t o = (size_t)node * 8 + side * 4;
      if (o + 4 > tree_size) return false;
      rec = (uint32_t(data[o]) << 24) | (uint32_t(data[o + 1]) << 16) | (uint32_t(data[o + 2]) << 8) | data[o + 3];
    }
    return true;
  }

  // returns 1 found, 0 not found, -1 error (corrupt tree)
  int lookup_v4(uint32_t bits, uint32_t& data_off, uint8_t& prefix) const {
    uint32_t node = 0; uint32_t depth = 0;
    if (ip_version == 6) {  // find_ipv4_start_node tree.rs:258-277
      for (int k = 0; k < 96; k++) {
        uint32_t rec;
        if (!read_record(node, 0, rec)) return -1;
        if (rec == node_count) break;
        else if (rec < node_count) node = rec;
        else break;
      }
      depth = 96;
    }
    for (int bi = 0; bi < 32; bi++) {
      int bit = (bits >> (31 - bi)) & 1;
      uint32_t rec;
      if (!read_record(node, bit, rec)) return -1;
      if (rec == node_count) return 0;
      else if (rec < node_count) { node = rec; depth++; }
      else {
        uint32_t o = rec - node_count;
        if (o < 16) return -1;
        data_off = o - 16;
        prefix = (uint8_t)(depth >= 96 ? depth - 96 + 1 : depth + 1);
        return 1;
      }
    }
    return 0;
  }

  int lookup_v6(const uint16_t seg[8], uint32_t& data_off, uint8_t& prefix) const {
    uint32_t node = 0; uint32_t depth = 0;
    for (int bi = 0; bi < 128; bi++) {
      int bit = (seg[bi >> 4] >> (15 - (bi & 15))) & 1;
      uint32_t rec;
      if (!read_record(node, bit, rec)) return -1;
      if (rec == node_count) return 0;
      else if (rec < node_count) { node = rec; depth = bi + 1; }
      else {
        uint32_t o = rec - node_count;
        if (o < 16) return -1;
        data_off = o - 16;
        prefix = (uint8_t)(depth + 1);
        return 1;
      }
    }
    return 0;
  }

  // ---- literal hash: matchy-literal-hash/src/lib.rs:467-575 ----
  static std::string ascii_lower(const uint8_t* s, size_t n) {
    std::string r((const char*)s, n);
    for (auto& ch : r) if (ch >= 'A' && ch <= 'Z') ch = char(ch + 32);
    return r;
  }

  bool literal_lookup(const uint8_t* q, size_t n, uint32_t& pattern_id) const {
    std::string norm;
    if (match_mode == 1) {
      // Unicode to_lowercase in the reference; ASCII-only here (non-ASCII + CI is documented unsupported)
      norm = ascii_lower(q, n); q = (const uint8_t*)norm.data();
    }
    uint64_t h = xxh64(q, n, 0);
    size_t shard = (size_t)(h % lh_num_shards);
    size_t s0 = lh_shard_offsets[shard], s1 = lh_shard_offsets[shard + 1];
    size_t cap = s1 - s0;
    if (cap == 0) return false;
    size_t mask = cap - 1;
    size_t slot = s0 + ((size_t)h & mask);
    for (size_t it = 0; it < cap; it++) {
      size_t eo = lh_table_start + slot * 16;
      if (eo + 16 > lh_len) return false;
      uint64_t eh = rd64(lh + eo);
      uint32_t so = rd32(lh + eo + 8), pid = rd32(lh + eo + 12);
      if (so == 0xFFFFFFFFu) return false;
      if (eh == h) {
        size_t abs = (size_t)lh_strings_offset + so;  // read_string :528-543
        if (abs + 2 <= lh_len) {
          size_t sl = rd16(lh + abs);
          if (abs + 2 + sl <= lh_len && valid_utf8(lh + abs + 2, sl) && sl == n && memcmp(lh + abs + 2, q, n) == 0) {
            pattern_id = pid; return true;
          }
        }
      }
      slot = s0 + ((slot + 1 - s0) & mask);
    }
    return false;
  }

  bool literal_data_offset(uint32_t pattern_id, uint32_t& off) const {  // :546-575 (linear scan)
    if (lh_mappings_start + 4 > lh_len) return false;
    uint32_t cnt = rd32(lh + lh_mappings_start);
    size_t base = lh_mappings_start + 4;
    if (lh_map_dense) {  // identical result to the scan: entry i has id i
      if (pattern_id >= cnt) return false;
      off = rd32(lh + base + (size_t)pattern_id * 8 + 4);
      return true;
    }
    for (uint32_t i = 0; i < cnt; i++) {
      size_t o = base + (size_t)i * 8;
      if (o + 8 > lh_len) return false;
      if (rd32(lh + o) == pattern_id) { off = rd32(lh + o + 4); return true; }
    }
    return false;
  }

  // ---- paraglob: matchy-paraglob/src/paraglob_offset.rs:1028-1639 ----
  // find_ac_transition :1271-1353
  static bool ac_transition(const uint8_t* ac, size_t acn, size_t node_off, uint8_t ch, size_t& next) {
    if (node_off + 20 > acn) return false;
    const uint8_t* nd = ac + node_off;
    uint8_t kind = nd[0];
    if (kind == 0) return false;
    if (kind == 1) {
      if (nd[1] == ch) { next = rd32(nd + 12); return true; }
      return false;
    }
    if (kind == 2) {
      size_t eo = rd32(nd + 12), cnt = nd[2];
      if (eo + cnt * 8 > acn) return false;
      for (size_t i = 0; i < cnt; i++) {
        uint8_t ec = ac[eo + i * 8];
        if (ec == ch) { next = rd32(ac + eo + i * 8 + 4); return true; }
        if (ec > ch) return false;
      }
      return false;
    }
    if (kind == 3) {
      size_t to = (size_t)rd32(nd + 12) + (size_t)ch * 4;
      if (to + 4 > acn) return false;
      uint32_t t = rd32(ac + to);
      if (t != 0) { next = t; return true; }
      return false;
    }
    return false;  // StateKind::from_u8 → None
  }

  // run_ac_matching_into_static :1186-1266
  void run_ac(const uint8_t* ac, size_t acn, const uint8_t* text, size_t n, std::vector<uint32_t>& lits) const {
    if (acn == 0 || n == 0) return;
    std::string lowered;
    if (match_mode == 1) { lowered = ascii_lower(text, n); text = (const uint8_t*)lowered.data(); }
    size_t cur = 0;
    for (size_t i = 0; i < n; i++) {
      uint8_t ch = text[i];
      for (;;) {
        size_t nx;
        if (ac_transition(ac, acn, cur, ch, nx)) { cur = nx; break; }
        if (cur == 0) break;
        if (cur + 20 > acn) break;
        cur = rd32(ac + cur + 8);
      }
      if (cur + 20 > acn) continue;
      const uint8_t* nd = ac + cur;
      uint8_t pc = nd[3];
      if (pc > 0) {
        size_t po = rd32(nd + 16);
        if (po + (size_t)pc * 4 <= acn) for (size_t k = 0; k < pc; k++) lits.push_back(rd32(ac + po + k * 4));
      }
    }
  }

  // decode one UTF-8 scalar of a valid &str; returns byte length
  static size_t char_at(const uint8_t* t, size_t n, size_t pos, uint32_t& cp) {
    uint8_t c = t[pos];
    if (c < 0x80) { cp = c; return 1; }
    if (c < 0xE0 && pos + 1 < n) { cp = ((c & 0x1F) << 6) | (t[pos + 1] & 0x3F); return 2; }
    if (c < 0xF0 && pos + 2 < n) { cp = ((c & 0x0F) << 12) | ((t[pos + 1] & 0x3F) << 6) | (t[pos + 2] & 0x3F); return 3; }
    if (pos + 3 < n) { cp = ((c & 0x07) << 18) | ((t[pos + 1] & 0x3F) << 12) | ((t[pos + 2] & 0x3F) << 6) | (t[pos + 3] & 0x3F); return 4; }
    cp = c; return 1;
  }
  static uint32_t lower_cp(uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }
  static bool valid_scalar(uint32_t c) { return c <= 0x10FFFF && !(c >= 0xD800 && c <= 0xDFFF); }

  // zerocopy::Ref::<_, T>::from_prefix fails unless the slice is aligned for T (all three glob structs hold u32s and are
  // not `Unaligned`, offset_format.rs:391-431).  The file is mmap'd page-aligned, so alignment follows from file offsets.
  bool aligned4(size_t pg_offset) const { return ((pg_file_off + pg_offset) & 3) == 0; }

  // match_segments_impl :1402-1639.  Returns 1 true, 0 false, -1 Err (propagates like `?`)
  int match_segments(const uint8_t* text, size_t tn, size_t first_seg, size_t seg_count, size_t tpos, size_t seg_idx,
                     size_t& steps) const {
    if (steps == 0) return 0;
    steps--;
    if (seg_idx >= seg_count) return tpos >= tn ? 1 : 0;
    size_t so = first_seg + seg_idx * 12;
    if (so + 12 > pg_len) return 0;
    if (!aligned4(so)) return -1;  // "Invalid GlobSegmentHeader" :1432-1435
    const uint8_t* sh = pg + so;
    uint8_t stype = sh[0], sflags = sh[1];
    size_t data_len = rd32(sh + 4), data_off = rd32(sh + 8);
    switch (stype) {
      case 0: {
        if (data_off + data_len > pg_len) return 0;
        const uint8_t* lit = pg + data_off;
        if (!valid_utf8(lit, data_len)) return -1;
        bool m; size_t adv;
        if (match_mode == 0) {
          m = tn - tpos >= data_len && memcmp(text + tpos, lit, data_len) == 0;
          adv = data_len;
        } else {  // :1456-1478
          size_t lp = 0, tp = tpos, matched = 0; m = true;
          while (tp < tn) {
            if (lp < data_len) {
              uint32_t lc, tc;
              size_t ll = char_at(lit, data_len, lp, lc), tl = char_at(text, tn, tp, tc);
              bool eq = (lc < 128 && tc < 128) ? lower_cp(lc) == lower_cp(tc) : lc == tc;
              if (!eq) { m = false; break; }
              lp += ll; tp += tl; matched += tl;
            } else break;
          }
          if (m && lp < data_len) m = false;
          adv = matched;
        }
        if (!m) return 0;
        return match_segments(text, tn, first_seg, seg_count, tpos + adv, seg_idx + 1, steps);
      }
      case 1: {
        if (seg_idx + 1 >= seg_count) return 1;
        size_t pos = tpos;
        for (;;) {
          int r = match_segments(text, tn, first_seg, seg_count, pos, seg_idx + 1, steps);
          if (r < 0) return r;
          if (r == 1) return 1;
          if (pos >= tn) break;
          uint32_t cp;
          pos += char_at(text, tn, pos, cp);
        }
        return 0;
      }
      case 2: {
        if (tpos >= tn) return 0;
        uint32_t cp;
        size_t l = char_at(text, tn, tpos, cp);
        return match_segments(text, tn, first_seg, seg_count, tpos + l, seg_idx + 1, steps);
      }
      case 3: {
        if (tpos >= tn) return 0;
        uint32_t ch;
        size_t l = char_at(text, tn, tpos, ch);
        uint32_t chn = match_mode == 1 ? lower_cp(ch) : ch;
        size_t item_count = data_len / 12;
        if (data_off + data_len > pg_len) return 0;
        bool negated = sflags & 1, in_class = false;
        for (size_t i = 0; i < item_count; i++) {
          if (!aligned4(data_off + i * 12)) return -1;  // "Invalid CharClassItemEncoded" :1572-1577
          const uint8_t* it = pg + data_off + i * 12;
          uint8_t itype = it[0];
          uint32_t c1 = rd32(it + 4), c2 = rd32(it + 8);
          bool mi = false;
          if (itype == 0) {
            if (valid_scalar(c1)) mi = chn == (match_mode == 1 ? lower_cp(c1) : c1);
          } else if (itype == 1) {
            if (valid_scalar(c1) && valid_scalar(c2)) {
              uint32_t a = match_mode == 1 ? lower_cp(c1) : c1, b = match_mode == 1 ? lower_cp(c2) : c2;
              mi = chn >= a && chn <= b;
            }
          }
          if (mi) { in_class = true; break; }
        }
        bool m = negated ? !in_class : in_class;
        if (!m) return 0;
        return match_segments(text, tn, first_seg, seg_count, tpos + l, seg_idx + 1, steps);
      }
      default: return 0;
    }
  }

  // match_glob_from_buffer :1364-1398
  bool match_glob(uint32_t pattern_id, const uint8_t* text, size_t tn) const {
    size_t gso = rd32(pg + 104);
    size_t io = gso + (size_t)pattern_id * 8;
    if (io + 8 > pg_len) return false;
    if (!aligned4(io)) return false;  // Err("Invalid GlobSegmentIndex") is not Ok(true)
    size_t first = rd32(pg + io), count = rd16(pg + io + 4);
    size_t steps = 100000;
    return match_segments(text, tn, first, count, 0, 0, steps) == 1;
  }

  // find_all :1028-1182
  void find_all(const uint8_t* text, size_t tn, std::vector<uint32_t>& result) const {
    result.clear();
    if (pg_len < 112) return;
    size_t ac_start = rd32(pg + 20), ac_size = rd32(pg + 24);
    // (the reference collects literal ids and candidate patterns in hash sets, :1046-1095; per-thread vectors that are sorted
    //  and deduplicated give the same sets without an allocation per token)
    static thread_local std::vector<uint32_t> lits, cands;
    lits.clear(); cands.clear();
    if (ac_size > 0 && ac_start + ac_size <= pg_len) {
      run_ac(pg + ac_start, ac_size, text, tn, lits);
      std::sort(lits.begin(), lits.end());
      lits.erase(std::unique(lits.begin(), lits.end()), lits.end());
      for (uint32_t lit : lits) {
        auto it = aclh.find(lit);
        if (it == aclh.end()) continue;
        for (uint32_t k = 0; k < it->second.second; k++) cands.push_back(rd32(pg + it->second.first + (size_t)k * 4));
      }
      std::sort(cands.begin(), cands.end());
      cands.erase(std::unique(cands.begin(), cands.end()), cands.end());
    }
    size_t unaligned = (size_t)rd32(pg + 40) + rd32(pg + 44);
    size_t wild_off = unaligned + ((8 - (unaligned % 8)) % 8);
    size_t wild_count = rd32(pg + 60);
    size_t patterns_offset = rd32(pg + 36);
    for (size_t i = 0; i < wild_count; i++) {
      size_t wo = wild_off + i * 8;
      if (wo + 8 > pg_len) continue;
      uint32_t pid = rd32(pg + wo);
      size_t eo = patterns_offset + (size_t)pid * 16;
      if (eo + 16 > pg_len) continue;
      if (match_glob(pid, text, tn)) result.push_back(pid);
    }
    for (uint32_t pid : cands) {
      size_t eo = patterns_offset + (size_t)pid * 16;
      if (eo + 16 > pg_len) continue;
      uint32_t entry_id = rd32(pg + eo);
      uint8_t ptype = pg[eo + 4];
      if (ptype == 0) result.push_back(entry_id);
      else if (match_glob(entry_id, text, tn)) result.push_back(entry_id);
    }
    std::sort(result.begin(), result.end());
    result.erase(std::unique(result.begin(), result.end()), result.end());
  }

  bool glob_data_offset(uint32_t pid, uint32_t& off) const {  // database.rs:212-228
    if (pid >= map_count) return false;
    size_t p = map_off + (size_t)pid * 4;
    if (p + 4 > len) return false;
    off = rd32(data + p);
    return true;
  }

  // lookup_string_uncached database.rs:911-981.  returns false for Ok(None)/NotFound
  bool lookup_string(const uint8_t* q, size_t n, std::vector<IdPair>& ids) const {
    ids.clear();
    if (has_literal) {
      uint32_t pid, off;
      if (literal_lookup(q, n, pid) && literal_data_offset(pid, off)) ids.push_back({pid, off});
    }
    if (has_glob) {
      std::vector<uint32_t> g;
      find_all(q, n, g);
      for (uint32_t pid : g) {
        uint32_t off;
        if (glob_data_offset(pid, off)) ids.push_back({pid, off});
        else ids.push_back({pid, NO_DATA});
      }
    }
    return !ids.empty();
  }

  Val decode_data(uint32_t off, bool& ok) const {  // decode_ip_data database.rs:1005-1020
    size_t ds = tree_size + 16;
    if (ds > len) { ok = false; return Val(); }
    Decoder dec(data + ds, len - ds);
    Val v = dec.decode(off);
    ok = dec.ok;
    return v;
  }
};

// ===========================================================================================
// Worker::process_bytes — crates/matchy/src/processing/mod.rs:353-448
// ===========================================================================================
struct Counters {  // WorkerStats mod.rs:86-128 (timing fields omitted)
  uint64_t lines = 0, bytes = 0, candidates = 0, matches = 0;
  uint64_t by_type[12] = {0};
};

struct Scan {
  std::vector<MatchRec> recs;
  std::vector<IdPair> ids;
  Counters c;
  bool error = false;
};

// The reference's thread-local LRU query cache (database.rs:32-37; lookup() :725-804, lookup_ip() :837-886): every Some(result),
// NotFound included, is stored under the query text (IP lookups: under the address).  `matchy match` opens its database with
// --cache-size 10000 by default (bin/matchy.rs).  Results are what the uncached lookups return, so only the time changes.
struct LruCache {
  struct Val { uint8_t kind; uint8_t prefix_len; uint32_t data_offset; std::vector<IdPair> ids; };  // kind 0 NotFound, 1 Ip, 2 Pattern
  size_t cap;
  std::list<std::pair<std::string, Val>> order;  // most recent first
  std::unordered_map<std::string, std::list<std::pair<std::string, Val>>::iterator> index;
  explicit LruCache(size_t c) : cap(c) {}
  const Val* get(const std::string& key) {
    auto it = index.find(key);
    if (it == index.end()) return nullptr;
    order.splice(order.begin(), order, it->second);
    return &it->second->second;
  }
  void put(const std::string& key, Val v) {
    auto it = index.find(key);
    if (it != index.end()) { it->second->second = std::move(v); order.splice(order.begin(), order, it->second); return; }
    order.emplace_front(key, std::move(v));
    index[key] = order.begin();
    if (order.size() > cap) { index.erase(order.back().first); order.pop_back(); }
  }
};

static void process_bytes(const Db& db, const Extractor& ex, const uint8_t* d, size_t n, uint64_t base, Scan& out, LruCache* cache = nullptr) {
  for (const uint8_t* q = d; (q = (const uint8_t*)memchr(q, '\n', (size_t)(d + n - q))) != nullptr; q++) out.c.lines++;  // memchr::memchr_iter(b'\n', data).count() (mod.rs:357)
  out.c.bytes += n;
  std::vector<Item> items;
  ex.extract_from_chunk(d, n, items);
  std::vector<IdPair> ids;
  for (const Item& it : items) {
    out.c.candidates++;
    out.c.by_type[it.type]++;
    MatchRec r{};
    r.offset = base + it.start; r.len = (uint32_t)(it.end - it.start); r.item_type = it.type;
    if (it.type == T_IPV4 || it.type == T_IPV6) {  // lookup_extracted database.rs:889-901 → lookup_ip
      if (!db.has_ip_header) continue;
      uint32_t off = 0; uint8_t pl = 0;
      int rc;
      std::string key;
      const LruCache::Val* hit = nullptr;
      if (cache) {  // (keyed by the address: one-to-one with the canonical text the reference keys on, database.rs:839)
        key = it.type == T_IPV4 ? std::string((const char*)&it.v4, 4) : std::string((const char*)it.v6, 16);
        hit = cache->get(key);
      }
      if (hit) { rc = hit->kind == 1; off = hit->data_offset; pl = hit->prefix_len; }
      else {
        rc = it.type == T_IPV4 ? db.lookup_v4(it.v4, off, pl) : db.lookup_v6(it.v6, off, pl);
        if (rc < 0) { out.error = true; return; }  // Err aborts the chunk (mod.rs:414-416)
        if (cache) cache->put(key, LruCache::Val{(uint8_t)(rc ? 1 : 0), pl, off, {}});
      }
      if (rc == 0) continue;
      r.kind = 1; r.prefix_len = pl; r.data_offset = off; r.n_ids = 0; r.ids_index = 0;
    } else {
      // lookup(): query.parse::<IpAddr>() first (database.rs:760).  Tokens never contain ':' so only the
      // IPv4 form could parse; PSL-validated domains / e-mails / hex hashes never do, kept for literalness.
      uint32_t v4;
      if (parse_ipv4_strict(d + it.start, it.end - it.start, v4)) {
        uint32_t off = 0; uint8_t pl = 0;
        int rc = db.lookup_v4(v4, off, pl);
        if (rc < 0) { out.error = true; return; }
        if (rc == 0) continue;
        r.kind = 1; r.prefix_len = pl; r.data_offset = off;
      } else {
        if (!db.has_literal && !db.has_glob) continue;  // Ok(None)
        if (cache) {
          const std::string key((const char*)d + it.start, it.end - it.start);
          if (const LruCache::Val* hit = cache->get(key)) {
            if (hit->kind != 2) continue;
            ids = hit->ids;
          } else {
            const bool found = db.lookup_string(d + it.start, it.end - it.start, ids);
            cache->put(key, LruCache::Val{(uint8_t)(found ? 2 : 0), 0, 0, found ? ids : std::vector<IdPair>()});
            if (!found) continue;
          }
        } else if (!db.lookup_string(d + it.start, it.end - it.start, ids)) continue;
        r.kind = 2; r.n_ids = (uint32_t)ids.size(); r.ids_index = (uint32_t)out.ids.size();
        r.data_offset = NO_DATA;
        out.ids.insert(out.ids.end(), ids.begin(), ids.end());
      }
    }
    out.c.matches++;
    out.recs.push_back(r);
  }
}

// FileReader::next_batch — processing/mod.rs:206-251, driven over an in-memory "file"
static void scan_stream(const Db& db, const Extractor& ex, const uint8_t* d, size_t n, uint64_t base, size_t chunk_size,
                        Scan& out, LruCache* cache = nullptr) {
  size_t rd = 0;           // bytes "read" so far
  size_t left_start = 0;   // leftover = d[left_start..rd)
  for (;;) {
    size_t got = std::min(chunk_size, n - rd);
    if (got == 0) {
      if (left_start < rd) process_bytes(db, ex, d + left_start, rd - left_start, base + left_start, out, cache);
      return;
    }
    rd += got;
    // memrchr('\n', combined)
    size_t pos = rd;
    bool found = false;
    while (pos > left_start) { if (d[pos - 1] == '\n') { found = true; break; } pos--; }
    if (!found) continue;  // leftover = combined; read more
    process_bytes(db, ex, d + left_start, pos - left_start, base + left_start, out, cache);
    if (out.error) return;
    left_start = pos;
  }
}

// format_cidr_into — bin/cli_utils.rs:107-141 (matched_text is re-parsed with IpAddr::from_str)
static std::string format_cidr(const uint8_t* t, size_t n, uint8_t prefix_len) {
  uint32_t v4; uint16_t v6[8]; char b[64];
  if (parse_ipv4_strict(t, n, v4)) {
    uint32_t mask = prefix_len == 0 ? 0u : (prefix_len >= 32 ? 0xFFFFFFFFu : (~0u << (32 - prefix_len)));
    uint32_t net = v4 & mask;
    snprintf(b, sizeof b, "%u.%u.%u.%u/%u", net >> 24, (net >> 16) & 255, (net >> 8) & 255, net & 255, prefix_len);
    return b;
  }
  if (memchr(t, '.', n) == nullptr && parse_ipv6(t, n, v6)) {
    for (int bit = 0; bit < 128; bit++) if (bit >= prefix_len) v6[bit >> 4] &= (uint16_t)~(1u << (15 - (bit & 15)));
    return ipv6_display(v6) + "/" + std::to_string(prefix_len);
  }
  return std::string((const char*)t, n) + "/" + std::to_string(prefix_len);
}

// library_match_to_cli_match + output_cli_match — bin/match_processor/parallel.rs:297-369
static bool format_ndjson(const Db& db, const Scan& sc, const MatchRec& r, const uint8_t* text, const char* source,
                          std::string& out) {
  std::string mt((const char*)text, r.len);
  out.push_back('{');
  if (r.kind == 1) {
    bool ok = true;
    Val v = db.decode_data(r.data_offset, ok);
    if (!ok) return false;
    out += "\"cidr\":"; json_escape(format_cidr(text, r.len, r.prefix_len), out);
    out += ",\"data\":"; json_val(v, out);
    out += ",\"match_type\":\"ip\",\"matched_text\":"; json_escape(mt, out);
    out += ",\"prefix_len\":" + std::to_string(r.prefix_len);
  } else {
    std::string arr;
    size_t present = 0;
    for (uint32_t k = 0; k < r.n_ids; k++) {
      const IdPair& p = sc.ids[r.ids_index + k];
      if (p.data_offset == NO_DATA) continue;
      bool ok = true;
      Val v = db.decode_data(p.data_offset, ok);
      if (!ok) return false;
      if (present) arr.push_back(',');
      json_val(v, arr);
      present++;
    }
    if (present) { out += "\"data\":[" + arr + "],"; }
    out += "\"match_type\":\"pattern\",\"matched_text\":"; json_escape(mt, out);
    out += ",\"pattern_count\":" + std::to_string(r.n_ids);
  }
  out += ",\"source\":"; json_escape(source, out);
  out += ",\"timestamp\":\"0.000\"}";
  return true;
}

}  // namespace orc

// ===========================================================================================
// C ABI for ctypes (tests/, bench cpu_baseline)
// ===========================================================================================
using namespace orc;

struct orc_handle {
  Db db;
  Extractor ex;
  Scan scan;
  std::string ndjson;
  size_t cache_capacity = 10000;  // `matchy match --cache-size` default; 0 = no cache (orc_set_cache)
};

extern "C" {

orc_handle* orc_open(const uint8_t* mxy, size_t len, const char* psl_path, int copy) {
  auto* h = new orc_handle();
  if (!h->db.psl.load(psl_path)) { h->db.error = "cannot load PSL"; return h; }
  const uint8_t* p = mxy;
  if (copy) { h->db.owned.assign(mxy, mxy + len); p = h->db.owned.data(); }
  h->db.open(p, len);
  h->ex.psl = &h->db.psl;
  return h;
}
const char* orc_error(orc_handle* h) { return h->db.error.c_str(); }
// digests of the crypto-address validators, for known-answer tests (what: 0 SHA-256, 1 Keccak-256)
void orc_digest(int what, const uint8_t* msg, size_t len, uint8_t out[32]) { if (what == 0) cryptoaddr::sha256(msg, len, out); else cryptoaddr::keccak256(msg, len, out); }
// 1 valid / 0 invalid (what: 0 bitcoin Base58Check, 1 bitcoin bech32, 2 ethereum "0x"+40 hex, 3 monero)
int orc_cryptoaddr(int what, const uint8_t* s, size_t n) {
  switch (what) { case 0: return cryptoaddr::bitcoin_base58(s, n); case 1: return cryptoaddr::bitcoin_bech32(s, n); case 2: return n == 42 && cryptoaddr::ethereum(s); default: return cryptoaddr::monero(s, n); }
}
void orc_close(orc_handle* h) { delete h; }
// per-thread LRU query cache of the multi-threaded scans (0 disables it), like `matchy match --cache-size N`
void orc_set_cache(orc_handle* h, size_t capacity) { h->cache_capacity = capacity; }

// default extractor flags as `matchy match` derives them from DB capabilities (match_cmd.rs:277-303),
// with the crypto extractors off (--extractors=-crypto)
uint32_t orc_default_flags(orc_handle* h) {
  uint32_t f = 0;
  if (h->db.has_ip_header) f |= X_IPV4 | X_IPV6;
  if (h->db.has_literal || h->db.has_glob) f |= X_DOMAINS | X_EMAILS | X_HASHES;
  return f;
}
int orc_has(orc_handle* h, int what) { return what == 0 ? h->db.has_ip_header : what == 1 ? h->db.has_literal : h->db.has_glob; }

// scan a buffer the way `matchy match` would scan a file with these bytes (chunk_size = reader block size)
int orc_scan(orc_handle* h, const uint8_t* data, size_t len, uint64_t base, uint32_t flags, size_t chunk_size) {
  h->scan = Scan();
  h->ex.flags = flags;
  if (chunk_size == 0) process_bytes(h->db, h->ex, data, len, base, h->scan);
  else scan_stream(h->db, h->ex, data, len, base, chunk_size, h->scan);
  return h->scan.error ? -1 : 0;
}
size_t orc_n_matches(orc_handle* h) { return h->scan.recs.size(); }
const MatchRec* orc_matches(orc_handle* h) { return h->scan.recs.data(); }
size_t orc_n_ids(orc_handle* h) { return h->scan.ids.size(); }
const IdPair* orc_ids(orc_handle* h) { return h->scan.ids.data(); }
// counters: lines, bytes, candidates, matches, by_type[12]
void orc_counters(orc_handle* h, uint64_t* out16) {
  out16[0] = h->scan.c.lines; out16[1] = h->scan.c.bytes; out16[2] = h->scan.c.candidates; out16[3] = h->scan.c.matches;
  for (int k = 0; k < 12; k++) out16[4 + k] = h->scan.c.by_type[k];
}

// Multi-threaded scan for the CPU baseline: the buffer is cut into `threads` newline-aligned shards (the way
// `matchy match -j N` works over N files: one worker per file, processing/parallel.rs:355-372, 416-433),
// each shard streamed through FileReader::next_batch in 128 KiB reads.  Only the summed counters are kept.
int orc_scan_mt(orc_handle* h, const uint8_t* data, size_t len, uint32_t flags, int threads, uint64_t* out16) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  std::vector<size_t> cuts{0};
  for (int t = 1; t < threads; t++) {
    size_t p = len / threads * t;
    if (p < cuts.back()) p = cuts.back();
    while (p < len && data[p] != '\n') p++;
    if (p < len) p++;
    cuts.push_back(p);
  }
  cuts.push_back(len);
  std::vector<Scan> scans(threads);
  std::vector<std::thread> th;
  Extractor ex = h->ex;
  ex.flags = flags;
  for (int t = 0; t < threads; t++) {
    th.emplace_back([&, t]() {
      LruCache cache(h->cache_capacity);
      if (cuts[t + 1] > cuts[t]) scan_stream(h->db, ex, data + cuts[t], cuts[t + 1] - cuts[t], cuts[t], 128 * 1024, scans[t], h->cache_capacity ? &cache : nullptr);
    });
  }
  for (auto& x : th) x.join();
  for (int k = 0; k < 16; k++) out16[k] = 0;
  int rc = 0;
  for (auto& s : scans) {
    if (s.error) rc = -1;
    out16[0] += s.c.lines; out16[1] += s.c.bytes; out16[2] += s.c.candidates; out16[3] += s.c.matches;
    for (int k = 0; k < 12; k++) out16[4 + k] += s.c.by_type[k];
  }
  return rc;
}

// The same sharded scan, keeping the records: h->scan afterwards holds every shard's records sorted by
// (offset, item_type, len) with the id pairs re-packed in record order and offsets made absolute with `base` — the layout
// mgpu_results() returns, so a multi-GiB record-exact comparison is two array compares (tests/, bench.py parity gate).
int orc_scan_mt_keep(orc_handle* h, const uint8_t* data, size_t len, uint64_t base, uint32_t flags, int threads, uint64_t* out16) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  std::vector<size_t> cuts{0};
  for (int t = 1; t < threads; t++) {
    size_t p = len / threads * t;
    if (p < cuts.back()) p = cuts.back();
    while (p < len && data[p] != '\n') p++;
    if (p < len) p++;
    cuts.push_back(p);
  }
  cuts.push_back(len);
  std::vector<Scan> scans(threads);
  std::vector<std::thread> th;
  Extractor ex = h->ex;
  ex.flags = flags;
  for (int t = 0; t < threads; t++) {
    th.emplace_back([&, t]() {
      LruCache cache(h->cache_capacity);
      if (cuts[t + 1] > cuts[t]) scan_stream(h->db, ex, data + cuts[t], cuts[t + 1] - cuts[t], base + cuts[t], 128 * 1024, scans[t], h->cache_capacity ? &cache : nullptr);
    });
  }
  for (auto& x : th) x.join();
  for (int k = 0; k < 16; k++) out16[k] = 0;
  int rc = 0;
  h->scan = Scan();
  std::vector<MatchRec> all;
  std::vector<IdPair> ids_all;
  for (auto& s : scans) {
    if (s.error) rc = -1;
    out16[0] += s.c.lines; out16[1] += s.c.bytes; out16[2] += s.c.candidates; out16[3] += s.c.matches;
    for (int k = 0; k < 12; k++) out16[4 + k] += s.c.by_type[k];
    const uint32_t rebase = (uint32_t)ids_all.size();
    for (MatchRec r : s.recs) { if (r.kind == 2) r.ids_index += rebase; all.push_back(r); }
    ids_all.insert(ids_all.end(), s.ids.begin(), s.ids.end());
  }
  std::stable_sort(all.begin(), all.end(), [](const MatchRec& x, const MatchRec& y) {
    if (x.offset != y.offset) return x.offset < y.offset;
    if (x.item_type != y.item_type) return x.item_type < y.item_type;
    return x.len < y.len;
  });
  for (MatchRec& r : all) {
    if (r.kind == 2) {
      const uint32_t at = (uint32_t)h->scan.ids.size();
      h->scan.ids.insert(h->scan.ids.end(), ids_all.begin() + r.ids_index, ids_all.begin() + r.ids_index + r.n_ids);
      r.ids_index = at;
    }
  }
  h->scan.recs.swap(all);
  h->scan.c.lines = out16[0]; h->scan.c.bytes = out16[1]; h->scan.c.candidates = out16[2]; h->scan.c.matches = out16[3];
  for (int k = 0; k < 12; k++) h->scan.c.by_type[k] = out16[4 + k];
  return rc;
}

// extraction only: returns count; items as (type, start, end) triples of uint64
size_t orc_extract(orc_handle* h, const uint8_t* data, size_t len, uint32_t flags, uint64_t* out, size_t cap) {
  h->ex.flags = flags;
  std::vector<Item> items;
  h->ex.extract_from_chunk(data, len, items);
  size_t n = std::min(cap, items.size());
  for (size_t k = 0; k < n; k++) { out[3 * k] = items[k].type; out[3 * k + 1] = items[k].start; out[3 * k + 2] = items[k].end; }
  return items.size();
}

// NDJSON for the last scan (records in reference order); `data` must be the scanned buffer, `base` its offset
const char* orc_ndjson(orc_handle* h, const uint8_t* data, uint64_t base, const char* source, size_t* out_len) {
  h->ndjson.clear();
  for (const MatchRec& r : h->scan.recs) {
    std::string line;
    if (!format_ndjson(h->db, h->scan, r, data + (r.offset - base), source, line)) continue;
    h->ndjson += line; h->ndjson.push_back('\n');
  }
  *out_len = h->ndjson.size();
  return h->ndjson.c_str();
}

// single-query helpers (for KAT tests of the lookup layers)
int orc_lookup_ip4(orc_handle* h, uint32_t addr, uint32_t* data_off, uint8_t* prefix) { return h->db.lookup_v4(addr, *data_off, *prefix); }
int orc_lookup_ip6(orc_handle* h, const uint16_t* seg, uint32_t* data_off, uint8_t* prefix) { return h->db.lookup_v6(seg, *data_off, *prefix); }
int orc_parse_ipv6(const uint8_t* s, size_t n, uint16_t* seg) { return parse_ipv6(s, n, seg) ? 1 : 0; }
// the record reader on bare tree bytes (pinned by the reference's vectors, mmdb/tree.rs:323-398); -1 = out of range
int64_t orc_tree_record(const uint8_t* tree, size_t tree_size, uint32_t node_count, int record_bits, uint32_t node, int side) {
  Db d;
  d.data = tree; d.tree_size = tree_size; d.node_count = node_count; d.record_bits = record_bits;
  uint32_t rec = 0;
  return d.read_record(node, side, rec) ? (int64_t)rec : -1;
}
size_t orc_lookup_string(orc_handle* h, const uint8_t* q, size_t n, uint32_t* out_pairs, size_t cap_pairs) {
  std::vector<IdPair> ids;
  h->db.lookup_string(q, n, ids);
  for (size_t k = 0; k < ids.size() && k < cap_pairs; k++) { out_pairs[2 * k] = ids[k].pattern_id; out_pairs[2 * k + 1] = ids[k].data_offset; }
  return ids.size();
}
const char* orc_data_json(orc_handle* h, uint32_t data_off, size_t* out_len) {
  bool ok = true;
  Val v = h->db.decode_data(data_off, ok);
  h->ndjson.clear();
  if (ok) json_val(v, h->ndjson);
  *out_len = h->ndjson.size();
  return h->ndjson.c_str();
}
uint64_t orc_xxh64(const uint8_t* p, size_t n) { return xxh64(p, n, 0); }
const char* orc_ipv6_display(const uint16_t* seg) { static thread_local std::string s; s = ipv6_display(seg); return s.c_str(); }

}  // extern "C"
