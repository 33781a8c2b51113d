// mxy_reader.cpp — see mxy_reader.h.  Host code for matched records and for locating sections at upload.
#include "mxy_reader.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/matchy_b200.h"
#include "mxy_builder.h"  // parse_ipv4_text / parse_ipv6_text

namespace mxy {

static inline uint32_t le32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

// ---------------------------------------------------------------------------------------------------------
// MMDB value decoding
// ---------------------------------------------------------------------------------------------------------
bool ValueReader::payload_size(size_t& cur, uint8_t low5, size_t& out) const {
  if (low5 < 29) { out = low5; return true; }
  size_t extra = low5 == 29 ? 1 : low5 == 30 ? 2 : 3;
  if (cur + extra > n_) return false;
  size_t v = 0;
  for (size_t k = 0; k < extra; k++) v = (v << 8) | p_[cur + k];
  cur += extra;
  out = v + (low5 == 29 ? 29 : low5 == 30 ? 285 : 65821);
  return true;
}

bool ValueReader::at(size_t& cur, Value& out, int depth) const {
  if (depth > 96 || cur >= n_) return false;
  uint8_t ctrl = p_[cur++];
  unsigned type = ctrl >> 5;
  uint8_t low5 = ctrl & 31;
  out.mmdb_type = (uint8_t)(type ? type : (cur < n_ ? 7u + p_[cur] : 0u));
  auto be = [&](size_t len, uint64_t& v) {
    v = 0;
    for (size_t k = 0; k < len; k++) v = (v << 8) | p_[cur + k];
    cur += len;
  };
  size_t len;
  if (type == 0) {  // extended: next byte + 7
    if (cur >= n_) return false;
    type = 7u + p_[cur++];
    switch (type) {
      case 8: {
        if (!payload_size(cur, low5, len) || len > 4 || cur + len > n_) return false;
        int32_t v = 0;
        if (len > 0) {
          if (p_[cur] & 0x80) v = -1;
          for (size_t k = 0; k < len; k++) v = (int32_t)(((uint32_t)v << 8) | p_[cur + k]);
        }
        cur += len;
        out.kind = Value::INT; out.i = v;
        return true;
      }
      case 9:
        if (!payload_size(cur, low5, len) || len > 8 || cur + len > n_) return false;
        out.kind = Value::UINT; be(len, out.u);
        return true;
      case 10: {
        if (!payload_size(cur, low5, len) || len > 16 || cur + len > n_) return false;
        unsigned __int128 b = 0;
        for (size_t k = 0; k < len; k++) b = (b << 8) | p_[cur + k];
        cur += len;
        out.kind = Value::U128; out.big = b;
        return true;
      }
      case 11: {
        if (!payload_size(cur, low5, len)) return false;
        out.kind = Value::ARR;
        for (size_t k = 0; k < len; k++) {
          out.items.emplace_back();
          if (!at(cur, out.items.back(), depth + 1)) return false;
        }
        return true;
      }
      case 14: out.kind = Value::BOOL; out.b = low5 != 0; return true;
      case 15: {
        if (low5 != 4 || cur + 4 > n_) return false;
        uint64_t bits; be(4, bits);
        uint32_t b32 = (uint32_t)bits; float f; memcpy(&f, &b32, 4);
        out.kind = Value::F32; out.f = (double)f;
        return true;
      }
      default: return false;
    }
  }
  switch (type) {
    case 1: {  // pointer
      unsigned sz = (low5 >> 3) & 3;
      uint32_t low3 = low5 & 7;
      uint64_t v;
      if (sz == 0) { if (cur >= n_) return false; be(1, v); out.u = (low3 << 8) | v; }
      else if (sz == 1) { if (cur + 1 >= n_) return false; be(2, v); out.u = 2048 + ((low3 << 16) | v); }
      else if (sz == 2) { if (cur + 2 >= n_) return false; be(3, v); out.u = 526336 + ((low3 << 24) | v); }
      else { if (cur + 3 >= n_) return false; be(4, v); out.u = v; }
      out.kind = Value::PTR;
      return true;
    }
    case 2: case 4:
      if (!payload_size(cur, low5, len) || cur + len > n_) return false;
      out.kind = type == 2 ? Value::STR : Value::BYTES;
      out.s.assign((const char*)p_ + cur, len);
      cur += len;
      return true;
    case 3: {
      if (cur + 8 > n_) return false;
      uint64_t bits; be(8, bits);
      out.kind = Value::F64; memcpy(&out.f, &bits, 8);
      return true;
    }
    case 5: case 6:
      if (!payload_size(cur, low5, len) || len > (type == 5 ? 2u : 4u) || cur + len > n_) return false;
      out.kind = Value::UINT; be(len, out.u);
      return true;
    default: {  // 7 map: keys are strings or pointers to strings
      if (!payload_size(cur, low5, len)) return false;
      out.kind = Value::MAP;
      for (size_t k = 0; k < len; k++) {
        Value key;
        if (!at(cur, key, depth + 1)) return false;
        if (key.kind == Value::PTR) {
          Value target;
          if (!read((uint32_t)key.u, target) || target.kind != Value::STR) return false;
          key = target;
        } else if (key.kind != Value::STR) return false;
        out.fields.emplace_back(key.s, Value());
        if (!at(cur, out.fields.back().second, depth + 1)) return false;
      }
      return true;
    }
  }
}

bool ValueReader::chase(Value& v, int depth) const {
  if (depth > 64) return false;
  if (v.kind == Value::PTR) {
    size_t cur = (size_t)v.u;
    Value t;
    if (!at(cur, t, 0)) return false;
    v = t;
    return chase(v, depth + 1);
  }
  if (v.kind == Value::MAP) { for (auto& f : v.fields) if (!chase(f.second, depth + 1)) return false; }
  if (v.kind == Value::ARR) { for (auto& e : v.items) if (!chase(e, depth + 1)) return false; }
  return true;
}

bool ValueReader::read(uint32_t offset, Value& out) const {
  size_t cur = offset;
  out = Value();
  if (!at(cur, out, 0)) return false;
  return chase(out, 0);
}

// ---------------------------------------------------------------------------------------------------------
// JSON (serde_json 1.0 compact writer; maps are BTreeMap => keys in byte order, duplicate keys: last wins)
// ---------------------------------------------------------------------------------------------------------
void json_string(const std::string& s, std::string& out) {
  static const char* H = "0123456789abcdef";
  out += '"';
  for (unsigned char c : s) {
    if (c == '"') out += "\\\"";
    else if (c == '\\') out += "\\\\";
    else if (c >= 0x20) out += (char)c;
    else if (c == '\n') out += "\\n";
    else if (c == '\r') out += "\\r";
    else if (c == '\t') out += "\\t";
    else if (c == '\b') out += "\\b";
    else if (c == '\f') out += "\\f";
    else { out += "\\u00"; out += H[c >> 4]; out += H[c & 15]; }
  }
  out += '"';
}

// ryu shortest round-trip; non-finite -> null.  as_f32: the value is an f32 and the digits are the shortest that round-trip
// as f32 (serde's serialize_f32, what serde_json::to_string(&DataValue) prints: 0.98, not 0.9800000190734863).
static void json_f64(double d, std::string& out, bool as_f32 = false) {
  if (d != d || d == 1.0 / 0.0 || d == -1.0 / 0.0) { out += "null"; return; }
  char b[40];
  int prec = 1;
  if (as_f32) { for (; prec <= 9; prec++) { snprintf(b, sizeof b, "%.*e", prec - 1, d); if (strtof(b, nullptr) == (float)d) break; } }
  else for (; prec <= 17; prec++) { snprintf(b, sizeof b, "%.*e", prec - 1, d); if (strtod(b, nullptr) == d) break; }
  // b = d.ddddde±XX ; re-lay the digits the way ryu's pretty printer does
  std::string digits; int exp10 = 0; bool neg = false;
  {
    const char* p = b;
    if (*p == '-') { neg = true; p++; }
    for (; *p && *p != 'e'; p++) if (*p != '.') digits += *p;
    exp10 = atoi(p + 1);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  }
  if (neg) out += '-';
  int nd = (int)digits.size();
  int k = exp10 + 1;  // decimal point position relative to digits
  if (0 < k && k <= 16 && nd <= k) { out += digits; out.append(k - nd, '0'); out += ".0"; }
  else if (0 < k && k <= 16) { out.append(digits, 0, k); out += '.'; out.append(digits, k, std::string::npos); }
  else if (-5 < k && k <= 0) { out += "0."; out.append(-k, '0'); out += digits; }
  else {
    out += digits[0];
    if (nd > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
    out += 'e'; out += std::to_string(k - 1);
  }
}

void render_json(const Value& v, std::string& out, bool f32_shortest) {
  switch (v.kind) {
    case Value::STR: json_string(v.s, out); break;
    case Value::F64: json_f64(v.f, out); break;
    case Value::F32: json_f64(v.f, out, f32_shortest); break;
    case Value::BYTES:
      out += '[';
      for (size_t k = 0; k < v.s.size(); k++) { if (k) out += ','; out += std::to_string((unsigned)(uint8_t)v.s[k]); }
      out += ']';
      break;
    case Value::UINT: out += std::to_string(v.u); break;
    case Value::INT: out += std::to_string(v.i); break;
    case Value::U128: {
      char b[48]; int n = 0; unsigned __int128 x = v.big;
      if (x == 0) b[n++] = '0';
      while (x) { b[n++] = char('0' + (int)(x % 10)); x /= 10; }
      out += '"'; while (n) out += b[--n]; out += '"';
      break;
    }
    case Value::BOOL: out += v.b ? "true" : "false"; break;
    case Value::MAP: {
      std::vector<const std::pair<std::string, Value>*> order;
      for (auto& f : v.fields) order.push_back(&f);
      std::stable_sort(order.begin(), order.end(), [](auto* a, auto* b) { return a->first < b->first; });
      out += '{';
      bool first = true;
      for (size_t k = 0; k < order.size(); k++) {
        if (k + 1 < order.size() && order[k + 1]->first == order[k]->first) continue;  // later duplicate replaces earlier
        if (!first) out += ',';
        first = false;
        json_string(order[k]->first, out); out += ':'; render_json(order[k]->second, out, f32_shortest);
      }
      out += '}';
      break;
    }
    case Value::ARR:
      out += '[';
      for (size_t k = 0; k < v.items.size(); k++) { if (k) out += ','; render_json(v.items[k], out, f32_shortest); }
      out += ']';
      break;
    case Value::PTR: out += "\"<pointer>\""; break;
    case Value::NUL: out += "null"; break;
  }
}

// ---------------------------------------------------------------------------------------------------------
// cidr text
// ---------------------------------------------------------------------------------------------------------
std::string ipv6_text(const uint16_t g[8]) {
  char b[64];
  if (g[0] == 0 && g[1] == 0 && g[2] == 0 && g[3] == 0 && g[4] == 0 && g[5] == 0xffff) {
    snprintf(b, sizeof b, "::ffff:%u.%u.%u.%u", g[6] >> 8, g[6] & 255, g[7] >> 8, g[7] & 255);
    return b;
  }
  // longest run of zero groups (length >= 2), leftmost on ties
  int zs = -1, zl = 0;
  for (int i = 0; i < 8;) {
    if (g[i] != 0) { i++; continue; }
    int j = i;
    while (j < 8 && g[j] == 0) j++;
    if (j - i > zl) { zs = i; zl = j - i; }
    i = j;
  }
  std::string out;
  auto groups = [&](int a, int e) { for (int i = a; i < e; i++) { if (i > a) out += ':'; snprintf(b, sizeof b, "%x", g[i]); out += b; } };
  if (zl >= 2) { groups(0, zs); out += "::"; groups(zs + zl, 8); }
  else groups(0, 8);
  return out;
}

std::string cidr_text(const uint8_t* t, size_t n, unsigned plen) {
  uint32_t v4; uint16_t v6[8];
  if (parse_ipv4_text((const char*)t, n, v4)) {
    uint32_t mask = plen == 0 ? 0u : (0xFFFFFFFFu << ((32 - plen) & 31));
    uint32_t net = v4 & mask;
    char b[40];
    snprintf(b, sizeof b, "%u.%u.%u.%u/%u", net >> 24, (net >> 16) & 255, (net >> 8) & 255, net & 255, plen);
    return b;
  }
  if (parse_ipv6_text((const char*)t, n, v6)) {
    for (unsigned bit = plen; bit < 128; bit++) v6[bit >> 4] &= (uint16_t)~(1u << (15 - (bit & 15)));
    return ipv6_text(v6) + "/" + std::to_string(plen);
  }
  return std::string((const char*)t, n) + "/" + std::to_string(plen);
}

// ---------------------------------------------------------------------------------------------------------
// section locator
// ---------------------------------------------------------------------------------------------------------
static const Value* field(const Value& m, const char* k) {
  const Value* r = nullptr;
  for (auto& f : m.fields) if (f.first == k) r = &f.second;
  return r;
}

bool read_metadata(const uint8_t* d, size_t n, Value& out) {
  static const uint8_t M[14] = {0xAB, 0xCD, 0xEF, 'M', 'a', 'x', 'M', 'i', 'n', 'd', '.', 'c', 'o', 'm'};
  if (n < 14) return false;
  size_t from = n > 128 * 1024 ? n - 128 * 1024 : 0, mk = (size_t)-1;
  for (size_t i = from; i + 14 <= n; i++) if (d[i] == 0xAB && memcmp(d + i, M, 14) == 0) mk = i;
  if (mk == (size_t)-1) return false;
  ValueReader mr(d + mk + 14, n - mk - 14);
  return mr.read(0, out) && out.kind == Value::MAP;
}

bool locate_sections(const uint8_t* d, size_t n, Layout& L, std::string& err) {
  static const uint8_t M[14] = {0xAB, 0xCD, 0xEF, 'M', 'a', 'x', 'M', 'i', 'n', 'd', '.', 'c', 'o', 'm'};
  if (n < 14) { err = "file too small"; return false; }
  size_t from = n > 128 * 1024 ? n - 128 * 1024 : 0, mk = (size_t)-1;
  for (size_t i = from; i + 14 <= n; i++) if (d[i] == 0xAB && memcmp(d + i, M, 14) == 0) mk = i;
  if (mk == (size_t)-1) { err = "MMDB metadata marker not found"; return false; }
  ValueReader mr(d + mk + 14, n - mk - 14);
  Value meta;
  if (!mr.read(0, meta) || meta.kind != Value::MAP) { err = "metadata is not a map"; return false; }
  auto want = [&](const char* k, uint64_t& v) { const Value* f = field(meta, k); if (!f || f->kind != Value::UINT) return false; v = f->u; return true; };
  uint64_t nc, rs, ipv;
  if (!want("node_count", nc) || !want("record_size", rs) || !want("ip_version", ipv)) { err = "metadata lacks node_count/record_size/ip_version"; return false; }
  if (rs != 24 && rs != 28 && rs != 32) { err = "unsupported record_size"; return false; }
  if (ipv != 4 && ipv != 6) { err = "unsupported ip_version"; return false; }
  // every offset / count below comes from the file: checks are written so that no addition or product can wrap
  const uint64_t node_bytes = rs == 24 ? 6 : rs == 28 ? 7 : 8;
  if (nc > 0xFFFFFFFFull || nc > n / node_bytes) { err = "search tree larger than file"; return false; }
  L.node_count = (uint32_t)nc; L.record_bits = (uint32_t)rs; L.ip_version = (uint32_t)ipv;
  L.tree_size = nc * node_bytes;
  if (n - L.tree_size < 16) { err = "search tree larger than file"; return false; }
  L.data_start = L.tree_size + 16;
  uint64_t v;
  L.match_mode = (want("match_mode", v) && v == 1) ? 1 : 0;
  if (want("literal_entry_count", v)) L.literal_count = (uint32_t)v;
  if (want("glob_entry_count", v)) L.glob_count = (uint32_t)v;
  if (want("pattern_section_offset", v) && v != 0) {
    if (v > n || n - v < 8) { err = "pattern section header out of range"; return false; }
    uint64_t pg_size = le32(d + v + 4), pg0 = v + 8;
    if (n - pg0 < pg_size || n - pg0 - pg_size < 4) { err = "paraglob buffer out of range"; return false; }
    uint64_t pg1 = pg0 + pg_size;
    uint64_t cnt = le32(d + pg1);
    if ((n - pg1 - 4) / 4 < cnt) { err = "glob data offsets out of range"; return false; }
    if (pg_size < 112 || memcmp(d + pg0, "PARAGLOB", 8) != 0) { err = "bad PARAGLOB magic"; return false; }
    L.has_glob = true; L.pg_off = pg0; L.pg_len = pg_size; L.map_off = pg1 + 4; L.map_count = cnt;
  }
  if (want("literal_section_offset", v) && v != 0) {
    if (v > n || n - v < 32) { err = "literal section out of range"; return false; }
    if (memcmp(d + v, "LHSH", 4) != 0 || le32(d + v + 4) != 1) { err = "bad LHSH header"; return false; }
    L.has_literal = true; L.lit_off = v; L.lit_len = n - v;
  }
  return true;
}

}  // namespace mxy

// =========================================================================================================
// C ABI: mxyr_*
// =========================================================================================================
using namespace mxy;

struct mxyr_db {
  const uint8_t* d; size_t n;
  Layout L;
  std::string text;
};

extern "C" {

mxyr_db* mxyr_open(const uint8_t* mxy, size_t len) {
  auto* h = new mxyr_db{mxy, len, Layout(), std::string()};
  std::string err;
  if (!locate_sections(mxy, len, h->L, err)) { delete h; return nullptr; }
  return h;
}
void mxyr_close(mxyr_db* h) { delete h; }

size_t mxyr_data_json(mxyr_db* h, uint32_t off, const char** out) {
  h->text.clear();
  ValueReader r(h->d + h->L.data_start, h->n - h->L.data_start);
  Value v;
  if (r.read(off, v)) render_json(v, h->text);
  *out = h->text.c_str();
  return h->text.size();
}

// One renderer for the two modes of `matchy match`:
//   parallel   (bin/match_processor/parallel.rs:297-369)   timestamp "0.000", matched_text = the raw bytes of the span
//   sequential (bin/match_processor/sequential.rs:205-390) timestamp = wall clock ("%.3f"), matched_text = item.as_value():
//              the canonical text of an address (Ipv4Addr / Ipv6Addr Display), the token itself for everything else
static size_t render_ndjson(mxyr_db* h, const mgpu_match* recs, size_t n, const mgpu_id_pair* ids, const uint8_t* log, uint64_t base,
                            const char* source, const char* timestamp, bool canonical, const char** out) {
  std::string& o = h->text;
  o.clear();
  ValueReader r(h->d + h->L.data_start, h->n - h->L.data_start);
  std::string src = source ? source : "";
  const std::string ts = timestamp ? timestamp : "0.000";
  for (size_t k = 0; k < n; k++) {
    const mgpu_match& m = recs[k];
    const uint8_t* t = log + (m.offset - base);
    std::string text((const char*)t, m.len);
    if (canonical && m.item_type == MGPU_T_IPV6) {
      uint16_t v6[8];
      if (parse_ipv6_text((const char*)t, m.len, v6)) text = ipv6_text(v6);
    }
    std::string line = "{";
    if (m.kind == MGPU_KIND_IP) {
      Value v;
      if (!r.read(m.data_offset, v)) continue;  // decode error: the reference drops the chunk; cannot occur on valid DBs
      line += "\"cidr\":"; json_string(cidr_text((const uint8_t*)text.data(), text.size(), m.prefix_len), line);
      line += ",\"data\":"; render_json(v, line);
      line += ",\"match_type\":\"ip\",\"matched_text\":"; json_string(text, line);
      line += ",\"prefix_len\":" + std::to_string((unsigned)m.prefix_len);
    } else {
      std::string arr; bool any = false, bad = false;
      for (uint32_t j = 0; j < m.n_ids; j++) {
        uint32_t off = ids[m.ids_index + j].data_offset;
        if (off == MGPU_NO_DATA) continue;
        Value v;
        if (!r.read(off, v)) { bad = true; break; }
        if (any) arr += ',';
        render_json(v, arr);
        any = true;
      }
      if (bad) continue;
      if (any) line += "\"data\":[" + arr + "],";
      line += "\"match_type\":\"pattern\",\"matched_text\":"; json_string(text, line);
      line += ",\"pattern_count\":" + std::to_string(m.n_ids);
    }
    line += ",\"source\":"; json_string(src, line);
    line += ",\"timestamp\":\"" + ts + "\"}\n";
    o += line;
  }
  *out = o.c_str();
  return o.size();
}

size_t mxyr_ndjson(mxyr_db* h, const mgpu_match* recs, size_t n, const mgpu_id_pair* ids, const uint8_t* log, uint64_t base,
                   const char* source, const char** out) {
  return render_ndjson(h, recs, n, ids, log, base, source, nullptr, false, out);
}
size_t mxyr_ndjson_sequential(mxyr_db* h, const mgpu_match* recs, size_t n, const mgpu_id_pair* ids, const uint8_t* log, uint64_t base,
                              const char* source, const char* timestamp, const char** out) {
  return render_ndjson(h, recs, n, ids, log, base, source, timestamp, true, out);
}

}  // extern "C"
