// matchy_capi.cpp — the reference's public C ABI (include/matchy/matchy.h) on top of the device engine.
//
// Every entry point follows the behaviour of crates/matchy/src/c_api/matchy.rs (line ranges cited per function) —
// NULL handling, return codes, which result of a multi-pattern match is exposed, who frees what — while the work
// itself runs on the GPU through the mgpu_* ABI (include/matchy_b200.h):
//   matchy_open*                  -> mgpu_create + mgpu_db_upload (sections byte-identical in HBM)
//   matchy_query                  -> Database::lookup semantics (database.rs:725-804) over mgpu_lookup_ip / mgpu_lookup_string
//   matchy_extractor_extract_chunk-> mgpu_extract (tokenizer + token kernels), reference item order (lib.rs:409-488)
//   matchy_builder_*              -> the from-scratch .mxy writer (mxy_builder.cpp)
// Host work here is limited to what the reference also does on the host around a lookup: parsing the query text,
// decoding the ONE matched MMDB value, JSON text.  No CPU matcher exists in this file; when no CUDA device is present
// matchy_open* and matchy_extractor_create return NULL.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/matchy_b200.h"
#include "../../include/matchy/matchy.h"
#include "mxy_builder.h"
#include "mxy_reader.h"
#include "db_prepare.h"

using namespace matchy;

namespace {

int env_device() {
  const char* e = getenv("MATCHY_B200_DEVICE");
  return e && *e ? atoi(e) : 0;
}

// public_suffix_list.dat: $MATCHY_B200_PSL, else <dir of this .so>/../data/public_suffix_list.dat
bool load_psl(std::string& out) {
  std::vector<std::string> tries;
  if (const char* e = getenv("MATCHY_B200_PSL")) tries.push_back(e);
  Dl_info info;
  if (dladdr((const void*)&env_device, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t s = p.rfind('/');
    std::string dir = s == std::string::npos ? "." : p.substr(0, s);
    tries.push_back(dir + "/../data/public_suffix_list.dat");
    tries.push_back(dir + "/public_suffix_list.dat");
  }
  for (auto& t : tries) {
    std::ifstream f(t, std::ios::binary);
    if (!f) continue;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    if (!out.empty()) return true;
  }
  return false;
}

char* dup_cstring(const std::string& s) {  // CString::new(..).into_raw(): NULL when the text holds a NUL byte
  if (memchr(s.data(), 0, s.size())) return nullptr;
  char* p = (char*)malloc(s.size() + 1);
  if (!p) return nullptr;
  memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  return p;
}

bool valid_utf8(const char* s, size_t n) {  // CStr::to_str
  size_t i = 0;
  auto cont = [&](size_t k) { return k < n && ((uint8_t)s[k] & 0xC0) == 0x80; };
  while (i < n) {
    uint8_t c = (uint8_t)s[i];
    if (c < 0x80) { i++; continue; }
    if (c >= 0xC2 && c <= 0xDF) { if (!cont(i + 1)) return false; i += 2; continue; }
    if (c >= 0xE0 && c <= 0xEF) {
      if (!cont(i + 1) || !cont(i + 2)) return false;
      uint8_t d = (uint8_t)s[i + 1];
      if ((c == 0xE0 && d < 0xA0) || (c == 0xED && d > 0x9F)) return false;
      i += 3; continue;
    }
    if (c >= 0xF0 && c <= 0xF4) {
      if (!cont(i + 1) || !cont(i + 2) || !cont(i + 3)) return false;
      uint8_t d = (uint8_t)s[i + 1];
      if ((c == 0xF0 && d < 0x90) || (c == 0xF4 && d > 0x8F)) return false;
      i += 4; continue;
    }
    return false;
  }
  return true;
}

// ---- JSON text -> DataValue with the reference's number typing (matchy-data-format/src/lib.rs:106-204) --------------
struct JsonParser {
  const char* p; const char* e; int depth = 0;
  void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
  bool lit(const char* t) { size_t n = strlen(t); if ((size_t)(e - p) < n || memcmp(p, t, n)) return false; p += n; return true; }
  static void put_utf8(uint32_t cp, std::string& o) {
    if (cp < 0x80) o += (char)cp;
    else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
    else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 63)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
  }
  bool hex4(uint32_t& v) {
    if (e - p < 4) return false;
    v = 0;
    for (int k = 0; k < 4; k++) {
      char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0'; else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10; else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10; else return false;
    }
    return true;
  }
  bool string(std::string& o) {
    if (p >= e || *p != '"') return false;
    p++;
    while (p < e) {
      uint8_t c = (uint8_t)*p++;
      if (c == '"') return true;
      if (c < 0x20) return false;
      if (c != '\\') { o += (char)c; continue; }
      if (p >= e) return false;
      char x = *p++;
      switch (x) {
        case '"': o += '"'; break; case '\\': o += '\\'; break; case '/': o += '/'; break;
        case 'b': o += '\b'; break; case 'f': o += '\f'; break; case 'n': o += '\n'; break; case 'r': o += '\r'; break; case 't': o += '\t'; break;
        case 'u': {
          uint32_t a;
          if (!hex4(a)) return false;
          if (a >= 0xDC00 && a <= 0xDFFF) return false;  // lone trail surrogate
          if (a >= 0xD800 && a <= 0xDBFF) {
            uint32_t b;
            if (e - p < 2 || p[0] != '\\' || p[1] != 'u') return false;
            p += 2;
            if (!hex4(b) || b < 0xDC00 || b > 0xDFFF) return false;
            a = 0x10000 + ((a - 0xD800) << 10) + (b - 0xDC00);
          }
          put_utf8(a, o);
          break;
        }
        default: return false;
      }
    }
    return false;
  }
  bool number(mxy::DataValue& v) {
    const char* s = p;
    bool neg = false, integral = true;
    if (p < e && *p == '-') { neg = true; p++; }
    if (p >= e) return false;
    if (*p == '0') p++;
    else if (*p >= '1' && *p <= '9') { while (p < e && *p >= '0' && *p <= '9') p++; }
    else return false;
    if (p < e && *p == '.') { integral = false; p++; if (p >= e || *p < '0' || *p > '9') return false; while (p < e && *p >= '0' && *p <= '9') p++; }
    if (p < e && (*p == 'e' || *p == 'E')) {
      integral = false; p++;
      if (p < e && (*p == '+' || *p == '-')) p++;
      if (p >= e || *p < '0' || *p > '9') return false;
      while (p < e && *p >= '0' && *p <= '9') p++;
    }
    std::string t(s, p);
    if (integral) {  // serde_json: u64 when it fits (visit_u64), negative i64 (visit_i64), else f64
      const char* d = t.c_str() + (neg ? 1 : 0);
      unsigned __int128 acc = 0; bool big = false;
      for (; *d; d++) { acc = acc * 10 + (unsigned)(*d - '0'); if (acc > ((unsigned __int128)1 << 64)) { big = true; break; } }
      if (!neg && !big && acc <= (unsigned __int128)UINT64_MAX) {
        uint64_t u = (uint64_t)acc;
        v = u <= 0xFFFF ? mxy::DataValue::Uint16((uint16_t)u) : u <= 0xFFFFFFFFull ? mxy::DataValue::Uint32((uint32_t)u) : mxy::DataValue::Uint64(u);
        return true;
      }
      if (neg && !big && acc <= ((unsigned __int128)1 << 63)) {
        if (acc == 0) { v = mxy::DataValue::Double(-0.0); return true; }  // serde_json parses "-0" as the float -0.0
        int64_t i = (int64_t)(0 - (uint64_t)acc);
        v = i >= INT32_MIN ? mxy::DataValue::Int32((int32_t)i) : mxy::DataValue::Double((double)i);
        return true;
      }
    }
    double dv = strtod(t.c_str(), nullptr);
    if (!std::isfinite(dv)) return false;  // serde_json: "number out of range"
    v = mxy::DataValue::Double(dv);
    return true;
  }
  bool value(mxy::DataValue& v) {
    if (++depth > 128) return false;  // serde_json's recursion limit
    ws();
    if (p >= e) return false;
    bool ok = false;
    if (*p == '{') {
      p++; v = mxy::DataValue::Map(); ws();
      if (p < e && *p == '}') { p++; ok = true; }
      else for (;;) {
        ws();
        std::string k;
        if (!string(k)) break;
        ws();
        if (p >= e || *p != ':') break;
        p++;
        mxy::DataValue x;
        if (!value(x)) break;
        v.map[k] = x;  // HashMap::insert: a later duplicate key replaces the earlier one
        ws();
        if (p < e && *p == ',') { p++; continue; }
        if (p < e && *p == '}') { p++; ok = true; }
        break;
      }
    } else if (*p == '[') {
      p++; v = mxy::DataValue::Array(); ws();
      if (p < e && *p == ']') { p++; ok = true; }
      else for (;;) {
        mxy::DataValue x;
        if (!value(x)) break;
        v.arr.push_back(x);
        ws();
        if (p < e && *p == ',') { p++; continue; }
        if (p < e && *p == ']') { p++; ok = true; }
        break;
      }
    } else if (*p == '"') { std::string s; ok = string(s); if (ok) v = mxy::DataValue::String(s); }
    else if (*p == 't') { ok = lit("true"); v = mxy::DataValue::Bool(true); }
    else if (*p == 'f') { ok = lit("false"); v = mxy::DataValue::Bool(false); }
    else if (*p == 'n') ok = false;  // null: the reference's visitor has no visit_unit -> error
    else ok = number(v);
    depth--;
    return ok;
  }
  bool parse(mxy::DataValue& v) { if (!value(v)) return false; ws(); return p == e; }
};

// ---- handles -----------------------------------------------------------------------------------------------------
struct Builder {
  mxy::DatabaseBuilder* b;
  Builder() : b(new mxy::DatabaseBuilder(mxy::MatchMode::CaseSensitive)) {}
  ~Builder() { delete b; }
  void reset() { delete b; b = new mxy::DatabaseBuilder(mxy::MatchMode::CaseSensitive); }  // mem::replace with a fresh builder (matchy.rs:524-528)
};

struct Cached { int kind; uint32_t data_offset; uint8_t prefix_len; };  // kind: 0 NotFound, 1 Ip, 2 Pattern (first data), 3 Pattern without data
struct Db {
  std::vector<uint8_t> bytes;
  mgpu_ctx* ctx = nullptr;
  mxyr_db* reader = nullptr;
  mxy::Layout L;
  mgpu_db_info info{};
  std::mutex mu;  // one host thread per device context (include/matchy_b200.h)
  matchy_stats_t st{};
  // the reference's query cache (database.rs:725-760): result-neutral, but visible through matchy_get_stats
  size_t cache_cap = 10000;
  std::list<std::pair<std::string, Cached>> lru;
  std::unordered_map<std::string, std::list<std::pair<std::string, Cached>>::iterator> idx;
  ~Db() { if (reader) mxyr_close(reader); if (ctx) mgpu_destroy(ctx); }
};

struct Extractor { mgpu_ctx* ctx = nullptr; uint32_t flags = 0; std::mutex mu; ~Extractor() { if (ctx) mgpu_destroy(ctx); } };
struct MatchesInternal { std::vector<matchy_match_t> items; std::vector<std::string> strings; };

Db* open_bytes(std::vector<uint8_t>&& bytes, size_t cache_cap) {
  if (bytes.empty()) return nullptr;
  Db* d = new Db();
  d->bytes = std::move(bytes);
  d->cache_cap = cache_cap;
  std::string err;
  if (!mxy::locate_sections(d->bytes.data(), d->bytes.size(), d->L, err)) { delete d; return nullptr; }
  d->ctx = mgpu_create(env_device(), (size_t)16 << 20);
  if (!d->ctx) { delete d; return nullptr; }
  if (mgpu_db_upload(d->ctx, d->bytes.data(), d->bytes.size()) != MGPU_OK || mgpu_db_info_get(d->ctx, &d->info) != MGPU_OK) { delete d; return nullptr; }
  d->reader = mxyr_open(d->bytes.data(), d->bytes.size());
  if (!d->reader) { delete d; return nullptr; }
  return d;
}

Db* open_path(const char* path, size_t cache_cap) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return nullptr;
  std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  return open_bytes(std::move(bytes), cache_cap);
}

bool decode_value(Db* d, uint32_t off, mxy::Value& out) {
  mxy::ValueReader r(d->bytes.data() + d->L.data_start, d->bytes.size() - d->L.data_start);
  return r.read(off, out);
}

// Rust `query.parse::<IpAddr>()`: v4 first, then v6
bool parse_ip(const char* q, size_t n, uint8_t ip16[16], bool& v6) {
  uint32_t a4; uint16_t seg[8];
  memset(ip16, 0, 16);
  if (mxy::parse_ipv4_text(q, n, a4)) { ip16[0] = a4 >> 24; ip16[1] = a4 >> 16; ip16[2] = a4 >> 8; ip16[3] = a4; v6 = false; return true; }
  if (mxy::parse_ipv6_text(q, n, seg)) { for (int k = 0; k < 8; k++) { ip16[2 * k] = seg[k] >> 8; ip16[2 * k + 1] = seg[k] & 0xFF; } v6 = true; return true; }
  return false;
}

matchy_result_t not_found() { return matchy_result_t{false, 0, nullptr, nullptr}; }

const mxy::Value* value_of(const matchy_entry_s* entry) {  // entry -> result -> cached value (matchy.rs:1806-1819)
  const matchy_result_t* r = (const matchy_result_t*)entry->data_ptr;
  if (!r || !r->_data_cache) return nullptr;
  return (const mxy::Value*)r->_data_cache;
}

const mxy::Value* map_get(const mxy::Value& m, const char* key) {
  const mxy::Value* r = nullptr;
  for (auto& f : m.fields) if (f.first == key) r = &f.second;  // later duplicate wins, like HashMap::insert
  return r;
}

bool to_entry_data(const mxy::Value& v, matchy_entry_data_t& o) {  // matchy_entry_data_t::from_data_value (matchy.rs:1586-1680)
  memset(&o, 0, sizeof o);
  o.has_data = true;
  switch (v.kind) {
    case mxy::Value::PTR: o.type_ = MATCHY_DATA_TYPE_POINTER; o.value.pointer = (uint32_t)v.u; break;
    case mxy::Value::STR:
      if (memchr(v.s.data(), 0, v.s.size())) return false;  // CString::new fails
      o.type_ = MATCHY_DATA_TYPE_UTF8_STRING; o.value.utf8_string = v.s.c_str(); o.data_size = (uint32_t)v.s.size(); break;
    case mxy::Value::F64: o.type_ = MATCHY_DATA_TYPE_DOUBLE; o.value.double_value = v.f; o.data_size = 8; break;
    case mxy::Value::F32: o.type_ = MATCHY_DATA_TYPE_FLOAT; o.value.float_value = (float)v.f; o.data_size = 4; break;
    case mxy::Value::BYTES: o.type_ = MATCHY_DATA_TYPE_BYTES; o.value.bytes = (const uint8_t*)v.s.data(); o.data_size = (uint32_t)v.s.size(); break;
    case mxy::Value::UINT:
      if (v.mmdb_type == 5) { o.type_ = MATCHY_DATA_TYPE_UINT16; o.value.uint16 = (uint16_t)v.u; o.data_size = 2; }
      else if (v.mmdb_type == 6) { o.type_ = MATCHY_DATA_TYPE_UINT32; o.value.uint32 = (uint32_t)v.u; o.data_size = 4; }
      else { o.type_ = MATCHY_DATA_TYPE_UINT64; o.value.uint64 = v.u; o.data_size = 8; }
      break;
    case mxy::Value::INT: o.type_ = MATCHY_DATA_TYPE_INT32; o.value.int32 = (int32_t)v.i; o.data_size = 4; break;
    case mxy::Value::U128:
      o.type_ = MATCHY_DATA_TYPE_UINT128; o.data_size = 16;
      for (int k = 0; k < 16; k++) o.value.uint128[k] = (uint8_t)(v.big >> (8 * (15 - k)));  // to_be_bytes
      break;
    case mxy::Value::MAP: {
      o.type_ = MATCHY_DATA_TYPE_MAP;
      std::vector<const std::string*> keys;
      for (auto& f : v.fields) keys.push_back(&f.first);
      std::sort(keys.begin(), keys.end(), [](auto* a, auto* b) { return *a < *b; });
      o.data_size = (uint32_t)(std::unique(keys.begin(), keys.end(), [](auto* a, auto* b) { return *a == *b; }) - keys.begin());
      break;
    }
    case mxy::Value::ARR: o.type_ = MATCHY_DATA_TYPE_ARRAY; o.data_size = (uint32_t)v.items.size(); break;
    case mxy::Value::BOOL: o.type_ = MATCHY_DATA_TYPE_BOOLEAN; o.value.boolean = v.b; o.data_size = 1; break;
    default: return false;
  }
  return true;
}

void flatten(const mxy::Value& v, matchy_entry_data_list_t*& head, matchy_entry_data_list_t*& tail) {  // matchy.rs:1920-1950
  matchy_entry_data_t ed;
  if (to_entry_data(v, ed)) {
    auto* node = new matchy_entry_data_list_t{ed, nullptr};
    if (!head) head = tail = node; else { tail->next = node; tail = node; }
  }
  if (v.kind == mxy::Value::MAP) {
    // the reference walks a HashMap (unspecified order); here: sorted keys, duplicates resolved like insert()
    std::vector<const std::pair<std::string, mxy::Value>*> order;
    for (auto& f : v.fields) order.push_back(&f);
    std::stable_sort(order.begin(), order.end(), [](auto* a, auto* b) { return a->first < b->first; });
    for (size_t k = 0; k < order.size(); k++) {
      if (k + 1 < order.size() && order[k + 1]->first == order[k]->first) continue;
      flatten(order[k]->second, head, tail);
    }
  } else if (v.kind == mxy::Value::ARR) {
    for (auto& x : v.items) flatten(x, head, tail);
  }
}

}  // namespace

namespace matchy {
extern "C" {

// ---- builder (matchy.rs:252-620) -----------------------------------------------------------------------------------
matchy_builder_t* matchy_builder_new(void) { return (matchy_builder_t*)new Builder(); }

int32_t matchy_builder_set_case_insensitive(matchy_builder_t* builder, bool ci) {
  if (!builder) return MATCHY_ERROR_INVALID_PARAM;
  ((Builder*)builder)->b->set_mode(ci ? mxy::MatchMode::CaseInsensitive : mxy::MatchMode::CaseSensitive);
  return MATCHY_SUCCESS;
}

int32_t matchy_builder_set_schema(matchy_builder_t* builder, const char* schema_name) {
  if (!builder || !schema_name) return MATCHY_ERROR_INVALID_PARAM;
  if (!valid_utf8(schema_name, strlen(schema_name))) return MATCHY_ERROR_INVALID_PARAM;
  return MATCHY_ERROR_UNKNOWN_SCHEMA;  // schema validation (schemas/*.json) is outside the scan path; no schema is "known" here
}

int32_t matchy_builder_add(matchy_builder_t* builder, const char* key, const char* json_data) {
  if (!builder || !key || !json_data) return MATCHY_ERROR_INVALID_PARAM;
  const size_t kn = strlen(key), jn = strlen(json_data);
  if (!valid_utf8(key, kn) || !valid_utf8(json_data, jn)) return MATCHY_ERROR_INVALID_PARAM;
  mxy::DataValue v;
  JsonParser jp{json_data, json_data + jn};
  if (!jp.parse(v)) return MATCHY_ERROR_INVALID_FORMAT;
  if (v.type != mxy::DataValue::MAP) {  // a bare value is wrapped as {"value": v} (matchy.rs:428-437)
    mxy::DataValue m = mxy::DataValue::Map();
    m.map["value"] = v;
    v = m;
  }
  return ((Builder*)builder)->b->add_entry(std::string(key, kn), v) ? MATCHY_SUCCESS : MATCHY_ERROR_INVALID_FORMAT;
}

int32_t matchy_builder_set_description(matchy_builder_t* builder, const char* description) {
  if (!builder || !description) return MATCHY_ERROR_INVALID_PARAM;
  if (!valid_utf8(description, strlen(description))) return MATCHY_ERROR_INVALID_PARAM;
  ((Builder*)builder)->b->set_description("en", description);  // with_description("en", ..) (matchy.rs:480-490)
  return MATCHY_SUCCESS;
}

static int32_t build_bytes(Builder* B, std::vector<uint8_t>& out) {
  bool ok = B->b->build(out);
  B->reset();  // the reference consumes the builder: afterwards the handle holds an empty one
  return ok ? MATCHY_SUCCESS : MATCHY_ERROR_INVALID_FORMAT;
}

int32_t matchy_builder_save(matchy_builder_t* builder, const char* filename) {
  if (!builder || !filename) return MATCHY_ERROR_INVALID_PARAM;
  if (!valid_utf8(filename, strlen(filename))) return MATCHY_ERROR_INVALID_PARAM;
  std::vector<uint8_t> out;
  int32_t rc = build_bytes((Builder*)builder, out);
  if (rc != MATCHY_SUCCESS) return rc;
  FILE* f = fopen(filename, "wb");
  if (!f) return MATCHY_ERROR_IO;
  bool ok = out.empty() || fwrite(out.data(), 1, out.size(), f) == out.size();
  ok = (fclose(f) == 0) && ok;
  return ok ? MATCHY_SUCCESS : MATCHY_ERROR_IO;
}

int32_t matchy_builder_build(matchy_builder_t* builder, uint8_t** buffer, uintptr_t* size) {
  if (!builder || !buffer || !size) return MATCHY_ERROR_INVALID_PARAM;
  std::vector<uint8_t> out;
  int32_t rc = build_bytes((Builder*)builder, out);
  if (rc != MATCHY_SUCCESS) return rc;
  uint8_t* p = (uint8_t*)malloc(out.size() ? out.size() : 1);
  if (!p) return MATCHY_ERROR_OUT_OF_MEMORY;
  memcpy(p, out.data(), out.size());
  *buffer = p; *size = out.size();
  return MATCHY_SUCCESS;
}

void matchy_builder_free(matchy_builder_t* builder) { delete (Builder*)builder; }

// ---- open / close (matchy.rs:739-1075) -----------------------------------------------------------------------------
void matchy_init_open_options(matchy_open_options_t* o) {
  if (!o) return;
  o->cache_capacity = 10000; o->auto_reload = false; o->reload_callback = nullptr; o->reload_callback_user_data = nullptr;
}

matchy_t* matchy_open_with_options(const char* filename, const matchy_open_options_t* options) {
  if (!filename || !options) return nullptr;
  if (!valid_utf8(filename, strlen(filename))) return nullptr;
  // auto_reload (file watching + re-upload) is not on the scan path: the database opens statically
  return (matchy_t*)open_path(filename, options->cache_capacity);
}

matchy_t* matchy_open(const char* filename) {
  if (!filename) return nullptr;
  if (!valid_utf8(filename, strlen(filename))) return nullptr;
  return (matchy_t*)open_path(filename, 10000);
}

matchy_t* matchy_open_buffer(const uint8_t* buffer, uintptr_t size) {
  if (!buffer || size == 0) return nullptr;
  return (matchy_t*)open_bytes(std::vector<uint8_t>(buffer, buffer + size), 10000);  // copies, like slice.to_vec()
}

void matchy_get_stats(const matchy_t* db, matchy_stats_t* stats) {
  if (!db || !stats) return;
  Db* d = (Db*)db;
  std::lock_guard<std::mutex> g(d->mu);
  *stats = d->st;
}

void matchy_clear_cache(const matchy_t* db) {
  if (!db) return;
  Db* d = (Db*)db;
  std::lock_guard<std::mutex> g(d->mu);
  d->lru.clear(); d->idx.clear();
}

void matchy_close(matchy_t* db) { delete (Db*)db; }

// ---- query (matchy.rs:1099-1245; Database::lookup database.rs:725-981) ---------------------------------------------
matchy_result_t matchy_query(const matchy_t* db, const char* query) {
  if (!db || !query) return not_found();
  const size_t n = strlen(query);
  if (!valid_utf8(query, n)) return not_found();
  Db* d = (Db*)db;
  std::lock_guard<std::mutex> g(d->mu);
  uint8_t ip16[16]; bool v6 = false;
  const bool is_ip = parse_ip(query, n, ip16, v6);
  Cached c{0, 0, 0};
  mxy::Value* missed = nullptr;  // cache miss with data: the value, decoded before anything is counted
  bool none = false;  // Ok(None): the database has no data of the queried class
  const std::string key(query, n);
  auto hit = d->cache_cap ? d->idx.find(key) : d->idx.end();
  if (hit != d->idx.end()) {
    c = hit->second->second;
    d->lru.splice(d->lru.begin(), d->lru, hit->second);
    d->st.total_queries++; d->st.cache_hits++;
    if (c.kind == 1) d->st.ip_queries++;
    else if (c.kind >= 2) d->st.string_queries++;
    else if (is_ip) d->st.ip_queries++; else d->st.string_queries++;
    if (c.kind) d->st.queries_with_match++; else d->st.queries_without_match++;
  } else {
    if (is_ip) {
      if (!d->info.has_ip) none = true;
      else {
        uint32_t off = 0; uint8_t pl = 0;
        int rc = mgpu_lookup_ip(d->ctx, ip16, v6 ? 1 : 0, &off, &pl);
        if (rc < 0) return not_found();  // Err(_) -> not found, stats untouched (the `?` leaves lookup early)
        if (rc > 0) c = Cached{1, off, pl};
      }
    } else {
      if (!d->info.has_literal && !d->info.has_glob) none = true;
      else {
        mgpu_id_pair first;
        int rc = mgpu_lookup_string(d->ctx, (const uint8_t*)query, n, &first, 1);
        if (rc < 0) return not_found();
        if (rc > 0) c = first.data_offset == MGPU_NO_DATA ? Cached{3, 0, 0} : Cached{2, first.data_offset, 0};
      }
    }
    // the reference decodes inside lookup_ip / lookup_string_uncached: a decode error leaves lookup() through `?` before the
    // statistics and the cache are touched (database.rs:725-804)
    if (c.kind == 1 || c.kind == 2) {
      missed = new mxy::Value();
      if (!decode_value(d, c.data_offset, *missed)) { delete missed; return not_found(); }
    }
    d->st.total_queries++;
    if (d->cache_cap) d->st.cache_misses++;
    if (c.kind == 1) { d->st.ip_queries++; d->st.queries_with_match++; }
    else if (c.kind >= 2) { d->st.string_queries++; d->st.queries_with_match++; }
    else if (!none) { d->st.string_queries++; d->st.queries_without_match++; }  // (sic: an IP miss counts as a string query, database.rs:790-795)
    else d->st.queries_without_match++;
    if (d->cache_cap && !none) {
      d->lru.emplace_front(key, c);
      d->idx[key] = d->lru.begin();
      if (d->lru.size() > d->cache_cap) { d->idx.erase(d->lru.back().first); d->lru.pop_back(); }
    }
  }
  if (c.kind != 1 && c.kind != 2) return not_found();  // Pattern whose first entry has no data -> found = false (matchy.rs:1143-1163)
  if (missed) return matchy_result_t{true, c.prefix_len, missed, db};
  auto* v = new mxy::Value();
  if (!decode_value(d, c.data_offset, *v)) { delete v; return not_found(); }
  return matchy_result_t{true, c.prefix_len, v, db};
}

void matchy_query_into(const matchy_t* db, const char* query, matchy_result_t* result) {
  if (!result) return;
  *result = matchy_query(db, query);
}

void matchy_free_result(matchy_result_t* result) {
  if (result && result->_data_cache) { delete (mxy::Value*)result->_data_cache; result->_data_cache = nullptr; }
}

void matchy_free_string(char* s) { free(s); }

const char* matchy_version(void) { return "1.2.2"; }  // CARGO_PKG_VERSION of the reference this ABI mirrors

const char* matchy_format(const matchy_t* db) {  // Database::format (database.rs:1072-1078, detect_format :1022-1068)
  if (!db) return nullptr;
  Db* d = (Db*)db;
  return (d->L.has_glob || d->L.has_literal) ? "Combined IP+Pattern database" : "IP database";
}

bool matchy_has_ip_data(const matchy_t* db) { return db && ((Db*)db)->info.has_ip; }
bool matchy_has_literal_data(const matchy_t* db) { return db && ((Db*)db)->info.has_literal; }
bool matchy_has_glob_data(const matchy_t* db) { return db && ((Db*)db)->info.has_glob; }
bool matchy_has_string_data(const matchy_t* db) { return matchy_has_literal_data(db) || matchy_has_glob_data(db); }
bool matchy_has_pattern_data(const matchy_t* db) { return matchy_has_string_data(db); }

char* matchy_metadata(const matchy_t* db) {  // database.rs:1113-1121
  if (!db) return nullptr;
  Db* d = (Db*)db;
  if (!d->info.has_ip) return nullptr;
  mxy::Value meta;
  if (!mxy::read_metadata(d->bytes.data(), d->bytes.size(), meta)) return nullptr;
  std::string s;
  mxy::render_json(meta, s);
  return dup_cstring(s);
}

static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

uintptr_t matchy_pattern_count(const matchy_t* db) {  // ParaglobHeader.pattern_count (paraglob_offset.rs:1763-1775)
  if (!db) return 0;
  Db* d = (Db*)db;
  if (!d->L.has_glob || d->L.pg_len < 112) return 0;
  return rd32(d->bytes.data() + d->L.pg_off + 32);
}

char* matchy_get_pattern_string(const matchy_t* db, uint32_t pattern_id) {  // Paraglob::get_pattern: PatternEntry[id] -> NUL-terminated text
  if (!db) return nullptr;
  Db* d = (Db*)db;
  if (!d->L.has_glob || d->L.pg_len < 112) return nullptr;
  const uint8_t* pg = d->bytes.data() + d->L.pg_off;
  const uint64_t count = rd32(pg + 32), entries = rd32(pg + 36);
  if (pattern_id >= count || entries + ((uint64_t)pattern_id + 1) * 16 > d->L.pg_len) return nullptr;
  const uint8_t* e = pg + entries + (uint64_t)pattern_id * 16;
  const uint64_t so = rd32(e + 8), sl = rd32(e + 12);
  if (so + sl > d->L.pg_len) return nullptr;
  return dup_cstring(std::string((const char*)pg + so, sl));
}

// ---- structured data access (matchy.rs:1734-2060) ------------------------------------------------------------------
int32_t matchy_result_get_entry(const matchy_result_t* result, matchy_entry_s* entry) {
  if (!result || !entry) return MATCHY_ERROR_INVALID_PARAM;
  if (!result->found) return MATCHY_ERROR_NO_DATA;
  entry->db = result->_db_ref;
  entry->data_ptr = result;
  return MATCHY_SUCCESS;
}

int32_t matchy_aget_value(const matchy_entry_s* entry, matchy_entry_data_t* entry_data, const char* const* path) {
  if (!entry || !entry_data || !path) return MATCHY_ERROR_INVALID_PARAM;
  for (size_t i = 0; path[i]; i++) if (!valid_utf8(path[i], strlen(path[i]))) return MATCHY_ERROR_INVALID_PARAM;
  const mxy::Value* v = value_of(entry);
  memset(entry_data, 0, sizeof *entry_data);
  if (!v) return MATCHY_ERROR_NO_DATA;
  for (size_t i = 0; path[i]; i++) {  // navigate_path (matchy.rs:1715-1732)
    if (v->kind == mxy::Value::MAP) { v = map_get(*v, path[i]); if (!v) return MATCHY_ERROR_LOOKUP_PATH_INVALID; }
    else if (v->kind == mxy::Value::ARR) {
      const char* s = path[i];  // usize::from_str: optional '+', then digits only
      if (*s == '+') s++;
      if (!*s) return MATCHY_ERROR_LOOKUP_PATH_INVALID;
      uint64_t idx = 0;
      for (; *s; s++) { if (*s < '0' || *s > '9' || idx > (UINT64_MAX - 9) / 10) return MATCHY_ERROR_LOOKUP_PATH_INVALID; idx = idx * 10 + (uint64_t)(*s - '0'); }
      if (idx >= v->items.size()) return MATCHY_ERROR_LOOKUP_PATH_INVALID;
      v = &v->items[idx];
    } else return MATCHY_ERROR_LOOKUP_PATH_INVALID;
  }
  if (!to_entry_data(*v, *entry_data)) { memset(entry_data, 0, sizeof *entry_data); return MATCHY_ERROR_DATA_PARSE; }
  return MATCHY_SUCCESS;  // string / byte pointers point into the result's cached value: valid until matchy_free_result
}

int32_t matchy_get_entry_data_list(const matchy_entry_s* entry, matchy_entry_data_list_t** list) {
  if (!entry || !list) return MATCHY_ERROR_INVALID_PARAM;
  const mxy::Value* v = value_of(entry);
  if (!v) return MATCHY_ERROR_NO_DATA;
  matchy_entry_data_list_t *head = nullptr, *tail = nullptr;
  flatten(*v, head, tail);
  *list = head;
  return MATCHY_SUCCESS;
}

void matchy_free_entry_data_list(matchy_entry_data_list_t* list) {
  while (list) { matchy_entry_data_list_t* next = list->next; delete list; list = next; }
}

char* matchy_result_to_json(const matchy_result_t* result) {
  if (!result || !result->found || !result->_data_cache) return nullptr;
  // serde_json::to_string(&DataValue): compact; object key order is the HashMap's (unspecified) in the reference, sorted here
  std::string s;
  mxy::render_json(*(const mxy::Value*)result->_data_cache, s, /*f32_shortest=*/true);
  return dup_cstring(s);
}

// ---- validation (matchy.rs:2072-2130): structural checks of the loader only ------------------------------------------
int32_t matchy_validate(const char* filename, int32_t level, char** error_message) {
  if (error_message) *error_message = nullptr;
  if (!filename) return MATCHY_ERROR_INVALID_PARAM;
  if (level != MATCHY_VALIDATION_STANDARD && level != MATCHY_VALIDATION_STRICT) return MATCHY_ERROR_INVALID_PARAM;
  std::ifstream f(filename, std::ios::binary);
  if (!f) { if (error_message) *error_message = dup_cstring(std::string("cannot open ") + filename); return MATCHY_ERROR_FILE_NOT_FOUND; }
  std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  // validate_database (crates/matchy/src/validation.rs:256-): metadata marker and required fields, section bounds, then the
  // structural pass the device upload runs anyway (db_prepare.h::prepare_db): every tree record is a node, the empty marker or
  // an in-range data pointer; LHSH / PARAGLOB / ACLH headers, every automaton offset in range and aligned, glob segments sane.
  mxy::Layout L; std::string err;
  if (!mxy::locate_sections(bytes.data(), bytes.size(), L, err)) { if (error_message) *error_message = dup_cstring(err); return MATCHY_ERROR_CORRUPT_DATA; }
  bytes.resize(bytes.size() + 64, 0);  // (the preparer reads whole words at the very end of sections)
  mgpu::PreparedDb P;
  if (!mgpu::prepare_db(bytes.data(), bytes.size() - 64, P, err)) { if (error_message) *error_message = dup_cstring(err); return MATCHY_ERROR_CORRUPT_DATA; }
  return MATCHY_SUCCESS;
}

// ---- extractor (matchy.rs:2270-2450) -------------------------------------------------------------------------------
matchy_extractor_t* matchy_extractor_create(uint32_t flags) {
  std::string psl;
  if (!load_psl(psl)) return nullptr;
  auto* x = new Extractor();
  x->ctx = mgpu_create(env_device(), (size_t)64 << 20);
  if (!x->ctx || mgpu_set_psl(x->ctx, (const uint8_t*)psl.data(), psl.size()) != MGPU_OK) { delete x; return nullptr; }
  // the public bit layout (matchy.h: domains 1, emails 2, ipv4 4, ipv6 8, hashes 16, btc 32, eth 64, xmr 128) -> engine bits
  uint32_t f = 0;
  if (flags & MATCHY_EXTRACT_DOMAINS) f |= MGPU_X_DOMAINS;
  if (flags & MATCHY_EXTRACT_EMAILS) f |= MGPU_X_EMAILS;
  if (flags & MATCHY_EXTRACT_IPV4) f |= MGPU_X_IPV4;
  if (flags & MATCHY_EXTRACT_IPV6) f |= MGPU_X_IPV6;
  if (flags & MATCHY_EXTRACT_HASHES) f |= MGPU_X_HASHES;
  if (flags & MATCHY_EXTRACT_BITCOIN) f |= MGPU_X_BITCOIN;
  if (flags & MATCHY_EXTRACT_ETHEREUM) f |= MGPU_X_ETHEREUM;
  if (flags & MATCHY_EXTRACT_MONERO) f |= MGPU_X_MONERO;
  x->flags = f;
  return (matchy_extractor_t*)x;
}

int32_t matchy_extractor_extract_chunk(const matchy_extractor_t* extractor, const uint8_t* data, uintptr_t len, matchy_matches_t* matches) {
  if (!extractor || !data || !matches) return MATCHY_ERROR_INVALID_PARAM;
  Extractor* x = (Extractor*)extractor;
  std::lock_guard<std::mutex> g(x->mu);
  std::vector<uint64_t> trip(3 * 1024);
  int64_t n = mgpu_extract(x->ctx, data, len, x->flags, trip.data(), trip.size() / 3);
  if (n < 0) return MATCHY_ERROR_INVALID_PARAM;
  if ((size_t)n > trip.size() / 3) {
    trip.resize((size_t)n * 3);
    n = mgpu_extract(x->ctx, data, len, x->flags, trip.data(), (size_t)n);
    if (n < 0) return MATCHY_ERROR_INVALID_PARAM;
  }
  // extract_from_chunk order: IPv6, IPv4, e-mail, domain, hashes, bitcoin, ethereum, monero; ascending offset inside a group
  static const int ORDER[12] = {3, 2, 1, 0, 4, 4, 4, 4, 4, 5, 6, 7};
  std::vector<size_t> ord((size_t)n);
  for (size_t k = 0; k < ord.size(); k++) ord[k] = k;
  std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
    const int oa = ORDER[trip[3 * a] % 12], ob = ORDER[trip[3 * b] % 12];
    return oa != ob ? oa < ob : trip[3 * a + 1] < trip[3 * b + 1];
  });
  auto* mi = new MatchesInternal();
  mi->strings.reserve(ord.size());
  for (size_t k : ord) {
    const uint8_t type = (uint8_t)trip[3 * k];
    const uint64_t s = trip[3 * k + 1], e = trip[3 * k + 2];
    std::string text((const char*)data + s, e - s);
    if (type == MATCHY_ITEM_TYPE_IPV4) {  // ExtractedItem::as_value: Display of the parsed address (lib.rs:300-311)
      uint32_t a;
      if (mxy::parse_ipv4_text(text.data(), text.size(), a)) text = std::to_string(a >> 24) + "." + std::to_string((a >> 16) & 255) + "." + std::to_string((a >> 8) & 255) + "." + std::to_string(a & 255);
    } else if (type == MATCHY_ITEM_TYPE_IPV6) {
      uint16_t seg[8];
      if (mxy::parse_ipv6_text(text.data(), text.size(), seg)) text = mxy::ipv6_text(seg);
    }
    if (memchr(text.data(), 0, text.size())) continue;  // CString::new fails -> the item is skipped
    mi->strings.push_back(std::move(text));
    mi->items.push_back(matchy_match_t{type, nullptr, (uintptr_t)s, (uintptr_t)e});
  }
  for (size_t k = 0; k < mi->items.size(); k++) mi->items[k].value = mi->strings[k].c_str();
  matches->items = mi->items.data();
  matches->count = mi->items.size();
  matches->_internal = mi;
  return MATCHY_SUCCESS;
}

void matchy_matches_free(matchy_matches_t* matches) {
  if (!matches || !matches->_internal) return;
  delete (MatchesInternal*)matches->_internal;
  matches->_internal = nullptr; matches->items = nullptr; matches->count = 0;
}

void matchy_extractor_free(matchy_extractor_t* extractor) { delete (Extractor*)extractor; }

const char* matchy_item_type_name(uint8_t t) {
  static const char* const N[12] = {"Domain", "Email", "IPv4", "IPv6", "MD5", "SHA1", "SHA256", "SHA384", "SHA512", "Bitcoin", "Ethereum", "Monero"};
  return t < 12 ? N[t] : "Unknown";
}

}  // extern "C"
}  // namespace matchy
