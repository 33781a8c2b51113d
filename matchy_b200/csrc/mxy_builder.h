// mxy_builder.h — host-side `.mxy` database writer (the `matchy build` / DatabaseBuilder side).
//
// The scan engine uploads an UNCHANGED .mxy file into HBM, so something has to produce
// reference-format files in an environment without the Rust toolchain.  This is a from-scratch
// C++ implementation of the reference's on-disk format as assembled by
//   crates/matchy-format/src/mmdb_builder.rs:392-760   (entry typing, section assembly, metadata)
//   crates/matchy-ip-trie/src/lib.rs:142-546            (IP search tree, 24/28/32-bit records)
//   crates/matchy-data-format/src/lib.rs:257-623        (MMDB data section: dedup + string interning)
//   crates/matchy-literal-hash/src/lib.rs:173-354,589-663 (sharded XXH64 open-addressing table)
//   crates/matchy-ac/src/lib.rs:201-516                 (Aho-Corasick node/edge/dense layout)
//   crates/matchy-paraglob/src/paraglob_offset.rs:524-888, glob.rs:307-451, literal_hash.rs:121-200
// It backs the mxyb_* C ABI (include/matchy_b200.h) and the synthetic DB generators.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace mxy {

typedef unsigned __int128 u128;

// MMDB data value (matchy-data-format/src/lib.rs:38-77)
struct DataValue {
  enum Type { STRING, DOUBLE, BYTES, UINT16, UINT32, MAP, INT32, UINT64, UINT128, ARRAY, BOOL, FLOAT } type = UINT16;
  std::string str;  // STRING / BYTES
  double dbl = 0;
  float flt = 0;
  uint64_t u = 0;  // UINT16/32/64, BOOL
  u128 big = 0;
  int32_t i32 = 0;
  std::map<std::string, DataValue> map;  // the reference sorts keys before encoding, so std::map is exact
  std::vector<DataValue> arr;

  static DataValue String(const std::string& s) { DataValue v; v.type = STRING; v.str = s; return v; }
  static DataValue Int32(int32_t x) { DataValue v; v.type = INT32; v.i32 = x; return v; }
  static DataValue Uint16(uint16_t x) { DataValue v; v.type = UINT16; v.u = x; return v; }
  static DataValue Uint32(uint32_t x) { DataValue v; v.type = UINT32; v.u = x; return v; }
  static DataValue Uint64(uint64_t x) { DataValue v; v.type = UINT64; v.u = x; return v; }
  static DataValue Double(double x) { DataValue v; v.type = DOUBLE; v.dbl = x; return v; }
  static DataValue Bool(bool x) { DataValue v; v.type = BOOL; v.u = x; return v; }
  static DataValue Map() { DataValue v; v.type = MAP; return v; }
  static DataValue Array() { DataValue v; v.type = ARRAY; return v; }
};

// DataEncoder (matchy-data-format/src/lib.rs:257-623)
class DataEncoder {
 public:
  uint32_t encode(const DataValue& v);  // whole-value dedup, then interned encoding
  std::vector<uint8_t>& bytes() { return buf_; }
  static void encode_plain(const DataValue& v, std::vector<uint8_t>& out);

 private:
  void encode_interned(const DataValue& v);
  void intern_string(const std::string& s);
  std::vector<uint8_t> buf_;
  std::unordered_map<std::string, uint32_t> dedup_;    // plain serialisation → offset
  std::unordered_map<std::string, uint32_t> strings_;  // interned strings → offset
};

enum class MatchMode { CaseSensitive = 0, CaseInsensitive = 1 };

struct IpKey { u128 bits; bool v6; uint8_t prefix; };  // v4: address in the low 32 bits

class DatabaseBuilder {
 public:
  explicit DatabaseBuilder(MatchMode mode = MatchMode::CaseSensitive) : mode_(mode) {}

  // mmdb_builder.rs:186-201 (auto-detect incl. literal:/glob:/ip: prefixes) — false + error() on invalid entries
  bool add_entry(const std::string& key, const DataValue& data_map);
  bool add_entry_at(const std::string& key, uint32_t data_offset);
  bool add_ip(const std::string& ip_or_cidr, uint32_t data_offset);
  void add_literal(const std::string& s, uint32_t data_offset) { literals_.push_back({s, data_offset}); }
  void add_glob(const std::string& s, uint32_t data_offset) { globs_.push_back({s, data_offset}); }
  void add_ip_raw(const IpKey& k, uint32_t data_offset) { ips_.push_back({k, data_offset}); }
  uint32_t encode_data(const DataValue& data_map) { return data_.encode(data_map); }

  void set_database_type(const std::string& t) { database_type_ = t; have_type_ = true; }
  void set_description(const std::string& lang, const std::string& text) { description_[lang] = text; }
  void set_build_epoch(uint64_t e) { build_epoch_ = e; have_epoch_ = true; }
  void set_mode(MatchMode m) { mode_ = m; }

  bool build(std::vector<uint8_t>& out);  // mmdb_builder.rs:432-760
  const std::string& error() const { return error_; }
  size_t ip_count() const { return ips_.size(); }
  size_t literal_count() const { return literals_.size(); }
  size_t glob_count() const { return globs_.size(); }

  static bool parse_ip_entry(const std::string& key, IpKey& out);  // mmdb_builder.rs:338-365

 private:
  struct IpEntry { IpKey k; uint32_t data_offset; };
  struct StrEntry { std::string s; uint32_t data_offset; };
  bool build_paraglob(std::vector<uint8_t>& out);
  bool build_literal_hash(std::vector<uint8_t>& out);

  MatchMode mode_;
  DataEncoder data_;
  std::vector<IpEntry> ips_;
  std::vector<StrEntry> literals_, globs_;
  std::string database_type_;
  bool have_type_ = false;
  std::map<std::string, std::string> description_;
  uint64_t build_epoch_ = 0;
  bool have_epoch_ = false;
  std::string error_;
};

// glob syntax (matchy-paraglob/src/glob.rs:307-451) — shared with tests
struct GlobSeg {
  enum T { LITERAL = 0, STAR = 1, QUESTION = 2, CLASS = 3 } type;
  std::string lit;                                     // LITERAL (UTF-8)
  bool negated = false;                                // CLASS
  struct Item { bool range; uint32_t a, b; };
  std::vector<Item> items;
};
bool parse_glob(const std::string& pattern, std::vector<GlobSeg>& out, std::string& err);
bool glob_is_glob(const std::string& pattern);                         // paraglob_offset.rs:93-107
std::vector<std::string> glob_extract_literals(const std::string& p);  // paraglob_offset.rs:109-159

uint64_t xxh64(const uint8_t* p, size_t len, uint64_t seed);
bool parse_ipv4_text(const char* s, size_t n, uint32_t& out);     // Rust Ipv4Addr::from_str
bool parse_ipv6_text(const char* s, size_t n, uint16_t out[8]);   // Rust Ipv6Addr::from_str (incl. embedded IPv4)

}  // namespace mxy
