// mxy_reader.h — host-side view of an `.mxy` file: section locator + MMDB value decoding + NDJSON rendering.
//
// Section discovery follows Database::from_storage (crates/matchy/src/database.rs:649-713, 1218-1278, 1315-1415)
// and MmdbHeader::from_file (crates/matchy-format/src/mmdb/format.rs:35-150).  Value decoding follows
// DataDecoder (crates/matchy-data-format/src/lib.rs:635-1047).  Rendering follows output_cli_match
// (crates/matchy/src/bin/match_processor/parallel.rs:297-369) and bin/cli_utils.rs:107-201.
// Only matched records ever reach this code; the scan itself runs on the device.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace mxy {

struct Layout {
  uint32_t node_count = 0, record_bits = 24, ip_version = 4, match_mode = 0;
  uint64_t tree_size = 0, data_start = 0;
  bool has_glob = false, has_literal = false;
  uint64_t pg_off = 0, pg_len = 0;       // PARAGLOB buffer
  uint64_t map_off = 0, map_count = 0;   // u32 data offsets indexed by glob id
  uint64_t lit_off = 0, lit_len = 0;     // LHSH section (runs to end of file as the reference reader sees it)
  uint32_t literal_count = 0, glob_count = 0;
};

bool locate_sections(const uint8_t* d, size_t n, Layout& out, std::string& err);
// the MMDB metadata map at the end of the file (mmdb/format.rs:41-48); false when there is no metadata marker
struct Value;
bool read_metadata(const uint8_t* d, size_t n, Value& out);

// decoded MMDB value
struct Value {
  enum Kind { NUL, STR, F64, F32, BYTES, UINT, INT, U128, MAP, ARR, BOOL, PTR } kind = NUL;
  uint8_t mmdb_type = 0;  // the MMDB type code the value was stored with (5/6/9 tell the UINT widths apart)
  std::string s;
  double f = 0;
  uint64_t u = 0;
  int64_t i = 0;
  unsigned __int128 big = 0;
  bool b = false;
  std::vector<std::pair<std::string, Value>> fields;  // MAP
  std::vector<Value> items;                            // ARR
};

class ValueReader {
 public:
  ValueReader(const uint8_t* base, size_t len) : p_(base), n_(len) {}
  bool read(uint32_t offset, Value& out) const;  // pointers resolved (relative to base)
 private:
  bool at(size_t& cur, Value& out, int depth) const;
  bool chase(Value& v, int depth) const;
  bool payload_size(size_t& cur, uint8_t low5, size_t& out) const;
  const uint8_t* p_;
  size_t n_;
};

// serde_json compact, object keys sorted.  f32_shortest = false: Float values widened to f64 first (`matchy match`: json!(f),
// bin/cli_utils.rs:177-201); true: shortest f32 digits (serde_json::to_string(&DataValue), the C API's matchy_result_to_json)
void render_json(const Value& v, std::string& out, bool f32_shortest = false);
void json_string(const std::string& s, std::string& out);    // serde_json string escaping
std::string cidr_text(const uint8_t* text, size_t n, unsigned prefix_len);  // format_cidr_into
std::string ipv6_text(const uint16_t seg[8]);                // Rust Display for Ipv6Addr (RFC 5952)

}  // namespace mxy
