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
#include "synth_gen.h"

struct mxyb_builder;
extern "C" mxyb_builder* mxyb_new(int);
extern "C" void mxyb_free(mxyb_builder*);
mxy::DatabaseBuilder& mxyb_inner(mxyb_builder*);  // defined in mxy_builder.cpp

namespace {

using mxy::DataValue;
using mxy::IpKey;
using namespace sgen;
typedef unsigned __int128 u128;

template <typename F>
std::string text_of(F&& f) {  // a generator's output as a string
  uint8_t b[256];
  Out o{b, 0, sizeof b};
  f(o);
  return std::string((const char*)b, o.n < sizeof b ? o.n : sizeof b);
}

DataValue meta_row(uint64_t i) {  // ~20 distinct rows: dedup of the data section is exercised
  uint32_t r = (uint32_t)(mix(i + 77) % 20);
  DataValue m = DataValue::Map();
  m.map["threat_level"] = DataValue::String(LEVELS(r & 3));
  m.map["category"] = DataValue::String(CATS((r * 7) & 15));
  m.map["source"] = DataValue::String(SOURCES(r % 5));
  return m;
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
    const Prefix4 q = ip4_prefix(i, cfg);
    IpKey k; k.bits = q.addr; k.v6 = false; k.prefix = (uint8_t)q.plen;
    if (!seen4.insert(((uint64_t)(uint32_t)k.bits << 8) | k.prefix).second) continue;
    b.add_ip_raw(k, off_of(i));
  }
  std::unordered_set<uint64_t> seen6;
  for (uint64_t i = 0; i < c.ip6; i++) {
    const Prefix6 q = ip6_prefix(i);
    IpKey k; k.bits = (u128)q.hi << 64; k.v6 = true; k.prefix = (uint8_t)q.plen;
    if (!seen6.insert((uint64_t)(k.bits >> 64) ^ ((uint64_t)k.prefix << 56)).second) continue;
    b.add_ip_raw(k, off_of(i + 1000003));
  }
  for (uint64_t i = 0; i < c.lit; i++) b.add_literal(text_of([&](Out& o) { lit_domain(o, i); }), off_of(i + 2000003));
  for (uint64_t i = 0; i < c.glob; i++) b.add_glob(text_of([&](Out& o) { glob_pattern(o, i); }), off_of(i + 3000017));
  for (uint64_t i = 0; i < c.hash; i++) b.add_literal(text_of([&](Out& o) { hash_text(o, i); }), off_of(i + 4000037));
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
    uint8_t line[LINE_CAP];
    for (size_t k = (size_t)t; k < nblocks; k += (size_t)threads) {
      size_t o = k * BLOCK;
      if (o + BLOCK <= len) gen_block(cfg, c, b0 + k, out + o, line);
      else {
        gen_block(cfg, c, b0 + k, tmp.data(), line);
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
