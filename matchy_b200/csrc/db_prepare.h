// db_prepare.h — host-side preparation of an `.mxy` file for the device: section bounds (mxy_reader.h), one-time
// validation, and the small derived indexes that sit beside the unchanged sections in HBM (device_fns.cuh DbView).
// Pure C++ (no CUDA) so that tests/host_emulation can build the very same view over host memory.
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "device_fns.cuh"
#include "mxy_reader.h"

namespace mgpu {

struct PslTable {
  std::vector<uint64_t> keys;
  std::vector<uint32_t> vals;
  std::vector<uint8_t> pool;
  uint32_t mask = 0, max_len = 0;
  std::vector<uint64_t> tld;  // last-label table for tld_class(): TLD_SLOTS entries
};

// lines().map(trim).filter(non-empty, not "//…") — matchy-extractor/src/lib.rs:1552-1563
inline bool build_psl(const uint8_t* text, size_t len, PslTable& t, std::string& err) {
  std::vector<std::pair<size_t, uint32_t>> ent;
  auto ws = [](uint8_t ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || ch == '\f' || ch == '\v'; };
  size_t i = 0;
  while (i < len) {
    size_t e = i;
    while (e < len && text[e] != '\n') e++;
    size_t s0 = i, s1 = e;
    while (s0 < s1 && ws(text[s0])) s0++;
    while (s1 > s0 && ws(text[s1 - 1])) s1--;
    if (s1 > s0 && !(s1 - s0 >= 2 && text[s0] == '/' && text[s0 + 1] == '/')) {
      if (s1 - s0 > 255) { err = "PSL entry longer than 255 bytes"; return false; }
      // the device relies on this: a PSL-validated domain can never parse as an IP address (database.rs:760)
      size_t ld = s1;
      while (ld > s0 && text[ld - 1] != '.') ld--;
      bool all_digit = true;
      for (size_t k = ld; k < s1; k++) all_digit &= text[k] >= '0' && text[k] <= '9';
      if (all_digit) { err = "PSL entry with an all-numeric last label is not supported"; return false; }
      ent.push_back({s0, (uint32_t)(s1 - s0)});
    }
    i = e + 1;
  }
  if (ent.empty()) { err = "empty PSL"; return false; }
  uint32_t cap = 1024;
  while (cap < ent.size() * 3) cap <<= 1;
  t.keys.assign(cap, 0); t.vals.assign(cap, 0); t.pool.clear(); t.mask = cap - 1; t.max_len = 0;
  for (auto& en : ent) {
    const uint8_t* s = text + en.first;
    uint64_t h = MGPU_FNV_BASIS;
    for (uint32_t k = en.second; k-- > 0;) h = psl_step(h, s[k]);
    if (h == 0) h = 1;
    uint32_t slot = psl_slot(h, t.mask);
    bool dup = false;
    while (t.keys[slot] != 0) {
      if (t.keys[slot] == h && (t.vals[slot] & 0xFF) == en.second && memcmp(t.pool.data() + (t.vals[slot] >> 8), s, en.second) == 0) { dup = true; break; }
      slot = (slot + 1) & t.mask;
    }
    if (dup) continue;
    t.keys[slot] = h;
    t.vals[slot] = ((uint32_t)t.pool.size() << 8) | en.second;
    t.pool.insert(t.pool.end(), s, s + en.second);
    t.max_len = std::max(t.max_len, en.second);
  }
  // Last-label table.  ACCEPT entries: single-label PSL entries of <= 7 bytes.  SLOW entries: labels of <= 7 bytes that end
  // some multi-label entry without being an entry themselves.  (Longer labels always take the general walk.)
  {
    t.tld.assign(TLD_SLOTS, 0);
    auto key_of = [&](const uint8_t* s, uint32_t n) { uint64_t k = 0; for (uint32_t i = 0; i < n; i++) k |= (uint64_t)s[i] << (8 * i); return k | ((uint64_t)n << 56); };
    auto find = [&](uint64_t key) -> uint64_t* {
      uint32_t slot = tld_slot(key);
      while (t.tld[slot] != 0 && (t.tld[slot] & ~TLD_SLOW) != key) slot = (slot + 1) & (TLD_SLOTS - 1);
      return &t.tld[slot];
    };
    size_t used = 0;
    for (int pass = 0; pass < 2; pass++)
      for (auto& en : ent) {
        const uint8_t* s = text + en.first;
        uint32_t n = en.second, ld = n;
        while (ld > 0 && s[ld - 1] != '.') ld--;
        bool single = ld == 0;
        uint32_t ll = n - ld;
        if (ll == 0 || ll > 7) continue;
        if ((pass == 0) != single) continue;
        uint64_t key = key_of(s + ld, ll);
        uint64_t* e = find(key);
        if (*e != 0) continue;  // pass 1: already an ACCEPT entry (or a SLOW duplicate)
        if (++used > TLD_SLOTS * 3 / 4) { err = "PSL has too many distinct last labels for the device table"; return false; }
        *e = single ? key : (key | TLD_SLOW);
      }
  }
  return true;
}

struct PreparedDb {
  mxy::Layout L;
  DbView view;                     // scalar members set; section pointers are the caller's business
  std::vector<uint32_t> lh_index;  // pattern_id -> data_offset
  std::vector<uint32_t> aclh;      // literal id -> (abs offset, count)
  std::vector<uint64_t> lh_bloom;  // blocked Bloom filter over the literal table's stored hashes
  std::vector<uint32_t> gram2, gram3;  // bitmaps over the first 2 / 3 bytes of every AC literal
  std::vector<uint64_t> pfx_keys;      // prefix map (see DbView::ac_pfx_*)
  std::vector<uint32_t> pfx_vals;
  std::vector<uint32_t> top16;         // IPv4 walk state after v4_top_bits bits (DbView::v4_top16)
  std::vector<uint8_t> top16_depth;
  std::vector<uint32_t> v6_top16;      // IPv6 walk state after 16 bits, IPv6 trees only (DbView::v6_top16)
  std::vector<uint8_t> v6_top16_depth;
  std::vector<uint32_t> gen2, gen3;    // first 2 / 3 bytes of every AC literal that leads to an UNANCHORED pattern (see string_filters)
  std::vector<uint32_t> hot;           // fast string path: hot (shared-memory) and cold (L2) Bloom filters
  std::vector<uint64_t> cold;
  uint32_t ac_node_count = 0;
};

struct FilterKey { uint32_t h, tag; };  // key hash (device_fns.cuh), tag class

inline uint32_t prep_le32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

inline bool prepare_db(const uint8_t* d, size_t n, PreparedDb& P, std::string& err) {
  mxy::Layout& L = P.L;
  if (!mxy::locate_sections(d, n, L, err)) return false;
  DbView& db = P.view;
  memset(&db, 0, sizeof db);
  db.node_count = L.node_count; db.record_bits = L.record_bits; db.ip_version = L.ip_version;
  db.match_mode = L.match_mode;
  db.has_ip = 1;  // the reference sets ip_header for every MMDB-format file (database.rs:664-700)
  // --- tree: validate once so that the per-lookup error paths of the reference cannot trigger
  {
    const uint64_t data_len = n - L.data_start;
    const size_t nb = L.record_bits == 24 ? 6 : L.record_bits == 28 ? 7 : 8;
    auto rec = [&](uint32_t node, int side) -> uint32_t {
      const uint8_t* b = d + (size_t)node * nb;
      if (L.record_bits == 24) { b += side * 3; return ((uint32_t)b[0] << 16) | ((uint32_t)b[1] << 8) | b[2]; }
      if (L.record_bits == 28) return side == 0 ? (((uint32_t)(b[3] >> 4) << 24) | ((uint32_t)b[0] << 16) | ((uint32_t)b[1] << 8) | b[2])
                                                 : (((uint32_t)(b[3] & 15) << 24) | ((uint32_t)b[4] << 16) | ((uint32_t)b[5] << 8) | b[6]);
      b += side * 4; return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    };
    bool any_data = false;
    for (uint32_t i = 0; i < L.node_count; i++)
      for (int s = 0; s < 2; s++) {
        uint32_t r = rec(i, s);
        any_data |= r > L.node_count;
        if (r > L.node_count && ((uint64_t)r < (uint64_t)L.node_count + 16 || (uint64_t)r - L.node_count - 16 >= data_len)) {
          err = "corrupt IP search tree: record is neither a node, the empty marker nor a data pointer";
          return false;
        }
      }
    uint32_t node = 0;
    if (L.ip_version == 6 && L.node_count > 0) {  // find_ipv4_start_node, once (tree.rs:258-277)
      for (int k = 0; k < 96; k++) { uint32_t r = rec(node, 0); if (r >= L.node_count) break; node = r; }
    }
    db.v4_start_node = node;
    if (L.node_count == 0) db.has_ip = 0;
    db.ip_empty = any_data ? 0 : 1;
    // Walk state after the first TB address bits, by depth-first expansion from the start node (2^TB leaves): IPv4 queries
    // (TB = 16, or 20 for trees with more than 2^16 nodes, where a /16 bucket still holds a subtree) and, in IPv6 trees, IPv6
    // queries (16 bits from the root).
    if (L.node_count > 0 && L.node_count < (1u << 28) && (uint64_t)L.node_count + 16 + data_len < (1ull << 28)) {
      auto expand = [&](uint32_t root, uint32_t tb, std::vector<uint32_t>& tab, std::vector<uint8_t>& dep) {
        tab.assign((size_t)1 << tb, 0);
        dep.assign((size_t)1 << tb, 0);
        struct Fr { uint32_t node, depth, prefix; };
        std::vector<Fr> st{{root, 0, 0}};
        while (!st.empty()) {
          Fr f = st.back(); st.pop_back();
          if (f.depth == tb) { tab[f.prefix] = f.node; continue; }
          for (int side = 0; side < 2; side++) {
            const uint32_t r = rec(f.node, side), pfx = f.prefix | ((uint32_t)side << (tb - 1 - f.depth));
            if (r < L.node_count) { st.push_back(Fr{r, f.depth + 1, pfx}); continue; }
            const uint32_t span = 1u << (tb - 1 - f.depth);  // every TB-bit prefix below this branch shares the outcome
            for (uint32_t k = 0; k < span; k++) {
              tab[pfx + k] = r == L.node_count ? (1u << 28) : ((2u << 28) | r);
              dep[pfx + k] = (uint8_t)(f.depth + 1);
            }
          }
        }
      };
      // (24 bits = a 64 MiB table for trees of half a million nodes and more: the IP-trie kernel moves ~32 bytes of L2 traffic per
      //  tree record read, and at a million prefixes the 4 levels this saves are a third of its loads)
      db.v4_top_bits = L.node_count > (1u << 19) ? 24u : (L.node_count > (1u << 16) ? 20u : 16u);
      expand(node, db.v4_top_bits, P.top16, P.top16_depth);
      if (L.ip_version == 6) expand(0, 16, P.v6_top16, P.v6_top16_depth);
    }
  }
  struct Anchor { const uint8_t* p; uint32_t m, tag; };  // anchor literal of a suffix- / prefix-anchored glob
  std::vector<Anchor> anchors;
  std::vector<FilterKey> lit_tail_keys, glob_keys;
  std::vector<uint32_t> lit_full_keys;
  bool fast = L.match_mode == 0;
  // --- literal hash
  if (L.has_literal) {
    const uint8_t* lh = d + L.lit_off;
    uint64_t len = L.lit_len;
    uint32_t strings_offset = prep_le32(lh + 16), strings_size = prep_le32(lh + 20), num_shards = prep_le32(lh + 24);
    if (num_shards == 0 || 32 + ((uint64_t)num_shards + 1) * 4 > len) { err = "literal hash: bad shard table"; return false; }
    uint64_t maps = (uint64_t)strings_offset + strings_size;
    if (maps + 4 <= len) {  // pattern_id -> data_offset, first entry wins (== the reference's linear scan, lib.rs:560-572)
      uint64_t cnt = prep_le32(lh + maps);
      if (maps + 4 + cnt * 8 > len) cnt = (len - maps - 4) / 8;  // the scan stops at the first out-of-bounds entry
      uint32_t max_id = 0; bool any = false;
      for (uint64_t k = 0; k < cnt; k++) { uint32_t id = prep_le32(lh + maps + 4 + k * 8); if (id != NONE32) { max_id = std::max(max_id, id); any = true; } }
      if (any) {
        if (max_id > 0x10000000u) { err = "literal hash: pattern ids too sparse"; return false; }
        P.lh_index.assign((size_t)max_id + 1, NONE32);
        for (uint64_t k = 0; k < cnt; k++) {
          uint32_t id = prep_le32(lh + maps + 4 + k * 8), off = prep_le32(lh + maps + 8 + k * 8);
          if (id <= max_id && P.lh_index[id] == NONE32) P.lh_index[id] = off;
        }
      }
    }
    {  // Bloom filter over every occupied slot's stored hash (the same XXH64 the lookup computes): no false negatives
      uint32_t table_size = prep_le32(lh + 12);
      uint64_t tstart = 32 + ((uint64_t)num_shards + 1) * 4;
      if (tstart + (uint64_t)table_size * 16 > len) table_size = (uint32_t)((len - tstart) / 16);
      uint64_t words = 1024;
      while (words * 4 < table_size) words <<= 1;  // >= 16 bits per slot
      P.lh_bloom.assign((size_t)words, 0);
      for (uint32_t k = 0; k < table_size; k++) {
        const uint8_t* e = lh + tstart + (uint64_t)k * 16;
        if (prep_le32(e + 8) == NONE32) continue;
        uint64_t h; memcpy(&h, e, 8);
        P.lh_bloom[(size_t)((uint32_t)(h >> 20) & (uint32_t)(words - 1))] |= (1ULL << (h & 63)) | (1ULL << ((h >> 6) & 63)) | (1ULL << ((h >> 12) & 63));
      }
      db.lh_bloom_mask = (uint32_t)(words - 1);
      // fast string path: (last min(len,8) bytes) -> hot key, (len, first 8, last 8) -> cold key, for every stored string
      for (uint32_t k = 0; k < table_size; k++) {
        const uint8_t* e = lh + tstart + (uint64_t)k * 16;
        uint32_t so = prep_le32(e + 8);
        if (so == NONE32) continue;
        uint64_t abs = (uint64_t)strings_offset + so;
        if (abs + 2 > len) continue;
        uint32_t sl = (uint32_t)lh[abs] | ((uint32_t)lh[abs + 1] << 8);
        if (sl == 0 || abs + 2 + sl > len) continue;
        const uint8_t* sp = lh + abs + 2;
        uint32_t th, hh, fh;
        lit_key_hashes(sp, sl, th, hh, fh);
        lit_tail_keys.push_back(FilterKey{th, (uint32_t)TAG_LIT_TAIL});
        lit_tail_keys.push_back(FilterKey{hh, (uint32_t)TAG_LIT_HEAD});
        lit_full_keys.push_back(fh);
      }
    }
    db.lh_len = len; db.has_literal = 1;
    db.lh_num_shards = num_shards; db.lh_strings_offset = strings_offset; db.lh_table_start = 32 + (num_shards + 1) * 4;
    db.lh_data_index_n = (uint32_t)P.lh_index.size();
  }
  // --- paraglob
  if (L.has_glob) {
    const uint8_t* pg = d + L.pg_off;
    uint32_t pg_len = (uint32_t)L.pg_len;
    db.ac_start = prep_le32(pg + 20); db.ac_size = prep_le32(pg + 24);
    if ((uint64_t)db.ac_start + db.ac_size > pg_len) { err = "paraglob: AC buffer out of range"; return false; }
    P.ac_node_count = prep_le32(pg + 16);
    db.patterns_offset = prep_le32(pg + 36);
    uint64_t un = (uint64_t)prep_le32(pg + 40) + prep_le32(pg + 44);
    db.wild_off = (uint32_t)(un + (8 - un % 8) % 8);
    db.wild_count = prep_le32(pg + 60);
    db.glob_segments_offset = prep_le32(pg + 104);
    uint32_t pattern_count = prep_le32(pg + 32);
    // ACLH -> dense literal-id index over every occupied slot (keeps the unpinned FxHash off the read side)
    uint32_t ao = prep_le32(pg + 96), acnt = prep_le32(pg + 100);
    if (ao != 0 && acnt != 0 && (uint64_t)ao + 24 <= pg_len && memcmp(pg + ao, "ACLH", 4) == 0 && prep_le32(pg + ao + 4) == 1) {
      uint32_t table_size = prep_le32(pg + ao + 12), lists = prep_le32(pg + ao + 16);
      uint32_t max_id = 0; bool any = false;
      for (uint32_t s = 0; s < table_size; s++) {
        uint64_t eo = (uint64_t)ao + 24 + (uint64_t)s * 16;
        if (eo + 16 > pg_len) break;
        uint32_t id = prep_le32(pg + eo);
        if (id != NONE32) { max_id = std::max(max_id, id); any = true; }
      }
      if (any) {
        if (max_id > 0x10000000u) { err = "paraglob: AC literal ids too sparse"; return false; }
        P.aclh.assign(((size_t)max_id + 1) * 2, 0);
        std::vector<uint8_t> seen((size_t)max_id + 1, 0);
        for (uint32_t s = 0; s < table_size; s++) {
          uint64_t eo = (uint64_t)ao + 24 + (uint64_t)s * 16;
          if (eo + 16 > pg_len) break;
          uint32_t id = prep_le32(pg + eo);
          if (id == NONE32 || seen[id]) continue;
          seen[id] = 1;
          uint64_t abs = (uint64_t)ao + lists + prep_le32(pg + eo + 4);
          uint32_t cnt = prep_le32(pg + eo + 8);
          if (abs + (uint64_t)cnt * 4 > pg_len) continue;  // read_pattern_list -> None -> no candidates
          P.aclh[2 * (size_t)id] = (uint32_t)abs; P.aclh[2 * (size_t)id + 1] = cnt;
        }
      }
    }
    // The device walks the automaton without per-step bounds/alignment checks (the reference checks bounds on every
    // access and tolerates any alignment).  Verify every offset once here instead; builder-written files always pass.
    {
      const uint8_t* ac = pg + db.ac_start;
      const uint64_t acn = db.ac_size;
      if (acn != 0 && (acn < 20 || (uint64_t)P.ac_node_count * 20 > acn)) { err = "paraglob: AC node table out of range"; return false; }
      const uint64_t nodes = P.ac_node_count;
      bool ok = true;
      auto node_ref = [&](uint32_t off) { if ((off % 20) != 0 || (uint64_t)off / 20 >= nodes) ok = false; };
      for (uint64_t i = 0; i < nodes && ok; i++) {
        const uint8_t* nd = ac + i * 20;
        uint32_t kind = nd[0], cnt = nd[2], pc = nd[3], eo = prep_le32(nd + 12), po = prep_le32(nd + 16);
        node_ref(prep_le32(nd + 8));
        if (pc && ((po & 3) || (uint64_t)po + (uint64_t)pc * 4 > acn)) ok = false;
        if (kind == 1) { node_ref(eo); if (eo == 0) ok = false; }
        else if (kind == 2) {
          if ((eo & 3) || (uint64_t)eo + (uint64_t)cnt * 8 > acn) { ok = false; break; }
          for (uint32_t k = 0; k < cnt; k++) { uint32_t t = prep_le32(ac + eo + k * 8 + 4); node_ref(t); if (t == 0) ok = false; }
        } else if (kind == 3) {
          if ((eo & 3) || (uint64_t)eo + 1024 > acn) { ok = false; break; }
          for (uint32_t k = 0; k < 256; k++) { uint32_t t = prep_le32(ac + eo + k * 4); if (t) node_ref(t); }
        } else if (kind != 0) ok = false;
      }
      for (size_t k = 0; k + 1 < P.aclh.size(); k += 2) if (P.aclh[k] & 3) ok = false;
      if (!ok) { err = "paraglob automaton with out-of-range or unaligned offsets is not supported on the device"; return false; }
    }
    // Prefix bitmaps for the anchored walk: every trie path of length 2 / 3 from the root.  The anchored walk is only
    // exact when no literal is shorter than 3 bytes (no outputs at depth < 3) and no node list is saturated at 255.
    {
      const uint8_t* ac = pg + db.ac_start;
      P.gram2.assign(65536 / 32, 0);
      P.gram3.assign((1u << 24) / 32, 0);
      bool anchored = P.ac_node_count > 0;
      auto children = [&](uint32_t off, std::vector<std::pair<uint8_t, uint32_t>>& out) {
        out.clear();
        const uint8_t* nd = ac + off;
        uint32_t kind = nd[0], cnt = nd[2], eo = prep_le32(nd + 12);
        if (kind == 1) out.push_back({nd[1], eo});
        else if (kind == 2) for (uint32_t k = 0; k < cnt; k++) out.push_back({ac[eo + k * 8], prep_le32(ac + eo + k * 8 + 4)});
        else if (kind == 3) for (uint32_t k = 0; k < 256; k++) { uint32_t t = prep_le32(ac + eo + k * 4); if (t) out.push_back({(uint8_t)k, t}); }
      };
      for (uint64_t i = 0; i < P.ac_node_count; i++) if (ac[i * 20 + 3] == 255) anchored = false;
      if (P.ac_node_count > 0) {
        if (ac[3] != 0) anchored = false;
        std::vector<std::pair<uint8_t, uint32_t>> c1, c2, c3;
        children(0, c1);
        for (auto& a1 : c1) {
          if (ac[a1.second + 3] != 0) anchored = false;
          children(a1.second, c2);
          for (auto& a2 : c2) {
            if (ac[a2.second + 3] != 0) anchored = false;
            uint32_t g2 = ((uint32_t)a1.first << 8) | a2.first;
            P.gram2[g2 >> 5] |= 1u << (g2 & 31);
            children(a2.second, c3);
            for (auto& a3 : c3) { uint32_t g3 = (g2 << 8) | a3.first; P.gram3[g3 >> 5] |= 1u << (g3 & 31); }
          }
        }
      }
      db.ac_anchored = anchored ? 1u : 0u;
      // Prefix map: every depth-8 trie node, and every node of depth 3..7 with outputs (a literal of that length or a
      // merged suffix literal), keyed by the path bytes.  Exact: a jump lands on the node the step-by-step walk reaches.
      if (anchored) {
        struct Fr { uint32_t off; uint32_t depth; uint64_t v; };
        std::vector<Fr> st{{0, 0, 0}};
        struct En { uint64_t v; uint32_t m, off; };
        std::vector<En> ents;
        std::vector<std::pair<uint8_t, uint32_t>> ch;
        uint32_t short_lens = 0;
        while (!st.empty()) {
          Fr f = st.back(); st.pop_back();
          if (f.depth >= 3 && f.depth < 8 && ac[f.off + 3] != 0) { ents.push_back({f.v, f.depth, f.off}); short_lens |= 1u << f.depth; }
          if (f.depth == 8) { ents.push_back({f.v, 8, f.off}); continue; }
          children(f.off, ch);
          for (auto& c1 : ch) st.push_back({c1.second, f.depth + 1, f.v | ((uint64_t)c1.first << (8 * f.depth))});
        }
        uint32_t cap = 1024;
        while ((uint64_t)cap < ents.size() * 2 + 2) cap <<= 1;
        P.pfx_keys.assign(cap, 0);
        P.pfx_vals.assign((size_t)cap * 2, 0);
        for (auto& e : ents) {
          uint32_t slot = prefix_slot(e.v, e.m, cap - 1);
          while (P.pfx_vals[2 * (size_t)slot + 1] != 0) slot = (slot + 1) & (cap - 1);
          P.pfx_keys[slot] = e.v; P.pfx_vals[2 * (size_t)slot] = e.off; P.pfx_vals[2 * (size_t)slot + 1] = e.m;
        }
        db.ac_pfx_mask = cap - 1;
        db.ac_short_lens = short_lens;
      }
    }
    // glob segments: the device matcher keeps one frame per '*'
    uint32_t gso = db.glob_segments_offset;
    for (uint32_t pid = 0; pid < pattern_count; pid++) {
      uint64_t io = (uint64_t)gso + (uint64_t)pid * 8;
      if (io + 8 > pg_len) break;
      uint32_t first = prep_le32(pg + io), cnt = (uint32_t)pg[io + 4] | ((uint32_t)pg[io + 5] << 8), stars = 0;
      for (uint32_t s = 0; s < cnt; s++) { uint64_t so = (uint64_t)first + (uint64_t)s * 12; if (so + 12 > pg_len) break; if (pg[so] == 1) stars++; }
      if (stars > MGPU_GLOB_MAX_STARS) { err = "glob pattern with more than 24 '*' segments is not supported on the device"; return false; }
    }
    db.pg_len = pg_len; db.has_glob = 1;
    db.pg_align = (uint32_t)(L.pg_off & 3);
    // the device reads these with aligned 32-bit loads (the reference tolerates any alignment here: raw pointer reads)
    if ((db.ac_start & 3) || (db.patterns_offset & 3) || (db.wild_off & 3) || (ao & 3) || (db.pg_align != 0)) {
      err = "paraglob section with unaligned tables is not supported on the device";
      return false;
    }
    db.aclh_n = (uint32_t)(P.aclh.size() / 2);
    db.glob_data_n = (uint32_t)L.map_count;
    // fast string path: classify every pattern some AC literal leads to (string_filters() in device_fns.cuh)
    if (db.wild_count != 0) fast = false;
    if (fast) {
      std::vector<uint8_t> reach(pattern_count, 0), generic(pattern_count, 0);
      bool any_generic = false;
      for (size_t lit = 0; lit + 1 < P.aclh.size(); lit += 2)
        for (uint32_t j = 0; j < P.aclh[lit + 1]; j++) { uint32_t pid = prep_le32(pg + P.aclh[lit] + j * 4); if (pid < pattern_count) reach[pid] = 1; else fast = false; }
      for (uint32_t pid = 0; pid < pattern_count && fast; pid++) {
        if (!reach[pid]) continue;
        uint64_t eo = (uint64_t)db.patterns_offset + (uint64_t)pid * 16;
        if (eo + 16 > pg_len) continue;                     // find_all skips it
        uint32_t entry_id = prep_le32(pg + eo);
        if (pg[eo + 4] == 0) { generic[pid] = 1; any_generic = true; continue; }  // literal-type pattern: substring semantics (quirk 12)
        uint64_t io = (uint64_t)gso + (uint64_t)entry_id * 8;
        if (io + 8 > pg_len) continue;                      // match_glob_from_buffer fails: never a match
        uint32_t first = prep_le32(pg + io), cnt = (uint32_t)pg[io + 4] | ((uint32_t)pg[io + 5] << 8);
        if (cnt == 0) continue;                             // matches the empty text only; tokens are never empty
        if ((uint64_t)first + (uint64_t)cnt * 12 > pg_len) { fast = false; break; }
        auto lit_of = [&](uint32_t idx, const uint8_t*& ptr, uint32_t& dl) -> int {  // 1 literal, 0 other, -1 unmatchable literal
          const uint8_t* sh = pg + first + (uint64_t)idx * 12;
          if (sh[0] != 0) return 0;
          dl = prep_le32(sh + 4);
          uint32_t doff = prep_le32(sh + 8);
          if ((uint64_t)doff + dl > pg_len) return -1;
          ptr = pg + doff;
          return dl > 0 ? 1 : 0;
        };
        const uint8_t* ptr = nullptr; uint32_t dl = 0;
        int r = lit_of(cnt - 1, ptr, dl);
        if (r < 0) continue;
        if (r == 1) { glob_keys.push_back(FilterKey{glob_key_hash(ptr, dl, TAG_GLOB_S), (uint32_t)TAG_GLOB_S}); anchors.push_back(Anchor{ptr, dl, (uint32_t)TAG_GLOB_S}); db.glob_s_lens |= 1u << glob_key_len(dl); continue; }
        r = lit_of(0, ptr, dl);
        if (r < 0) continue;
        if (r == 1) { glob_keys.push_back(FilterKey{glob_key_hash(ptr, dl, TAG_GLOB_P), (uint32_t)TAG_GLOB_P}); anchors.push_back(Anchor{ptr, dl, (uint32_t)TAG_GLOB_P}); db.glob_p_lens |= 1u << glob_key_len(dl); continue; }
        generic[pid] = 1; any_generic = true;               // neither end is a literal: any position can match
      }
      // Unanchored patterns: a pattern is only ever a candidate through one of its AC literals (find_all), so a token can
      // match one only if some literal that leads to an unanchored pattern OCCURS in it.  Collect the first 2 / 3 bytes of
      // those literals: in the (breadth-first) trie walk the first node that lists a literal id is the node whose path IS
      // the literal (deeper nodes list it as a merged suffix output).  Needs complete output lists (db.ac_anchored).
      if (fast && any_generic) {
        if (!db.ac_anchored) fast = false;
        else {
          std::vector<uint8_t> lit_generic(P.aclh.size() / 2, 0);
          for (size_t lit = 0; lit + 1 < P.aclh.size(); lit += 2)
            for (uint32_t j = 0; j < P.aclh[lit + 1]; j++) { uint32_t pid = prep_le32(pg + P.aclh[lit] + j * 4); if (pid < pattern_count && generic[pid]) lit_generic[lit / 2] = 1; }
          P.gen2.assign(65536 / 32, 0);
          P.gen3.assign((1u << 24) / 32, 0);
          const uint8_t* ac = pg + db.ac_start;
          struct Fr { uint32_t off, depth, first3; };
          std::vector<Fr> cur{{0, 0, 0}}, nxt;
          std::vector<uint8_t> seen(lit_generic.size(), 0);
          while (!cur.empty()) {
            nxt.clear();
            for (const Fr& f : cur) {
              const uint8_t* nd = ac + f.off;
              uint32_t pc = nd[3], po = prep_le32(nd + 16);
              for (uint32_t k = 0; k < pc; k++) {
                uint32_t lit = prep_le32(ac + po + k * 4);
                if (lit >= seen.size() || seen[lit]) continue;
                seen[lit] = 1;
                if (!lit_generic[lit]) continue;
                uint32_t g3 = f.first3 & 0xFFFFFFu, g2 = g3 >> 8;  // (depth >= 3: every literal has at least 3 bytes)
                P.gen2[g2 >> 5] |= 1u << (g2 & 31);
                P.gen3[g3 >> 5] |= 1u << (g3 & 31);
              }
              uint32_t kind = nd[0], cnt = nd[2], eo = prep_le32(nd + 12);
              auto push = [&](uint8_t ch, uint32_t t) { nxt.push_back(Fr{t, f.depth + 1, f.depth < 3 ? ((f.first3 << 8) | ch) : f.first3}); };
              if (kind == 1) push(nd[1], eo);
              else if (kind == 2) for (uint32_t k = 0; k < cnt; k++) push(ac[eo + k * 8], prep_le32(ac + eo + k * 8 + 4));
              else if (kind == 3) for (uint32_t k = 0; k < 256; k++) { uint32_t t = prep_le32(ac + eo + k * 4); if (t) push((uint8_t)k, t); }
            }
            cur.swap(nxt);
          }
          db.has_generic = 1;
        }
      }
    }
  }
  // --- fast string path filters
  db.fast_ok = fast ? 1u : 0u;
  if (fast) {
    P.hot.assign(HOT_WORDS, 0);
    auto try_tag = [&](const std::vector<FilterKey>& keys, uint32_t tag) {
      // a tag class joins the hot filter only while the filter stays selective (<= 30 % of its bits set)
      std::vector<uint32_t> h = P.hot;
      bool any = false;
      for (auto& key : keys) {
        if (key.tag != tag) continue;
        hot_set(h.data(), key.h);
        any = true;
      }
      uint64_t bits = 0;
      for (uint32_t w : h) bits += (uint64_t)__builtin_popcount(w);
      if (any && bits * 10 > (uint64_t)HOT_WORDS * 32 * 3) return;
      P.hot.swap(h);
      db.hot_tags |= 1u << tag;
    };
    // hot keys of the glob classes (string_gate): keys of 1..3 bytes as they are; for the longer ones the gate key = the
    // last / first G bytes of the anchor literal, G = the shortest key length >= 4 of the class
    auto lowest_len = [](uint32_t lens) { lens &= ~0xFu; return lens ? (uint32_t)__builtin_ctz(lens) : 0u; };
    db.glob_s_gate = lowest_len(db.glob_s_lens);
    db.glob_p_gate = lowest_len(db.glob_p_lens);
    std::vector<FilterKey> glob_hot_keys;
    for (const Anchor& an : anchors) {
      const uint32_t k = glob_key_len(an.m);
      const uint32_t h = k < 4 ? glob_key_hash(an.p, an.m, an.tag) : glob_gate_hash(an.p, an.m, an.tag == TAG_GLOB_S ? db.glob_s_gate : db.glob_p_gate, an.tag);
      glob_hot_keys.push_back(FilterKey{h, an.tag});
    }
    try_tag(glob_hot_keys, TAG_GLOB_S);
    try_tag(glob_hot_keys, TAG_GLOB_P);
    try_tag(lit_tail_keys, TAG_LIT_TAIL);
    try_tag(lit_tail_keys, TAG_LIT_HEAD);
    if (db.has_literal && !((db.hot_tags >> TAG_LIT_TAIL) & 1u) && !((db.hot_tags >> TAG_LIT_HEAD) & 1u)) db.gate_inline |= G_LIT;
    if (db.glob_s_lens && !((db.hot_tags >> TAG_GLOB_S) & 1u)) db.gate_inline |= G_S;
    if (db.glob_p_lens && !((db.hot_tags >> TAG_GLOB_P) & 1u)) db.gate_inline |= G_P;
    if (db.has_generic) db.gate_inline |= G_GEN;
    uint64_t nkeys = glob_keys.size() + lit_full_keys.size();
    uint64_t words = 1024;
    while (words * 4 < nkeys) words <<= 1;  // >= 16 bits per key
    P.cold.assign((size_t)words, 0);
    for (auto& key : glob_keys) cold_set(P.cold.data(), (uint32_t)(words - 1), key.h);
    for (uint32_t h : lit_full_keys) cold_set(P.cold.data(), (uint32_t)(words - 1), h);
    db.cold_mask = (uint32_t)(words - 1);
  }
  return true;
}

}  // namespace mgpu
