// emu.cpp — HOST EMULATION OF THE DEVICE LOGIC.  TEST INFRASTRUCTURE ONLY, never part of the product.
//
// There is no GPU in the development container, so the kernel logic is written as host/device functions
// (matchy_b200/csrc/device_fns.cuh, tokenize.cuh, db_prepare.h) and this file re-wires the kernels of engine.cu
// around them with plain loops: a "warp" is a loop over 32 lanes, ballots and shuffles are arrays.  The CPU test
// suite (-m "not gpu") checks this emulation against the oracle so that logic bugs are found before GPU time is
// spent; the -m gpu suite then checks the real kernels.  Nothing under matchy_b200/ loads or links this file,
// and the product has no CPU execution path: libmatchy_b200.so fails without a CUDA device.
#include <algorithm>
#include <array>
#include <chrono>
#include <memory>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/matchy_b200.h"
#include "../../matchy_b200/csrc/db_prepare.h"
#include "../../matchy_b200/csrc/tokenize.cuh"
#include "../../matchy_b200/csrc/crypto_addr.cuh"
#include "../../matchy_b200/csrc/host_sort.h"

using namespace mgpu;

struct Cand { uint32_t start, len; };
struct StrTok { uint32_t start, len, type; };
struct IpTok { uint32_t start, len, type; uint32_t w[4]; };

struct emu_ctx {
  std::vector<uint8_t> file;
  PreparedDb P;
  PslTable psl;
  DbView db;
  bool loaded = false;
  bool use_anchored = true;  // false: force the Aho-Corasick formulation (both are exercised by the tests)
  bool force_generic = false;  // true: never take the fast string path (string_filters), like mgpu_set_ac_mode(2)
  std::vector<mgpu_match> recs;
  std::vector<mgpu_id_pair> ids;
  mgpu_counters counters;
  std::vector<StrTok> str;
  std::vector<IpTok> ip;
  std::string err;
  uint64_t filter_pass[3] = {0, 0, 0};  // fast path: tokens that passed the literal / glob filters, tokens tested
  uint64_t gate_pass[4] = {0, 0, 0, 0};  // stage 1: tokens owing stage 2 for the literal / suffix / prefix class, any class
};

// ---- K1: tokenize_kernel, one "warp" per range of tiles -------------------------------------------------
static void emu_tokenize(const uint8_t* buf, uint64_t lo, uint64_t n, uint32_t flags, uint64_t nwarps,
                         std::vector<Cand>& qd, std::vector<Cand>& qh, std::vector<uint32_t>& qa, std::vector<uint32_t>& qc,
                         std::vector<Cand>& qn, std::vector<Cand>& ql, uint64_t& lines) {
  const uint64_t tiles = (n + TILE_BYTES - 1) / TILE_BYTES;
  const uint64_t tpw = (tiles + nwarps - 1) / nwarps;
  const bool want_dot = (flags & (MGPU_X_IPV4 | MGPU_X_DOMAINS)) != 0, want_hash = (flags & MGPU_X_HASHES) != 0;
  const bool want_at = (flags & MGPU_X_EMAILS) != 0, want_c2 = (flags & MGPU_X_IPV6) != 0, want_long = (flags & MGPU_X_CRYPTO) != 0;
  for (uint64_t w = 0; w < nwarps; w++) {
    uint64_t t0 = w * tpw, t1 = std::min(t0 + tpw, tiles);
    if (t0 >= t1) continue;
    TileCarry cy = range_prologue(buf, lo, t0 * TILE_BYTES);
    for (uint64_t t = t0; t < t1; t++) {
      const uint64_t tile_base = t * TILE_BYTES;
      LaneMasks m[32];
      for (uint32_t lane = 0; lane < 32; lane++) {
        uint64_t p = tile_base + (uint64_t)lane * 32;
        LaneMasks x{0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t i = 0; i < 32; i++) {
          uint64_t q = p + i;
          uint32_t c = (q >= lo && q < n) ? class_bits(buf[q]) : (1u << CLS_B);
          x.B |= ((c >> CLS_B) & 1u) << i; x.DOT |= ((c >> CLS_DOT) & 1u) << i; x.AT |= ((c >> CLS_AT) & 1u) << i;
          x.CL |= ((c >> CLS_CL) & 1u) << i; x.NL |= ((c >> CLS_NL) & 1u) << i; x.DM |= ((c >> CLS_DM) & 1u) << i;
          x.HX |= ((c >> CLS_HX) & 1u) << i; x.DASH |= ((c >> CLS_DASH) & 1u) << i;
        }
        m[lane] = x;
        lines += (uint64_t)__builtin_popcount(x.NL);
      }
      if (!(cy.prev & PV_T)) cy.open_start = tile_base;
      uint32_t pT[32], pv[32], S[32], bad[32], bad_end[32], T[32];
      for (uint32_t lane = 0; lane < 32; lane++) {
        pv[lane] = lane ? prev_bits_of(m[lane - 1]) : cy.prev;
        pT[lane] = pv[lane] & PV_T;
        T[lane] = ~m[lane].B;
        S[lane] = T[lane] & ~((T[lane] << 1) | pT[lane]);
        domain_rule_masks(m[lane], S[lane], pv[lane], bad[lane], bad_end[lane]);
      }
      // the three "word contains ..." chains of tokenize_kernel: ballots are bit loops here
      uint32_t H[4][32];
      uint32_t pb = 0;
      for (uint32_t lane = 0; lane < 32; lane++) pb |= (T[lane] == 0xFFFFFFFFu ? 1u : 0u) << lane;
      for (int cls = 0; cls < 4; cls++) {
        uint32_t gen = 0, Y[32];
        for (uint32_t lane = 0; lane < 32; lane++) {
          Y[lane] = cls == 0 ? (T[lane] & (~m[lane].DM | bad[lane])) : cls == 1 ? m[lane].DOT : cls == 2 ? (T[lane] & ~m[lane].HX) : (T[lane] & ~m[lane].HX & ~m[lane].DOT);
          gen |= chain_gen(T[lane], Y[lane]) << lane;
        }
        uint32_t& c0 = cls == 0 ? cy.cBad : cls == 1 ? cy.cDot : cls == 2 ? cy.cNhx : cy.cNhd;
        uint32_t co, cv = carry_chain(gen, pb & ~gen, c0, co);
        for (uint32_t lane = 0; lane < 32; lane++) H[cls][lane] = chain_ends(T[lane], Y[lane], (cv >> lane) & 1u, m[lane].B);
        c0 = co;
      }
      uint32_t hasB = 0;
      for (uint32_t lane = 0; lane < 32; lane++) hasB |= (m[lane].B != 0 ? 1u : 0u) << lane;
      for (uint32_t lane = 0; lane < 32; lane++) {
        uint64_t p = tile_base + (uint64_t)lane * 32;
        uint32_t E = m[lane].B & ((T[lane] << 1) | pT[lane]);
        uint32_t candDot = want_dot ? (H[1][lane] & ~H[0][lane] & ~bad_end[lane]) : 0u;
        uint32_t candHex = (want_hash && pT[lane]) ? (E & ~H[2][lane] & (m[lane].B & (0u - m[lane].B))) : 0u;
        uint32_t Bprev = lane ? m[lane - 1].B : cy.prevB;
        if (candHex && (Bprev >> __builtin_ctz(candHex)) != 0) candHex = 0;  // shorter than 32 bytes
        uint32_t candAt = want_at ? m[lane].AT : 0u;
        uint32_t cl1 = (m[lane].CL << 1) | ((pv[lane] >> 3) & 1u), cl2 = (m[lane].CL << 2) | (((pv[lane] >> 3) & 1u) << 1) | ((pv[lane] >> 4) & 1u);
        uint32_t candC2 = want_c2 ? (m[lane].CL & cl1 & ~cl2) : 0u;
        uint32_t src = lane_below_with_boundary(hasB, lane);
        uint64_t lane_open = src < 32u ? tile_base + (uint64_t)src * 32 + top_bit(m[src].B) + 1 : cy.open_start;
        for (uint32_t mm = candDot; mm; mm &= mm - 1) {
          uint32_t bit = (uint32_t)__builtin_ctz(mm);
          uint64_t s = word_start_in_lane(m[lane].B, bit, p, lane_open);
          ((H[3][lane] >> bit) & 1u ? qd : qn).push_back(Cand{(uint32_t)s, (uint32_t)(p + bit - s)});
        }
        if (candHex) {
          uint32_t bit = (uint32_t)__builtin_ctz(candHex);
          uint64_t s = word_start_in_lane(m[lane].B, bit, p, lane_open);
          if (is_hash_len(p + bit - s)) qh.push_back(Cand{(uint32_t)s, (uint32_t)(p + bit - s)});
        }
        uint32_t candLong = want_long ? long_word_ends(T[lane], ~Bprev, E) : 0u;
        for (uint32_t mm = candLong; mm; mm &= mm - 1) {
          uint32_t bit = (uint32_t)__builtin_ctz(mm);
          uint64_t s = word_start_in_lane(m[lane].B, bit, p, lane_open);
          if (is_crypto_len(p + bit - s)) ql.push_back(Cand{(uint32_t)s, (uint32_t)(p + bit - s)});
        }
        for (uint32_t mm = candAt; mm; mm &= mm - 1) qa.push_back((uint32_t)(p + (uint32_t)__builtin_ctz(mm)));
        for (uint32_t mm = candC2; mm; mm &= mm - 1) qc.push_back((uint32_t)(p + (uint32_t)__builtin_ctz(mm) - 1));
      }
      if (hasB) {
        uint32_t ll = top_bit(hasB);
        cy.open_start = tile_base + (uint64_t)ll * 32 + top_bit(m[ll].B) + 1;
      }
      cy.prev = prev_bits_of(m[31]);
      cy.prevB = m[31].B;
    }
  }
}

static void emu_piece(emu_ctx* c, const uint8_t* buf, uint64_t lo, uint64_t n, uint64_t base, uint32_t flags, uint64_t nwarps, bool lookups) {
  std::vector<Cand> qd, qh, qn, ql;
  std::vector<uint32_t> qa, qc;
  uint64_t lines = 0;
  emu_tokenize(buf, lo, n, flags, nwarps, qd, qh, qa, qc, qn, ql, lines);
  const DbView& db = c->db;
  size_t s0 = c->str.size(), i0 = c->ip.size();
  // K2 token kernel: dotted queue = domains only; numeric queue = IPv4, else maybe a domain
  for (int numeric = 0; numeric < 2; numeric++)
    for (auto& cd : (numeric ? qn : qd)) {
      const uint8_t* wp = buf + cd.start;
      uint32_t addr;
      bool high = false;  // the kernel passes a per-window flag; the promise "false => pure ASCII" is what matters
      for (uint32_t k = 0; k < cd.len; k++) high |= wp[k] >= 0x80;
      if (numeric && (flags & MGPU_X_IPV4) && parse_ipv4_word(wp, cd.len, addr)) { c->ip.push_back(IpTok{cd.start, cd.len, MGPU_T_IPV4, {addr, 0, 0, 0}}); continue; }
      if ((flags & MGPU_X_DOMAINS) && domain_word_fast(db, db.psl_tld, wp, cd.len, high)) c->str.push_back(StrTok{cd.start, cd.len, MGPU_T_DOMAIN});
    }
  for (auto& cd : qh) {
    uint32_t ty = cd.len == 32 ? MGPU_T_MD5 : cd.len == 40 ? MGPU_T_SHA1 : cd.len == 64 ? MGPU_T_SHA256 : cd.len == 96 ? MGPU_T_SHA384 : MGPU_T_SHA512;
    c->str.push_back(StrTok{cd.start, cd.len, ty});
  }
  for (auto& cd : ql) {  // crypto_kernel
    uint32_t ty = crypto_word_type(buf + cd.start, cd.len, flags);
    if (ty != NONE32) c->str.push_back(StrTok{cd.start, cd.len, ty});
  }
  for (uint32_t at : qa) {
    size_t s, e;
    if (email_at(db, buf, (size_t)lo, (size_t)n, at, s, e)) c->str.push_back(StrTok{(uint32_t)s, (uint32_t)(e - s), MGPU_T_EMAIL});
  }
  for (uint32_t at : qc) {
    size_t s, e;
    IpTok t{0, 0, MGPU_T_IPV6, {0, 0, 0, 0}};
    if ((uint64_t)at + 2 <= n && ipv6_at_words(buf, (size_t)lo, (size_t)n, at, s, e, t.w)) {  // (the kernel's call)
      t.start = (uint32_t)s; t.len = (uint32_t)(e - s);
      c->ip.push_back(t);
    }
  }
  c->counters.lines += lines;
  c->counters.bytes += n - lo;
  for (size_t k = s0; k < c->str.size(); k++) { c->counters.by_type[c->str[k].type]++; c->counters.candidates++; }
  for (size_t k = i0; k < c->ip.size(); k++) { c->counters.by_type[c->ip[k].type]++; c->counters.candidates++; }
  if (lookups) {
    // K3
    for (size_t k = i0; k < c->ip.size() && db.has_ip; k++) {
      const IpTok& t = c->ip[k];
      uint32_t off = 0; uint8_t pl = 0; bool hit;
      if (t.type == MGPU_T_IPV4) hit = trie_lookup_v4(db, t.w[0], off, pl);
      else {
        uint16_t seg[8];
        for (int j = 0; j < 4; j++) { seg[2 * j] = (uint16_t)(t.w[j] >> 16); seg[2 * j + 1] = (uint16_t)t.w[j]; }
        hit = trie_lookup_v6(db, seg, off, pl);
      }
      if (!hit) continue;
      mgpu_match r{};
      r.offset = base + t.start; r.len = t.len; r.item_type = (uint8_t)t.type; r.kind = MGPU_KIND_IP; r.prefix_len = pl;
      r.data_offset = off;
      c->recs.push_back(r);
    }
    // K4 + K5 (generic) or the filters of K2 + the exact kernel (fast path)
    const bool fast = db.fast_ok && !c->force_generic && (db.has_literal || db.has_glob);
    for (size_t k = s0; k < c->str.size(); k++) {
      const StrTok& t = c->str[k];
      const uint8_t* text = buf + t.start;
      uint32_t f = fast ? string_filters(db, db.hot, text, t.len) : (uint32_t)(F_LIT | F_GLOB);
      if (fast) c->filter_pass[0] += (f & F_LIT) != 0, c->filter_pass[1] += (f & F_GLOB) != 0, c->filter_pass[2]++;
      if (fast) {  // stage-1 statistics: how many tokens owe stage 2, per class
        KeyWords kw; load_head_words(text, kw.h); load_tail_words(text, t.len, kw.t);
        const uint32_t g = string_gate(db, HotPtr{db.hot, db.gen_gram2}, kw, t.len);
        c->gate_pass[0] += (g & G_LIT) != 0; c->gate_pass[1] += (g & G_S) != 0; c->gate_pass[2] += (g & G_P) != 0; c->gate_pass[3] += g != 0;
      }
      if (!f) continue;
      uint32_t lit_pid = NONE32, lit_off = 0;
      if (!(f & F_LIT) || !db.has_literal || !lh_lookup(db, text, t.len, lit_pid)) lit_pid = NONE32;
      bool lit_ok = lit_pid != NONE32 && lh_data_offset(db, lit_pid, lit_off);
      std::vector<uint32_t> g;
      AcAccel acc;
      acc.root_tab = ac_root_table(db); acc.gram2 = c->use_anchored ? db.ac_gram2 : nullptr;
      if (db.has_glob && (f & F_GLOB)) {
        if (fast && db.ac_anchored && db.wild_count == 0 && db.ac_size >= 20)  // exact_kernel: the lanes share these positions
          for (uint32_t pos = 0; pos + 3 <= t.len; pos++) anchored_visit_at(db, text, t.len, pos, db.ac_gram2, [&](uint32_t pid) { g.push_back(pid); });
        else find_all_visit(db, text, t.len, acc, [&](uint32_t pid) { g.push_back(pid); });
      }
      if (!lit_ok && g.empty()) continue;
      std::sort(g.begin(), g.end());
      g.erase(std::unique(g.begin(), g.end()), g.end());
      mgpu_match r{};
      r.offset = base + t.start; r.len = t.len; r.item_type = (uint8_t)t.type; r.kind = MGPU_KIND_PATTERN;
      r.ids_index = (uint32_t)c->ids.size(); r.data_offset = MGPU_NO_DATA;
      if (lit_ok) c->ids.push_back(mgpu_id_pair{lit_pid, lit_off});
      for (uint32_t pid : g) { uint32_t off; c->ids.push_back(mgpu_id_pair{pid, glob_data_offset(db, pid, off) ? off : MGPU_NO_DATA}); }
      r.n_ids = (uint32_t)c->ids.size() - r.ids_index;
      c->recs.push_back(r);
    }
  }
  for (size_t k = s0; k < c->str.size(); k++) c->str[k].start += (uint32_t)base;
  for (size_t k = i0; k < c->ip.size(); k++) c->ip[k].start += (uint32_t)base;
}

extern "C" {

emu_ctx* emu_create(const uint8_t* psl, size_t psl_len) {
  auto* c = new emu_ctx();
  if (!build_psl(psl, psl_len, c->psl, c->err)) { fprintf(stderr, "emu: %s\n", c->err.c_str()); delete c; return nullptr; }
  return c;
}
void emu_destroy(emu_ctx* c) { delete c; }
const char* emu_error(emu_ctx* c) { return c->err.c_str(); }

int emu_db_upload(emu_ctx* c, const uint8_t* mxy, size_t len) {
  c->file.assign(mxy, mxy + len);
  c->file.resize(len + 64, 0);
  if (!prepare_db(c->file.data(), len, c->P, c->err)) return MGPU_E_FORMAT;
  c->db = c->P.view;
  const mxy::Layout& L = c->P.L;
  c->db.tree = c->file.data();
  if (!c->P.top16.empty()) { c->db.v4_top16 = c->P.top16.data(); c->db.v4_top16_depth = c->P.top16_depth.data(); }
  if (!c->P.v6_top16.empty()) { c->db.v6_top16 = c->P.v6_top16.data(); c->db.v6_top16_depth = c->P.v6_top16_depth.data(); }
  if (L.has_literal) { c->db.lh = c->file.data() + L.lit_off; c->db.lh_data_index = c->P.lh_index.data(); c->db.lh_bloom = c->P.lh_bloom.data(); }
  if (L.has_glob) {
    c->db.pg = c->file.data() + L.pg_off; c->db.aclh_index = c->P.aclh.data();
    c->db.ac_gram2 = c->P.gram2.data(); c->db.ac_gram3 = c->P.gram3.data();
    c->db.ac_pfx_keys = c->P.pfx_keys.empty() ? nullptr : c->P.pfx_keys.data();
    c->db.ac_pfx_vals = c->P.pfx_vals.empty() ? nullptr : c->P.pfx_vals.data();
    c->db.glob_data = reinterpret_cast<const uint32_t*>(c->file.data() + L.map_off);
  }
  c->db.psl_keys = c->psl.keys.data(); c->db.psl_vals = c->psl.vals.data(); c->db.psl_pool = c->psl.pool.data();
  c->db.psl_mask = c->psl.mask; c->db.psl_max_len = c->psl.max_len; c->db.psl_tld = c->psl.tld.data();
  if (c->db.fast_ok) { c->db.hot = c->P.hot.data(); c->db.cold = c->P.cold.data(); }
  if (c->db.fast_ok && c->db.has_generic) { c->db.gen_gram2 = c->P.gen2.data(); c->db.gen_gram3 = c->P.gen3.data(); }
  c->loaded = true;
  return MGPU_OK;
}

void emu_set_anchored(emu_ctx* c, int on) { c->use_anchored = on != 0; }
void emu_set_generic(emu_ctx* c, int on) { c->force_generic = on != 0; }
int emu_is_fast(emu_ctx* c) { return (int)c->db.fast_ok; }
void emu_filter_stats(emu_ctx* c, uint64_t out[3]) { for (int k = 0; k < 3; k++) out[k] = c->filter_pass[k]; }
void emu_gate_stats(emu_ctx* c, uint64_t out[4]) { for (int k = 0; k < 4; k++) out[k] = c->gate_pass[k]; }
int emu_is_anchored_exact(emu_ctx* c) { return (int)c->db.ac_anchored; }

uint32_t emu_default_flags(emu_ctx* c) {
  uint32_t f = 0;
  if (c->db.has_ip) f |= MGPU_X_IPV4 | MGPU_X_IPV6;
  if (c->db.has_literal || c->db.has_glob) f |= MGPU_X_DOMAINS | MGPU_X_EMAILS | MGPU_X_HASHES;
  return f;
}

// chunk_bytes: the scan is cut into newline-aligned pieces of at most this size, the way mgpu_scan does;
// nwarps: how many tile ranges each piece is divided into; misalign: 0..15 alignment padding in front of the data.
int emu_scan(emu_ctx* c, const uint8_t* data, size_t len, uint64_t base, uint32_t flags, size_t chunk_bytes, uint64_t nwarps, uint32_t misalign,
             int lookups) {
  c->recs.clear(); c->ids.clear(); c->str.clear(); c->ip.clear();
  memset(&c->counters, 0, sizeof c->counters);
  c->filter_pass[0] = c->filter_pass[1] = c->filter_pass[2] = 0;
  c->gate_pass[0] = c->gate_pass[1] = c->gate_pass[2] = c->gate_pass[3] = 0;
  if (!c->loaded && lookups) return MGPU_E_NODB;
  if (chunk_bytes == 0) chunk_bytes = len ? len : 1;
  size_t pos = 0;
  while (pos < len) {
    size_t want = std::min(len - pos, chunk_bytes);
    if (pos + want < len) {
      const uint8_t* q = (const uint8_t*)memrchr(data + pos, '\n', want);
      if (!q) return MGPU_E_OVERFLOW;
      want = (size_t)(q - (data + pos)) + 1;
    }
    // private padded copy: [misalign garbage][piece][garbage up to the tile boundary]
    std::vector<uint8_t> buf(misalign + want + 2 * TILE_BYTES, 'x');
    memcpy(buf.data() + misalign, data + pos, want);
    emu_piece(c, buf.data(), misalign, misalign + want, base + pos - misalign, flags, nwarps ? nwarps : 1, lookups != 0);
    pos += want;
  }
  c->counters.matches = c->recs.size();
  std::sort(c->recs.begin(), c->recs.end(), [](const mgpu_match& x, const mgpu_match& y) {
    if (x.offset != y.offset) return x.offset < y.offset;
    if (x.item_type != y.item_type) return x.item_type < y.item_type;
    return x.len < y.len;
  });
  return MGPU_OK;
}

int emu_results(emu_ctx* c, const mgpu_match** recs, size_t* n_recs, const mgpu_id_pair** ids, size_t* n_ids) {
  *recs = c->recs.data(); *n_recs = c->recs.size(); *ids = c->ids.data(); *n_ids = c->ids.size();
  return 0;
}
int emu_counters_get(emu_ctx* c, mgpu_counters* out) { *out = c->counters; return 0; }

// the kernels' tree record reader on bare tree bytes
uint32_t emu_tree_record(const uint8_t* tree, uint32_t node_count, uint32_t record_bits, uint32_t node, uint32_t side) {
  DbView db;
  memset(&db, 0, sizeof db);
  db.tree = tree; db.node_count = node_count; db.record_bits = record_bits;
  return tree_record(db, node, side);
}
// the kernels' mask-arithmetic IPv6 parser on a bare run (s must be readable 8 bytes past n); 1 = parsed
int emu_parse_ipv6_masks(const uint8_t* s, uint32_t n, uint32_t w[4]) { return parse_ipv6_run_masks(s, n, w) ? 1 : 0; }

// scan_kernel's byte classification: category planes of byte b (bytes >= 0x80 take the row of 'g') -> class bits
uint32_t emu_class_bits_via_planes(uint32_t b) { return class_bits_from_planes(category_planes((uint8_t)(b >= 0x80 ? 'g' : b))); }
uint32_t emu_class_bits(uint32_t b) { return class_bits((uint8_t)b); }

// both IPv4 parsers on one word of hex digits and dots (s readable 20 bytes past n): bit 0 = general parser accepted,
// bit 1 = hex-dot parser accepted, *out = their addresses
uint32_t emu_parse_ipv4_both(const uint8_t* s, uint32_t n, uint32_t out[2]) {
  uint32_t h[4];
  load_head_words(s, h);
  out[0] = out[1] = 0;
  return (parse_ipv4_words(h, n, out[0]) ? 1u : 0u) | (parse_ipv4_hexdot(h, n, out[1]) ? 2u : 0u);
}

// the engine's result sort (host_sort.h) on n records, with `threads` pool workers (0: no pool); returns microseconds
double emu_sort_records(mgpu_match* r, size_t n, uint64_t lo, uint64_t hi, unsigned threads, int repeat) {
  std::vector<mgpu_match> tmp, copy(r, r + n);
  std::unique_ptr<mgpu::WorkerPool> pool(threads ? new mgpu::WorkerPool(threads) : nullptr);
  double best = 1e30;
  for (int k = 0; k < repeat; k++) {
    std::copy(copy.begin(), copy.end(), r);
    auto t0 = std::chrono::steady_clock::now();
    mgpu::sort_records(r, n, lo, hi, tmp, pool.get());
    best = std::min(best, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
  }
  return best;
}

// the engine's order check (what decides whether the host still sorts after the device has)
int emu_records_sorted(const mgpu_match* r, size_t n, unsigned threads) {
  std::unique_ptr<mgpu::WorkerPool> pool(threads ? new mgpu::WorkerPool(threads) : nullptr);
  return mgpu::records_sorted(r, n, pool.get()) ? 1 : 0;
}

// the engine's id re-pack on n sorted records; out must hold every pair; returns the number of pairs
size_t emu_repack_ids(mgpu_match* r, size_t n, const mgpu_id_pair* ids, mgpu_id_pair* out, unsigned threads) {
  std::vector<mgpu_id_pair> packed;
  std::unique_ptr<mgpu::WorkerPool> pool(threads ? new mgpu::WorkerPool(threads) : nullptr);
  const size_t np = mgpu::repack_ids(r, n, ids, packed, pool.get());
  std::copy(packed.begin(), packed.begin() + np, out);
  return np;
}

int64_t emu_tokens(emu_ctx* c, uint64_t* out, size_t cap) {
  std::vector<std::array<uint64_t, 3>> items;
  for (auto& t : c->str) items.push_back({t.type, t.start, (uint64_t)t.start + t.len});
  for (auto& t : c->ip) items.push_back({t.type, t.start, (uint64_t)t.start + t.len});
  std::sort(items.begin(), items.end(), [](auto& x, auto& y) { return x[1] != y[1] ? x[1] < y[1] : x[0] < y[0]; });
  for (size_t k = 0; k < items.size() && k < cap; k++) { out[3 * k] = items[k][0]; out[3 * k + 1] = items[k][1]; out[3 * k + 2] = items[k][2]; }
  return (int64_t)items.size();
}

}  // extern "C"
