// device_fns.cuh — per-token device functions of the scan path: token validation (the second half of the
// extractor), IP-trie walk, literal-hash probe, Aho-Corasick walk + ACLH map + glob verification.
//
// Everything here is a plain function of byte pointers into the UNCHANGED .mxy sections as uploaded to HBM
// (plus a few derived, result-neutral indexes built at upload, see DbView).  The functions are MGPU_HD so that
// tests/host_emulation can compile the very same code with g++ and check it against the oracle without a GPU;
// the product only ever runs them inside CUDA kernels (engine.cu).
//
// Reference behaviour restated (crates/… paths relative to the reference root):
//   try_parse_ipv4                matchy-extractor/src/lib.rs:813-869
//   extract_domains_chunk/is_valid_domain/find_valid_tld_suffix_bytes   :537-689, 1671-1692
//   extract_email_at              :891-950
//   extract_ipv6_chunk (+ core::net::parser Ipv6Addr::from_str)  :1044-1116, 1425-1456
//   SearchTree::lookup_v4/v6      matchy-format/src/mmdb/tree.rs:46-277
//   LiteralHash::lookup           matchy-literal-hash/src/lib.rs:467-575
//   Paraglob::find_all & co       matchy-paraglob/src/paraglob_offset.rs:1028-1639
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define MGPU_HD __host__ __device__ __forceinline__
#define MGPU_HDN __host__ __device__
#else
#define MGPU_HD inline
#define MGPU_HDN inline
#endif

namespace mgpu {

static const uint32_t NONE32 = 0xFFFFFFFFu;

// View of one uploaded database.  Raw sections are byte-identical to the file; "derived" members are small
// indexes computed once at upload that return exactly what the reference's slower lookup would.
struct DbView {
  // --- IP search tree (file offset 0) ---
  const uint8_t* tree;
  uint32_t node_count, record_bits, ip_version, has_ip;
  uint32_t v4_start_node;  // derived: result of find_ipv4_start_node (tree.rs:258-277), identical for every query
  uint32_t ip_empty;       // derived: no record of the tree points into the data section — every lookup is a miss (string-only databases)
  // derived: the state of the IPv4 walk after its first v4_top_bits (16 or 20) address bits, for every such prefix:
  // bits 0..27 value, bits 28..31 kind — 0: continue at node `value`; 1: not found; 2: data record `value` found at depth
  // v4_top16_depth[prefix] (1..v4_top_bits).  Exact: the entry is what the bit-by-bit walk of tree.rs:46-90 reaches.
  const uint32_t* v4_top16;
  const uint8_t* v4_top16_depth;
  uint32_t v4_top_bits;
  // the same for IPv6 queries in an IPv6 tree: state after the first 16 address bits, from the root (tree.rs:92-125)
  const uint32_t* v6_top16;
  const uint8_t* v6_top16_depth;
  // --- literal hash section ("LHSH") ---
  const uint8_t* lh;
  uint64_t lh_len;
  uint32_t has_literal, lh_num_shards, lh_strings_offset, lh_table_start;
  const uint32_t* lh_data_index;  // derived: pattern_id -> data_offset of the FIRST mapping entry with that id (== linear scan :560-572)
  uint32_t lh_data_index_n;
  const uint64_t* lh_bloom;       // derived: blocked Bloom filter over the stored 64-bit hashes (no false negatives): a miss
  uint32_t lh_bloom_mask;         //          costs one load instead of a walk over a ~38-slot probe cluster
  // --- paraglob buffer ("PARAGLOB") ---
  const uint8_t* pg;
  uint32_t pg_len, has_glob;
  uint32_t pg_align;              // (file offset of the PARAGLOB buffer) & 3: the reference reads glob structs through
                                  // zerocopy::Ref, which fails on misaligned addresses (page-aligned mmap + file offset)
  uint32_t ac_start, ac_size, patterns_offset, wild_off, wild_count, glob_segments_offset;
  const uint32_t* aclh_index;     // derived: literal_id -> {abs offset of its pattern list in pg, count} (2 words each)
  uint32_t aclh_n;
  const uint32_t* ac_gram2;       // derived: bitmap over the first 2 bytes of every AC literal (2^16 bits)
  const uint32_t* ac_gram3;       // derived: bitmap over the first 3 bytes of every AC literal (2^24 bits)
  uint32_t ac_anchored;           // 1: every literal has >= 3 bytes and no node list is saturated -> anchored walk is exact
  // derived: exact open-addressing map  (m, first m bytes of a trie path) -> node offset, holding every depth-8 node and
  // every node of depth 3..7 that has outputs.  An anchored walk jumps straight to depth 8 instead of taking 8 dependent steps.
  const uint64_t* ac_pfx_keys;    // the m bytes, little-endian, zero-padded
  const uint32_t* ac_pfx_vals;    // 2 words per slot: node offset, m (0 = empty slot)
  uint32_t ac_pfx_mask;
  uint32_t ac_short_lens;         // bit m set: some node of depth m (3 <= m < 8) has outputs
  const uint32_t* glob_data;      // [data_offset × count] array that follows the paraglob buffer (database.rs:212-228)
  uint32_t glob_data_n;
  uint32_t match_mode;            // 0 case-sensitive, 1 case-insensitive (ASCII folding)
  // --- Public Suffix List as an open-addressing set (derived from the reference's PSL text) ---
  const uint64_t* psl_keys;       // 0 = empty
  const uint32_t* psl_vals;       // pool offset << 8 | length
  const uint8_t* psl_pool;
  uint32_t psl_mask, psl_max_len;
  const uint64_t* psl_tld;        // derived: last-label table (TLD_SLOTS entries), see tld_class()
  // --- fast string path (derived, result-neutral: necessary conditions for a match, see string_filters()) ---
  uint32_t fast_ok;               // 1: case-sensitive, no pure-wildcard patterns (unanchored globs are handled through gen_gram2/3)
  uint32_t has_generic;           // 1: some reachable glob is unanchored (or literal-typed): its literals' first bytes are in gen_gram2/3
  const uint32_t* gen_gram2;      // bitmaps (2^16 / 2^24 bits) over the first 2 / 3 bytes of the AC literals that lead to such a glob;
  const uint32_t* gen_gram3;
  uint32_t glob_s_gate, glob_p_gate;  // shortest key length >= 4 of the suffix / prefix class (0: none): the hot filter's gate keys
  uint32_t gate_inline;               // G_* bits of the classes whose gate cannot reject (class not in the hot filter, unanchored globs)
  uint32_t glob_s_lens, glob_p_lens;  // bit K (K in 1,2,3,4,8,12,16): some suffix- / prefix-anchored glob has a key of K bytes
  uint32_t hot_tags;              // bit t: keys of tag class t are in the hot filter (else that class skips the hot test)
  const uint32_t* hot;            // HOT_WORDS-word blocked Bloom filter; the token kernel keeps a copy in shared memory
  const uint64_t* cold;           // blocked Bloom filter in L2 (>= 16 bits per key) over the same keys + full literal keys
  uint32_t cold_mask;
};

// ---- fast string path: filter geometry and key hashing (shared by db_prepare.h and the kernels) ----
static const uint32_t TLD_SLOTS = 4096;      // 32 KiB of shared memory
static const uint32_t HOT_WORDS = 32768;     // 128 KiB of shared memory
static const uint64_t TLD_SLOW = 1ULL << 63; // entry flag: the label only ends multi-label PSL entries -> general PSL walk
enum { TAG_GLOB_S = 0, TAG_GLOB_P = 1, TAG_LIT_TAIL = 2, TAG_LIT_HEAD = 3, TAG_LIT_FULL = 4 };
enum { F_LIT = 0x100u, F_GLOB = 0x200u };    // StrTok.type flag bits: which exact lookups the token still needs

MGPU_HD uint32_t tld_slot(uint64_t key) {
  uint32_t h = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA77u);
  return (h ^ (h >> 15)) & (TLD_SLOTS - 1);
}
// Keys of the string filters.  A key is (tag class, K, K bytes of text); the bytes are fed to the hash as little-endian
// 32-bit WORDS: K in {4, 8, 12, 16} is K/4 whole words, K in {1, 2, 3} one word masked to K bytes.  Head keys feed the words
// front to back, tail keys back to front, and the hash is incremental (key_mix per word, key_fin per key), so a token's
// 16-byte key extends its 12-byte key by one step.
//   glob anchor literal of m bytes -> K = glob_key_len(m) of its first (prefix-anchored) / last (suffix-anchored) bytes
//   stored literal of n bytes      -> K = lit_key_len(n) of its last bytes (TAG_LIT_TAIL) and of its first bytes (TAG_LIT_HEAD);
//                                     (n, first K, last K bytes) for the cold filter (TAG_LIT_FULL)
MGPU_HD uint32_t glob_key_len(uint32_t m) { return m < 4 ? m : (m >= 16 ? 16u : (m & ~3u)); }
MGPU_HD uint32_t lit_key_len(uint32_t n) { return n >= 8 ? 8u : (n >= 4 ? 4u : n); }
// The running state depends only on the DIRECTION the words are fed in (tail keys: TAG_GLOB_S, TAG_LIT_TAIL; head keys:
// TAG_GLOB_P, TAG_LIT_HEAD), the tag class enters in key_fin: a token's suffix-gate key and its literal tail key share their
// key_mix steps, and so do the prefix-gate key and the literal head key.
MGPU_HD uint32_t key_seed(uint32_t tag) {
  return tag == TAG_LIT_FULL ? 0x811C9DC5u ^ (4u * 0x632BE5ABu) : ((tag == TAG_GLOB_S || tag == TAG_LIT_TAIL) ? 0x811C9DC5u : 0x811C9DC5u ^ 0x632BE5ABu);
}
MGPU_HD uint32_t key_mix(uint32_t s, uint32_t w) { s = (s ^ w) * 0x9E3779B1u; return s ^ (s >> 15); }
// one more multiply-xorshift round per key; k separates keys of different lengths that feed identical words
MGPU_HD uint32_t key_fin(uint32_t s, uint32_t k, uint32_t tag) { s = (s ^ (k * 0x7F4A7C15u + tag * 0x2C1B3C6Du)) * 0x85EBCA77u; return s ^ (s >> 15); }
MGPU_HD uint32_t low_bytes32(uint32_t w, uint32_t k) { return k >= 4 ? w : (w & ((1u << (8 * k)) - 1u)); }
// hash of a whole key given its words in feeding order (host side: database preparation)
MGPU_HD uint32_t key_hash_words(const uint32_t* words, uint32_t k, uint32_t tag) {
  uint32_t s = key_seed(tag);
  for (uint32_t i = 0; i < (k + 3) / 4; i++) s = key_mix(s, words[i]);
  return key_fin(s, k, tag);
}
// hot filter: blocked Bloom, 2 bits in one 32-bit word.  cold filter: blocked Bloom, 3 bits in one 64-bit word; its word
// index and bit positions come from a second mix of the same key hash, so one hash per key serves both filters.
// Where the hot filter's words come from: a plain pointer (host emulation, database preparation) or the token kernel's
// shared-memory copy addressed with ld.shared (a generic pointer would cost the generic-address path on every test).
struct HotPtr {
  const uint32_t* p; const uint32_t* g2;  // hot filter words; gen_gram2 words (or nullptr)
  MGPU_HD uint32_t word(uint32_t i) const { return p[i]; }
  MGPU_HD uint32_t gen2(uint32_t i) const { return g2[i]; }
};
struct BytesPtr { const uint8_t* p; MGPU_HD uint32_t at(uint32_t i) const { return p[i]; } };  // a token's bytes
#ifdef __CUDACC__
struct HotShared {
  uint32_t base;        // shared-space byte address of word 0 of the hot filter
  const uint32_t* g2;   // gen_gram2 in global memory (8 KiB, L1-resident)
  __device__ __forceinline__ uint32_t word(uint32_t i) const { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + i * 4)); return v; }
  __device__ __forceinline__ uint32_t gen2(uint32_t i) const { return __ldg(g2 + i); }
};
struct HotGlobal {    // the hot filter in global memory, read through L1 (scan_kernel's wide block keeps no copy)
  const uint32_t* p; const uint32_t* g2;
  __device__ __forceinline__ uint32_t word(uint32_t i) const { return __ldg(p + i); }
  __device__ __forceinline__ uint32_t gen2(uint32_t i) const { return __ldg(g2 + i); }
};
struct BytesShared {  // a token in a shared-memory window
  uint32_t saddr;
  __device__ __forceinline__ uint32_t at(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr + i) : "memory"); return v; }
};
#endif
// two independent bits per key inside one 32-bit word: bits (h & 31) and ((h >> 5) & 31)
MGPU_HD uint32_t hot_mask(uint32_t h) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(1u, 1u, h) | __funnelshift_l(1u, 1u, h >> 5);  // (a rotate takes the shift modulo 32)
#else
  return (1u << (h & 31u)) | (1u << ((h >> 5) & 31u));
#endif
}
template <typename H>
MGPU_HD bool hot_test(const H& hot, uint32_t h) {
  const uint32_t m = hot_mask(h);
  return (hot.word(h >> 17) & m) == m;
}
MGPU_HD void hot_set(uint32_t* hot, uint32_t h) { hot[h >> 17] |= hot_mask(h); }
MGPU_HD uint32_t cold_word(uint32_t h, uint32_t mask) { return ((h * 0x2545F491u) ^ (h >> 11)) & mask; }
MGPU_HD uint64_t cold_bits(uint32_t h) {
  uint32_t g = (h ^ 0x5BD1E995u) * 0x846CA68Bu;
  g ^= g >> 15;
  return (1ULL << (g & 63)) | (1ULL << ((g >> 6) & 63)) | (1ULL << ((g >> 12) & 63));
}
MGPU_HD bool cold_test(const uint64_t* cold, uint32_t mask, uint32_t h) {
  uint64_t need = cold_bits(h);
  return (cold[cold_word(h, mask)] & need) == need;
}
MGPU_HD void cold_set(uint64_t* cold, uint32_t mask, uint32_t h) { cold[cold_word(h, mask)] |= cold_bits(h); }

MGPU_HD uint32_t ld32(const uint8_t* p) {  // little-endian; 4-byte aligned on the device (layout chosen at upload)
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const uint32_t*>(p);
#else
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);  // host emulation: any alignment
#endif
}
MGPU_HD uint32_t ld32u(const uint8_t* p) {  // unaligned
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
MGPU_HD uint64_t ld64u(const uint8_t* p) { return (uint64_t)ld32u(p) | ((uint64_t)ld32u(p + 4) << 32); }
MGPU_HD uint8_t lc(uint8_t c, bool fold) { return (fold && c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

// ---- character classes (matchy-extractor/src/lib.rs:1568-1717) ----
MGPU_HD bool is_digit(uint8_t b) { return (uint8_t)(b - '0') < 10; }
MGPU_HD bool is_alpha(uint8_t b) { return (uint8_t)((b | 32) - 'a') < 26; }
MGPU_HD bool is_hex(uint8_t b) { return is_digit(b) || (uint8_t)((b | 32) - 'a') < 6; }
MGPU_HD bool is_boundary(uint8_t b) {
  switch (b) {
    case ' ': case '\t': case '\n': case '\r': case '/': case ',': case ';': case ':': case '(': case ')':
    case '[': case ']': case '{': case '}': case '<': case '>': case '"': case '\'': case '@': case '=':
      return true;
    default: return false;
  }
}
MGPU_HD bool is_domain_ascii(uint8_t b) { return is_digit(b) || is_alpha(b) || b == '-' || b == '.'; }
MGPU_HD bool is_domain_fast(uint8_t b) { return is_domain_ascii(b) || b >= 0x80; }
MGPU_HD bool is_local_char(uint8_t b) { return is_digit(b) || is_alpha(b) || b == '.' || b == '-' || b == '_' || b == '+'; }

// ---- XXH64, seed 0 (public specification) ----
#define MGPU_P1 11400714785074694791ULL
#define MGPU_P2 14029467366897019727ULL
#define MGPU_P3 1609587929392839161ULL
#define MGPU_P4 9650029242287828579ULL
#define MGPU_P5 2870177450012600261ULL
MGPU_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
MGPU_HD uint64_t xx_round(uint64_t acc, uint64_t in) { acc += in * MGPU_P2; acc = rotl64(acc, 31); return acc * MGPU_P1; }
MGPU_HD uint64_t xx_merge(uint64_t acc, uint64_t v) { v = xx_round(0, v); acc ^= v; return acc * MGPU_P1 + MGPU_P4; }
// Unaligned little-endian loads built from aligned 32-bit loads + funnel shifts (3 loads instead of 8 byte loads).
// They may touch up to 4 bytes past the value: every buffer the scan reads from has at least 16 bytes of slack.
MGPU_HD uint64_t ldu64_fast(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  uintptr_t a = (uintptr_t)p;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
  return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
#else
  return ld64u(p);
#endif
}
MGPU_HD uint32_t ldu32_fast(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  uintptr_t a = (uintptr_t)p;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  return __funnelshift_r(q[0], q[1], (uint32_t)(a & 3) * 8);
#else
  return ld32u(p);
#endif
}
MGPU_HD uint64_t rd64f(const uint8_t* p, bool fold) {
  if (!fold) return ldu64_fast(p);
  uint64_t v = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) v |= (uint64_t)lc(p[k], fold) << (8 * k);
  return v;
}
MGPU_HDN uint64_t xxh64_fold(const uint8_t* p, size_t len, bool fold) {
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = MGPU_P1 + MGPU_P2, v2 = MGPU_P2, v3 = 0, v4 = 0ULL - MGPU_P1;
    const uint8_t* lim = end - 32;
    do {
      v1 = xx_round(v1, rd64f(p, fold)); v2 = xx_round(v2, rd64f(p + 8, fold));
      v3 = xx_round(v3, rd64f(p + 16, fold)); v4 = xx_round(v4, rd64f(p + 24, fold));
      p += 32;
    } while (p <= lim);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xx_merge(h, v1); h = xx_merge(h, v2); h = xx_merge(h, v3); h = xx_merge(h, v4);
  } else {
    h = MGPU_P5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) { h ^= xx_round(0, rd64f(p, fold)); h = rotl64(h, 27) * MGPU_P1 + MGPU_P4; p += 8; }
  if (p + 4 <= end) {
    uint32_t w = fold ? ((uint32_t)lc(p[0], fold) | ((uint32_t)lc(p[1], fold) << 8) | ((uint32_t)lc(p[2], fold) << 16) | ((uint32_t)lc(p[3], fold) << 24))
                      : ldu32_fast(p);
    h ^= (uint64_t)w * MGPU_P1; h = rotl64(h, 23) * MGPU_P2 + MGPU_P3; p += 4;
  }
  while (p < end) { h ^= (uint64_t)lc(*p, fold) * MGPU_P5; h = rotl64(h, 11) * MGPU_P1; p++; }
  h ^= h >> 33; h *= MGPU_P2; h ^= h >> 29; h *= MGPU_P3; h ^= h >> 32;
  return h;
}

// ---- UTF-8 (Rust str::from_utf8) ----
MGPU_HDN bool valid_utf8(const uint8_t* s, size_t n) {
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
MGPU_HD uint32_t utf8_len(uint8_t c) { return c < 0x80 ? 1u : c < 0xE0 ? 2u : c < 0xF0 ? 3u : 4u; }
MGPU_HD uint32_t utf8_decode(const uint8_t* t, uint32_t n, uint32_t pos, uint32_t& cp) {  // t is valid UTF-8
  uint8_t c = t[pos];
  if (c < 0x80) { cp = c; return 1; }
  if (c < 0xE0 && pos + 1 < n) { cp = ((c & 0x1Fu) << 6) | (t[pos + 1] & 0x3Fu); return 2; }
  if (c < 0xF0 && pos + 2 < n) { cp = ((c & 0x0Fu) << 12) | ((t[pos + 1] & 0x3Fu) << 6) | (t[pos + 2] & 0x3Fu); return 3; }
  if (pos + 3 < n) { cp = ((c & 0x07u) << 18) | ((t[pos + 1] & 0x3Fu) << 12) | ((t[pos + 2] & 0x3Fu) << 6) | (t[pos + 3] & 0x3Fu); return 4; }
  cp = c; return 1;
}

// =================================================================================================
// Public Suffix List membership.  Key = 64-bit FNV-1a over the suffix bytes taken RIGHT TO LEFT, so one
// backward pass over a token yields the key of every dot-suffix.  Hits are confirmed byte-for-byte.
// =================================================================================================
#define MGPU_FNV_BASIS 0xcbf29ce484222325ULL
#define MGPU_FNV_PRIME 0x100000001b3ULL
MGPU_HD uint64_t psl_step(uint64_t h, uint8_t b) { return (h ^ b) * MGPU_FNV_PRIME; }
// FNV-1a mixes short strings poorly in any fixed bit window; finish with a multiply-xorshift before taking slot bits
MGPU_HD uint32_t psl_slot(uint64_t key, uint32_t mask) {
  key ^= key >> 29; key *= 0xBF58476D1CE4E5B9ULL; key ^= key >> 32;
  return (uint32_t)key & mask;
}
MGPU_HDN bool psl_contains(const DbView& db, uint64_t key, const uint8_t* suffix, uint32_t len) {
  if (key == 0) key = 1;
  uint32_t slot = psl_slot(key, db.psl_mask);
  for (;;) {
    uint64_t k = db.psl_keys[slot];
    if (k == 0) return false;
    if (k == key) {
      uint32_t v = db.psl_vals[slot];
      if ((v & 0xFF) == len) {
        const uint8_t* e = db.psl_pool + (v >> 8);
        bool eq = true;
        uint32_t i = 0;
        for (; i + 8 <= len && eq; i += 8) eq = ldu64_fast(e + i) == ldu64_fast(suffix + i);
        if (eq && i < len) {  // 1..7 bytes left (both buffers have >= 16 bytes of slack)
          uint64_t m = (1ULL << (8 * (len - i))) - 1;
          eq = ((ldu64_fast(e + i) ^ ldu64_fast(suffix + i)) & m) == 0;
        }
        if (eq) return true;
      }
    }
    slot = (slot + 1) & db.psl_mask;
  }
}
// find_valid_tld_suffix_bytes(...).is_some(): does ANY dot-suffix of d[0..n) belong to the PSL?  (lib.rs:1671-1692)
MGPU_HDN bool psl_any_suffix(const DbView& db, const uint8_t* d, uint32_t n) {
  uint64_t h = MGPU_FNV_BASIS;
  for (uint32_t k = n; k-- > 0;) {
    uint8_t b = d[k];
    if (b == '.') {
      uint32_t sl = n - k - 1;
      if (sl > db.psl_max_len) return false;  // longer suffixes cannot be entries either
      if (sl > 0 && psl_contains(db, h, d + k + 1, sl)) return true;
    }
    h = psl_step(h, b);
  }
  return false;
}

// =================================================================================================
// Token validation (second half of the extractor; candidates come from the tokenizer kernel)
// =================================================================================================
// First 16 / last 16 bytes of a token as little-endian words.  h[i] = bytes [4i, 4i+4) (bytes past the token are whatever
// follows it in the buffer: every buffer has 16 bytes of slack).  t[i] = bytes [n-16+4i, n-12+4i), loaded only when they lie
// inside the token (n >= 16-4i), else 0 — keys never use bytes outside the token.
struct KeyWords { uint32_t h[4], t[4]; };
MGPU_HD void load_head_words(const uint8_t* w, uint32_t h[4]) {
#ifdef __CUDA_ARCH__
  const uintptr_t a = (uintptr_t)w;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
  h[0] = __funnelshift_r(w0, w1, sh); h[1] = __funnelshift_r(w1, w2, sh); h[2] = __funnelshift_r(w2, w3, sh); h[3] = __funnelshift_r(w3, w4, sh);
#else
  for (int i = 0; i < 4; i++) h[i] = ld32u(w + 4 * i);
#endif
}
MGPU_HD void load_tail_words(const uint8_t* w, uint32_t n, uint32_t t[4]) {
  t[3] = n >= 4 ? ldu32_fast(w + n - 4) : 0u;
  t[2] = n >= 8 ? ldu32_fast(w + n - 8) : 0u;
  t[1] = n >= 12 ? ldu32_fast(w + n - 12) : 0u;
  t[0] = n >= 16 ? ldu32_fast(w + n - 16) : 0u;
}

// The tail words of a token of n <= 16 bytes from its head words (same contract as load_tail_words: t[i] = bytes
// [n-16+4i, n-12+4i) when they lie inside the token, else 0).
MGPU_HD uint32_t head_word_at(const uint32_t h[4], uint32_t s) {  // bytes [s, s+4) of the 16 head bytes, s <= 12
  const uint32_t i = s >> 2, sh = (s & 3) * 8;
  const uint32_t a = i == 0 ? h[0] : (i == 1 ? h[1] : (i == 2 ? h[2] : h[3]));
  const uint32_t b = i == 0 ? h[1] : (i == 1 ? h[2] : (i == 2 ? h[3] : 0u));
#ifdef __CUDA_ARCH__
  return __funnelshift_r(a, b, sh);
#else
  return sh ? ((a >> sh) | (b << (32 - sh))) : a;
#endif
}
MGPU_HD void tail_words_from_head(const uint32_t h[4], uint32_t n, uint32_t t[4]) {
  t[3] = n >= 4 ? head_word_at(h, n - 4) : 0u;
  t[2] = n >= 8 ? head_word_at(h, n - 8) : 0u;
  t[1] = n >= 12 ? head_word_at(h, n - 12) : 0u;
  t[0] = n >= 16 ? h[0] : 0u;
}

#ifdef __CUDACC__
// the same two loaders for a token that lies in a shared-memory window (saddr = shared-space byte address of its first byte)
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_unaligned32(uint32_t saddr) {
  const uint32_t q = saddr & ~3u;
  return __funnelshift_r(lds_u32(q), lds_u32(q + 4), (saddr & 3u) * 8);
}
__device__ __forceinline__ void load_head_words_shared(uint32_t saddr, uint32_t h[4]) {
  const uint32_t q = saddr & ~3u, sh = (saddr & 3u) * 8;
  const uint32_t w0 = lds_u32(q), w1 = lds_u32(q + 4), w2 = lds_u32(q + 8), w3 = lds_u32(q + 12), w4 = lds_u32(q + 16);
  h[0] = __funnelshift_r(w0, w1, sh); h[1] = __funnelshift_r(w1, w2, sh); h[2] = __funnelshift_r(w2, w3, sh); h[3] = __funnelshift_r(w3, w4, sh);
}
// (e = shared-space address just past the token's last byte)
__device__ __forceinline__ void load_tail_words_shared_end(uint32_t e, uint32_t n, uint32_t t[4]) {
  // the four words end at e; all share one alignment: five aligned words, the lower ones only when inside the token
  const uint32_t q = e & ~3u, sh = (e & 3u) * 8;
  const uint32_t w4 = sh ? lds_u32(q) : 0u;  // (the word that holds the bytes just below an unaligned end)
  const uint32_t w3 = n >= 1 ? lds_u32(q - 4) : 0u, w2 = n >= 5 ? lds_u32(q - 8) : 0u, w1 = n >= 9 ? lds_u32(q - 12) : 0u, w0 = n >= 13 ? lds_u32(q - 16) : 0u;
  t[3] = n >= 4 ? __funnelshift_r(w3, w4, sh) : 0u;
  t[2] = n >= 8 ? __funnelshift_r(w2, w3, sh) : 0u;
  t[1] = n >= 12 ? __funnelshift_r(w1, w2, sh) : 0u;
  t[0] = n >= 16 ? __funnelshift_r(w0, w1, sh) : 0u;
}
__device__ __forceinline__ void load_tail_words_shared(uint32_t saddr, uint32_t n, uint32_t t[4]) { load_tail_words_shared_end(saddr + n, n, t); }
#endif

// try_parse_ipv4 (lib.rs:813-869) on a whole boundary-delimited word of n bytes held in h[0..4) (only the first n bytes count):
// all of it must be consumed — four groups of 1..3 digits, value <= 255, no leading zero in a multi-digit group, single dots
// between.  Straight-line SWAR: per-byte digit / dot flags -> 16-bit masks -> dot positions -> group values.
MGPU_HD uint32_t swar_flags4(uint32_t m) { return (((m >> 7) * 0x00204081u) >> 21) & 0xFu; }  // 0x80-per-byte flags of 4 bytes -> 4 bits
MGPU_HD uint32_t swar_is_digit(uint32_t v) {  // 0x80 in every byte that is '0'..'9'
  uint32_t x = v ^ 0x30303030u;
  return ~(((x & 0x7F7F7F7Fu) + 0x76767676u) | x) & 0x80808080u;
}
MGPU_HD uint32_t swar_is_dot(uint32_t v) {  // 0x80 in every byte that is '.'
  uint32_t d = v ^ 0x2E2E2E2Eu;
  return ~(((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
}
MGPU_HD uint32_t word_at(const uint32_t h[4], uint32_t s) {  // bytes [s, s+4) of the 16 (zero past the end), s < 16
  const uint32_t i = s >> 2, sh = (s & 3) * 8;
  const uint32_t a = i == 0 ? h[0] : (i == 1 ? h[1] : (i == 2 ? h[2] : h[3]));
  const uint32_t b = i == 0 ? h[1] : (i == 1 ? h[2] : (i == 2 ? h[3] : 0u));
#ifdef __CUDA_ARCH__
  return __funnelshift_r(a, b, sh);
#else
  return sh ? ((a >> sh) | (b << (32 - sh))) : a;
#endif
}
MGPU_HDN bool parse_ipv4_words(const uint32_t h[4], uint32_t n, uint32_t& addr_out) {
  if (n < 7 || n > 15) return false;
  const uint32_t D = swar_flags4(swar_is_digit(h[0])) | (swar_flags4(swar_is_digit(h[1])) << 4) | (swar_flags4(swar_is_digit(h[2])) << 8) |
                     (swar_flags4(swar_is_digit(h[3])) << 12);
  const uint32_t P = swar_flags4(swar_is_dot(h[0])) | (swar_flags4(swar_is_dot(h[1])) << 4) | (swar_flags4(swar_is_dot(h[2])) << 8) |
                     (swar_flags4(swar_is_dot(h[3])) << 12);
  const uint32_t all = (1u << n) - 1u;
  if (((D | P) & all) != all) return false;  // digits and dots only
  const uint32_t Pn = P & all;
#ifdef __CUDA_ARCH__
  if (__popc(Pn) != 3) return false;
  const uint32_t p1 = (uint32_t)__ffs((int)Pn) - 1u, P2 = Pn & (Pn - 1), p2 = (uint32_t)__ffs((int)P2) - 1u, p3 = (uint32_t)__ffs((int)(P2 & (P2 - 1))) - 1u;
#else
  if (__builtin_popcount(Pn) != 3) return false;
  const uint32_t p1 = (uint32_t)__builtin_ctz(Pn), P2 = Pn & (Pn - 1), p2 = (uint32_t)__builtin_ctz(P2), p3 = (uint32_t)__builtin_ctz(P2 & (P2 - 1));
#endif
  const uint32_t st[4] = {0u, p1 + 1, p2 + 1, p3 + 1}, ln[4] = {p1, p2 - p1 - 1, p3 - p2 - 1, n - p3 - 1};
  uint32_t addr = 0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t l = ln[k];
    ok = ok && l >= 1 && l <= 3;
    const uint32_t dv = word_at(h, st[k] & 15u) ^ 0x30303030u;  // the group's digits, first digit in the low byte
    const uint32_t d0 = dv & 0xFF, d1 = (dv >> 8) & 0xFF, d2 = (dv >> 16) & 0xFF;
    const uint32_t v = l == 1 ? d0 : (l == 2 ? d0 * 10 + d1 : d0 * 100 + d1 * 10 + d2);
    ok = ok && v <= 255 && !(l > 1 && d0 == 0);
    addr = (addr << 8) | (v & 0xFF);
  }
  if (!ok) return false;
  addr_out = addr;
  return true;
}
// The same parser for a word that is KNOWN to consist of hex digits and '.' only (the numeric queue's words: the tokenizer's
// fourth carry chain established it).  Then a byte is a decimal digit iff its bit 4 is set ('0'..'9' = 0x30..0x39, '.' = 0x2E,
// letters 0x41.. / 0x61..) and a dot iff bits 4 and 6 are both clear, the digit values are the low nibbles, and the four
// groups are cut out of the 16 nibbles by shifts: about half the instructions of the general version.  The fused scan
// kernel runs this one; tests/test_emulation_vs_oracle.py checks it against parse_ipv4_words on every such word it can make up.
MGPU_HD uint32_t pack_flags4(uint32_t x) { return (x * 0x01020408u) >> 24; }  // bit 0 of each of the four bytes -> four bits
MGPU_HDN bool parse_ipv4_hexdot(const uint32_t h[4], uint32_t n, uint32_t& addr_out) {
  if (n < 7 || n > 15) return false;
  const uint32_t D = pack_flags4((h[0] >> 4) & 0x01010101u) | (pack_flags4((h[1] >> 4) & 0x01010101u) << 4) | (pack_flags4((h[2] >> 4) & 0x01010101u) << 8) |
                     (pack_flags4((h[3] >> 4) & 0x01010101u) << 12);
  const uint32_t P = pack_flags4(~((h[0] >> 4) | (h[0] >> 6)) & 0x01010101u) | (pack_flags4(~((h[1] >> 4) | (h[1] >> 6)) & 0x01010101u) << 4) |
                     (pack_flags4(~((h[2] >> 4) | (h[2] >> 6)) & 0x01010101u) << 8) | (pack_flags4(~((h[3] >> 4) | (h[3] >> 6)) & 0x01010101u) << 12);
  const uint32_t all = (1u << n) - 1u;
  if (((D | P) & all) != all) return false;  // a letter
  const uint32_t Pn = P & all;
#ifdef __CUDA_ARCH__
  if (__popc(Pn) != 3) return false;
  const uint32_t p1 = (uint32_t)__ffs((int)Pn) - 1u, P2 = Pn & (Pn - 1), p2 = (uint32_t)__ffs((int)P2) - 1u, p3 = (uint32_t)__ffs((int)(P2 & (P2 - 1))) - 1u;
#else
  if (__builtin_popcount(Pn) != 3) return false;
  const uint32_t p1 = (uint32_t)__builtin_ctz(Pn), P2 = Pn & (Pn - 1), p2 = (uint32_t)__builtin_ctz(P2), p3 = (uint32_t)__builtin_ctz(P2 & (P2 - 1));
#endif
  // the 16 low nibbles, byte i's in nibble i
  uint32_t nlo, nhi;
  {
    const uint32_t a0 = h[0] & 0x0F0F0F0Fu, a1 = h[1] & 0x0F0F0F0Fu, a2 = h[2] & 0x0F0F0F0Fu, a3 = h[3] & 0x0F0F0F0Fu;
    const uint32_t b0 = a0 | (a0 >> 4), b1 = a1 | (a1 >> 4), b2 = a2 | (a2 >> 4), b3 = a3 | (a3 >> 4);  // bytes 0 and 2 hold two nibbles each
#ifdef __CUDA_ARCH__
    nlo = __byte_perm(b0, b1, 0x6420); nhi = __byte_perm(b2, b3, 0x6420);
#else
    nlo = (b0 & 0xFFu) | ((b0 >> 8) & 0xFF00u) | ((b1 & 0xFFu) << 16) | ((b1 & 0xFF0000u) << 8);
    nhi = (b2 & 0xFFu) | ((b2 >> 8) & 0xFF00u) | ((b3 & 0xFFu) << 16) | ((b3 & 0xFF0000u) << 8);
#endif
  }
  const uint32_t st[4] = {0u, p1 + 1, p2 + 1, p3 + 1}, ln[4] = {p1, p2 - p1 - 1, p3 - p2 - 1, n - p3 - 1};
  uint32_t addr = 0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t l = ln[k], sh = 4 * st[k];
    ok = ok && l >= 1 && l <= 3;
    // the group's nibbles, first digit lowest; then right-aligned in three nibbles: (0, 0, d0) / (0, d0, d1) / (d0, d1, d2)
#ifdef __CUDA_ARCH__
    const uint32_t x = sh < 32 ? __funnelshift_r(nlo, nhi, sh) : (nhi >> (sh - 32));
#else
    const uint64_t nn = ((uint64_t)nhi << 32) | nlo;
    const uint32_t x = (uint32_t)(nn >> sh);
#endif
    const uint32_t y = (x << (4 * ((3 - l) & 3))) & 0xFFFu;
    const uint32_t v = (y & 15u) * 100u + ((y >> 4) & 15u) * 10u + (y >> 8);
    ok = ok && v <= 255 && !(l > 1 && (x & 15u) == 0);
    addr = (addr << 8) | (v & 0xFF);
  }
  if (!ok) return false;
  addr_out = addr;
  return true;
}
MGPU_HDN bool parse_ipv4_word(const uint8_t* w, uint32_t n, uint32_t& addr_out) {
  if (n < 7 || n > 15) return false;
  uint32_t h[4];
  load_head_words(w, h);
  return parse_ipv4_words(h, n, addr_out);
}

// A boundary-delimited word made only of domain characters (incl. bytes >= 0x80), containing a '.', whose labels are
// non-empty and neither start nor end with '-' (all of that is established by the tokenizer's mask arithmetic):
// is it a domain?  Remaining rules: some dot-suffix is in the PSL, and the bytes are valid UTF-8.  (lib.rs:537-689)
MGPU_HDN bool domain_word_psl_utf8(const DbView& db, const uint8_t* w, uint32_t n) {
  // any byte >= 0x80?  word-wise scan, done first while the lanes of a warp are still in step
  uint32_t high = 0;
  for (uint32_t k = 0; k < n; k += 4) {
    uint32_t v = ldu32_fast(w + k);
    if (k + 4 > n) v &= 0xFFFFFFFFu >> (8 * (k + 4 - n));
    high |= v;
  }
  // PSL: shortest suffix first; any hit accepts.  The last label is hashed in a loop of its own so that the lanes probe
  // the table together; longer suffixes (only needed after a miss) follow in the general loop.
  uint64_t h = MGPU_FNV_BASIS;
  uint32_t k = n;
  while (k > 1 && w[k - 1] != '.') { h = psl_step(h, w[k - 1]); k--; }
  // w[k-1] == '.' (k >= 2) or k == 1 (no dot past index 0: cannot happen for tokenizer output, handled as a miss)
  bool hit = false;
  while (k >= 2) {
    uint32_t sl = n - k;  // suffix = w[k .. n)
    if (sl > db.psl_max_len) break;
    if (psl_contains(db, h, w + k, sl)) { hit = true; break; }
    h = psl_step(h, '.');
    k--;
    while (k > 1 && w[k - 1] != '.') { h = psl_step(h, w[k - 1]); k--; }
    if (k == 1) break;  // the remaining dot-free prefix is the first label: no further suffix
  }
  if (!hit) return false;
  if ((high & 0x80808080u) && !valid_utf8(w, n)) return false;
  return true;
}

// ---- fast path of the two functions above and of the string lookups: constant work per token ----
// The last min(n,8) bytes of w[0..n) with the LAST byte in bits 56..63 (missing leading bytes are zero).  n >= 1.
MGPU_HD uint64_t load_tail8(const uint8_t* w, uint32_t n) { return n >= 8 ? ldu64_fast(w + n - 8) : (ldu64_fast(w) << (8 * (8 - n))); }

enum { TLD_REJECT = 0, TLD_ACCEPT = 1, TLD_GENERAL = 2 };
// PSL decision from the last label alone.  find_valid_tld_suffix_bytes (lib.rs:1671-1692) accepts iff ANY dot-suffix of
// the word is a PSL entry, and every dot-suffix ends with the word's last label L.  So: L itself an entry -> accept;
// L neither an entry nor the last label of any multi-label entry -> no suffix can be an entry -> reject; otherwise
// (or when L is longer than 7 bytes) the general right-to-left walk decides.  `tld` holds (label bytes | len << 56),
// flagged TLD_SLOW for the third kind.  tail8 = load_tail8 of a word made of domain characters only ('/' cannot occur,
// which makes the borrow trick below exact).
MGPU_HD int tld_class(const uint64_t* tld, uint64_t tail8) {
  uint64_t x = tail8 ^ 0x2E2E2E2E2E2E2E2EULL;
  uint64_t z = (x - 0x0101010101010101ULL) & ~x & 0x8080808080808080ULL;  // 0x80 in every byte that is '.'
  if (z == 0) return TLD_GENERAL;                                          // no dot within the last 8 bytes
#ifdef __CUDA_ARCH__
  uint32_t L = (uint32_t)__clzll((long long)z) >> 3;                       // bytes after the last dot
#else
  uint32_t L = (uint32_t)__builtin_clzll(z) >> 3;
#endif
  if (L == 0) return TLD_GENERAL;
  uint64_t key = (tail8 >> (8 * (8 - L))) | ((uint64_t)L << 56);
  uint32_t slot = tld_slot(key);
  for (;;) {
    uint64_t k = tld[slot];
    if (k == 0) return TLD_REJECT;
    if ((k & ~TLD_SLOW) == key) return (k & TLD_SLOW) ? TLD_GENERAL : TLD_ACCEPT;
    slot = (slot + 1) & (TLD_SLOTS - 1);
  }
}
// domain_word_psl_utf8 with the constant-time front end.  maybe_high = false promises that w[0..n) is pure ASCII.
MGPU_HDN bool domain_word_fast(const DbView& db, const uint64_t* tld, const uint8_t* w, uint32_t n, bool maybe_high, uint64_t tail8) {
  int c = tld_class(tld, tail8);
  if (c == TLD_REJECT) return false;
  if (c == TLD_GENERAL) return domain_word_psl_utf8(db, w, n);
  if (!maybe_high) return true;
  uint32_t high = 0;
  for (uint32_t k = 0; k < n; k += 4) {
    uint32_t v = ldu32_fast(w + k);
    if (k + 4 > n) v &= 0xFFFFFFFFu >> (8 * (k + 4 - n));
    high |= v;
  }
  return !(high & 0x80808080u) || valid_utf8(w, n);
}
MGPU_HDN bool domain_word_fast(const DbView& db, const uint64_t* tld, const uint8_t* w, uint32_t n, bool maybe_high) {
  return domain_word_fast(db, tld, w, n, maybe_high, load_tail8(w, n));
}

// Which exact lookups can a string token still need?  Necessary conditions only (no false negatives):
//  * literal hash (LiteralHash::lookup is an exact string match): some stored literal has the same last K and the same first
//    K bytes (hot filter; K = lit_key_len(n)) and the same (length, first K, last K bytes) (cold filter);
//  * globs, when db.fast_ok: a pattern whose LAST segment is a literal can only match a text that ends with it, one whose
//    FIRST segment is a literal only a text that starts with it (match_segments_impl anchors segment 0 at position 0 and
//    requires the whole text to be consumed, paraglob_offset.rs:1402-1639); keys are the last / first glob_key_len(len)
//    bytes of that literal.  Patterns with neither anchor and literal-type patterns (substring semantics) are covered by
//    generic_literal_scan (one of THEIR literals must occur somewhere in the text); pure wildcards and case-insensitive
//    databases clear fast_ok and every token takes the exact path.
// Hot tests first (shared memory, `hot` may point at a copy of db.hot); at most one cold test (L2) per class afterwards.
// Returns F_LIT | F_GLOB bits.
// Unanchored (or literal-typed) globs: does any AC literal that leads to one of them START somewhere in the token?  Exact
// 2-byte bitmap (shared memory) then exact 3-byte bitmap (L2) per position; literals have at least 3 bytes.
#ifdef __CUDACC__
#define MGPU_NOINLINE_HD __host__ __device__ __noinline__
#else
#define MGPU_NOINLINE_HD inline
#endif
template <typename H, typename B>
MGPU_NOINLINE_HD bool generic_literal_scan(const DbView& db, const H& hot, const B& bytes, uint32_t n) {
  if (n < 3) return false;
  uint32_t g = (bytes.at(0) << 8) | bytes.at(1);
  for (uint32_t i = 2; i < n; i++) {
    const uint32_t c = bytes.at(i);
    if ((hot.gen2(g >> 5) >> (g & 31)) & 1u) {
      const uint32_t g3 = (g << 8) | c;
      if ((db.gen_gram3[g3 >> 5] >> (g3 & 31)) & 1u) return true;
    }
    g = ((g << 8) | c) & 0xFFFFu;
  }
  return false;
}

// The filters run in two stages so that the per-token cost is the cheap one.
//  Stage 1, string_gate: ONE hot-filter test per key class, in shared memory.  For the glob classes the hot filter holds,
//    for every anchored glob whose key has K >= 4 bytes, the last (suffix class) / first (prefix class) G bytes of the anchor
//    literal, G = the shortest such K in the database (db.glob_s_gate / glob_p_gate): a text that ends with the K-byte key ends
//    with its last G bytes, so "the G-byte key is in the hot filter" is necessary for ANY of those globs; keys of 1..3 bytes are
//    tested as they are.  For literals the gate is the tail key AND the head key.  Returns G_* bits: the classes that still owe
//    their exact tests.
//  Stage 2, string_filters_full: only for tokens with gate bits (a few per cent; the token kernel collects them and runs this
//    stage with full warps): the cold filter (L2, >= 16 bits per key) on the literal class's full key and on every glob key
//    length present.  Returns F_LIT | F_GLOB.
// string_filters() = stage 2 of stage 1: the decision function the tests pin against the oracle.
enum { G_LIT = 1u, G_S = 2u, G_P = 4u, G_GEN = 8u };
template <typename H>
MGPU_HDN uint32_t string_gate(const DbView& db, const H& hot, const KeyWords& kw, uint32_t n) {
  uint32_t g = 0;
  // the two running states every key of this stage is cut from: last words back to front, first words front to back
  const uint32_t c1 = key_mix(key_seed(TAG_GLOB_S), kw.t[3]), c2 = key_mix(c1, kw.t[2]);
  const uint32_t d1 = key_mix(key_seed(TAG_GLOB_P), kw.h[0]), d2 = key_mix(d1, kw.h[1]);
  if (db.has_literal) {  // some stored literal has the same last K and the same first K bytes (K = lit_key_len(n))
    const uint32_t k = lit_key_len(n);
    uint32_t st, sh;  // tail / head states
    if (k == 8) { st = c2; sh = d2; }
    else if (k == 4) { st = c1; sh = d1; }
    else { const uint32_t x = low_bytes32(kw.h[0], k); st = key_mix(key_seed(TAG_LIT_TAIL), x); sh = key_mix(key_seed(TAG_LIT_HEAD), x); }  // n < 4: head == tail
    bool pass = !((db.hot_tags >> TAG_LIT_TAIL) & 1u) || hot_test(hot, key_fin(st, k, TAG_LIT_TAIL));
    pass = pass && (!((db.hot_tags >> TAG_LIT_HEAD) & 1u) || hot_test(hot, key_fin(sh, k, TAG_LIT_HEAD)));
    if (pass) g |= G_LIT;
  }
  if (db.has_glob) {
    if (db.glob_s_lens) {
      if (!((db.hot_tags >> TAG_GLOB_S) & 1u)) g |= G_S;
      else {
        const uint32_t lens = db.glob_s_lens;
        if (lens & 0xEu) {  // K = 1, 2, 3: the last K bytes sit in the top bytes of t[3] (n >= 4) or are the head word's low bytes
          for (uint32_t k = 1; k <= 3; k++) {
            if (!((lens >> k) & 1u) || k > n) continue;
            const uint32_t x = n >= 4 ? (kw.t[3] >> (8 * (4 - k))) : (low_bytes32(kw.h[0], n) >> (8 * (n - k)));
            if (hot_test(hot, key_fin(key_mix(key_seed(TAG_GLOB_S), x), k, TAG_GLOB_S))) g |= G_S;
          }
        }
        const uint32_t gs = db.glob_s_gate;
        if (gs && n >= gs) {
          uint32_t s = gs == 4 ? c1 : c2;
          if (gs >= 12) s = key_mix(s, kw.t[1]);
          if (gs >= 16) s = key_mix(s, kw.t[0]);
          if (hot_test(hot, key_fin(s, gs, TAG_GLOB_S))) g |= G_S;
        }
      }
    }
    if (db.glob_p_lens) {
      if (!((db.hot_tags >> TAG_GLOB_P) & 1u)) g |= G_P;
      else {
        const uint32_t lens = db.glob_p_lens;
        if (lens & 0xEu) {
          for (uint32_t k = 1; k <= 3; k++) {
            if (!((lens >> k) & 1u) || k > n) continue;
            if (hot_test(hot, key_fin(key_mix(key_seed(TAG_GLOB_P), low_bytes32(kw.h[0], k)), k, TAG_GLOB_P))) g |= G_P;
          }
        }
        const uint32_t gp = db.glob_p_gate;
        if (gp && n >= gp) {
          uint32_t s = gp == 4 ? d1 : d2;
          if (gp >= 12) s = key_mix(s, kw.h[2]);
          if (gp >= 16) s = key_mix(s, kw.h[3]);
          if (hot_test(hot, key_fin(s, gp, TAG_GLOB_P))) g |= G_P;
        }
      }
    }
    if (db.has_generic) g |= G_GEN;
  }
  return g;
}

template <typename H, typename B>
MGPU_HDN uint32_t string_filters_full(const DbView& db, const H& hot, const KeyWords& kw, const B& bytes, uint32_t n, uint32_t g) {
  uint32_t flags = 0;
  if (g & G_LIT) {
    const uint32_t k = lit_key_len(n);
    uint32_t sh;  // head state
    if (k == 8) sh = key_mix(key_mix(key_seed(TAG_LIT_HEAD), kw.h[0]), kw.h[1]);
    else if (k == 4) sh = key_mix(key_seed(TAG_LIT_HEAD), kw.h[0]);
    else sh = key_mix(key_seed(TAG_LIT_HEAD), low_bytes32(kw.h[0], k));
    // (n, first K, last K bytes): continue the head state with the tail words
    uint32_t s = sh ^ key_seed(TAG_LIT_FULL);
    if (k == 8) s = key_mix(key_mix(s, kw.t[3]), kw.t[2]);
    else if (k == 4) s = key_mix(s, kw.t[3]);
    if (cold_test(db.cold, db.cold_mask, key_fin(s, n, TAG_LIT_FULL))) flags |= F_LIT;
  }
  if (g & G_S) {  // every key length present in the database, shortest first, straight to the cold filter
    const uint32_t lens = db.glob_s_lens;
    if (lens & 0xEu) {
      for (uint32_t k = 1; k <= 3; k++) {
        if (!((lens >> k) & 1u) || k > n) continue;
        const uint32_t x = n >= 4 ? (kw.t[3] >> (8 * (4 - k))) : (low_bytes32(kw.h[0], n) >> (8 * (n - k)));
        if (cold_test(db.cold, db.cold_mask, key_fin(key_mix(key_seed(TAG_GLOB_S), x), k, TAG_GLOB_S))) flags |= F_GLOB;
      }
    }
    uint32_t s = key_seed(TAG_GLOB_S);
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
      const uint32_t k = 4 * (j + 1);
      if ((lens >> k) == 0 || k > n || (flags & F_GLOB)) break;
      s = key_mix(s, kw.t[3 - j]);
      if (((lens >> k) & 1u) && cold_test(db.cold, db.cold_mask, key_fin(s, k, TAG_GLOB_S))) flags |= F_GLOB;
    }
  }
  if ((g & G_P) && !(flags & F_GLOB)) {
    const uint32_t lens = db.glob_p_lens;
    if (lens & 0xEu) {
      for (uint32_t k = 1; k <= 3; k++) {
        if (!((lens >> k) & 1u) || k > n) continue;
        if (cold_test(db.cold, db.cold_mask, key_fin(key_mix(key_seed(TAG_GLOB_P), low_bytes32(kw.h[0], k)), k, TAG_GLOB_P))) flags |= F_GLOB;
      }
    }
    uint32_t s = key_seed(TAG_GLOB_P);
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
      const uint32_t k = 4 * (j + 1);
      if ((lens >> k) == 0 || k > n || (flags & F_GLOB)) break;
      s = key_mix(s, kw.h[j]);
      if (((lens >> k) & 1u) && cold_test(db.cold, db.cold_mask, key_fin(s, k, TAG_GLOB_P))) flags |= F_GLOB;
    }
  }
  if ((g & G_GEN) && !(flags & F_GLOB) && generic_literal_scan(db, hot, bytes, n)) flags |= F_GLOB;
  return flags;
}

template <typename H, typename B>
MGPU_HDN uint32_t string_filters(const DbView& db, const H& hot, const KeyWords& kw, const B& bytes, uint32_t n) {
  const uint32_t g = string_gate(db, hot, kw, n);
  return g ? string_filters_full(db, hot, kw, bytes, n, g) : 0u;
}
MGPU_HDN uint32_t string_filters(const DbView& db, const uint32_t* hot, const uint8_t* w, uint32_t n) {
  KeyWords kw;
  load_head_words(w, kw.h);
  load_tail_words(w, n, kw.t);
  return string_filters(db, HotPtr{hot, db.gen_gram2}, kw, BytesPtr{w}, n);
}

// The same key hashes from a literal's bytes (database preparation; must mirror string_filters exactly).
MGPU_HD uint32_t bytes_le32(const uint8_t* p, uint32_t k) { uint32_t v = 0; for (uint32_t i = 0; i < k && i < 4; i++) v |= (uint32_t)p[i] << (8 * i); return v; }
MGPU_HD uint32_t glob_key_hash(const uint8_t* lit, uint32_t m, uint32_t tag) {  // tag: TAG_GLOB_S (last bytes) or TAG_GLOB_P (first bytes)
  const uint32_t k = glob_key_len(m);
  uint32_t s = key_seed(tag);
  if (k < 4) return key_fin(key_mix(s, bytes_le32(tag == TAG_GLOB_S ? lit + m - k : lit, k)), k, tag);
  for (uint32_t j = 0; j < k / 4; j++) s = key_mix(s, tag == TAG_GLOB_S ? bytes_le32(lit + m - 4 * (j + 1), 4) : bytes_le32(lit + 4 * j, 4));
  return key_fin(s, k, tag);
}
// the gate key of a glob anchor literal whose own key has K >= 4 bytes: its last / first g bytes (g in {4, 8, 12, 16}, g <= K)
MGPU_HD uint32_t glob_gate_hash(const uint8_t* lit, uint32_t m, uint32_t g, uint32_t tag) {
  uint32_t s = key_seed(tag);
  for (uint32_t j = 0; j < g / 4; j++) s = key_mix(s, tag == TAG_GLOB_S ? bytes_le32(lit + m - 4 * (j + 1), 4) : bytes_le32(lit + 4 * j, 4));
  return key_fin(s, g, tag);
}
MGPU_HD void lit_key_hashes(const uint8_t* str, uint32_t n, uint32_t& tail_h, uint32_t& head_h, uint32_t& full_h) {
  const uint32_t k = lit_key_len(n);
  uint32_t st = key_seed(TAG_LIT_TAIL), sh = key_seed(TAG_LIT_HEAD);
  if (k == 8) { st = key_mix(key_mix(st, bytes_le32(str + n - 4, 4)), bytes_le32(str + n - 8, 4)); sh = key_mix(key_mix(sh, bytes_le32(str, 4)), bytes_le32(str + 4, 4)); }
  else if (k == 4) { st = key_mix(st, bytes_le32(str + n - 4, 4)); sh = key_mix(sh, bytes_le32(str, 4)); }
  else { const uint32_t x = bytes_le32(str, k); st = key_mix(st, x); sh = key_mix(sh, x); }
  tail_h = key_fin(st, k, TAG_LIT_TAIL); head_h = key_fin(sh, k, TAG_LIT_HEAD);
  uint32_t s = sh ^ key_seed(TAG_LIT_FULL);
  if (k == 8) s = key_mix(key_mix(s, bytes_le32(str + n - 4, 4)), bytes_le32(str + n - 8, 4));
  else if (k == 4) s = key_mix(s, bytes_le32(str + n - 4, 4));
  full_h = key_fin(s, n, TAG_LIT_FULL);
}

// extract_email_at: buf[lo..n) is the chunk, at = position of '@'.
MGPU_HDN bool email_at(const DbView& db, const uint8_t* buf, size_t lo, size_t n, size_t at, size_t& s_out, size_t& e_out) {
  size_t start = at;
  while (start > lo && is_local_char(buf[start - 1])) start--;
  if (start == at) return false;
  if (start > lo && !is_boundary(buf[start - 1])) return false;
  size_t end = at + 1;
  while (end < n && is_domain_ascii(buf[end])) end++;
  if (end == at + 1) return false;
  if (end < n && !is_boundary(buf[end])) return false;
  bool has_letter = false;
  for (size_t k = start; k < at; k++) {
    if (buf[k] == '.' && k + 1 < at && buf[k + 1] == '.') return false;
    has_letter |= is_alpha(buf[k]);
  }
  if (!has_letter) return false;
  // "domain part contains a dot" is implied by a PSL suffix hit
  if (end - at - 1 > 0xFFFFFFFFull) return false;
  if (!psl_any_suffix(db, buf + at + 1, (uint32_t)(end - at - 1))) return false;
  s_out = start; e_out = end;
  return true;
}

// core::net::parser Ipv6Addr::from_str for inputs made of hex digits and ':' only.
MGPU_HD uint32_t v6_read_groups(const uint8_t* s, uint32_t n, uint32_t& pos, uint16_t* groups, uint32_t limit) {
  for (uint32_t i = 0; i < limit; i++) {
    uint32_t p = pos;
    if (i > 0) { if (p < n && s[p] == ':') p++; else return i; }
    uint32_t v = 0, digits = 0;
    while (p < n) {
      uint8_t c = s[p]; uint32_t d;
      if (is_digit(c)) d = c - '0';
      else if ((uint8_t)((c | 32) - 'a') < 6) d = (c | 32) - 'a' + 10;
      else break;
      v = v * 16 + d; digits++; p++;
      if (digits > 4) return i;
    }
    if (digits == 0) return i;
    groups[i] = (uint16_t)v;
    pos = p;
  }
  return limit;
}
MGPU_HDN bool parse_ipv6_run(const uint8_t* s, uint32_t n, uint16_t out[8]) {
  uint32_t pos = 0;
  uint16_t head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t hs = v6_read_groups(s, n, pos, head, 8);
  if (hs == 8) {
    if (pos != n) return false;
  } else {
    if (!(pos + 1 < n && s[pos] == ':' && s[pos + 1] == ':')) return false;
    pos += 2;
    uint16_t tail[7] = {0, 0, 0, 0, 0, 0, 0};
    uint32_t limit = 8 - (hs + 1);
    uint32_t ts = v6_read_groups(s, n, pos, tail, limit);
    if (pos != n) return false;
    for (uint32_t k = 0; k < ts; k++) head[8 - ts + k] = tail[k];
  }
  for (int k = 0; k < 8; k++) out[k] = head[k];
  return true;
}
// The same parser as straight-line mask arithmetic (one code path for every lane of a warp; the loop version above costs a
// three-way divergent branch per character).  For a run of 8..39 bytes made of hex digits and ':' only, the rules of
// parse_ipv6_run collapse to: the text holds "::" (the caller anchors on one), so the eight-group form cannot consume it all;
// it is valid iff no single ':' stands at either end, exactly one position starts a "::" (":::" counts twice), no hex
// field is longer than 4 digits, and there are at most 7 fields; the fields before the "::" fill groups 0.., those after it
// fill the groups up to 7.  C = colon mask (bit i = byte i).  Output as IpTok words: w[k] = seg[2k] << 16 | seg[2k+1].
MGPU_HD uint32_t swar_is_colon(uint32_t v) {  // 0x80 in every byte that is ':'
  uint32_t d = v ^ 0x3A3A3A3Au;
  return ~(((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
}
MGPU_HD uint32_t bit_count64(uint64_t x) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__popcll(x);
#else
  return (uint32_t)__builtin_popcountll(x);
#endif
}
MGPU_HD uint32_t low_bit64(uint64_t x) {  // x != 0
#ifdef __CUDA_ARCH__
  return (uint32_t)__ffsll((long long)x) - 1u;
#else
  return (uint32_t)__builtin_ctzll(x);
#endif
}
MGPU_HDN bool parse_ipv6_run_masks(const uint8_t* s, uint32_t n, uint32_t w[4]) {
  if (n < 2 || n > 39) return false;
  uint64_t C = 0;
#pragma unroll
  for (uint32_t j = 0; j < 10; j++) {
    const uint32_t v = 4 * j < n ? ldu32_fast(s + 4 * j) : 0u;
    C |= (uint64_t)swar_flags4(swar_is_colon(v)) << (4 * j);
  }
  const uint64_t all = (1ULL << n) - 1;
  C &= all;
  const uint64_t H = ~C & all;
  const uint64_t dc = C & (C >> 1);
  if (bit_count64(dc) != 1) return false;
  if (((C & 3) == 1) || ((C >> (n - 2)) & 3) == 2) return false;  // a single ':' at either end (a "::" there is fine)
  if (H & (H >> 1) & (H >> 2) & (H >> 3) & (H >> 4)) return false;  // a field of five or more digits
  uint64_t starts = H & ~(H << 1);
  const uint32_t fields = bit_count64(starts);
  if (fields > 7) return false;
  const uint32_t gap_at = low_bit64(dc);  // fields that start beyond the "::" are right-aligned in the eight groups
  w[0] = w[1] = w[2] = w[3] = 0u;
#pragma unroll 1
  for (uint32_t k = 0; k < fields; k++) {
    const uint32_t at = low_bit64(starts);
    starts &= starts - 1;
    const uint32_t len = low_bit64(~(H >> at));  // 1..4
    const uint32_t v = ldu32_fast(s + at);       // up to four digits, first one in the low byte (the buffer has 16 bytes of slack)
    const uint32_t nib = (v & 0x0F0F0F0Fu) + 9u * ((v >> 6) & 0x01010101u);  // '0'-'9' -> 0-9, 'a'-'f' / 'A'-'F' -> 10-15
    const uint32_t val = (((nib & 0xFu) << 12) | (((nib >> 8) & 0xFu) << 8) | (((nib >> 16) & 0xFu) << 4) | ((nib >> 24) & 0xFu)) >> (4 * (4 - len));
    const uint32_t g = at > gap_at ? 8 - fields + k : k;
    const uint32_t word = val << ((g & 1u) ? 0 : 16);
    w[0] |= (g >> 1) == 0 ? word : 0u; w[1] |= (g >> 1) == 1 ? word : 0u; w[2] |= (g >> 1) == 2 ? word : 0u; w[3] |= (g >> 1) == 3 ? word : 0u;
  }
  return true;
}

// extract_ipv6_chunk for the "::" whose first colon is at `at`.  Returns true with the run span and address.
MGPU_HDN bool ipv6_at(const uint8_t* buf, size_t lo, size_t n, size_t at, size_t& s_out, size_t& e_out, uint16_t segs[8]) {
  // maximal [0-9A-Fa-f:] run around the anchor; a valid address is at most 39 bytes, longer runs cannot parse
  size_t start = at;
  while (start > lo && at - start <= 40) { uint8_t c = buf[start - 1]; if (!is_hex(c) && c != ':') break; start--; }
  if (at - start > 40) return false;
  size_t end = at + 2;
  while (end < n && end - at <= 42) { uint8_t c = buf[end]; if (!is_hex(c) && c != ':') break; end++; }
  uint32_t len = (uint32_t)(end - start);
  if (len > 39 || len < 8) return false;
  const uint8_t* c = buf + start;
  if ((c[0] == ':' && c[1] == ':') || (c[len - 2] == ':' && c[len - 1] == ':')) return false;
  if ((c[0] | 32) == 'f' && (c[1] | 32) == 'e') {  // link-local fe80::/10 by text prefix (lib.rs:1425-1456)
    uint8_t t = c[2] | 32;
    if (c[2] == '8' || c[2] == '9' || t == 'a' || t == 'b') return false;
  }
  if (!parse_ipv6_run(c, len, segs)) return false;
  s_out = start; e_out = end;
  return true;
}
// The same, for the kernels: run bounds by the same byte loops, address through parse_ipv6_run_masks, as IpTok words.
MGPU_HDN bool ipv6_at_words(const uint8_t* buf, size_t lo, size_t n, size_t at, size_t& s_out, size_t& e_out, uint32_t w[4]) {
  size_t start = at;
  while (start > lo && at - start <= 40) { uint8_t c = buf[start - 1]; if (!is_hex(c) && c != ':') break; start--; }
  if (at - start > 40) return false;
  size_t end = at + 2;
  while (end < n && end - at <= 42) { uint8_t c = buf[end]; if (!is_hex(c) && c != ':') break; end++; }
  uint32_t len = (uint32_t)(end - start);
  if (len > 39 || len < 8) return false;
  const uint8_t* c = buf + start;
  if ((c[0] == ':' && c[1] == ':') || (c[len - 2] == ':' && c[len - 1] == ':')) return false;
  if ((c[0] | 32) == 'f' && (c[1] | 32) == 'e') {
    uint8_t t = c[2] | 32;
    if (c[2] == '8' || c[2] == '9' || t == 'a' || t == 'b') return false;
  }
  if (!parse_ipv6_run_masks(c, len, w)) return false;
  s_out = start; e_out = end;
  return true;
}

// =================================================================================================
// IP search tree — tree.rs:46-277.  Tree validity (every record < node_count, == node_count or >= node_count+16)
// is checked at upload, so the Err paths of the reference cannot occur here.
// =================================================================================================
MGPU_HD uint32_t tree_record(const DbView& db, uint32_t node, uint32_t side) {
  if (db.record_bits == 24) {
    const uint8_t* p = db.tree + (size_t)node * 6 + side * 3;
    return ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2];
  } else if (db.record_bits == 28) {
    const uint8_t* b = db.tree + (size_t)node * 7;
    if (side == 0) return ((uint32_t)(b[3] >> 4) << 24) | ((uint32_t)b[0] << 16) | ((uint32_t)b[1] << 8) | b[2];
    return ((uint32_t)(b[3] & 15) << 24) | ((uint32_t)b[4] << 16) | ((uint32_t)b[5] << 8) | b[6];
  } else {
    const uint8_t* p = db.tree + (size_t)node * 8 + side * 4;
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
}
// The walk as a resumable state machine: trie_walk_begin() consults the jump table, trie_walk_step() reads ONE record.
// The IP-trie kernel advances all lanes of a warp in lock step (a few steps per round), so that no lane sits in a
// 100-step serial walk while the rest of its warp waits; trie_lookup_v4 / _v6 below are the same machine run to the end.
// Address bits as IpTok words: bit bi (0 = most significant) is bit 31 - (bi & 31) of w[bi >> 5]; an IPv4 address is w[0].
enum { TRIE_MISS = 0, TRIE_HIT = 1, TRIE_MORE = 2 };
struct TrieWalk { uint32_t node, bi, end; };
MGPU_HD uint32_t trie_addr_bit(const uint32_t w[4], uint32_t bi) {
  const uint32_t k = bi >> 5;
  const uint32_t word = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
  return (word >> (31u - (bi & 31u))) & 1u;
}
MGPU_HD int trie_walk_begin(const DbView& db, const uint32_t w[4], bool v6, TrieWalk& s, uint32_t& data_off, uint8_t& prefix) {
  if (!v6) {
    s.node = db.ip_version == 6 ? db.v4_start_node : 0; s.bi = 0; s.end = 32;
    if (db.v4_top16) {  // sixteen / twenty dependent node reads in one table read
      const uint32_t tb = db.v4_top_bits, idx = w[0] >> (32 - tb);
      const uint32_t e = db.v4_top16[idx], kind = e >> 28;
      if (kind == 1) return TRIE_MISS;
      if (kind == 2) { data_off = (e & 0x0FFFFFFFu) - db.node_count - 16; prefix = db.v4_top16_depth[idx]; return TRIE_HIT; }
      s.node = e & 0x0FFFFFFFu; s.bi = tb;
    }
  } else {
    s.node = 0; s.bi = 0; s.end = 128;
    if (db.v6_top16) {
      const uint32_t idx = w[0] >> 16;
      const uint32_t e = db.v6_top16[idx], kind = e >> 28;
      if (kind == 1) return TRIE_MISS;
      if (kind == 2) { data_off = (e & 0x0FFFFFFFu) - db.node_count - 16; prefix = db.v6_top16_depth[idx]; return TRIE_HIT; }
      s.node = e & 0x0FFFFFFFu; s.bi = 16;
    }
  }
  return s.bi < s.end ? TRIE_MORE : TRIE_MISS;
}
MGPU_HD int trie_walk_step(const DbView& db, const uint32_t w[4], TrieWalk& s, uint32_t& data_off, uint8_t& prefix) {
  const uint32_t rec = tree_record(db, s.node, trie_addr_bit(w, s.bi));
  if (rec == db.node_count) return TRIE_MISS;
  if (rec < db.node_count) { s.node = rec; s.bi++; return s.bi < s.end ? TRIE_MORE : TRIE_MISS; }
  data_off = rec - db.node_count - 16;
  prefix = (uint8_t)(s.bi + 1);  // depth counts one per address bit in both tree kinds (96 is subtracted again, tree.rs:76-80)
  return TRIE_HIT;
}
MGPU_HDN bool trie_lookup_words(const DbView& db, const uint32_t w[4], bool v6, uint32_t& data_off, uint8_t& prefix) {
  TrieWalk s;
  int r = trie_walk_begin(db, w, v6, s, data_off, prefix);
  while (r == TRIE_MORE) r = trie_walk_step(db, w, s, data_off, prefix);
  return r == TRIE_HIT;
}
MGPU_HDN bool trie_lookup_v4(const DbView& db, uint32_t bits, uint32_t& data_off, uint8_t& prefix) {
  const uint32_t w[4] = {bits, 0, 0, 0};
  return trie_lookup_words(db, w, false, data_off, prefix);
}
MGPU_HDN bool trie_lookup_v6(const DbView& db, const uint16_t seg[8], uint32_t& data_off, uint8_t& prefix) {
  uint32_t w[4];
  for (int k = 0; k < 4; k++) w[k] = ((uint32_t)seg[2 * k] << 16) | seg[2 * k + 1];
  return trie_lookup_words(db, w, true, data_off, prefix);
}

// =================================================================================================
// Literal hash — matchy-literal-hash/src/lib.rs:467-575
// =================================================================================================
MGPU_HDN bool lh_lookup(const DbView& db, const uint8_t* q, uint32_t n, uint32_t& pattern_id) {
  bool fold = db.match_mode == 1;
  uint64_t h = xxh64_fold(q, n, fold);
  if (db.lh_bloom) {
    uint64_t need = (1ULL << (h & 63)) | (1ULL << ((h >> 6) & 63)) | (1ULL << ((h >> 12) & 63));
    if ((db.lh_bloom[(uint32_t)(h >> 20) & db.lh_bloom_mask] & need) != need) return false;  // certainly absent
  }
  uint32_t shard = (uint32_t)(h % db.lh_num_shards);
  uint32_t s0 = ld32(db.lh + 32 + (size_t)shard * 4), s1 = ld32(db.lh + 32 + (size_t)shard * 4 + 4);
  uint32_t cap = s1 - s0;
  if (cap == 0) return false;
  uint32_t mask = cap - 1;
  uint32_t slot = s0 + ((uint32_t)h & mask);
  for (uint32_t it = 0; it < cap; it++) {
    uint64_t eo = (uint64_t)db.lh_table_start + (uint64_t)slot * 16;
    if (eo + 16 > db.lh_len) return false;
    const uint8_t* e = db.lh + eo;
#ifdef __CUDA_ARCH__
    const uint4 ent = *reinterpret_cast<const uint4*>(e);  // entries are 16-byte aligned in the device copy
    const uint32_t so = ent.z, e_pid = ent.w;
    const uint64_t eh = (uint64_t)ent.x | ((uint64_t)ent.y << 32);
#else
    const uint32_t so = ld32u(e + 8), e_pid = ld32u(e + 12);
    const uint64_t eh = ld64u(e);
#endif
    if (so == NONE32) return false;
    if (eh == h) {
      uint64_t abs = (uint64_t)db.lh_strings_offset + so;
      if (abs + 2 <= db.lh_len) {
        uint32_t sl = (uint32_t)db.lh[abs] | ((uint32_t)db.lh[abs + 1] << 8);
        if (abs + 2 + sl <= db.lh_len && sl == n) {
          // (stored strings are valid UTF-8 — checked at upload — so read_string cannot fail here)
          const uint8_t* sp = db.lh + abs + 2;
          bool eq = true;
          for (uint32_t i = 0; i < n; i++) if (sp[i] != lc(q[i], fold)) { eq = false; break; }
          if (eq) { pattern_id = e_pid; return true; }
        }
      }
    }
    slot = s0 + ((slot + 1 - s0) & mask);
  }
  return false;
}
MGPU_HD bool lh_data_offset(const DbView& db, uint32_t pattern_id, uint32_t& off) {
  if (pattern_id >= db.lh_data_index_n) return false;
  uint32_t v = db.lh_data_index[pattern_id];
  // NONE32 doubles as "no mapping entry"; a genuine data offset of 0xFFFFFFFF cannot exist in a < 4 GiB file
  if (v == NONE32) return false;
  off = v;
  return true;
}

// =================================================================================================
// Paraglob — paraglob_offset.rs:1028-1639
// =================================================================================================
// match_segments_impl :1402-1639 as an explicit-stack machine.  Every reference call (including failed ones and
// the calls a Star makes for each candidate position) decrements the 100 000-step budget exactly once.
#define MGPU_GLOB_MAX_STARS 24  // patterns with more '*' segments are refused at upload
MGPU_HDN bool glob_match(const DbView& db, uint32_t pattern_id, const uint8_t* text, uint32_t tn) {
  uint64_t io = (uint64_t)db.glob_segments_offset + (uint64_t)pattern_id * 8;
  if (io + 8 > db.pg_len) return false;
  // zerocopy::Ref::from_prefix needs 4-byte alignment for GlobSegmentIndex / GlobSegmentHeader / CharClassItemEncoded
  // (offset_format.rs:391-431); a misaligned read is an Err, which `?` carries out of the whole match => not a match.
  if (((db.pg_align + io) & 3) != 0) return false;
  const uint32_t first = ld32(db.pg + io);
  const uint32_t count = (uint32_t)db.pg[io + 4] | ((uint32_t)db.pg[io + 5] << 8);
  const bool ci = db.match_mode == 1;
  uint32_t steps = 100000;
  uint32_t st_idx[MGPU_GLOB_MAX_STARS], st_pos[MGPU_GLOB_MAX_STARS];
  int sp = 0;
  uint32_t pos = 0, idx = 0;
  for (;;) {
    bool result;
    // ---- one call of match_segments_impl(pos, idx) ----
    for (;;) {
      if (steps == 0) { result = false; break; }
      steps--;
      if (idx >= count) { result = pos >= tn; break; }
      uint64_t so = (uint64_t)first + (uint64_t)idx * 12;
      if (so + 12 > db.pg_len) { result = false; break; }
      if (((db.pg_align + so) & 3) != 0) return false;  // Err("Invalid GlobSegmentHeader")
      const uint8_t* sh = db.pg + so;
      uint32_t stype = sh[0], sflags = sh[1], dlen = ld32(sh + 4), doff = ld32(sh + 8);
      if (stype == 0) {
        if ((uint64_t)doff + dlen > db.pg_len) { result = false; break; }
        const uint8_t* lit = db.pg + doff;
        uint32_t adv;
        bool m;
        if (!ci) {
          m = tn - pos >= dlen;
          for (uint32_t i = 0; m && i < dlen; i++) m = text[pos + i] == lit[i];
          adv = dlen;
        } else {  // chars of literal vs chars of text, ASCII-insensitive (:1456-1478)
          uint32_t lp = 0, tp = pos;
          m = true;
          while (tp < tn && lp < dlen) {
            uint32_t lcp, tcp;
            uint32_t ll = utf8_decode(lit, dlen, lp, lcp), tl = utf8_decode(text, tn, tp, tcp);
            bool eq = (lcp < 128 && tcp < 128) ? lc((uint8_t)lcp, true) == lc((uint8_t)tcp, true) : lcp == tcp;
            if (!eq) { m = false; break; }
            lp += ll; tp += tl;
          }
          if (m && lp < dlen) m = false;
          adv = tp - pos;
        }
        if (!m) { result = false; break; }
        pos += adv; idx++;
        continue;  // tail call
      } else if (stype == 1) {
        if (idx + 1 >= count) { result = true; break; }
        if (sp >= MGPU_GLOB_MAX_STARS) { result = false; break; }  // unreachable: checked at upload
        st_idx[sp] = idx; st_pos[sp] = pos; sp++;
        idx++;
        continue;  // first candidate position: call (pos, idx+1)
      } else if (stype == 2) {
        if (pos >= tn) { result = false; break; }
        pos += utf8_len(text[pos]); idx++;
        continue;
      } else if (stype == 3) {
        if (pos >= tn) { result = false; break; }
        uint32_t ch;
        uint32_t l = utf8_decode(text, tn, pos, ch);
        if (ci && ch < 128) ch = lc((uint8_t)ch, true);
        if ((uint64_t)doff + dlen > db.pg_len) { result = false; break; }
        uint32_t items = dlen / 12;
        if (items && ((db.pg_align + doff) & 3) != 0) return false;  // Err("Invalid CharClassItemEncoded")
        bool in_class = false;
        for (uint32_t i = 0; i < items; i++) {
          const uint8_t* it = db.pg + doff + i * 12;
          uint32_t c1 = ld32(it + 4), c2 = ld32(it + 8);
          bool v1 = c1 <= 0x10FFFF && !(c1 >= 0xD800 && c1 <= 0xDFFF), v2 = c2 <= 0x10FFFF && !(c2 >= 0xD800 && c2 <= 0xDFFF);
          if (ci && c1 < 128) c1 = lc((uint8_t)c1, true);
          if (ci && c2 < 128) c2 = lc((uint8_t)c2, true);
          bool mi = false;
          if (it[0] == 0) mi = v1 && ch == c1;
          else if (it[0] == 1) mi = v1 && v2 && ch >= c1 && ch <= c2;
          if (mi) { in_class = true; break; }
        }
        if (((sflags & 1) != 0) == in_class) { result = false; break; }
        pos += l; idx++;
        continue;
      } else { result = false; break; }
    }
    // ---- unwind: the innermost call returned `result` ----
    for (;;) {
      if (result || sp == 0) return result;  // a true result propagates through every enclosing Star
      uint32_t p = st_pos[sp - 1];
      if (p >= tn) { sp--; continue; }  // this Star is exhausted -> it returns false to its caller
      p += utf8_len(text[p]);
      st_pos[sp - 1] = p;
      pos = p; idx = st_idx[sp - 1] + 1;
      break;  // next candidate position for the innermost Star
    }
  }
}

// Automaton node held in registers: the 20-byte ACNodeHot minus one_target (== edges_offset for kind One).
struct AcNode { uint32_t w0, fail, eo, po; };
MGPU_HD AcNode ac_fetch(const uint8_t* ac, uint32_t off) {
  AcNode n;
  n.w0 = ld32(ac + off); n.fail = ld32(ac + off + 8); n.eo = ld32(ac + off + 12); n.po = ld32(ac + off + 16);
  return n;
}
// goto function of node `nd` on byte ch (find_ac_transition :1271-1353); 0 = no transition (offset 0 is the root and is
// never a goto target).  Offsets were bounds- and alignment-checked at upload (db_prepare.h), so no per-step checks.
MGPU_HD uint32_t ac_goto(const uint8_t* ac, const AcNode& nd, uint8_t ch) {
  uint32_t kind = nd.w0 & 0xFF;
  if (kind == 1) return ((nd.w0 >> 8) & 0xFF) == ch ? nd.eo : 0u;
  if (kind == 2) {
    uint32_t cnt = (nd.w0 >> 16) & 0xFF;
    for (uint32_t i = 0; i < cnt; i++) {
      uint32_t e0 = ld32(ac + nd.eo + i * 8);
      uint8_t ec = (uint8_t)e0;
      if (ec == ch) return ld32(ac + nd.eo + i * 8 + 4);
      if (ec > ch) return 0u;
    }
    return 0u;
  }
  if (kind == 3) return ld32(ac + nd.eo + (uint32_t)ch * 4);
  return 0u;
}

// slot hash of an m-byte trie path held in the low bytes of v (m <= 8)
MGPU_HD uint32_t prefix_slot(uint64_t v, uint32_t m, uint32_t mask) {
  uint64_t k = (v ^ ((uint64_t)m * 0x9E3779B97F4A7C15ULL)) * 0xFF51AFD7ED558CCDULL;
  k ^= k >> 33; k *= 0xC4CEB9FE1A85EC53ULL; k ^= k >> 29;
  return (uint32_t)k & mask;
}
// node reached from the root by the m bytes in v, or 0 if the map has no such path (0 is the root: never a target)
MGPU_HD uint32_t prefix_node(const DbView& db, uint64_t v, uint32_t m) {
  uint32_t slot = prefix_slot(v, m, db.ac_pfx_mask);
  for (;;) {
    uint32_t len = db.ac_pfx_vals[2 * slot + 1];
    if (len == 0) return 0u;
    if (len == m && db.ac_pfx_keys[slot] == v) return db.ac_pfx_vals[2 * slot];
    slot = (slot + 1) & db.ac_pfx_mask;
  }
}
MGPU_HD uint64_t load_prefix8(const uint8_t* text, uint32_t pos, bool fold) {
  if (!fold) return ldu64_fast(text + pos);
  uint64_t v = 0;
  for (uint32_t k = 0; k < 8; k++) v |= (uint64_t)lc(text[pos + k], true) << (8 * k);  // may read past the token: buffer slack
  return v;
}
MGPU_HD uint64_t low_bytes(uint64_t v, uint32_t m) { return m >= 8 ? v : (v & ((1ULL << (8 * m)) - 1)); }

struct AcAccel {
  const uint32_t* root_tab;  // root's 256-entry dense table (maybe a shared-memory copy), or nullptr if the root is not dense
  const uint32_t* gram2;     // 2-byte-prefix bitmap (maybe a shared-memory copy), or nullptr
};

// candidate patterns of one literal hit -> verification -> emit (paraglob_offset.rs:1136-1171)
template <typename F>
MGPU_HD void ac_outputs(const DbView& db, const uint8_t* ac, const AcNode& nd, const uint8_t* text, uint32_t tn, F&& emit) {
  uint32_t pc = nd.w0 >> 24;
  for (uint32_t k = 0; k < pc; k++) {
    uint32_t lit = ld32(ac + nd.po + k * 4);
    if (lit >= db.aclh_n) continue;
    uint32_t lo = db.aclh_index[2 * lit], lcnt = db.aclh_index[2 * lit + 1];
    for (uint32_t j = 0; j < lcnt; j++) {
      uint32_t pid = ld32(db.pg + lo + j * 4);
      uint64_t eo = (uint64_t)db.patterns_offset + (uint64_t)pid * 16;
      if (eo + 16 > db.pg_len) continue;
      uint32_t entry_id = ld32(db.pg + eo);
      if (db.pg[eo + 4] == 0) emit(entry_id);                 // literal-type pattern: accepted on the AC hit alone
      else if (glob_match(db, entry_id, text, tn)) emit(entry_id);
    }
  }
}

// Visit every glob id find_all would return for `text` (before sort+dedup; duplicates possible).
// run_ac_matching_into_static :1186-1266 + ACLH map + verification :1136-1171, pure wildcards :1096-1134.
//
// Two result-identical ways to find the literal ids that occur in the text:
//  * the reference's Aho-Corasick walk (goto / failure links, suffix outputs merged into every node);
//  * ANCHORED walks (when db.ac_anchored): for every start position whose first 2 and 3 bytes begin some literal
//    (exact bitmaps built at upload), follow goto edges only, from the root, until they stop.  A literal that occurs
//    at [i, j) is found by the walk anchored at i; every id listed at a node reached from i is a suffix of
//    text[i..j], hence occurs too.  Same id SET as the AC walk (duplicates differ, they are removed later anyway).
//    Benign tokens leave the first shared-memory bitmap test at almost every position, and lanes do not diverge
//    in failure-link loops.
template <typename F>
MGPU_HDN void find_all_visit(const DbView& db, const uint8_t* text, uint32_t tn, const AcAccel& acc, F&& emit) {
  if (db.pg_len < 112) return;
  // pure wildcards are tested on every query
  for (uint32_t i = 0; i < db.wild_count; i++) {
    uint64_t wo = (uint64_t)db.wild_off + (uint64_t)i * 8;
    if (wo + 8 > db.pg_len) continue;
    uint32_t pid = ld32(db.pg + wo);
    if ((uint64_t)db.patterns_offset + (uint64_t)pid * 16 + 16 > db.pg_len) continue;
    if (glob_match(db, pid, text, tn)) emit(pid);
  }
  if (db.ac_size < 20 || tn == 0) return;
  const uint8_t* ac = db.pg + db.ac_start;
  const bool fold = db.match_mode == 1;
  const AcNode root = ac_fetch(ac, 0);
  if (db.ac_anchored && acc.gram2) {
    if (tn < 3) return;  // every literal has at least 3 bytes
    // One event per loop iteration (either test the next start position or take one goto step of the current walk), so the
    // lanes of a warp stay in one loop instead of serialising each other's inner walks.
    uint32_t g = ((uint32_t)lc(text[0], fold) << 8) | lc(text[1], fold);
    uint32_t i = 0, j = 0;
    bool walking = false;
    AcNode nd = root;
    for (;;) {
      if (!walking) {
        if (i + 2 >= tn) break;
        uint32_t g2 = g & 0xFFFF;
        g = (g << 8) | lc(text[i + 2], fold);
        if ((acc.gram2[g2 >> 5] >> (g2 & 31)) & 1u) {
          uint32_t g3 = g & 0xFFFFFF;
          if ((db.ac_gram3[g3 >> 5] >> (g3 & 31)) & 1u) {
            // outputs at depth 3..7, then jump to the depth-8 node
            uint64_t v = load_prefix8(text, i, fold);
            uint32_t avail = tn - i;
            for (uint32_t lens = db.ac_short_lens; lens; lens &= lens - 1) {
#ifdef __CUDA_ARCH__
              uint32_t m = (uint32_t)__ffs((int)lens) - 1u;
#else
              uint32_t m = (uint32_t)__builtin_ctz(lens);
#endif
              if (m > avail) break;
              uint32_t off = prefix_node(db, low_bytes(v, m), m);
              if (off) { AcNode sn = ac_fetch(ac, off); ac_outputs(db, ac, sn, text, tn, emit); }
            }
            if (avail >= 8) {
              uint32_t off = prefix_node(db, v, 8);
              if (off) {
                nd = ac_fetch(ac, off);
                if (nd.w0 >> 24) ac_outputs(db, ac, nd, text, tn, emit);
                j = i + 8;
                walking = j < tn;
              }
            }
          }
        }
        i++;
      } else {
        uint32_t nx = ac_goto(ac, nd, lc(text[j], fold));
        if (!nx) { walking = false; continue; }
        nd = ac_fetch(ac, nx);
        if (nd.w0 >> 24) ac_outputs(db, ac, nd, text, tn, emit);
        if (++j >= tn) walking = false;
      }
    }
    return;
  }
  // reference formulation, one automaton event per iteration (transition, stay at root, or failure hop)
  uint32_t cur = 0;
  AcNode nd = root;
  uint32_t i = 0;
  uint8_t ch = lc(text[0], fold);
  while (i < tn) {
    uint32_t nx = (cur == 0 && acc.root_tab) ? acc.root_tab[ch] : ac_goto(ac, nd, ch);
    if (!nx && cur != 0) {  // failure hop: retry the same byte from the fallback state
      cur = nd.fail;
      nd = cur ? ac_fetch(ac, cur) : root;
      continue;
    }
    if (nx) { cur = nx; nd = ac_fetch(ac, cur); }
    if (nd.w0 >> 24) ac_outputs(db, ac, nd, text, tn, emit);
    i++;
    if (i < tn) ch = lc(text[i], fold);
  }
}
// The anchored formulation of find_all_visit, one START POSITION at a time (db.ac_anchored, no pure wildcards): every
// pattern reached through a literal occurrence that starts at text[i].  The union over all i is what find_all_visit
// emits (as a set); positions are independent, which is what the warp-cooperative exact kernel exploits.
template <typename F>
MGPU_HD void anchored_visit_at(const DbView& db, const uint8_t* text, uint32_t tn, uint32_t i, const uint32_t* gram2, F&& emit) {
  if (i + 3 > tn) return;  // every literal has at least 3 bytes
  const bool fold = db.match_mode == 1;
  const uint8_t* ac = db.pg + db.ac_start;
  const uint32_t g2 = ((uint32_t)lc(text[i], fold) << 8) | lc(text[i + 1], fold);
  if (!((gram2[g2 >> 5] >> (g2 & 31)) & 1u)) return;
  const uint32_t g3 = (g2 << 8) | lc(text[i + 2], fold);
  if (!((db.ac_gram3[g3 >> 5] >> (g3 & 31)) & 1u)) return;
  const uint64_t v = load_prefix8(text, i, fold);
  const uint32_t avail = tn - i;
  for (uint32_t lens = db.ac_short_lens; lens; lens &= lens - 1) {
#ifdef __CUDA_ARCH__
    uint32_t m = (uint32_t)__ffs((int)lens) - 1u;
#else
    uint32_t m = (uint32_t)__builtin_ctz(lens);
#endif
    if (m > avail) break;
    uint32_t off = prefix_node(db, low_bytes(v, m), m);
    if (off) { AcNode sn = ac_fetch(ac, off); ac_outputs(db, ac, sn, text, tn, emit); }
  }
  if (avail < 8) return;
  uint32_t off = prefix_node(db, v, 8);
  if (!off) return;
  AcNode nd = ac_fetch(ac, off);
  if (nd.w0 >> 24) ac_outputs(db, ac, nd, text, tn, emit);
  for (uint32_t j = i + 8; j < tn; j++) {
    uint32_t nx = ac_goto(ac, nd, lc(text[j], fold));
    if (!nx) return;
    nd = ac_fetch(ac, nx);
    if (nd.w0 >> 24) ac_outputs(db, ac, nd, text, tn, emit);
  }
}

// the root's dense table, if the root is a Dense state
MGPU_HD const uint32_t* ac_root_table(const DbView& db) {
  if (!db.has_glob || db.ac_size < 20) return nullptr;
  const uint8_t* ac = db.pg + db.ac_start;
  if ((ld32(ac) & 0xFF) != 3) return nullptr;
  return reinterpret_cast<const uint32_t*>(ac + ld32(ac + 12));
}
MGPU_HD bool glob_data_offset(const DbView& db, uint32_t pid, uint32_t& off) {
  if (pid >= db.glob_data_n) return false;
  off = db.glob_data[pid];
  return true;
}

}  // namespace mgpu
