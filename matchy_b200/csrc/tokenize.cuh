// tokenize.cuh — first half of the extractor: find every candidate token in a log chunk.
//
// Restates, as bit-parallel mask arithmetic, what the reference does with per-anchor byte loops
// (crates/matchy-extractor/src/lib.rs):
//   find_word_boundaries_into :1742-1782   words = maximal runs of non-boundary bytes (BOUNDARY_LOOKUP :1568-1593)
//   hashes   :1212-1250   word of length 32/40/64/96/128, all hex
//   domains  :537-628     maximal run of DOMAIN_CHAR_LOOKUP bytes that holds a '.', boundary (or chunk edge) both sides
//                         == a WORD made only of domain bytes with a '.' in it
//   IPv4     :1120-1179   maximal [0-9.] run with boundaries both sides == a word of digits and dots (subset of the above)
//   e-mail   :1182-1196   anchored at every '@'
//   IPv6     :1044-1116   anchored at every "::"
// The tokenizer emits CANDIDATES (exact word extents / anchor positions, conservative class filters); the
// validate kernel (device_fns.cuh) applies the remaining rules.  Layout: one warp owns a contiguous range of
// 1 KiB tiles and walks it front to back; lane L owns bytes [32L, 32L+32) of the tile and holds one bit per byte
// in 32-bit class masks.  A word belongs to the lane that holds the boundary byte FOLLOWING it, so all state
// flows forward: lane → lane by shuffle/ballot carry resolution, tile → tile in registers.
#pragma once
#include "device_fns.cuh"

namespace mgpu {

enum { CLS_B = 0, CLS_DOT = 1, CLS_AT = 2, CLS_CL = 3, CLS_NL = 4, CLS_DM = 5, CLS_HX = 6, CLS_DASH = 7 };
enum { Q_DOTTED = 0, Q_HASH = 1, Q_AT = 2, Q_COLON2 = 3, Q_NUMERIC = 4, Q_LONG = 5, Q_COUNT = 6 };  // Q_NUMERIC: dotted words of hex digits and dots only (IPv4 candidates)
static const uint32_t TILE_BYTES = 1024;
static const uint32_t SLICE_BYTES = 32;

MGPU_HD uint32_t class_bits(uint8_t b) {
  uint32_t c = 0;
  if (is_boundary(b)) c |= 1u << CLS_B;
  if (b == '.') c |= 1u << CLS_DOT;
  if (b == '@') c |= 1u << CLS_AT;
  if (b == ':') c |= 1u << CLS_CL;
  if (b == '\n') c |= 1u << CLS_NL;
  if (is_domain_fast(b)) c |= 1u << CLS_DM;
  if (is_hex(b)) c |= 1u << CLS_HX;
  if (b == '-') c |= 1u << CLS_DASH;
  return c;
}

// byte category as four bit planes in the four bytes of a word (bit 0 of byte c = plane c):
//   plane 3: boundary byte; among them planes 1,0 = 00 other, 01 '\n', 10 '@', 11 ':'
//   plane 2: domain character;       planes 1,0 = 00 other, 01 '.',  10 '-', 11 hex digit
MGPU_HD uint32_t category_planes(uint8_t b) {
  const uint32_t c = class_bits(b);
  uint32_t p3 = 0, p2 = 0, lo = 0;
  if (c & (1u << CLS_B)) { p3 = 1; lo = (c & (1u << CLS_NL)) ? 1u : (c & (1u << CLS_AT)) ? 2u : (c & (1u << CLS_CL)) ? 3u : 0u; }
  else if (c & (1u << CLS_DM)) { p2 = 1; lo = (c & (1u << CLS_DOT)) ? 1u : (c & (1u << CLS_DASH)) ? 2u : (c & (1u << CLS_HX)) ? 3u : 0u; }
  return (lo & 1u) | ((lo >> 1) << 8) | (p2 << 16) | (p3 << 24);
}
// the eight class bits back from the planes of ONE byte (what the kernel does per 32-byte mask with three LOP3 each)
MGPU_HD uint32_t class_bits_from_planes(uint32_t e) {
  const uint32_t P0 = e & 1u, P1 = (e >> 8) & 1u, P2 = (e >> 16) & 1u, P3 = (e >> 24) & 1u;
  return (P3 << CLS_B) | ((P2 & ~P1 & P0) << CLS_DOT) | ((P3 & P1 & ~P0 & 1u) << CLS_AT) | ((P3 & P1 & P0) << CLS_CL) | ((P3 & ~P1 & P0 & 1u) << CLS_NL) |
         (P2 << CLS_DM) | ((P2 & P1 & P0) << CLS_HX) | ((P2 & P1 & ~P0 & 1u) << CLS_DASH);
}

struct LaneMasks { uint32_t B, DOT, AT, CL, NL, DM, HX, DASH; };

// State carried into a tile (identical in all lanes of the warp).
struct TileCarry {
  uint32_t prev;        // facts about the bytes just before the tile: PV_* bits
  uint32_t cBad, cDot, cNhx;  // the open word so far holds a byte that rules out a domain / a '.' / a non-hex byte
  uint32_t cNhd;        // ... a byte that is neither a hex digit nor a '.' (such a word cannot be an IPv4 address)
  uint32_t prevB;       // boundary mask of the 32 bytes before the tile (bit i = byte i - 32 of the tile)
  uint64_t open_start;  // chunk offset where the open word starts (valid when prev & PV_T)
};
// bits of the "previous bytes" word that travels lane -> lane (one shuffle) and tile -> tile
enum { PV_T = 1, PV_DOT = 2, PV_DASH = 4, PV_CL1 = 8, PV_CL2 = 16 };  // byte -1 is a word byte / '.' / '-' / ':' ; byte -2 is ':'
MGPU_HD uint32_t prev_bits_of(const LaneMasks& m) {  // the same facts about a slice's own last bytes, for its successor
  return ((~m.B) >> 31) | ((m.DOT >> 31) << 1) | ((m.DASH >> 31) << 2) | ((m.CL >> 31) << 3) | (((m.CL >> 30) & 1u) << 4);
}
// Domain label rules as byte-adjacency facts (is_valid_domain / is_valid_label, lib.rs:637-689): no empty label ("..",
// leading or trailing '.'), no label starting or ending with '-'.  `bad` marks bytes that break a rule when looking back;
// `bad_end` marks boundary positions whose preceding byte may not end a domain.
// (prevDOT / prevDASH: bit i = byte i-1 is '.' / '-')
MGPU_HD void domain_rule_masks_prev(const LaneMasks& m, uint32_t S, uint32_t prevDOT, uint32_t prevDASH, uint32_t& bad, uint32_t& bad_end) {
  bad = (m.DOT & (prevDOT | prevDASH)) | (m.DASH & prevDOT) | (S & (m.DOT | m.DASH));
  bad_end = prevDOT | prevDASH;
}
MGPU_HD void domain_rule_masks(const LaneMasks& m, uint32_t S, uint32_t pv, uint32_t& bad, uint32_t& bad_end) {
  domain_rule_masks_prev(m, S, (m.DOT << 1) | ((pv >> 1) & 1u), (m.DASH << 1) | ((pv >> 2) & 1u), bad, bad_end);
}

// "Does the word that ends here contain a byte of Y?" for every word at once.  T = ~B marks word bytes, Y is a subset
// of T.  In T + Y a carry is born at the lowest Y bit of a word, ripples through the rest of the word (all ones in T)
// and lands on the boundary bit that follows it; boundaries stop it, so words do not influence each other.  Hence
// bit e of (T + Y + cin) & B  <=>  e is a boundary and the word ending at e-1 holds a Y byte (cin: the word that was open
// at the start of the slice already held one).  Across lanes the carries resolve like a 32-bit adder over the
// per-lane (generate, propagate) pairs: generate = carry out of T + Y, propagate = T is all ones (shared by all chains).
MGPU_HD uint32_t chain_gen(uint32_t T, uint32_t Y) { return (uint32_t)(((uint64_t)T + Y) >> 32); }
MGPU_HD uint32_t chain_ends(uint32_t T, uint32_t Y, uint32_t cin, uint32_t B) { return (T + Y + cin) & B; }
// Given per-lane generate/propagate ballots and the carry into lane 0, the carry into every lane (bit i) and
// out of lane 31: c[i+1] = g[i] | (p[i] & c[i]) is exactly the carry chain of (g|p) + g + c0.
MGPU_HD uint32_t carry_chain(uint32_t gen, uint32_t prop, uint32_t c0, uint32_t& cout) {
  uint64_t a = (uint64_t)(gen | prop), b = (uint64_t)gen;
  uint64_t x = a + b + c0;
  uint64_t c = x ^ a ^ b;
  cout = (uint32_t)(c >> 32) & 1u;
  return (uint32_t)c;
}

// Chunk offset where the word that is open at the START of lane `lane` begins: just after the last boundary of the
// lanes below (hasB = ballot of "my slice has a boundary", Bsrc = boundary mask of the highest such lane below, fetched by
// the caller with one shuffle from lane src_lane), or the tile's own open_start when there is none.
MGPU_HD uint32_t lane_below_with_boundary(uint32_t hasB, uint32_t lane) {  // 32 = none
  uint32_t lower = hasB & ((1u << lane) - 1u);
#ifdef __CUDA_ARCH__
  return lower ? 31u - (uint32_t)__clz((int)lower) : 32u;
#else
  return lower ? 31u - (uint32_t)__builtin_clz(lower) : 32u;
#endif
}
MGPU_HD uint32_t top_bit(uint32_t x) {  // x != 0
#ifdef __CUDA_ARCH__
  return 31u - (uint32_t)__clz((int)x);
#else
  return 31u - (uint32_t)__builtin_clz(x);
#endif
}
// start of the word that ends at boundary bit `bit` of a lane whose slice starts at chunk offset p
MGPU_HD uint64_t word_start_in_lane(uint32_t B, uint32_t bit, uint64_t p, uint64_t lane_open_start) {
  uint32_t below = bit ? (B & (0xFFFFFFFFu >> (32 - bit))) : 0u;
  return below ? p + top_bit(below) + 1 : lane_open_start;
}

// Q_LONG: words of 26..62 or 90..110 bytes (crypto-address candidates, lib.rs:1269-1409).
MGPU_HD bool is_crypto_len(uint64_t len) { return (len >= 26 && len <= 62) || (len >= 90 && len <= 110); }
// Boundaries of a 32-byte slice that end a word of AT LEAST 26 bytes.  T = word bytes of the slice, Tprev = of the 32 bytes
// before it, E = boundaries of the slice that end a word.  Runs of ones are grown by doubling: r_k has bit j set when bits
// j-k+1..j are all set.
MGPU_HD uint32_t long_word_ends(uint32_t T, uint32_t Tprev, uint32_t E) {
  const uint64_t X = ((uint64_t)T << 32) | Tprev;
  const uint64_t r2 = X & (X << 1), r4 = r2 & (r2 << 2), r8 = r4 & (r4 << 4), r16 = r8 & (r8 << 8), r24 = r16 & (r8 << 16), r26 = r24 & (r2 << 24);
  return E & (uint32_t)(r26 >> 31);  // boundary bit i needs the run to reach bit 32 + i - 1
}
MGPU_HD bool is_hash_len(uint64_t len) { return len == 32 || len == 40 || len == 64 || len == 96 || len == 128; }

// Carry state for a range that starts at offset `a` of a chunk whose first byte is at `lo`: look back over the
// word that is open there.
MGPU_HDN TileCarry range_prologue(const uint8_t* buf, uint64_t lo, uint64_t a) {
  TileCarry c;
  c.prev = 0; c.cBad = 0; c.cDot = 0; c.cNhx = 0; c.cNhd = 0; c.open_start = a; c.prevB = 0xFFFFFFFFu;
  if (a <= lo) return c;
  c.prevB = 0;
  for (uint32_t k = 0; k < 32; k++) if (a < lo + 32 - k || is_boundary(buf[a - 32 + k])) c.prevB |= 1u << k;  // bytes before the chunk count as boundaries
  uint8_t prev = buf[a - 1];
  c.prev = (prev == '.' ? (uint32_t)PV_DOT : 0u) | (prev == '-' ? (uint32_t)PV_DASH : 0u) | (prev == ':' ? (uint32_t)PV_CL1 : 0u) |
           ((a >= lo + 2 && buf[a - 2] == ':') ? (uint32_t)PV_CL2 : 0u);
  if (is_boundary(prev)) return c;
  c.prev |= PV_T;
  uint32_t bad = 0, dot = 0, nhx = 0, nhd = 0;
  uint64_t s = a;
  uint8_t right = 0;  // the byte to the right of b inside the open word (0 = none yet)
  while (s > lo) {
    uint8_t b = buf[s - 1];
    if (is_boundary(b)) break;
    bad |= !is_domain_fast(b); dot |= b == '.'; nhx |= !is_hex(b); nhd |= !is_hex(b) && b != '.';
    if (right == '.' && (b == '.' || b == '-')) bad = 1;  // "..", "-."
    if (right == '-' && b == '.') bad = 1;                // ".-"
    right = b;
    s--;
  }
  if (right == '.' || right == '-') bad = 1;  // the word starts with '.' or '-'
  c.cBad = bad; c.cDot = dot; c.cNhx = nhx; c.cNhd = nhd; c.open_start = s;
  return c;
}

}  // namespace mgpu
