// engine.cu — the B200 log-scan engine: CUDA kernels + host driver + C ABI (include/matchy_b200.h).
//
// Replaces the body of processing::Worker::process_bytes (crates/matchy/src/processing/mod.rs:353-448):
//   tokenize_kernel  (K1)  candidates: words / '@' / "::" anchors          extractor lib.rs:409-488 (first half)
//   token_kernel     (K2)  candidate -> typed token (IPv4/IPv6/domain/e-mail/hash)   (second half)
//                          + constant-time string filters that decide which tokens need the exact lookups at all
//   iptrie_kernel    (K3)  SearchTree::lookup per IP token + record emission      mmdb/tree.rs:46-125
//   lithash_kernel   (K4)  LiteralHash::lookup per string token                    matchy-literal-hash lib.rs:467-575
//   acglob_kernel    (K5)  Paraglob::find_all per string token + record emission   paraglob_offset.rs:1028-1182
// Records are appended with aggregated atomics (the "stream compaction" step) and sorted on the host.
// sm_100a only.  No CPU fallback: every entry point fails when there is no CUDA device.
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <cub/device/device_radix_sort.cuh>  // (the result sort below: library code, off the scan path)
#include "../../include/matchy_b200.h"
#include "db_prepare.h"
#include "mxy_reader.h"
#include "tokenize.cuh"
#include "crypto_addr.cuh"
#include "host_sort.h"

namespace cg = cooperative_groups;

namespace mgpu {

static_assert(sizeof(mgpu_match) == 32, "mgpu_match is copied as two 16-byte words");
struct Cand { uint32_t start, len; };
struct StrTok { uint32_t start, len, type; };
struct IpTok { uint32_t start, len, type, pad; uint32_t w[4]; };  // v4: w[0]; v6: w[k] = seg[2k] << 16 | seg[2k+1]

struct ScanTotals { uint32_t n_rec, n_ids; };  // records / id pairs appended so far to the shared result buffers (all pieces of a batch)

struct DevCounters {  // one per piece
  uint32_t n_str, n_ip;
  uint32_t n_rec, n_ids;      // snapshot of ScanTotals when the piece finished (piece_end_kernel)
  uint32_t overflow;          // bit q (0..5): candidate queue q, 8: str tokens, 9: ip tokens, 10: records, 11: ids
  uint32_t pad[3];
  unsigned long long lines;
  unsigned long long by_type[12];
  unsigned long long matches;
};

// Candidate queues are SEGMENTED: tokenizer warp w owns slots [w * seg_cap[q], (w+1) * seg_cap[q]) of queue q and
// writes how many it filled to seg_cnt[q * nseg_max + w].  No atomics, no padding, and every segment is sorted by log
// position, which is what lets the token kernel read tokens through one shared-memory window per warp.
struct ScanArgs {
  const uint8_t* buf;  // 16-byte aligned, readable up to round_up(n, 1024); the chunk is buf[lo .. n)
  uint64_t lo;         // 0..15: bytes before the chunk (alignment padding, treated like a chunk edge)
  uint64_t n;          // end of the chunk relative to buf
  uint64_t base;       // absolute log offset of buf[0]
  uint32_t flags;
  uint32_t fast;       // 1: the token kernel applies string_filters() and only flagged string tokens reach the exact kernel
  uint32_t lookups;    // 0: extraction only (mgpu_extract): tokens are produced, nothing is looked up
  uint32_t ip_skip;    // 1: the database has no IP entry (DbView::ip_empty): addresses are validated and counted, not listed, and iptrie_kernel is not launched
  DbView db;
  Cand* q_dotted; Cand* q_hash; uint32_t* q_at; uint32_t* q_c2; Cand* q_numeric; Cand* q_long;
  uint32_t seg_cap[Q_COUNT];
  uint32_t* seg_cnt;   // [Q_COUNT][nseg_max]
  uint32_t nseg, nseg_max;
  StrTok* str; uint32_t cap_str;
  StrTok* defer;       // per-warp queues (DEFER_CAP slots each) of tokens that passed string_gate and owe stage 2 (scan_kernel / tokenize path)
  StrTok* defer_tk;    // the same for token_kernel, which runs beside the next piece's scan_kernel
  IpTok* ip; uint32_t cap_ip;
  uint32_t* lh_res;
  mgpu_match* recs; uint32_t cap_rec;
  mgpu_id_pair* ids; uint32_t cap_ids;
  DevCounters* ctr;
  ScanTotals* tot;
  uint32_t tok_unit;   // slots a token-kernel warp takes from a token list per trip to its counter (TOK_RESERVE unless a test overrides it)
  unsigned long long* dbg;  // audit accumulators (mgpu_set_option "verify_tokens"), else nullptr
  uint32_t variant;         // experiment switch (mgpu_set_option "variant"), 0 in production
  uint32_t sub_block;       // token_kernel: bytes of a segment handled per dotted / numeric round (0: whole segment; must be 0 when the queues are not sorted)
};

static const int K1_THREADS = 512;
static const int K1_WARPS = K1_THREADS / 32;

// ---------------------------------------------------------------------------------------------------------
// K1 tokenize
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// inclusive prefix sum over the lanes; the shuffle's own "source lane in range" predicate guards the add (no compare per step)
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t /*lane*/) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    asm volatile("{ .reg .pred p; .reg .u32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t; }" : "+r"(v) : "r"(d));
  }
  return v;
}

// exclusive prefix over the lanes of a per-lane count < 16, and the warp total, from four ballots (no shuffle chain)
__device__ __forceinline__ uint32_t warp_excl_count4(uint32_t cnt, uint32_t lt_mask, uint32_t& total) {
  const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, cnt & 1u), b1 = __ballot_sync(0xFFFFFFFFu, cnt & 2u);
  const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, cnt & 4u), b3 = __ballot_sync(0xFFFFFFFFu, cnt & 8u);
  total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
  return __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask) + 8 * __popc(b3 & lt_mask);
}

// FIXED = 0: the extractor flags come from a.flags; FIXED != 0: they are that compile-time constant (the `matchy match` default,
// everything but the crypto-address words), which turns the per-tile flag tests and the branches behind them into nothing.
static const uint32_t K1_DEFAULT_FLAGS = MGPU_X_DOMAINS | MGPU_X_EMAILS | MGPU_X_IPV4 | MGPU_X_IPV6 | MGPU_X_HASHES;
template <uint32_t FIXED>
__global__ void __launch_bounds__(K1_THREADS) tokenize_kernel(ScanArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  // class table replicated per lane: entry (b, lane) at b*256 + lane*8 -> every LDS.64 of a warp is conflict-free
  uint2* lut = reinterpret_cast<uint2*>(smem);
  for (uint32_t i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
    uint32_t c = class_bits((uint8_t)(i >> 5));
    uint2 e;
    e.x = (c & 1u) | (((c >> 1) & 1u) << 8) | (((c >> 2) & 1u) << 16) | (((c >> 3) & 1u) << 24);
    e.y = ((c >> 4) & 1u) | (((c >> 5) & 1u) << 8) | (((c >> 6) & 1u) << 16) | (((c >> 7) & 1u) << 24);
    lut[i] = e;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t lane8 = lane * 8;
  const uint64_t nwarps = (uint64_t)gridDim.x * K1_WARPS;  // == a.nseg
  const uint64_t w = (uint64_t)blockIdx.x * K1_WARPS + warp;
  const uint64_t tiles = (a.n + TILE_BYTES - 1) / TILE_BYTES;
  const uint64_t tpw = (tiles + nwarps - 1) / nwarps;
  const uint64_t t0 = w * tpw;
  uint64_t t1 = t0 + tpw;
  if (t1 > tiles) t1 = tiles;
  // this warp's queue segments and fill counts (warp-uniform)
  Cand* const qd = a.q_dotted + w * a.seg_cap[Q_DOTTED];
  Cand* const qh = a.q_hash + w * a.seg_cap[Q_HASH];
  uint32_t* const qa = a.q_at + w * a.seg_cap[Q_AT];
  uint32_t* const qc = a.q_c2 + w * a.seg_cap[Q_COLON2];
  Cand* const qn = a.q_numeric + w * a.seg_cap[Q_NUMERIC];
  Cand* const ql = a.q_long + w * a.seg_cap[Q_LONG];
  uint32_t nd = 0, nh = 0, na = 0, nc = 0, nn = 0, nl = 0, ovf = 0;
  uint32_t lines = 0;
  if (t0 < t1) {
    TileCarry cy = range_prologue(a.buf, a.lo, t0 * TILE_BYTES);
    // masks of the 32 bytes before the tile, as far as the tile's first bytes can see them (bit 31 = byte -1, bit 30 = byte -2)
    uint32_t pDOT = (cy.prev & PV_DOT) ? 0x80000000u : 0u, pDASH = (cy.prev & PV_DASH) ? 0x80000000u : 0u;
    uint32_t pCL = ((cy.prev & PV_CL1) ? 0x80000000u : 0u) | ((cy.prev & PV_CL2) ? 0x40000000u : 0u);
    const uint32_t xflags = FIXED ? FIXED : a.flags;
    const bool want_dot = (xflags & (MGPU_X_IPV4 | MGPU_X_DOMAINS)) != 0;
    const bool want_hash = (xflags & MGPU_X_HASHES) != 0;
    const bool want_at = (xflags & MGPU_X_EMAILS) != 0;
    const bool want_c2 = (xflags & MGPU_X_IPV6) != 0;
    const bool want_long = (xflags & MGPU_X_CRYPTO) != 0;
    // the next tile's slice is loaded into registers while the current one is processed (the buffer is readable one tile
    // past the chunk, see mgpu_scan_device), so the loads' latency never sits in front of the classification
    uint4 nx0 = ld_stream(a.buf + (uint32_t)(t0 * TILE_BYTES) + lane * SLICE_BYTES), nx1 = ld_stream(a.buf + (uint32_t)(t0 * TILE_BYTES) + lane * SLICE_BYTES + 16);
    const uint32_t t_end = (uint32_t)t1, t_last = (uint32_t)(tiles - 1);  // (tile indices fit 32 bits: a chunk has at most 2^21 tiles)
    for (uint32_t t = (uint32_t)t0; t < t_end; t++) {
      const uint32_t tile_base = t * TILE_BYTES;  // chunks are at most 2 GiB: positions fit 32 bits
      const uint32_t p = tile_base + lane * SLICE_BYTES;
      const uint4 v0 = nx0, v1 = nx1;
      if (t + 1 < t_end) { nx0 = ld_stream(a.buf + p + TILE_BYTES); nx1 = ld_stream(a.buf + p + TILE_BYTES + 16); }
      if (t + 4 < t_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.buf + p + 4 * TILE_BYTES));  // my slice of the tile four steps ahead
      uint32_t wds[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      uint32_t accLo[4] = {0, 0, 0, 0}, accHi[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint32_t off = __byte_perm(wds[j], lane8, 0x5504u | ((uint32_t)k << 4));  // (byte k of the word) * 256 + lane * 8
          uint2 e = *reinterpret_cast<const uint2*>(smem + off);
          const uint32_t sh = (uint32_t)((j & 1) * 4 + k);
          accLo[j >> 1] += e.x << sh;
          accHi[j >> 1] += e.y << sh;
        }
      }
      // byte c of acc[q] = class c of bytes 8q..8q+7: a 4x4 byte transpose (two PRMT stages) turns them into one mask per class
      LaneMasks m;
      {
        const uint32_t r0 = __byte_perm(accLo[0], accLo[1], 0x5140), r1 = __byte_perm(accLo[0], accLo[1], 0x7362);
        const uint32_t r2 = __byte_perm(accLo[2], accLo[3], 0x5140), r3 = __byte_perm(accLo[2], accLo[3], 0x7362);
        m.B = __byte_perm(r0, r2, 0x5410); m.DOT = __byte_perm(r0, r2, 0x7632);
        m.AT = __byte_perm(r1, r3, 0x5410); m.CL = __byte_perm(r1, r3, 0x7632);
      }
      {
        const uint32_t r0 = __byte_perm(accHi[0], accHi[1], 0x5140), r1 = __byte_perm(accHi[0], accHi[1], 0x7362);
        const uint32_t r2 = __byte_perm(accHi[2], accHi[3], 0x5140), r3 = __byte_perm(accHi[2], accHi[3], 0x7362);
        m.NL = __byte_perm(r0, r2, 0x5410); m.DM = __byte_perm(r0, r2, 0x7632);
        m.HX = __byte_perm(r1, r3, 0x5410); m.DASH = __byte_perm(r1, r3, 0x7632);
      }
      if (t == 0 || t == t_last) {  // only the chunk's first (lo < 16) and last tile can hold bytes outside [lo, n): they behave like a boundary (chunk edge)
        uint64_t valid = a.n > p ? a.n - p : 0;
        uint32_t keep = valid >= 32 ? 0xFFFFFFFFu : ((1u << (uint32_t)valid) - 1u);
        if (p < a.lo) keep &= (a.lo - p >= 32) ? 0u : (0xFFFFFFFFu << (uint32_t)(a.lo - p));
        m.B |= ~keep; m.DOT &= keep; m.AT &= keep; m.CL &= keep; m.NL &= keep; m.DM &= keep; m.HX &= keep; m.DASH &= keep;
      }
      if ((int32_t)cy.prevB < 0) cy.open_start = tile_base;  // no word is open (the byte before the tile is a boundary): a word that fills the tile from its first byte starts here
      lines += __popc(m.NL);

      const uint32_t T = ~m.B;
      // "the byte before byte i is ..." masks: the previous lane's mask funnel-shifted into mine (lane 0 looks into the
      // previous tile) — one SHF per mask instead of packing five facts into a word on one side and unpacking them on the other
      uint32_t Bprev = __shfl_up_sync(0xFFFFFFFFu, m.B, 1), DOTp = __shfl_up_sync(0xFFFFFFFFu, m.DOT, 1);
      uint32_t DASHp = __shfl_up_sync(0xFFFFFFFFu, m.DASH, 1), CLp = __shfl_up_sync(0xFFFFFFFFu, m.CL, 1);
      if (lane == 0) { Bprev = cy.prevB; DOTp = pDOT; DASHp = pDASH; CLp = pCL; }
      const uint32_t prevT = ~__funnelshift_l(Bprev, m.B, 1);  // bit i: byte i-1 is a word byte
      const bool pT = (int32_t)Bprev >= 0;                      // the byte before my slice is a word byte
      const uint32_t S = T & ~prevT;
      uint32_t bad, bad_end;
      domain_rule_masks_prev(m, S, __funnelshift_l(DOTp, m.DOT, 1), __funnelshift_l(DASHp, m.DASH, 1), bad, bad_end);

      // four "does the word contain ..." chains: a byte that rules out a domain, a '.', a non-hex byte, a byte that is
      // neither a hex digit nor a '.'
      // (three chains: "the word holds a non-hex byte" is "it holds a '.' or a byte that is neither hex nor '.'")
      const uint32_t Y1 = T & (~m.DM | bad), Y2 = m.DOT, Y4 = T & ~m.HX & ~m.DOT;
      const uint32_t g1 = chain_gen(T, Y1), g2 = chain_gen(T, Y2), g4 = chain_gen(T, Y4);
      const uint32_t pb = __ballot_sync(0xFFFFFFFFu, T == 0xFFFFFFFFu);
      const uint32_t G1 = __ballot_sync(0xFFFFFFFFu, g1), G2 = __ballot_sync(0xFFFFFFFFu, g2), G4 = __ballot_sync(0xFFFFFFFFu, g4);
      uint32_t co1, co2, co4;
      const uint32_t cv1 = carry_chain(G1, pb & ~G1, cy.cBad, co1), cv2 = carry_chain(G2, pb & ~G2, cy.cDot, co2), cv4 = carry_chain(G4, pb & ~G4, cy.cNhd, co4);
      cy.cBad = co1; cy.cDot = co2; cy.cNhd = co4;
      const uint32_t E = m.B & prevT;  // boundaries that end a word
      const uint32_t hasBad = chain_ends(T, Y1, (cv1 >> lane) & 1u, m.B), hasDot = chain_ends(T, Y2, (cv2 >> lane) & 1u, m.B);
      const uint32_t hasNhd = chain_ends(T, Y4, (cv4 >> lane) & 1u, m.B), hasNhx = hasNhd | hasDot;
      const uint32_t candDot = want_dot ? (hasDot & ~hasBad & ~bad_end) : 0u;
      // a word of >= 32 bytes that ends in my slice started in an earlier one: only my first boundary qualifies
      uint32_t candHex = (want_hash && pT) ? (E & ~hasNhx & (m.B & (0u - m.B))) : 0u;
      // ... and at least 32 bytes long: the previous slice has no boundary at or above the bit position where the word ends here
      if (candHex && (Bprev >> (__ffs(candHex) - 1)) != 0) candHex = 0;
      const uint32_t candLong = want_long ? long_word_ends(T, ~Bprev, E) : 0u;  // words of >= 26 bytes (crypto-address candidates)
      const uint32_t candAt = want_at ? m.AT : 0u;
      // second colon of the FIRST "::" of a colon run: ':' at i and i-1, not at i-2
      const uint32_t cl1 = __funnelshift_l(CLp, m.CL, 1), cl2 = __funnelshift_l(CLp, m.CL, 2);
      const uint32_t candC2 = want_c2 ? (m.CL & cl1 & ~cl2) : 0u;

      // where the word that is open at the start of my slice begins
      const uint32_t hasB = __ballot_sync(0xFFFFFFFFu, m.B != 0);
      const uint32_t src = lane_below_with_boundary(hasB, lane);
      const uint32_t Bsrc = __shfl_sync(0xFFFFFFFFu, m.B, src & 31u);
      const uint32_t lane_open = src < 32u ? tile_base + src * 32 + top_bit(Bsrc) + 1 : (uint32_t)cy.open_start;

      // ---- emission ----
      {
        // dotted words of hex digits and dots only go to the numeric queue (IPv4 candidates), the rest to the dotted queue;
        // one scan serves both (counts packed in the two halves of a word)
        const uint32_t candNum = candDot & ~hasNhd;
        const uint32_t cnt_n = __popc(candNum), cnt = cnt_n | ((__popc(candDot) - cnt_n) << 16), incl = warp_incl_scan(cnt, lane);
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31), excl = incl - cnt;
        if (total) {
          const uint32_t tn = total & 0xFFFFu, td = total >> 16;
          const bool okn = nn + tn <= a.seg_cap[Q_NUMERIC], okd = nd + td <= a.seg_cap[Q_DOTTED];
          Cand* dn = qn + nn + (excl & 0xFFFFu);
          Cand* dd = qd + nd + (excl >> 16);
          if (okn && okd) {  // (the usual case: no capacity test per store)
            for (uint32_t mm = candDot; mm; mm &= mm - 1) {
              const uint32_t low = mm & (0u - mm), bit = __ffs(mm) - 1;
              const uint32_t below = m.B & (low - 1u);
              const uint32_t s = below ? p + top_bit(below) + 1 : lane_open;
              const Cand c{s, p + bit - s};
              if (candNum & low) *dn++ = c; else *dd++ = c;
            }
          } else {
            for (uint32_t mm = candDot; mm; mm &= mm - 1) {
              const uint32_t bit = __ffs(mm) - 1;
              const uint32_t below = m.B & ~(0xFFFFFFFFu << bit);
              const uint32_t s = below ? p + top_bit(below) + 1 : lane_open;
              const Cand c{s, p + bit - s};
              if ((candNum >> bit) & 1u) { if (okn) *dn++ = c; }
              else if (okd) *dd++ = c;
            }
            if (!okn) ovf |= 1u << Q_NUMERIC;
            if (!okd) ovf |= 1u << Q_DOTTED;
          }
          nn += tn; nd += td;
        }
      }
      if (__any_sync(0xFFFFFFFFu, (candHex | candAt | candC2 | candLong) != 0)) {  // rare in most logs: one vote covers the four
        {
          bool keep = false; Cand c{0, 0};
          if (candHex) {
            const uint32_t bit = __ffs(candHex) - 1;
            const uint32_t below = m.B & ~(0xFFFFFFFFu << bit);
            const uint32_t s = below ? p + top_bit(below) + 1 : lane_open;
            keep = is_hash_len(p + bit - s);
            c.start = s; c.len = p + bit - s;
          }
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
          if (bal) {
            const uint32_t total = __popc(bal);
            if (nh + total <= a.seg_cap[Q_HASH]) { if (keep) qh[nh + __popc(bal & lt_mask)] = c; }
            else ovf |= 1u << Q_HASH;
            nh += total;
          }
        }
        if (__any_sync(0xFFFFFFFFu, candAt != 0)) {
          // up to 32 '@' per slice: prefix by shuffle scan
          const uint32_t cnt = __popc(candAt), incl = warp_incl_scan(cnt, lane);
          const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          if (na + total <= a.seg_cap[Q_AT]) {
            uint32_t idx = na + incl - cnt;
            for (uint32_t mm = candAt; mm; mm &= mm - 1) qa[idx++] = (uint32_t)(p + __ffs(mm) - 1);
          } else ovf |= 1u << Q_AT;
          na += total;
        }
        if (__any_sync(0xFFFFFFFFu, candLong != 0)) {
          // at most two per slice; keep those whose exact length is in one of the address ranges
          Cand c2[2]; uint32_t k = 0;
          for (uint32_t mm = candLong; mm; mm &= mm - 1) {
            const uint32_t bit = __ffs(mm) - 1;
            const uint32_t below = m.B & ~(0xFFFFFFFFu << bit);
            const uint32_t s = below ? p + top_bit(below) + 1 : lane_open;
            if (is_crypto_len(p + bit - s) && k < 2) c2[k++] = Cand{s, p + bit - s};
          }
          const uint32_t incl = warp_incl_scan(k, lane), total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          if (total) {
            if (nl + total <= a.seg_cap[Q_LONG]) { for (uint32_t j = 0; j < k; j++) ql[nl + incl - k + j] = c2[j]; }
            else ovf |= 1u << Q_LONG;
            nl += total;
          }
        }
        if (__any_sync(0xFFFFFFFFu, candC2 != 0)) {
          const uint32_t cnt = __popc(candC2), incl = warp_incl_scan(cnt, lane);
          const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          if (nc + total <= a.seg_cap[Q_COLON2]) {
            uint32_t idx = nc + incl - cnt;
            for (uint32_t mm = candC2; mm; mm &= mm - 1) qc[idx++] = (uint32_t)(p + __ffs(mm) - 2);
          } else ovf |= 1u << Q_COLON2;
          nc += total;
        }
      }

      // ---- carry into the next tile ----
      if (hasB) {
        const uint32_t ll = top_bit(hasB);
        const uint32_t Bl = __shfl_sync(0xFFFFFFFFu, m.B, ll);
        cy.open_start = tile_base + ll * 32 + top_bit(Bl) + 1;
      }
      cy.prevB = __shfl_sync(0xFFFFFFFFu, m.B, 31);
      pDOT = __shfl_sync(0xFFFFFFFFu, m.DOT, 31); pDASH = __shfl_sync(0xFFFFFFFFu, m.DASH, 31); pCL = __shfl_sync(0xFFFFFFFFu, m.CL, 31);
    }
  }
  if (lane == 0) {
    uint32_t* sc = a.seg_cnt + w;
    sc[Q_DOTTED * a.nseg_max] = ovf & (1u << Q_DOTTED) ? 0u : nd;
    sc[Q_HASH * a.nseg_max] = ovf & (1u << Q_HASH) ? 0u : nh;
    sc[Q_AT * a.nseg_max] = ovf & (1u << Q_AT) ? 0u : na;
    sc[Q_COLON2 * a.nseg_max] = ovf & (1u << Q_COLON2) ? 0u : nc;
    sc[Q_NUMERIC * a.nseg_max] = ovf & (1u << Q_NUMERIC) ? 0u : nn;
    sc[Q_LONG * a.nseg_max] = ovf & (1u << Q_LONG) ? 0u : nl;
    if (ovf) atomicOr(&a.ctr->overflow, ovf);
  }
  for (int d = 16; d; d >>= 1) lines += __shfl_down_sync(0xFFFFFFFFu, lines, d);
  if (lane == 0 && lines) atomicAdd(&a.ctr->lines, (unsigned long long)lines);
}

// ---------------------------------------------------------------------------------------------------------
// Token bytes (generic lookup kernels): a warp handles 32 consecutive tokens, which lie close together in the log.
// The warp copies the covering window into shared memory with coalesced 16-byte loads, and every lane then reads
// its token from there instead of issuing 32 scattered sector requests per byte.
// ---------------------------------------------------------------------------------------------------------
static const uint32_t WIN_BYTES = 4096;
static const int KT_THREADS = 256;  // lithash / acglob block size
static const int KT_WARPS = KT_THREADS / 32;

// Returns p with p[pos] readable for pos in [lo, hi + 16).  Falls back to the global buffer for wide windows.
__device__ __forceinline__ const uint8_t* stage_window(const uint8_t* buf, uint8_t* win, uint32_t lo, uint32_t hi, uint32_t lane,
                                                       uint32_t win_bytes = WIN_BYTES) {
  if (lo >= hi) return buf;
  const uint32_t alo = lo & ~15u, ahi = (hi + 31u) & ~15u;
  if (ahi - alo > win_bytes) return buf;
  __syncwarp();
  for (uint32_t o = alo + lane * 16; o < ahi; o += 512) *reinterpret_cast<uint4*>(win + (o - alo)) = *reinterpret_cast<const uint4*>(buf + o);
  __syncwarp();
  return win - alo;
}

// ---------------------------------------------------------------------------------------------------------
// K2 = the token kernel: candidate -> typed token (second half of the extractor), then — on the fast string path —
// the constant-time filters that decide which string tokens need the exact lookups at all.  One thread per candidate;
// token-kernel warp w reads tokenizer segment w.  Shared memory holds the PSL last-label table, the hot string filter and
// one 2 KiB log window per warp.  Tokens are appended through per-warp reservations of TOK_RESERVE slots (unused slots
// are marked invalid), so a warp's tokens stay in log order for the generic lookup kernels.
// ---------------------------------------------------------------------------------------------------------
static const uint32_t TOK_RESERVE = 256;  // (config 3 at 8 GB: 578 GB/s with 128, 619 with 256, 618 with 512 — the wait for the counter's atomic was 10 % of the token kernel's stalls)
static const uint32_t TOK_INVALID = 0xFFu;  // StrTok.type / IpTok.type of a padding slot
static const uint32_t TOK_RAW4 = 0xFEu;     // IpTok.type: a numeric-queue word of 7..15 bytes, not parsed yet; w[] = its first 16 bytes (iptrie_kernel parses it)
struct QueueCursor { uint32_t base, left; };

__device__ __forceinline__ uint32_t tok_reserve(uint32_t* counter, uint32_t cap, uint32_t need, uint32_t lane, QueueCursor& c,
                                                 StrTok* str, IpTok* ip, uint32_t* overflow, uint32_t ovf_bit, uint32_t unit) {
  if (need <= c.left) { uint32_t b = c.base; c.base += need; c.left -= need; return b; }
  for (uint32_t i = lane; i < c.left; i += 32) { if (str) str[c.base + i].type = TOK_INVALID; else ip[c.base + i].type = TOK_INVALID; }
  uint32_t sz = need > unit ? need : unit;
  uint32_t b = 0;
  if (lane == 0) b = atomicAdd(counter, sz);
  b = __shfl_sync(0xFFFFFFFFu, b, 0);
  if ((uint64_t)b + sz > cap) {
    if (lane == 0) atomicOr(overflow, ovf_bit);
    for (uint32_t i = b + lane; i < cap; i += 32) { if (str) str[i].type = TOK_INVALID; else ip[i].type = TOK_INVALID; }
    c.base = 0; c.left = 0;
    return NONE32;
  }
  c.base = b + need; c.left = sz - need;
  return b;
}

static const int TK_THREADS = 1024;  // one persistent block per SM
static const int TK_WARPS = TK_THREADS / 32;
static const uint32_t TK_WIN = 2048;
// hot filter + windows = 193 KiB: stays inside the 196 KiB shared-memory carve-out, which leaves 60 KiB of L1 for the PSL
// last-label table, the cold filter and gen_gram2 (the next carve-out step, 228 KiB, would leave 28 KiB and thrash them)
static const size_t TOKEN_SMEM = (size_t)HOT_WORDS * 4 + (size_t)TK_WARPS * (TK_WIN + 32);

static const uint32_t DEFER_CAP = 64;  // a warp drains its queue as soon as it holds 32 tokens, so 63 is the most it ever holds
struct TokenWarp {  // per-warp state of the token kernel
  QueueCursor cs, ci;
  StrTok* defer; uint32_t n_def;
  uint32_t n_dom, n_mail, n_v4, n_v6, n_md5, n_sha1, n_sha256, n_sha384, n_sha512;
};

__device__ __forceinline__ uint32_t hash_type_of(uint32_t len) {
  return len == 32 ? MGPU_T_MD5 : len == 40 ? MGPU_T_SHA1 : len == 64 ? MGPU_T_SHA256 : len == 96 ? MGPU_T_SHA384 : MGPU_T_SHA512;
}

// Append this iteration's tokens (ws: my lane has a string token st, wi: an IP token it).
__device__ __forceinline__ void append_tokens(const ScanArgs& a, TokenWarp& tw, uint32_t lane, bool ws, const StrTok& st, bool wi, const IpTok& it) {
  const uint32_t bs = __ballot_sync(0xFFFFFFFFu, ws), bi = __ballot_sync(0xFFFFFFFFu, wi);
  if (bs && a.fast) {  // fast path: flagged tokens are rare and their order is irrelevant -> one aggregated atomic, no padding
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(&a.ctr->n_str, (uint32_t)__popc(bs));
    b = __shfl_sync(0xFFFFFFFFu, b, 0);
    if ((uint64_t)b + __popc(bs) > a.cap_str) { if (lane == 0) atomicOr(&a.ctr->overflow, 1u << 8); }
    else if (ws) a.str[b + __popc(bs & ((1u << lane) - 1u))] = st;
  } else if (bs) {
    uint32_t b = tok_reserve(&a.ctr->n_str, a.cap_str, __popc(bs), lane, tw.cs, a.str, nullptr, &a.ctr->overflow, 1u << 8, a.tok_unit);
    if (ws && b != NONE32) a.str[b + __popc(bs & ((1u << lane) - 1u))] = st;
  }
  if (bi && !a.ip_skip) {
    uint32_t b = tok_reserve(&a.ctr->n_ip, a.cap_ip, __popc(bi), lane, tw.ci, nullptr, a.ip, &a.ctr->overflow, 1u << 9, a.tok_unit);
    if (wi && b != NONE32) a.ip[b + __popc(bi & ((1u << lane) - 1u))] = it;
  }
}

// A string token passed validation: count it, then (fast path) let the filters decide whether anything can match it.
// Returns "append it now"; `gate` != 0 means the token passed stage 1 of the filters and owes stage 2 (defer_push).
template <typename B>
__device__ __forceinline__ bool string_token(const ScanArgs& a, TokenWarp& tw, bool fast, const HotShared& s_hot, const KeyWords& kw, const B& bytes, StrTok& st,
                                             uint32_t& gate) {
  tw.n_dom += st.type == MGPU_T_DOMAIN; tw.n_mail += st.type == MGPU_T_EMAIL; tw.n_md5 += st.type == MGPU_T_MD5; tw.n_sha1 += st.type == MGPU_T_SHA1;
  tw.n_sha256 += st.type == MGPU_T_SHA256; tw.n_sha384 += st.type == MGPU_T_SHA384; tw.n_sha512 += st.type == MGPU_T_SHA512;
  gate = 0;
  if (!fast) return true;
  const uint32_t g = string_gate(a.db, s_hot, kw, st.len);
  if (g & a.db.gate_inline) {  // a class whose gate rejects nothing (not in the hot filter / unanchored globs): decide here, as one stage
    const uint32_t f = string_filters_full(a.db, s_hot, kw, bytes, st.len, g);
    st.type |= f;
    return f != 0;
  }
  gate = g;
  return false;
}

// Stage 2 of the string filters for `count` (<= 32) queued tokens, one per lane, read back from the log buffer.
template <typename H>
__device__ __forceinline__ void defer_drain(const ScanArgs& a, TokenWarp& tw, uint32_t lane, const StrTok* src, uint32_t count, const H& s_hot) {
  const bool have = lane < count;
  StrTok t{0, 0, 0};
  if (have) { t.start = __ldcg(&src[lane].start); t.len = __ldcg(&src[lane].len); t.type = __ldcg(&src[lane].type); }
  const uint32_t g = t.type >> 16;
  t.type &= 0xFFFFu;
  bool ws = false;
  if (have) {
    KeyWords kw;
    const uint8_t* w = a.buf + t.start;
    load_head_words(w, kw.h);
    load_tail_words(w, t.len, kw.t);
    const uint32_t f = string_filters_full(a.db, s_hot, kw, BytesPtr{w}, t.len, g);
    t.type |= f;
    ws = f != 0;
  }
  __syncwarp();
  append_tokens(a, tw, lane, ws, t, false, IpTok{0, 0, 0, 0, {0, 0, 0, 0}});
}

// Queue the lanes' gate-passing tokens; run stage 2 with a full warp once 32 are waiting (flush: with whatever is waiting).
// (One call of defer_drain in the code: the filters are large, and the fused kernel's loop has to stay in the instruction cache.)
template <typename H>
__device__ __forceinline__ void defer_push(const ScanArgs& a, TokenWarp& tw, uint32_t lane, uint32_t gate, const StrTok& st, const H& s_hot, bool flush = false) {
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, gate != 0);
  if (!bal && !flush) return;
  if (gate) {
    StrTok* d = tw.defer + tw.n_def + __popc(bal & ((1u << lane) - 1u));
    __stcg(&d->start, st.start); __stcg(&d->len, st.len); __stcg(&d->type, (st.type & 0xFFFFu) | (gate << 16));
  }
  tw.n_def += __popc(bal);
  __syncwarp();
  while (tw.n_def >= 32 || (flush && tw.n_def)) {
    const uint32_t cnt = tw.n_def >= 32 ? 32u : tw.n_def;
    tw.n_def -= cnt;
    defer_drain(a, tw, lane, tw.defer + tw.n_def, cnt, s_hot);
  }
}

// streaming 16-byte load that does not allocate in L1 (L1 is left to the PSL last-label table and the filter words)
__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ Cand ld_cand(const Cand* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return Cand{r.x, r.y};
}

// One segment of a word queue.  W_DOTTED: dotted domain-character words that hold a byte which is neither a hex digit nor a
// '.' (domains; never an IPv4 address).  W_NUMERIC: dotted words of hex digits and dots only (IPv4 addresses, and the odd
// domain such as "cafe.de").  W_HASH: hash-length hex words.  The segment is sorted by position; every iteration takes as
// many consecutive candidates (at most 32) as fit in the warp's window, stages the window with coalesced 16-byte loads and
// lets one lane handle one candidate.
enum { W_DOTTED = 0, W_NUMERIC = 1, W_HASH = 2 };
// Resumable: starts at candidate i0 and stops in front of the first candidate whose start is >= limit (the segment being sorted,
// that is a prefix); returns the index it stopped at.  token_kernel alternates the dotted and the numeric queue in sub-blocks of
// a warp's range, so that the numeric pass finds its bytes still in L2.
template <int MODE>
__device__ __forceinline__ uint32_t token_words(const ScanArgs& a, TokenWarp& tw, const Cand* q, uint32_t i0, uint32_t n, uint32_t limit, uint8_t* s_win,
                                                const HotShared& s_hot, bool fast, uint32_t lane) {
  const bool want_dom = (a.flags & MGPU_X_DOMAINS) != 0, want_v4 = (a.flags & MGPU_X_IPV4) != 0;
  if (MODE == W_DOTTED && !want_dom) return n;
  // IPv4 candidates are sparse in most logs (a 2 KiB window would hold a handful): 32 per iteration straight from the log
  // buffer — all that is needed of each is its first 16 bytes.  (Choosing per segment at run time cost 4 % on every config.)
  constexpr bool direct = MODE == W_NUMERIC;
  Cand cn{0xFFFFFFFFu, 0};  // prefetched: candidate i0 + lane of the NEXT iteration (assuming a full group of 32)
  if (i0 + lane < n) cn = ld_cand(q + i0 + lane);
  while (i0 < n) {
    const uint32_t idx = i0 + lane;
    const Cand c = cn;
    const bool have = idx < n && c.start < limit;
    const uint32_t hv = __ballot_sync(0xFFFFFFFFu, have);
    if (hv == 0) break;  // the next candidate belongs to a later sub-block
    if (idx + 32 < n) cn = ld_cand(q + idx + 32); else cn = Cand{0xFFFFFFFFu, 0};
    // (lowest start of the 32: the tokenizer's segments are sorted by position, the fused kernel's slow-path entries are not)
    const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, have ? c.start : 0xFFFFFFFFu);
    const uint32_t alo = lo & ~15u;
    const bool fits = !have || (uint64_t)c.start + c.len + 16 <= (uint64_t)alo + TK_WIN;
    const uint32_t nf = (direct ? 0u : __ballot_sync(0xFFFFFFFFu, !fits)) | ~hv;  // (lanes past the limit or the end do not "fit")
    uint32_t g = nf ? (uint32_t)__ffs((int)nf) - 1u : 32u;  // candidates of this iteration: lanes [0, g)
    const uint8_t* p = a.buf;
    bool high = true;  // "the token bytes may hold bytes >= 0x80"
    if (direct) {
    } else if (g == 0) g = 1;  // a single word longer than the window: read it from global memory
    else {
      const bool mine = have && lane < g;
      const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, mine ? c.start + c.len : 0u);
      const uint32_t ahi = (hi + 31u) & ~15u;
      __syncwarp();
      // all (up to four) 16-byte loads of a lane are issued before the first store, so their latencies overlap
      uint4 v[4];
      const uint32_t o0 = alo + lane * 16;
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = (o0 + j * 512 < ahi) ? ld_stream16(a.buf + o0 + j * 512) : make_uint4(0, 0, 0, 0);
      // the next iteration's window starts about where this one ends: pull it towards L2 meanwhile (64-byte line per lane)
      if ((uint64_t)ahi + lane * 64 < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.buf + ahi + lane * 64));
      uint32_t hb = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (o0 + j * 512 < ahi) *reinterpret_cast<uint4*>(s_win + (o0 - alo) + j * 512) = v[j];
        hb |= v[j].x | v[j].y | v[j].z | v[j].w;
      }
      high = __any_sync(0xFFFFFFFFu, (hb & 0x80808080u) != 0);  // (also orders the window writes before the reads below)
      p = s_win - alo;
    }
    const bool active = have && lane < g && (uint64_t)c.start + c.len <= a.n;
    bool ws = false;
    StrTok st{c.start, c.len, 0};
    KeyWords kw;
    const uint8_t* wp = p + (active ? c.start : lo);
    bool wi = false;
    IpTok it{c.start, c.len, MGPU_T_IPV4, 0, {0, 0, 0, 0}};
    if (MODE == W_NUMERIC) {
      // straight from the log buffer: the first 16 bytes decide IPv4; the last 16 are only needed by the few words that
      // are not an address (domain rules), and for words of at most 16 bytes they are a shift of the first 16
      load_head_words(wp, kw.h);
      if (active && want_v4) wi = parse_ipv4_words(kw.h, c.len, it.w[0]);
      __syncwarp();
      if (wi) tw.n_v4++;
      kw.t[0] = kw.t[1] = kw.t[2] = kw.t[3] = 0u;
      const bool need_tail = active && want_dom && !wi;
      if (__any_sync(0xFFFFFFFFu, need_tail)) {
        if (need_tail) {
          if (c.len <= 16) tail_words_from_head(kw.h, c.len, kw.t);
          else load_tail_words(wp, c.len, kw.t);
        }
        __syncwarp();
      }
    } else if (p != a.buf) {  // (warp-uniform) the token lies in the shared-memory window: ld.shared instead of generic loads
      const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_win) + ((active ? c.start : lo) - alo);
      load_head_words_shared(saddr, kw.h);
      load_tail_words_shared(saddr, active ? c.len : 0u, kw.t);
    } else {
      load_head_words(wp, kw.h);
      load_tail_words(wp, active ? c.len : 0u, kw.t);
    }
    if (MODE != W_HASH) {
      // a valid IPv4 address is never a domain: its last label is numeric, and no PSL entry ends in one (checked at upload)
      if (active && want_dom && !wi) {
        st.type = MGPU_T_DOMAIN;
        const uint64_t tail8 = c.len >= 8 ? (((uint64_t)kw.t[3] << 32) | kw.t[2]) : ((((uint64_t)kw.h[1] << 32) | kw.h[0]) << (8 * (8 - c.len)));  // == load_tail8(wp, len), from registers
        ws = domain_word_fast(a.db, a.db.psl_tld, wp, c.len, high, tail8);
      }
      __syncwarp();
    } else if (active) { st.type = hash_type_of(c.len); ws = true; }
    uint32_t gate = 0;
    if (__any_sync(0xFFFFFFFFu, ws)) {
      if (ws) ws = string_token(a, tw, fast, s_hot, kw, BytesPtr{wp}, st, gate);  // (one call site: the filters are the bulk of the kernel's code)
      __syncwarp();
    }
    append_tokens(a, tw, lane, ws, st, wi, it);
    defer_push(a, tw, lane, gate, st, s_hot);
    if (g != 32) {  // a short group: the prefetch assumed 32; fetch the right candidates again
      cn = Cand{0xFFFFFFFFu, 0};
      if (i0 + g + lane < n) cn = ld_cand(q + i0 + g + lane);
    }
    i0 += g;
  }
  return i0;
}

__global__ void __launch_bounds__(TK_THREADS, 1) token_kernel(ScanArgs a) {
  extern __shared__ __align__(16) uint8_t tk_smem[];
  uint32_t* s_hot_words = reinterpret_cast<uint32_t*>(tk_smem);
  const HotShared s_hot{(uint32_t)__cvta_generic_to_shared(tk_smem), a.db.gen_gram2};
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* s_win = tk_smem + (size_t)HOT_WORDS * 4 + (size_t)warp * (TK_WIN + 32);
  const bool fast = a.fast != 0;
  // an earlier stage of this piece ran out of room: the host redoes the piece in smaller parts.  (One thread reads the flag for
  // the block: other blocks may set it while this one starts, and the exit has to be uniform in front of the barrier below.)
  __shared__ uint32_t s_ovf;
  if (threadIdx.x == 0) s_ovf = a.ctr->overflow;
  __syncthreads();
  if (s_ovf) return;
  if (fast) {
    const uint4* src = reinterpret_cast<const uint4*>(a.db.hot);
    for (uint32_t i = threadIdx.x; i < HOT_WORDS / 4; i += blockDim.x) reinterpret_cast<uint4*>(s_hot_words)[i] = src[i];

  }
  __syncthreads();
  TokenWarp tw;
  memset(&tw, 0, sizeof tw);
  tw.defer = a.defer_tk + (size_t)(blockIdx.x * TK_WARPS + warp) * DEFER_CAP;
  const uint32_t nwarps = gridDim.x * TK_WARPS;
  for (uint32_t seg = blockIdx.x * TK_WARPS + warp; seg < a.nseg; seg += nwarps) {
    const uint32_t* sc = a.seg_cnt + seg;
    // Dotted and numeric words can alternate in sub-blocks of a.sub_block bytes of the segment: the numeric pass reads its words
    // straight from the log buffer, and after a dotted pass over the WHOLE segment (all warps together: the whole piece) those
    // bytes have left L2 again — the top stall of the kernel and one DRAM byte per log byte.  Built, measured, and off by
    // default (sub_block = 0: one round): see mgpu_ctx::sub_block.
    {
      const Cand* qD = a.q_dotted + (size_t)seg * a.seg_cap[Q_DOTTED];
      const Cand* qN = a.q_numeric + (size_t)seg * a.seg_cap[Q_NUMERIC];
      const uint32_t nD = (a.flags & MGPU_X_DOMAINS) ? sc[Q_DOTTED * a.nseg_max] : 0u, nN = sc[Q_NUMERIC * a.nseg_max];
      uint32_t iD = 0, iN = 0;
      while (iD < nD || iN < nN) {
        uint32_t limit = 0xFFFFFFFFu;
        if (a.sub_block) {
          const uint32_t sD = iD < nD ? ld_cand(qD + iD).start : 0xFFFFFFFFu, sN = iN < nN ? ld_cand(qN + iN).start : 0xFFFFFFFFu;
          limit = min(sD, sN) + a.sub_block;  // (positions inside a piece are < 2^31)
        }
        iD = token_words<W_DOTTED>(a, tw, qD, iD, nD, limit, s_win, s_hot, fast, lane);
        iN = token_words<W_NUMERIC>(a, tw, qN, iN, nN, limit, s_win, s_hot, fast, lane);
      }
    }
    token_words<W_HASH>(a, tw, a.q_hash + (size_t)seg * a.seg_cap[Q_HASH], 0u, sc[Q_HASH * a.nseg_max], 0xFFFFFFFFu, s_win, s_hot, fast, lane);
    // '@' and "::" anchors: rare, straight from the log buffer
    const uint32_t nA = sc[Q_AT * a.nseg_max], nC = sc[Q_COLON2 * a.nseg_max];
    const uint32_t* qa = a.q_at + (size_t)seg * a.seg_cap[Q_AT];
    const uint32_t* qc = a.q_c2 + (size_t)seg * a.seg_cap[Q_COLON2];
    for (uint32_t i0 = 0; i0 < nA + nC; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool ws = false, wi = false;
      StrTok st{0, 0, 0};
      IpTok it{0, 0, 0, 0, {0, 0, 0, 0}};
      if (i < nA) {
        const uint32_t at = qa[i];
        size_t s, e;
        if (at < a.n && email_at(a.db, a.buf, (size_t)a.lo, (size_t)a.n, at, s, e)) { ws = true; st.start = (uint32_t)s; st.len = (uint32_t)(e - s); st.type = MGPU_T_EMAIL; }
      } else if (i < nA + nC) {
        const uint32_t at = qc[i - nA];
        size_t s, e;
        if ((uint64_t)at + 2 <= a.n && ipv6_at_words(a.buf, (size_t)a.lo, (size_t)a.n, at, s, e, it.w)) {
          wi = true; it.start = (uint32_t)s; it.len = (uint32_t)(e - s); it.type = MGPU_T_IPV6;
          tw.n_v6++;
        }
      }
      __syncwarp();
      uint32_t gate = 0;
      if (ws) {
        KeyWords kw;
        load_head_words(a.buf + st.start, kw.h);
        load_tail_words(a.buf + st.start, st.len, kw.t);
        ws = string_token(a, tw, fast, s_hot, kw, BytesPtr{a.buf + st.start}, st, gate);
      }
      __syncwarp();
      append_tokens(a, tw, lane, ws, st, wi, it);
      defer_push(a, tw, lane, gate, st, s_hot);
    }
  }
  if (tw.n_def) { defer_drain(a, tw, lane, tw.defer, tw.n_def, s_hot); tw.n_def = 0; }  // (n_def is warp-uniform)
  for (uint32_t i = lane; i < tw.cs.left; i += 32) a.str[tw.cs.base + i].type = TOK_INVALID;
  for (uint32_t i = lane; i < tw.ci.left; i += 32) a.ip[tw.ci.base + i].type = TOK_INVALID;
  // per-type candidate counters (WorkerStats): warp-reduce, one atomic per warp and type
  uint32_t cnt[9] = {tw.n_dom, tw.n_mail, tw.n_v4, tw.n_v6, tw.n_md5, tw.n_sha1, tw.n_sha256, tw.n_sha384, tw.n_sha512};
#pragma unroll
  for (int t = 0; t < 9; t++) {
    uint32_t v = __reduce_add_sync(0xFFFFFFFFu, cnt[t]);
    if (lane == 0 && v) atomicAdd(&a.ctr->by_type[t], (unsigned long long)v);
  }
}

// ---------------------------------------------------------------------------------------------------------
// K1+K2 fused: scan_kernel — ONE pass over the log (round 2).
//
// tokenize_kernel + token_kernel read the log twice (the token kernel through 2 KiB windows, then again for the numeric
// queue) and pass ~0.2 candidates per log byte through queues in HBM: 3.35 DRAM bytes per log byte between them.  Here a
// warp owns a contiguous range of 1 KiB tiles as before, but the tiles arrive in a private 4 KiB shared-memory ring by
// bulk asynchronous copies (cp.async.bulk + mbarrier: one elected lane issues, all lanes wait on the tile's barrier;
// UBLKCP / SYNCS in SASS), and everything that follows reads them from there:
//   * classification: one 4-byte table entry per byte (128 rows x 32 lanes, conflict-free) holds a 4-bit byte category as
//     four bit planes; three LOP3 per mask turn the planes into the eight class masks of tokenize.cuh;
//   * the carry-chain word logic of tokenize_kernel, unchanged;
//   * dotted / numeric candidates go into a 64-entry queue in shared memory instead of HBM; whenever 32 are waiting (or
//     the tiles they lie in are about to be overwritten) one lane per candidate reads the word's first / last 16 bytes
//     from the ring and decides: IPv4 by SWAR, domain by the PSL last-label table, then the string gate;
//   * whatever needs a loop over the token's bytes (multi-label PSL suffixes, UTF-8 validation, words that began before
//     the ring's reach) and the rare anchors ('@', "::", hash-length and crypto-length words) still go to the segmented
//     queues in HBM and are handled by token_kernel / crypto_kernel afterwards, exactly as in round 1.
// Ring discipline: while tile t is processed, tiles t-1 and t-2 are intact and t+1 is in flight; the load of t+2 (into
// the slot of t-2) is issued at the end of tile t, after every queued candidate that ended before tile t has been
// handled.  A candidate that ends in tile t is queued only if it starts at or after the first byte of tile t-1.
// The 32 bytes behind the ring mirror its first 32, so a 20-byte read may start anywhere in the ring.
// ---------------------------------------------------------------------------------------------------------
static const int SK_WARPS = 16;       // with the 128 KiB hot string filter in shared memory
static const int SK_WARPS_NOHOT = 24; // when the database puts nothing into the hot filter (no strings, or only classes too big for it)
static const uint32_t SK_RING = 4096;
static const uint32_t SK_PAD = 32;
static const uint32_t SK_QCAP = 64;
static const uint32_t SK_LUT_BYTES = 128 * 32 * 4;  // byte values 0..127 (bytes >= 0x80 are mapped onto 'g' first: same category)
struct SkWarp { uint8_t ring[SK_RING + SK_PAD]; uint2 q[SK_QCAP]; unsigned long long mbar[4]; };
static_assert(sizeof(SkWarp) % 16 == 0, "rings must stay 16-byte aligned");
// (the wide block has room for the PSL last-label table as well: 32 KiB)
static constexpr size_t scan_smem(bool hot, int warps) { return (hot ? (size_t)HOT_WORDS * 4 : (size_t)TLD_SLOTS * 8) + SK_LUT_BYTES + (size_t)warps * sizeof(SkWarp); }

__device__ __forceinline__ void mbar_init(uint32_t mbar_s, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s), "r"(count) : "memory"); }
// one lane: arm the tile's barrier with the byte count, then start the bulk copy global -> shared that completes on it
__device__ __forceinline__ void tile_load(uint32_t dst_s, const uint8_t* src, uint32_t bytes, uint32_t mbar_s) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(src), "r"(bytes), "r"(mbar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar_s, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SK_DONE;\n"
      "bra SK_WAIT;\n"
      "SK_DONE:\n"
      "}\n" ::"r"(mbar_s), "r"(parity) : "memory");
}
__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
  return r;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t saddr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr) : "memory");
  return r;
}
__device__ __forceinline__ void sts_u64(uint32_t saddr, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(saddr), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(x) : "memory"); }

// per-warp state of the fused kernel's candidate queue and its global overflow queues (all warp-uniform)
struct SkQueues {
  uint32_t ring_s, q_s;            // shared-space addresses of the ring and of the candidate queue
  uint32_t qhead, qcount, q_old;   // queue cursor; q_old = entries at the head that ended before the current tile
  uint32_t high;                   // "some byte >= 0x80 in the last three tiles"
  Cand* qd; Cand* qn; Cand* qh;    // this warp's segments of the dotted / numeric / hash queues in HBM (slow path)
  uint32_t nd, nn, nh, ovf;
};
enum { SK_DOTTED = 0, SK_NUMERIC = 1, SK_HASH = 2 };  // kind of a queued word (bits 31, 30 of the entry's second word: numeric, hash)

// Slow path: the lanes with `slow` append their word to the dotted / numeric / hash queue segment in HBM (token_kernel's input).
__device__ __forceinline__ void sk_push_slow(const ScanArgs& a, SkQueues& k, uint32_t lane, bool slow, uint32_t kind, uint32_t start, uint32_t len) {
  const uint32_t bd = __ballot_sync(0xFFFFFFFFu, slow && kind == SK_DOTTED), bn = __ballot_sync(0xFFFFFFFFu, slow && kind == SK_NUMERIC);
  const uint32_t bh = __ballot_sync(0xFFFFFFFFu, slow && kind == SK_HASH);
  const uint32_t lt = (1u << lane) - 1u;
  if (bd) {
    if (k.nd + __popc(bd) <= a.seg_cap[Q_DOTTED]) { if (slow && kind == SK_DOTTED) k.qd[k.nd + __popc(bd & lt)] = Cand{start, len}; }
    else k.ovf |= 1u << Q_DOTTED;
    k.nd += __popc(bd);
  }
  if (bn) {
    if (k.nn + __popc(bn) <= a.seg_cap[Q_NUMERIC]) { if (slow && kind == SK_NUMERIC) k.qn[k.nn + __popc(bn & lt)] = Cand{start, len}; }
    else k.ovf |= 1u << Q_NUMERIC;
    k.nn += __popc(bn);
  }
  if (bh) {
    if (k.nh + __popc(bh) <= a.seg_cap[Q_HASH]) { if (slow && kind == SK_HASH) k.qh[k.nh + __popc(bh & lt)] = Cand{start, len}; }
    else k.ovf |= 1u << Q_HASH;
    k.nh += __popc(bh);
  }
}

// One lane per queued candidate, g (<= 32) of them from the head of the queue.
template <typename H>
__device__ __forceinline__ void sk_group(const ScanArgs& a, TokenWarp& tw, SkQueues& k, const H& s_hot, const uint64_t* tld, uint32_t g, uint32_t lane, uint32_t xflags, bool fast,
                                         bool final) {
  __syncwarp();  // the queue entries were written by other lanes
  const bool have = lane < g;
  uint2 e = make_uint2(0u, 0u);
  if (have) e = lds_u64(k.q_s + ((k.qhead + lane) & (SK_QCAP - 1)) * 8);
  k.qhead = (k.qhead + g) & (SK_QCAP - 1); k.qcount -= g; k.q_old = k.q_old > g ? k.q_old - g : 0u;
  const uint32_t start = e.x, len = e.y & 0x3FFFFFFFu;
  const bool numeric = (e.y >> 31) != 0, is_hash = ((e.y >> 30) & 1u) != 0;
  KeyWords kw;
  load_head_words_shared(k.ring_s + (start & (SK_RING - 1)), kw.h);
  load_tail_words_shared_end(k.ring_s + ((start + len - 16u) & (SK_RING - 1)) + 16u, have ? len : 0u, kw.t);
  // A word of hex digits and dots, 7..15 bytes long, may be an IPv4 address: it goes to the IP token list UNPARSED, with its
  // first 16 bytes, and iptrie_kernel parses it (one thread per token there, every lane busy, in a kernel that waits on memory
  // anyway; here a quarter of the lanes at best would run the parser).  Should it not be an address, iptrie_kernel gives it
  // the domain treatment (numeric_word_as_domain).
  const bool wi = have && numeric && len >= 7 && len <= 15 && (xflags & MGPU_X_IPV4);
  const IpTok it{start, len, TOK_RAW4, 0, {kw.h[0], kw.h[1], kw.h[2], kw.h[3]}};
  bool ws = false, slow = false;
  StrTok st{start, len, MGPU_T_DOMAIN};
  if (have && is_hash) {  // a hash-length word of hex digits is a token as it stands (lib.rs:1212-1250)
    st.type = hash_type_of(len);
    ws = true;
  } else if (have && !wi && (xflags & MGPU_X_DOMAINS)) {
    // a valid IPv4 address is never a domain: its last label is numeric, and no PSL entry ends in one (checked at upload)
    const uint64_t tail8 = len >= 8 ? (((uint64_t)kw.t[3] << 32) | kw.t[2]) : ((((uint64_t)kw.h[1] << 32) | kw.h[0]) << (8 * (8 - len)));
    const int c = tld_class(tld, tail8);
    if (c == TLD_ACCEPT && !k.high) ws = true;
    else if (c != TLD_REJECT) slow = true;  // multi-label suffix walk or UTF-8 validation: loops over the bytes -> token_kernel
  }
  __syncwarp();
  if (__any_sync(0xFFFFFFFFu, slow)) sk_push_slow(a, k, lane, slow, SK_DOTTED, start, len);
  uint32_t gate = 0;
  if (__any_sync(0xFFFFFFFFu, ws)) {
    if (ws) {
      tw.n_dom += st.type == MGPU_T_DOMAIN; tw.n_md5 += st.type == MGPU_T_MD5; tw.n_sha1 += st.type == MGPU_T_SHA1;
      tw.n_sha256 += st.type == MGPU_T_SHA256; tw.n_sha384 += st.type == MGPU_T_SHA384; tw.n_sha512 += st.type == MGPU_T_SHA512;
      if (fast) { gate = string_gate(a.db, s_hot, kw, len); ws = false; }  // (both filter stages of a class run in defer_drain: no per-lane loops here)
    }
    __syncwarp();
  }
  append_tokens(a, tw, lane, ws, st, wi, it);
  defer_push(a, tw, lane, gate, st, s_hot, final);
}

// HOT = false: the database's string filters use no hot-filter class (DbView::hot_tags == 0, or no string lookups at all), so
// the 128 KiB are not staged and the block runs more warps instead.
template <uint32_t FIXED, bool HOT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) scan_kernel(ScanArgs a) {
  extern __shared__ __align__(128) uint8_t sk_smem[];
  constexpr size_t HOT_BYTES = HOT ? (size_t)HOT_WORDS * 4 : (size_t)TLD_SLOTS * 8;  // hot string filter, or (wide block) the PSL last-label table
  uint32_t* s_hot_words = reinterpret_cast<uint32_t*>(sk_smem);
  uint32_t* lut = reinterpret_cast<uint32_t*>(sk_smem + HOT_BYTES);
  SkWarp* sw = reinterpret_cast<SkWarp*>(sk_smem + HOT_BYTES + SK_LUT_BYTES);
  // the hot string filter: the block's shared-memory copy, or (wide block) the 128 KiB in global memory behind L1
  typename std::conditional<HOT, HotShared, HotGlobal>::type s_hot;
  if constexpr (HOT) { s_hot.base = (uint32_t)__cvta_generic_to_shared(sk_smem); s_hot.g2 = a.db.gen_gram2; }
  else { s_hot.p = a.db.hot; s_hot.g2 = a.db.gen_gram2; }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const bool fast = a.fast != 0;
  for (uint32_t i = threadIdx.x; i < 128 * 32; i += blockDim.x) lut[i] = category_planes((uint8_t)(i >> 5));
  if (HOT && fast) {
    const uint4* src = reinterpret_cast<const uint4*>(a.db.hot);
    for (uint32_t i = threadIdx.x; i < HOT_WORDS / 4; i += blockDim.x) reinterpret_cast<uint4*>(s_hot_words)[i] = src[i];
  }
  const uint64_t* tld_tab = a.db.psl_tld;
  if (!HOT && a.db.psl_tld) {
    const uint4* src = reinterpret_cast<const uint4*>(a.db.psl_tld);
    for (uint32_t i = threadIdx.x; i < TLD_SLOTS / 2; i += blockDim.x) reinterpret_cast<uint4*>(sk_smem)[i] = src[i];
    tld_tab = reinterpret_cast<const uint64_t*>(sk_smem);
  }
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(&sw[warp].mbar[0]);
  if (lane == 0) { for (uint32_t j = 0; j < 4; j++) mbar_init(mbar_s + 8 * j, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const uint32_t lut_lane_s = (uint32_t)__cvta_generic_to_shared(lut) + lane * 4;
  const uint64_t nwarps = (uint64_t)gridDim.x * WARPS;  // == a.nseg
  const uint64_t w = (uint64_t)blockIdx.x * WARPS + warp;
  const uint64_t tiles = (a.n + TILE_BYTES - 1) / TILE_BYTES;
  const uint64_t tpw = (tiles + nwarps - 1) / nwarps;
  const uint64_t t0 = w * tpw;
  uint64_t t1 = t0 + tpw;
  if (t1 > tiles) t1 = tiles;
  uint32_t* const qa = a.q_at + w * a.seg_cap[Q_AT];
  uint32_t* const qc = a.q_c2 + w * a.seg_cap[Q_COLON2];
  Cand* const ql = a.q_long + w * a.seg_cap[Q_LONG];
  SkQueues k;
  k.ring_s = (uint32_t)__cvta_generic_to_shared(&sw[warp].ring[0]);
  k.q_s = (uint32_t)__cvta_generic_to_shared(&sw[warp].q[0]);
  k.qhead = 0; k.qcount = 0; k.q_old = 0; k.high = 0;
  k.qd = a.q_dotted + w * a.seg_cap[Q_DOTTED]; k.qn = a.q_numeric + w * a.seg_cap[Q_NUMERIC]; k.qh = a.q_hash + w * a.seg_cap[Q_HASH];
  k.nd = 0; k.nn = 0; k.nh = 0; k.ovf = 0;
  uint32_t na = 0, nc = 0, nl = 0;
  uint32_t lines = 0;
  TokenWarp tw;
  memset(&tw, 0, sizeof tw);
  tw.defer = a.defer + (size_t)w * DEFER_CAP;
  const uint32_t xflags = FIXED ? FIXED : a.flags;
  if (t0 < t1) {
    TileCarry cy = range_prologue(a.buf, a.lo, t0 * TILE_BYTES);
    uint32_t pDOT = (cy.prev & PV_DOT) ? 0x80000000u : 0u, pDASH = (cy.prev & PV_DASH) ? 0x80000000u : 0u;
    uint32_t pCL = ((cy.prev & PV_CL1) ? 0x80000000u : 0u) | ((cy.prev & PV_CL2) ? 0x40000000u : 0u);
    const bool want_dot = (xflags & (MGPU_X_IPV4 | MGPU_X_DOMAINS)) != 0;
    const bool want_hash = (xflags & MGPU_X_HASHES) != 0;
    const bool want_at = (xflags & MGPU_X_EMAILS) != 0;
    const bool want_c2 = (xflags & MGPU_X_IPV6) != 0;
    const bool want_long = (xflags & MGPU_X_CRYPTO) != 0;
    const uint32_t t_begin = (uint32_t)t0, t_end = (uint32_t)t1, t_last = (uint32_t)(tiles - 1);
    const uint32_t range_lo = t_begin * TILE_BYTES;
    if (lane == 0) {
      tile_load(k.ring_s + (t_begin & 3u) * TILE_BYTES, a.buf + (size_t)t_begin * TILE_BYTES, TILE_BYTES, mbar_s + (t_begin & 3u) * 8);
      if (t_begin + 1 < t_end) tile_load(k.ring_s + ((t_begin + 1) & 3u) * TILE_BYTES, a.buf + (size_t)(t_begin + 1) * TILE_BYTES, TILE_BYTES, mbar_s + ((t_begin + 1) & 3u) * 8);
    }
    uint32_t high_hist = 0;
    bool flushed = false;
    // One iteration per tile, plus a last one (t == t_end) that only empties the queues: sk_group() is large, so it is
    // called from ONE place and everything that has to reach it goes through the loop below.
    for (uint32_t t = t_begin; t <= t_end; t++) {
      const bool last = t == t_end;
      const uint32_t tile_base = t * TILE_BYTES;  // chunks are at most 2 GiB: positions fit 32 bits
      const uint32_t p = tile_base + lane * SLICE_BYTES;
      uint32_t rem = 0, candNum = 0, candHashK = 0, rounds = 0, Bm = 0, lane_open = 0, ring_lo = 0, total_dot = 0, excl_dot = 0;
      bool may_slow = false;
      if (!last) {
      const uint32_t slot = t & 3u;
      mbar_wait(mbar_s + slot * 8, ((t - t_begin) >> 2) & 1u);
      if (slot == 0) {  // refresh the mirror of the ring's first bytes
        if (lane < SK_PAD / 4) sts_u32(k.ring_s + SK_RING + lane * 4, lds_u32(k.ring_s + lane * 4));
        __syncwarp();
      }
      const uint4 v0 = lds_u128(k.ring_s + slot * TILE_BYTES + lane * SLICE_BYTES), v1 = lds_u128(k.ring_s + slot * TILE_BYTES + lane * SLICE_BYTES + 16);
      uint32_t wds[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      const uint32_t hb = (v0.x | v0.y | v0.z | v0.w | v1.x | v1.y | v1.z | v1.w) & 0x80808080u;
      const bool tile_high = __any_sync(0xFFFFFFFFu, hb != 0);
      high_hist = ((high_hist << 1) | (tile_high ? 1u : 0u)) & 7u;
      k.high = high_hist;
      if (tile_high) {  // (warp-uniform, rare) bytes >= 0x80 are domain characters of the 'g' kind: give them that row of the table
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const uint32_t mk = ((wds[j] & 0x80808080u) >> 7) * 0xFFu;
          wds[j] = (wds[j] & ~mk) | (0x67676767u & mk);
        }
      }
      uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const uint32_t bv = __byte_perm(wds[j], 0u, 0x4440u | (uint32_t)kk);  // byte kk of the word
          const uint32_t ev = lds_u32(bv * 128u + lut_lane_s);
          acc[j >> 1] += ev << ((j & 1) * 4 + kk);
        }
      }
      // byte c of acc[q] = plane c of bytes 8q..8q+7: a 4x4 byte transpose gives one 32-bit mask per plane
      LaneMasks m;
      {
        const uint32_t r0 = __byte_perm(acc[0], acc[1], 0x5140), r1 = __byte_perm(acc[0], acc[1], 0x7362);
        const uint32_t r2 = __byte_perm(acc[2], acc[3], 0x5140), r3 = __byte_perm(acc[2], acc[3], 0x7362);
        const uint32_t P0 = __byte_perm(r0, r2, 0x5410), P1 = __byte_perm(r0, r2, 0x7632);
        const uint32_t P2 = __byte_perm(r1, r3, 0x5410), P3 = __byte_perm(r1, r3, 0x7632);
        m.B = P3; m.NL = P3 & ~P1 & P0; m.AT = P3 & P1 & ~P0; m.CL = P3 & P1 & P0;
        m.DM = P2; m.DOT = P2 & ~P1 & P0; m.DASH = P2 & P1 & ~P0; m.HX = P2 & P1 & P0;
      }
      if (t == 0 || t == t_last) {  // only the chunk's first (lo < 16) and last tile can hold bytes outside [lo, n): they behave like a boundary (chunk edge)
        uint64_t valid = a.n > p ? a.n - p : 0;
        uint32_t keep = valid >= 32 ? 0xFFFFFFFFu : ((1u << (uint32_t)valid) - 1u);
        if (p < a.lo) keep &= (a.lo - p >= 32) ? 0u : (0xFFFFFFFFu << (uint32_t)(a.lo - p));
        m.B |= ~keep; m.DOT &= keep; m.AT &= keep; m.CL &= keep; m.NL &= keep; m.DM &= keep; m.HX &= keep; m.DASH &= keep;
      }
      if ((int32_t)cy.prevB < 0) cy.open_start = tile_base;  // no word is open: a word that fills the tile from its first byte starts here
      lines += __popc(m.NL);
      // oldest byte a queued word may start at; only the word that is open at the start of the tile can begin before it
      ring_lo = tile_base >= range_lo + TILE_BYTES ? tile_base - TILE_BYTES : range_lo;
      may_slow = (uint32_t)cy.open_start < ring_lo;

      const uint32_t T = ~m.B;
      uint32_t Bprev = __shfl_up_sync(0xFFFFFFFFu, m.B, 1), DOTp = __shfl_up_sync(0xFFFFFFFFu, m.DOT, 1);
      uint32_t DASHp = __shfl_up_sync(0xFFFFFFFFu, m.DASH, 1), CLp = __shfl_up_sync(0xFFFFFFFFu, m.CL, 1);
      if (lane == 0) { Bprev = cy.prevB; DOTp = pDOT; DASHp = pDASH; CLp = pCL; }
      const uint32_t prevT = ~__funnelshift_l(Bprev, m.B, 1);  // bit i: byte i-1 is a word byte
      const bool pT = (int32_t)Bprev >= 0;                      // the byte before my slice is a word byte
      const uint32_t S = T & ~prevT;
      uint32_t bad, bad_end;
      domain_rule_masks_prev(m, S, __funnelshift_l(DOTp, m.DOT, 1), __funnelshift_l(DASHp, m.DASH, 1), bad, bad_end);
      // "does the word contain ..." chains (tokenize.cuh): a byte that rules out a domain, a '.', a byte that is neither a hex
      // digit nor a '.'; "a non-hex byte" is the OR of the last two
      const uint32_t Y1 = T & (~m.DM | bad), Y2 = m.DOT, Y4 = T & ~m.HX & ~m.DOT;
      const uint32_t g1 = chain_gen(T, Y1), g2 = chain_gen(T, Y2), g4 = chain_gen(T, Y4);
      const uint32_t pb = __ballot_sync(0xFFFFFFFFu, T == 0xFFFFFFFFu);
      const uint32_t G1 = __ballot_sync(0xFFFFFFFFu, g1), G2 = __ballot_sync(0xFFFFFFFFu, g2), G4 = __ballot_sync(0xFFFFFFFFu, g4);
      uint32_t co1, co2, co4;
      const uint32_t cv1 = carry_chain(G1, pb & ~G1, cy.cBad, co1), cv2 = carry_chain(G2, pb & ~G2, cy.cDot, co2), cv4 = carry_chain(G4, pb & ~G4, cy.cNhd, co4);
      cy.cBad = co1; cy.cDot = co2; cy.cNhd = co4;
      const uint32_t E = m.B & prevT;  // boundaries that end a word
      const uint32_t hasBad = chain_ends(T, Y1, (cv1 >> lane) & 1u, m.B), hasDot = chain_ends(T, Y2, (cv2 >> lane) & 1u, m.B);
      const uint32_t hasNhd = chain_ends(T, Y4, (cv4 >> lane) & 1u, m.B);
      const uint32_t hasNhx = hasNhd | hasDot;
      const uint32_t candDot = want_dot ? (hasDot & ~hasBad & ~bad_end) : 0u;
      uint32_t candHex = (want_hash && pT) ? (E & ~hasNhx & (m.B & (0u - m.B))) : 0u;
      if (candHex && (Bprev >> (__ffs(candHex) - 1)) != 0) candHex = 0;
      const uint32_t candLong = want_long ? long_word_ends(T, ~Bprev, E) : 0u;
      const uint32_t candAt = want_at ? m.AT : 0u;
      const uint32_t cl1 = __funnelshift_l(CLp, m.CL, 1), cl2 = __funnelshift_l(CLp, m.CL, 2);
      const uint32_t candC2 = want_c2 ? (m.CL & cl1 & ~cl2) : 0u;

      // where the word that is open at the start of my slice begins
      const uint32_t hasB = __ballot_sync(0xFFFFFFFFu, m.B != 0);
      const uint32_t src = lane_below_with_boundary(hasB, lane);
      const uint32_t Bsrc = __shfl_sync(0xFFFFFFFFu, m.B, src & 31u);
      lane_open = src < 32u ? tile_base + src * 32 + top_bit(Bsrc) + 1 : (uint32_t)cy.open_start;

      if (__any_sync(0xFFFFFFFFu, (candHex | candAt | candC2 | candLong) != 0)) {  // rare in most logs: one vote covers the four
        if (candHex) {  // an all-hex word of >= 32 bytes ends here: a hash iff its length is one of the five (joins the dotted words below)
          const uint32_t bit = __ffs(candHex) - 1;
          const uint32_t below = m.B & ~(0xFFFFFFFFu << bit);
          const uint32_t s = below ? p + top_bit(below) + 1 : lane_open;
          if (is_hash_len(p + bit - s)) candHashK = candHex;
        }
        if (__any_sync(0xFFFFFFFFu, candAt != 0)) {
          const uint32_t cnt = __popc(candAt), incl = warp_incl_scan(cnt, lane);
          const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          if (na + total <= a.seg_cap[Q_AT]) {
            uint32_t idx = na + incl - cnt;
            for (uint32_t mm = candAt; mm; mm &= mm - 1) qa[idx++] = (uint32_t)(p + __ffs(mm) - 1);
          } else k.ovf |= 1u << Q_AT;
          na += total;
        }
        if (__any_sync(0xFFFFFFFFu, candLong != 0)) {
          Cand c2[2]; uint32_t kq = 0;
          for (uint32_t mm = candLong; mm; mm &= mm - 1) {
            const uint32_t bit = __ffs(mm) - 1;
            const uint32_t below = m.B & ~(0xFFFFFFFFu << bit);
            const uint32_t s = below ? p + top_bit(below) + 1 : lane_open;
            if (is_crypto_len(p + bit - s) && kq < 2) c2[kq++] = Cand{s, p + bit - s};
          }
          const uint32_t incl = warp_incl_scan(kq, lane), total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          if (total) {
            if (nl + total <= a.seg_cap[Q_LONG]) { for (uint32_t j = 0; j < kq; j++) ql[nl + incl - kq + j] = c2[j]; }
            else k.ovf |= 1u << Q_LONG;
            nl += total;
          }
        }
        if (__any_sync(0xFFFFFFFFu, candC2 != 0)) {
          const uint32_t cnt = __popc(candC2), incl = warp_incl_scan(cnt, lane);
          const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          if (nc + total <= a.seg_cap[Q_COLON2]) {
            uint32_t idx = nc + incl - cnt;
            for (uint32_t mm = candC2; mm; mm &= mm - 1) qc[idx++] = (uint32_t)(p + __ffs(mm) - 2);
          } else k.ovf |= 1u << Q_COLON2;
          nc += total;
        }
      }

      // ---- carry into the next tile (nothing below needs the class masks any more, only the boundary mask) ----
      if (hasB) {
        const uint32_t ll = top_bit(hasB);
        const uint32_t Bl = __shfl_sync(0xFFFFFFFFu, m.B, ll);
        cy.open_start = tile_base + ll * 32 + top_bit(Bl) + 1;
      }
      cy.prevB = __shfl_sync(0xFFFFFFFFu, m.B, 31);
      pDOT = __shfl_sync(0xFFFFFFFFu, m.DOT, 31); pDASH = __shfl_sync(0xFFFFFFFFu, m.DASH, 31); pCL = __shfl_sync(0xFFFFFFFFu, m.CL, 31);
      Bm = m.B;
      rem = candDot | candHashK;    // (disjoint: a hash word holds no '.')
      candNum = candDot & ~hasNhd;  // hex digits and dots only: IPv4 candidates
      {
        const uint32_t cnt = (uint32_t)__popc(rem), incl = warp_incl_scan(cnt, lane);
        total_dot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        excl_dot = incl - cnt;
        rounds = total_dot ? 32u : 0u;  // (upper bound for the round-by-round path, which stops when no lane has a word left)
      }
      }

      // ---- dotted words: into the shared-memory queue; a group of 32 whenever that many wait ----
      // Usual case: all of the tile's words fit behind what is queued -> one prefix sum over the lanes' counts and every
      // lane stores its own.  Otherwise (a tile full of short words, or a word that began behind the ring) one word per lane
      // and round, with groups in between.
      if (total_dot && !may_slow && k.qcount + total_dot <= SK_QCAP) {
        uint32_t idx = k.qhead + k.qcount + excl_dot;
        for (uint32_t mm = rem; mm; mm &= mm - 1) {
          const uint32_t low = mm & (0u - mm), bit = __ffs(mm) - 1;
          const uint32_t below = Bm & (low - 1u);
          const uint32_t s = below ? p + top_bit(below) + 1 : lane_open;
          sts_u64(k.q_s + (idx++ & (SK_QCAP - 1)) * 8, s, (p + bit - s) | ((candNum & low) ? 0x80000000u : 0u) | ((candHashK & low) ? 0x40000000u : 0u));
        }
        k.qcount += total_dot;
        rem = 0; rounds = 0;
      }
      for (;;) {
        while (rounds && k.qcount < 32) {
          if (!__any_sync(0xFFFFFFFFu, rem != 0)) { rounds = 0; break; }
          const bool have = rem != 0;
          uint32_t s = 0, len = 0, kind = SK_DOTTED;
          if (have) {
            const uint32_t low = rem & (0u - rem), bit = __ffs(rem) - 1;
            const uint32_t below = Bm & (low - 1u);
            s = below ? p + top_bit(below) + 1 : lane_open;
            len = p + bit - s;
            kind = (candNum & low) ? (uint32_t)SK_NUMERIC : ((candHashK & low) ? (uint32_t)SK_HASH : (uint32_t)SK_DOTTED);
            rem &= rem - 1;
          }
          bool slow = false;
          if (may_slow) {  // (warp-uniform) the word that was open when the tile began lies partly behind the ring
            slow = have && s < ring_lo;
            if (__any_sync(0xFFFFFFFFu, slow)) sk_push_slow(a, k, lane, slow, kind, s, len);
          }
          const uint32_t bq = __ballot_sync(0xFFFFFFFFu, have && !slow);
          if (have && !slow) sts_u64(k.q_s + ((k.qhead + k.qcount + __popc(bq & lt_mask)) & (SK_QCAP - 1)) * 8, s, len | (kind == SK_NUMERIC ? 0x80000000u : 0u) | (kind == SK_HASH ? 0x40000000u : 0u));
          k.qcount += __popc(bq);
        }
        uint32_t g;
        if (k.qcount >= 32) g = 32;
        else if (rounds == 0 && (last ? !flushed : k.q_old != 0)) g = k.qcount;  // what ended before this tile must go now; at the very end, everything
        else break;
        const bool fin = last && k.qcount == g;
        sk_group(a, tw, k, s_hot, tld_tab, g, lane, xflags, fast, fin);
        if (fin) flushed = true;
      }
      if (last) break;
      // ---- the next load: tile t+2 into the slot of tile t-2, whose words have all been handled ----
      k.q_old = k.qcount;
      __syncwarp();
      if (lane == 0 && t + 2 < t_end) tile_load(k.ring_s + ((t + 2) & 3u) * TILE_BYTES, a.buf + (size_t)(t + 2) * TILE_BYTES, TILE_BYTES, mbar_s + ((t + 2) & 3u) * 8);
    }
  }
  for (uint32_t i = lane; i < tw.cs.left; i += 32) a.str[tw.cs.base + i].type = TOK_INVALID;
  for (uint32_t i = lane; i < tw.ci.left; i += 32) a.ip[tw.ci.base + i].type = TOK_INVALID;
  if (lane == 0) {
    uint32_t* sc = a.seg_cnt + w;
    sc[Q_DOTTED * a.nseg_max] = k.ovf & (1u << Q_DOTTED) ? 0u : k.nd;
    sc[Q_HASH * a.nseg_max] = k.ovf & (1u << Q_HASH) ? 0u : k.nh;
    sc[Q_AT * a.nseg_max] = k.ovf & (1u << Q_AT) ? 0u : na;
    sc[Q_COLON2 * a.nseg_max] = k.ovf & (1u << Q_COLON2) ? 0u : nc;
    sc[Q_NUMERIC * a.nseg_max] = k.ovf & (1u << Q_NUMERIC) ? 0u : k.nn;
    sc[Q_LONG * a.nseg_max] = k.ovf & (1u << Q_LONG) ? 0u : nl;
    if (k.ovf) atomicOr(&a.ctr->overflow, k.ovf);
  }
  for (int d = 16; d; d >>= 1) lines += __shfl_down_sync(0xFFFFFFFFu, lines, d);
  if (lane == 0 && lines) atomicAdd(&a.ctr->lines, (unsigned long long)lines);
  uint32_t cnt[9] = {tw.n_dom, tw.n_mail, tw.n_v4, tw.n_v6, tw.n_md5, tw.n_sha1, tw.n_sha256, tw.n_sha384, tw.n_sha512};
#pragma unroll
  for (int tt = 0; tt < 9; tt++) {
    uint32_t v = __reduce_add_sync(0xFFFFFFFFu, cnt[tt]);
    if (lane == 0 && v) atomicAdd(&a.ctr->by_type[tt], (unsigned long long)v);
  }
}

// K2b crypto kernel: the long-word queue -> Bitcoin / Ethereum / Monero address tokens (lib.rs:1269-1409).  One thread per
// candidate; candidates are rare and the validators (base58 big-number decode, SHA-256, Keccak-f) are long, so this is a
// kernel of its own with its own register budget.  Valid tokens are counted and go through the same string filters /
// token list as every other string token.
__global__ void __launch_bounds__(128) crypto_kernel(ScanArgs a) {
  if (a.ctr->overflow) return;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t seg = blockIdx.x * (blockDim.x >> 5) + warp; seg < a.nseg; seg += nwarps) {
    const uint32_t n = a.seg_cnt[Q_LONG * a.nseg_max + seg];
    const Cand* q = a.q_long + (size_t)seg * a.seg_cap[Q_LONG];
    for (uint32_t i = lane; i < n; i += 32) {
      const Cand c = q[i];
      if ((uint64_t)c.start + c.len > a.n) continue;
      const uint8_t* w = a.buf + c.start;
      const uint32_t type = crypto_word_type(w, c.len, a.flags);
      if (type == NONE32) continue;
      atomicAdd(&a.ctr->by_type[type], 1ULL);
      uint32_t f = 0;
      if (a.fast) { f = string_filters(a.db, a.db.hot, w, c.len); if (!f) continue; }
      const uint32_t k = atomicAdd(&a.ctr->n_str, 1u);
      if (k >= a.cap_str) { atomicOr(&a.ctr->overflow, 1u << 8); continue; }
      a.str[k] = StrTok{c.start, c.len, type | f};
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// record emission (aggregated across the currently active lanes)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t agg_add(uint32_t* ctr, uint32_t v) {
  cg::coalesced_group g = cg::coalesced_threads();
  uint32_t incl = cg::inclusive_scan(g, v);
  uint32_t base = 0;
  if (g.thread_rank() == g.size() - 1) base = atomicAdd(ctr, incl);
  base = g.shfl(base, g.size() - 1);
  return base + incl - v;
}

// K3: IP tokens -> SearchTree::lookup (mmdb/tree.rs:46-125) -> records.
//  * Lock step: a warp takes 32 consecutive tokens and advances all their walks together, IPT_STEPS tree records per round,
//    until none is left walking (trie_walk_begin / trie_walk_step).  Round 1 let every lane run its own walk to the end;
//    at full size, with larger token reservation units, that version lost 0..5 of 8 million records from run to run: ONE lane in a deep IPv6 walk (~50 us of
//    dependent DRAM reads) next to idle lanes was left behind by its warp — the warp's other 31 lanes, and with them the
//    block's barriers and the warp-uniform registers, ran 16 loop iterations ahead of it (measured with the audit
//    counters of mgpu_set_option("verify_tokens"); profiles/README.md "the lost records of round 1").  Here no lane is
//    ever more than IPT_STEPS loads away from the rest of its warp, and there is no block barrier after the first line.
//  * A hit rate of a few per cent over 10^8 tokens means millions of records: appended one warp's worth at a time they
//    serialise on the single record counter (same-address atomics retire about one per nanosecond).  So every warp stages
//    its records in shared memory and takes one slot range per IPT_STAGE records, written out as coalesced 16-byte stores.
static const uint32_t IPT_STAGE = 96;
static const int IPT_STEPS = 4;
__device__ __forceinline__ void iptrie_flush(const ScanArgs& a, const mgpu_match* stage, uint32_t staged, uint32_t lane) {
  __syncwarp();
  uint32_t b = 0;
  if (lane == 0) {
    if (a.dbg) atomicAdd(&a.dbg[10], (unsigned long long)staged);
    b = atomicAdd(&a.tot->n_rec, staged);
    if ((uint64_t)b + staged > a.cap_rec) atomicOr(&a.ctr->overflow, 1u << 10);
  }
  b = __shfl_sync(0xFFFFFFFFu, b, 0);
  const uint4* src = reinterpret_cast<const uint4*>(stage);
  uint4* dst = reinterpret_cast<uint4*>(a.recs);
  for (uint32_t j = lane; j < 2 * staged; j += 32) if ((uint64_t)b + (j >> 1) < a.cap_rec) dst[2 * (size_t)b + j] = src[j];
  __syncwarp();
}
// A numeric-queue word (hex digits and dots, 7..15 bytes, all of it in t.w) that is not an IPv4 address may still be a
// domain ("cafe.de"): PSL check, then the string path like any other domain token.  Rare; one thread, plain atomics.
__device__ __noinline__ void numeric_word_as_domain(const ScanArgs& a, const IpTok& t) {
  uint32_t tw[4];
  tail_words_from_head(t.w, t.len, tw);
  const uint64_t tail8 = t.len >= 8 ? (((uint64_t)tw[3] << 32) | tw[2]) : ((((uint64_t)t.w[1] << 32) | t.w[0]) << (8 * (8 - t.len)));
  if (!domain_word_fast(a.db, a.db.psl_tld, a.buf + t.start, t.len, /*maybe_high=*/false, tail8)) return;
  atomicAdd(&a.ctr->by_type[MGPU_T_DOMAIN], 1ULL);
  uint32_t f = 0;
  if (a.fast) { f = string_filters(a.db, a.db.hot, a.buf + t.start, t.len); if (!f) return; }
  const uint32_t k = atomicAdd(&a.ctr->n_str, 1u);
  if (k >= a.cap_str) { atomicOr(&a.ctr->overflow, 1u << 8); return; }
  a.str[k] = StrTok{t.start, t.len, (uint32_t)MGPU_T_DOMAIN | f};
}

// (MINB: resident blocks per SM the register budget is cut for — measured, more warps at the price of spills do not pay;
//  grid constant: the out-of-line fallback takes &a without a per-thread copy of the 0.7 KB of arguments)
template <int MINB>
__global__ void __launch_bounds__(256, MINB) iptrie_kernel(const __grid_constant__ ScanArgs a) {
  __shared__ __align__(16) mgpu_match s_rec[8][IPT_STAGE];
  __shared__ uint32_t s_ovf;
  const uint32_t n = min(a.ctr->n_ip, a.cap_ip);
  if (threadIdx.x == 0) s_ovf = a.ctr->overflow;  // (read once per block: the exit must be uniform, other blocks may set the flag)
  __syncthreads();
  if (s_ovf) return;
  const bool walk = a.lookups && a.db.has_ip;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nw = gridDim.x * 8;
  uint32_t staged = 0;  // records in s_rec[warp] (warp-uniform)
  uint32_t n_v4 = 0;    // raw numeric words that turned out to be addresses (WorkerStats.ipv4_count)
  for (uint32_t t0 = (blockIdx.x * 8 + warp) * 32u; t0 < n; t0 += nw * 32u) {
    const uint32_t i = t0 + lane;
    IpTok t;
    t.type = TOK_INVALID;
    if (i < n) t = a.ip[i];
    if (t.type == TOK_RAW4) {  // the fused scan kernel's numeric words: try_parse_ipv4 (lib.rs:813-869) happens here
      uint32_t addr = 0;
      if (parse_ipv4_hexdot(t.w, t.len, addr)) { t.type = MGPU_T_IPV4; t.w[0] = addr; n_v4++; }
      else { t.type = TOK_INVALID; if (a.flags & MGPU_X_DOMAINS) numeric_word_as_domain(a, t); }
      if (!a.lookups) { a.ip[i].type = t.type; a.ip[i].w[0] = t.w[0]; }  // extraction only: the list itself is the result
    }
    if (!walk) continue;
    uint32_t off = 0; uint8_t pl = 0;
    TrieWalk w;
    int r = TRIE_MISS;
    if (t.type == MGPU_T_IPV4 || t.type == MGPU_T_IPV6) r = trie_walk_begin(a.db, t.w, t.type == MGPU_T_IPV6, w, off, pl);
    while (__any_sync(0xFFFFFFFFu, r == TRIE_MORE)) {
#pragma unroll
      for (int k = 0; k < IPT_STEPS; k++) if (r == TRIE_MORE) r = trie_walk_step(a.db, t.w, w, off, pl);
    }
    const bool hit = r == TRIE_HIT;
    if (a.dbg) {  // audit: hits counted per thread; sub-groups of a warp that arrive here on their own
      if (hit) atomicAdd(&a.dbg[8], 1ULL);
      const uint32_t am = __activemask();
      if (am != 0xFFFFFFFFu && (am & (0u - am)) == (1u << lane)) atomicAdd(&a.dbg[11], 1ULL);
    }
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
    if (bal) {
      const uint32_t cnt = (uint32_t)__popc(bal);
      if (staged + cnt > IPT_STAGE) { iptrie_flush(a, s_rec[warp], staged, lane); staged = 0; }
      if (hit) {
        mgpu_match m;
        m.offset = a.base + t.start; m.len = t.len; m.item_type = (uint8_t)t.type; m.kind = MGPU_KIND_IP; m.prefix_len = pl; m.reserved = 0;
        m.n_ids = 0; m.ids_index = 0; m.data_offset = off; m.pad = 0;
        s_rec[warp][staged + __popc(bal & ((1u << lane) - 1u))] = m;
      }
      staged += cnt;
    }
  }
  if (staged) iptrie_flush(a, s_rec[warp], staged, lane);
  n_v4 = __reduce_add_sync(0xFFFFFFFFu, n_v4);
  if (lane == 0 && n_v4) atomicAdd(&a.ctr->by_type[MGPU_T_IPV4], (unsigned long long)n_v4);
}

// tokens of one warp iteration -> staged window pointer
__device__ __forceinline__ const uint8_t* stage_tokens(const ScanArgs& a, uint8_t* win, bool valid, const StrTok& t, uint32_t lane,
                                                       uint32_t win_bytes = WIN_BYTES) {
  uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, valid ? t.start : 0xFFFFFFFFu);
  uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, valid ? t.start + t.len : 0u);
  return stage_window(a.buf, win, lo, hi, lane, win_bytes);
}

// One string token's result: the literal-hash id first, then every glob id find_all returns, sorted and deduplicated
// (database.rs:916-965, paraglob_offset.rs:1173-1181); a record is written iff there is at least one id.
__device__ __forceinline__ void emit_string_match(const ScanArgs& a, const StrTok& t, const uint8_t* text, bool lit_ok, uint32_t lit_pid,
                                                  uint32_t lit_off, bool exact, const AcAccel& acc) {
  uint32_t cnt = 0;
  if (exact) find_all_visit(a.db, text, t.len, acc, [&](uint32_t) { cnt++; });
  if (!lit_ok && cnt == 0) return;
  uint32_t total = cnt + (lit_ok ? 1u : 0u);
  uint32_t b = agg_add(&a.tot->n_ids, total);
  if ((uint64_t)b + total > a.cap_ids) { atomicOr(&a.ctr->overflow, 1u << 11); return; }
  uint32_t k = b;
  if (lit_ok) { a.ids[k].pattern_id = lit_pid; a.ids[k].data_offset = lit_off; k++; }
  const uint32_t g0 = k;
  if (cnt) {
    find_all_visit(a.db, text, t.len, acc, [&](uint32_t pid) { if (k < b + total) a.ids[k++].pattern_id = pid; });
    for (uint32_t x = g0 + 1; x < k; x++) {
      uint32_t v = a.ids[x].pattern_id, y = x;
      while (y > g0 && a.ids[y - 1].pattern_id > v) { a.ids[y].pattern_id = a.ids[y - 1].pattern_id; y--; }
      a.ids[y].pattern_id = v;
    }
    uint32_t u = g0;
    for (uint32_t x = g0; x < k; x++) if (x == g0 || a.ids[x].pattern_id != a.ids[u - 1].pattern_id) a.ids[u++].pattern_id = a.ids[x].pattern_id;
    k = u;
    for (uint32_t x = g0; x < k; x++) {
      uint32_t off;
      a.ids[x].data_offset = glob_data_offset(a.db, a.ids[x].pattern_id, off) ? off : MGPU_NO_DATA;
    }
  }
  uint32_t r_i = agg_add(&a.tot->n_rec, 1u);
  if (r_i >= a.cap_rec) { atomicOr(&a.ctr->overflow, 1u << 10); return; }
  mgpu_match r;
  r.offset = a.base + t.start; r.len = t.len; r.item_type = (uint8_t)t.type; r.kind = MGPU_KIND_PATTERN; r.prefix_len = 0; r.reserved = 0;
  r.n_ids = k - b; r.ids_index = b; r.data_offset = MGPU_NO_DATA; r.pad = 0;
  a.recs[r_i] = r;
}

// K4: literal hash probe per string token
__global__ void __launch_bounds__(KT_THREADS) lithash_kernel(ScanArgs a) {
  __shared__ __align__(16) uint8_t s_win[KT_WARPS][WIN_BYTES + 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = min(a.ctr->n_str, a.cap_str);
  const uint32_t nround = (n + 31u) & ~31u;
  if (a.ctr->overflow) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
    bool valid = i < n;
    StrTok t{0, 0, 0};
    if (valid) { t = a.str[i]; valid = t.type != TOK_INVALID; }
    const uint8_t* p = stage_tokens(a, s_win[warp], valid, t, lane);
    if (!valid) continue;
    uint32_t pid = NONE32;
    if (!lh_lookup(a.db, p + t.start, t.len, pid)) pid = NONE32;
    a.lh_res[i] = pid;
  }
}

// K5: Paraglob::find_all per string token, merge with the literal result, emit.
//
// Almost every token matches nothing, so the kernel first answers, for the 32 tokens of a warp, "does ANY automaton
// literal occur in this token?" and only flagged tokens (plus every token when the database has pure-wildcard
// patterns) take the exact find_all path.  The filter is the anchored search of device_fns.cuh split in two so that lanes
// stay busy:  (1) every lane scans its own token for start positions whose first 2 bytes begin some literal (bitmap in
// shared memory) and pushes them into a warp-shared queue;  (2) the queue is dealt round-robin to the lanes, which apply
// the exact 3-byte bitmap and the 8-byte prefix filter (L2) and then follow goto edges from the root, one step per loop
// iteration, until a node with outputs is reached (flag the token) or the path ends.  No failure links, no per-lane
// inner loops.
static const uint32_t SCAN_ROUND = 8;     // positions a lane scans per round
static const uint32_t ANCHOR_CAP = 384;   // queue A: positions that passed the 2-byte bitmap (drained above 128 entries)
static const uint32_t WALK_CAP = 128;     // queue B: positions that also passed the 3-byte bitmap and the 8-byte prefix filter

// The prefilter handles case-sensitive databases only (the reference's default and all BASELINE configs); with a
// case-insensitive database every token takes the exact path, which folds ASCII case itself.

// queue A -> queue B: exact 3-byte prefix bitmap, then the prefix map (both in L2): a node with outputs at depth 3..7 flags
// the token at once; a depth-8 node goes into queue B as (token lane, position, node offset).  Every lane takes entries
// lane, lane+32, ... of the slice [from, to).
__device__ __forceinline__ void filter_anchors(const ScanArgs& a, const uint8_t* p, const uint32_t* s_start, const uint32_t* s_len,
                                               const uint32_t* s_anchor, uint32_t from, uint32_t to, uint2* s_walk, uint32_t* s_wcnt,
                                               uint32_t* s_hit, uint32_t lane) {
  for (uint32_t k = from + lane; k < to; k += 32) {
    uint32_t an = s_anchor[k];
    uint32_t tl = an & 31u, j = an >> 5;
    if ((*(volatile const uint32_t*)s_hit >> tl) & 1u) continue;  // already flagged: it takes the exact path anyway
    const uint8_t* text = p + s_start[tl];
    uint32_t g3 = ((uint32_t)text[j] << 16) | ((uint32_t)text[j + 1] << 8) | text[j + 2];
    if (!((a.db.ac_gram3[g3 >> 5] >> (g3 & 31)) & 1u)) continue;
    const uint32_t avail = s_len[tl] - j;
    const uint64_t v = ldu64_fast(text + j);
    bool hit = false;
    for (uint32_t lens = a.db.ac_short_lens; lens && !hit; lens &= lens - 1) {
      uint32_t m = (uint32_t)__ffs((int)lens) - 1u;
      if (m > avail) break;
      hit = prefix_node(a.db, low_bytes(v, m), m) != 0;  // such nodes have outputs by construction
    }
    if (hit) { atomicOr(s_hit, 1u << tl); continue; }
    if (avail < 8) continue;
    uint32_t off = prefix_node(a.db, v, 8);
    if (!off) continue;
    s_walk[atomicAdd(s_wcnt, 1u)] = make_uint2(an, off);
  }
}

// queue B: continue from the depth-8 node along goto edges, one step per loop iteration, until a node with outputs
// (flag the token) or the end of the path.  Entries are dealt round-robin, so with >= 32 entries every lane walks.
__device__ __forceinline__ void walk_anchors(const ScanArgs& a, const uint8_t* p, const uint32_t* s_start, const uint32_t* s_len,
                                             const uint2* s_walk, uint32_t cnt, uint32_t* s_hit, uint32_t lane) {
  const uint8_t* ac = a.db.pg + a.db.ac_start;
  uint32_t my = lane, j = 0, tn = 0, tl = 0;
  const uint8_t* text = nullptr;
  bool walking = false;
  AcNode nd{0, 0, 0, 0};
  for (;;) {
    if (!walking) {
      if (my >= cnt) break;
      uint2 e = s_walk[my];
      my += 32;
      tl = e.x & 31u;
      if ((*(volatile uint32_t*)s_hit >> tl) & 1u) continue;
      j = (e.x >> 5) + 8; tn = s_len[tl]; text = p + s_start[tl];
      nd = ac_fetch(ac, e.y);
      if (nd.w0 >> 24) { atomicOr(s_hit, 1u << tl); continue; }
      walking = j < tn;
      continue;
    }
    uint32_t nx = ac_goto(ac, nd, text[j]);
    if (!nx) { walking = false; continue; }
    nd = ac_fetch(ac, nx);
    if (nd.w0 >> 24) { atomicOr(s_hit, 1u << tl); walking = false; continue; }
    if (++j >= tn) walking = false;
  }
}

__device__ __forceinline__ void drain_anchors(const ScanArgs& a, const uint8_t* p, const uint32_t* s_start,
                                              const uint32_t* s_len, const uint32_t* s_anchor, uint32_t cnt, uint2* s_walk, uint32_t* s_wcnt,
                                              uint32_t* s_hit, uint32_t lane) {
  for (uint32_t from = 0; from < cnt; from += WALK_CAP) {  // a slice of A can put at most WALK_CAP entries into B
    uint32_t to = min(from + WALK_CAP, cnt);
    filter_anchors(a, p, s_start, s_len, s_anchor, from, to, s_walk, s_wcnt, s_hit, lane);
    __syncwarp();
    walk_anchors(a, p, s_start, s_len, s_walk, *(volatile uint32_t*)s_wcnt, s_hit, lane);
    __syncwarp();
    if (lane == 0) *s_wcnt = 0;
    __syncwarp();
  }
}

static const uint32_t ACG_WIN = 2048;  // token window of this kernel (smaller than WIN_BYTES: shared memory buys occupancy here)
static const size_t ACGLOB_SMEM = KT_WARPS * (ACG_WIN + 32) + (256 + 2048 + KT_WARPS * (ANCHOR_CAP + 2 * WALK_CAP) + 2 * KT_WARPS * 32 + 3 * KT_WARPS) * 4;

__global__ void __launch_bounds__(KT_THREADS) acglob_kernel(ScanArgs a) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];  // ACGLOB_SMEM bytes, carved below
  uint8_t (*s_win)[ACG_WIN + 32] = reinterpret_cast<uint8_t (*)[ACG_WIN + 32]>(dyn_smem);
  uint32_t* s_root = reinterpret_cast<uint32_t*>(dyn_smem + KT_WARPS * (ACG_WIN + 32));
  uint32_t* s_gram2 = s_root + 256;
  uint32_t (*s_anchor)[ANCHOR_CAP] = reinterpret_cast<uint32_t (*)[ANCHOR_CAP]>(s_gram2 + 2048);
  uint2 (*s_walk)[WALK_CAP] = reinterpret_cast<uint2 (*)[WALK_CAP]>(s_gram2 + 2048 + KT_WARPS * ANCHOR_CAP);
  uint32_t (*s_start)[32] = reinterpret_cast<uint32_t (*)[32]>(s_gram2 + 2048 + KT_WARPS * (ANCHOR_CAP + 2 * WALK_CAP));
  uint32_t (*s_len)[32] = s_start + KT_WARPS;
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_len + KT_WARPS);
  uint32_t* s_hit = s_cnt + KT_WARPS;
  uint32_t* s_wcnt = s_hit + KT_WARPS;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (a.ctr->overflow) return;
  // hot state in shared memory: the root's dense table (every walk starts there) and the 2-byte prefix bitmap that
  // almost every token position fails
  AcAccel acc;
  acc.root_tab = ac_root_table(a.db);
  acc.gram2 = nullptr;
  if (acc.root_tab) { s_root[threadIdx.x] = acc.root_tab[threadIdx.x]; acc.root_tab = s_root; }
  if (a.db.has_glob && a.db.ac_gram2) {
    for (uint32_t k = threadIdx.x; k < 2048; k += blockDim.x) s_gram2[k] = a.db.ac_gram2[k];
    acc.gram2 = s_gram2;
  }
  __syncthreads();
  const bool use_filter = a.db.has_glob && a.db.ac_anchored && acc.gram2 && a.db.wild_count == 0 && a.db.ac_size >= 20 && a.db.match_mode == 0;
  const uint32_t n = min(a.ctr->n_str, a.cap_str);
  const uint32_t nround = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
    bool valid = i < n;
    StrTok t{0, 0, 0};
    if (valid) { t = a.str[i]; valid = t.type != TOK_INVALID; }
    const uint8_t* p = stage_tokens(a, s_win[warp], valid, t, lane, ACG_WIN);
    bool exact = valid && a.db.has_glob;
    if (use_filter) {
      // ---- (1) anchors ----
      s_start[warp][lane] = t.start; s_len[warp][lane] = valid ? t.len : 0u;
      if (lane == 0) { s_cnt[warp] = 0; s_hit[warp] = 0; s_wcnt[warp] = 0; }
      __syncwarp();
      const uint8_t* text = p + t.start;
      const uint32_t tn = valid ? t.len : 0u;
      const uint32_t npos = tn >= 3 ? tn - 2 : 0u;  // start positions with at least 3 bytes left
      uint32_t qn = 0;                              // entries in queue A (warp-uniform)
      uint32_t g = npos ? (uint32_t)text[0] : 0u;
      for (uint32_t pos0 = 0; __any_sync(0xFFFFFFFFu, pos0 < npos); pos0 += SCAN_ROUND) {
        // one round: SCAN_ROUND positions per lane, branch-free; hits are collected in a register mask
        uint32_t hits = 0;
#pragma unroll
        for (uint32_t r = 0; r < SCAN_ROUND; r++) {
          uint32_t pos = pos0 + r;
          if (pos < npos) {
            g = ((g << 8) | text[pos + 1]) & 0xFFFFu;  // bytes pos, pos+1
            hits |= ((acc.gram2[g >> 5] >> (g & 31)) & 1u) << r;
          }
        }
        uint32_t cnt = __popc(hits), incl = warp_incl_scan(cnt, lane);
        uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t at = qn + incl - cnt;
        while (hits) { uint32_t r = __ffs(hits) - 1; hits &= hits - 1; s_anchor[warp][at++] = ((pos0 + r) << 5) | lane; }
        qn += total;
        __syncwarp();
        if (qn + 32 * SCAN_ROUND > ANCHOR_CAP) {  // ---- (2) filters + walks, when the next round might not fit ----
          drain_anchors(a, p, s_start[warp], s_len[warp], s_anchor[warp], qn, s_walk[warp], &s_wcnt[warp], &s_hit[warp], lane);
          qn = 0;
        }
      }
      if (lane == 0) s_cnt[warp] = qn;
      __syncwarp();
      drain_anchors(a, p, s_start[warp], s_len[warp], s_anchor[warp], s_cnt[warp], s_walk[warp], &s_wcnt[warp], &s_hit[warp], lane);
      __syncwarp();
      exact = valid && ((s_hit[warp] >> lane) & 1u);
      __syncwarp();
    }
    if (!valid) continue;
    const uint8_t* text = p + t.start;
    uint32_t lit_pid = a.db.has_literal ? a.lh_res[i] : NONE32, lit_off = 0;
    bool lit_ok = lit_pid != NONE32 && lh_data_offset(a.db, lit_pid, lit_off);
    if (!lit_ok && !exact) continue;
    emit_string_match(a, t, text, lit_ok, lit_pid, lit_off, exact, acc);
  }
}

// All 32 lanes of warp `w` of the block meet here, however they arrive (one at a time is fine), and their shared-memory
// writes are visible afterwards: a named hardware barrier (barrier.sync without .aligned), ids 1..15.  Unlike __syncwarp()
// it cannot be optimised away on the strength of an assumed reconvergence.
__device__ __forceinline__ void warp_rendezvous(uint32_t w) { asm volatile("barrier.sync %0, 32;" ::"r"(w + 1) : "memory"); }

// K3 on the fast path: the few string tokens whose filters passed (StrTok.type carries F_LIT / F_GLOB) get the exact
// LiteralHash::lookup and Paraglob::find_all, straight from the log buffer.  One WARP per token: the lanes share the
// start positions of the anchored literal search (positions are independent, anchored_visit_at), pattern ids meet in a
// small shared-memory list, lane 0 sorts, deduplicates and writes the record.
static const uint32_t EX_IDS = 96;
__global__ void __launch_bounds__(256, 5) exact_kernel(ScanArgs a) {
  __shared__ uint32_t s_ids[8][EX_IDS];
  __shared__ uint32_t s_n[8];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = min(a.ctr->n_str, a.cap_str);
  if (a.ctr->overflow) return;
  AcAccel acc;
  acc.root_tab = ac_root_table(a.db);
  acc.gram2 = a.db.ac_gram2;
  const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + warp; i < n; i += nwarps) {
    StrTok t = a.str[i];
    if ((t.type & 0xFFu) == TOK_INVALID) continue;
    const uint32_t f = t.type;
    t.type &= 0xFFu;
    const uint8_t* text = a.buf + t.start;
    uint32_t lit_pid = NONE32, lit_off = 0;  // every lane computes the same probe: the loads coalesce into broadcasts
    const bool lit_ok = (f & F_LIT) && a.db.has_literal && lh_lookup(a.db, text, t.len, lit_pid) && lh_data_offset(a.db, lit_pid, lit_off);
    const bool exact = (f & F_GLOB) && a.db.has_glob;
    if (!lit_ok && !exact) continue;
    uint32_t cnt = 0;
    bool serial = exact && !(a.db.ac_anchored && a.db.wild_count == 0 && a.db.ac_size >= 20);
    if (exact && !serial) {
      // (warp_rendezvous, not __syncwarp: the lanes come out of divergent walks, and what follows only needs every lane's
      // shared-memory writes to have happened — see iptrie_kernel)
      if (lane == 0) s_n[warp] = 0;
      warp_rendezvous(warp);
      for (uint32_t pos = lane; pos + 3 <= t.len; pos += 32)
        anchored_visit_at(a.db, text, t.len, pos, acc.gram2, [&](uint32_t pid) { uint32_t k = atomicAdd(&s_n[warp], 1u); if (k < EX_IDS) s_ids[warp][k] = pid; });
      warp_rendezvous(warp);
      cnt = s_n[warp];
      if (cnt > EX_IDS) serial = true;  // more ids than the list holds: let one lane redo it with the two-pass emitter
      warp_rendezvous(warp);  // every lane has read the count before lane 0 goes on (and the next token resets it)
    }
    if (lane != 0) continue;
    if (serial) { emit_string_match(a, t, text, lit_ok, lit_pid, lit_off, true, acc); continue; }
    // sort_unstable + dedup (paraglob_offset.rs:1173-1181)
    uint32_t* ids = s_ids[warp];
    for (uint32_t x = 1; x < cnt; x++) {
      uint32_t v = ids[x], y = x;
      while (y > 0 && ids[y - 1] > v) { ids[y] = ids[y - 1]; y--; }
      ids[y] = v;
    }
    uint32_t u = 0;
    for (uint32_t x = 0; x < cnt; x++) if (x == 0 || ids[x] != ids[u - 1]) ids[u++] = ids[x];
    const uint32_t total = u + (lit_ok ? 1u : 0u);
    if (total == 0) continue;
    const uint32_t b = atomicAdd(&a.tot->n_ids, total);
    if ((uint64_t)b + total > a.cap_ids) { atomicOr(&a.ctr->overflow, 1u << 11); continue; }
    uint32_t k = b;
    if (lit_ok) { a.ids[k].pattern_id = lit_pid; a.ids[k].data_offset = lit_off; k++; }
    for (uint32_t x = 0; x < u; x++, k++) {
      uint32_t off;
      a.ids[k].pattern_id = ids[x];
      a.ids[k].data_offset = glob_data_offset(a.db, ids[x], off) ? off : MGPU_NO_DATA;
    }
    const uint32_t r_i = atomicAdd(&a.tot->n_rec, 1u);
    if (r_i >= a.cap_rec) { atomicOr(&a.ctr->overflow, 1u << 10); continue; }
    mgpu_match r;
    r.offset = a.base + t.start; r.len = t.len; r.item_type = (uint8_t)t.type; r.kind = MGPU_KIND_PATTERN; r.prefix_len = 0; r.reserved = 0;
    r.n_ids = total; r.ids_index = b; r.data_offset = MGPU_NO_DATA; r.pad = 0;
    a.recs[r_i] = r;
  }
}

// end of a piece: snapshot the running totals into the piece's counter block
// Result order.  The C ABI hands out records sorted by (offset, item_type, len); the kernels append them with atomics.  A piece's
// records are sorted on the device before they leave it (pieces follow each other in offset order, so the batch arrives sorted):
// 64-bit key = (offset - piece start) << 4 | item_type (a position yields at most one token per type, so len never decides),
// value = record index, cub::DeviceRadixSort over the <= 35 bits in use, then the 32-byte records gathered in that order.  The host only checks the order (host_sort.h: records_sorted)
// and keeps its own sort for what did not come this way (pieces redone after an overflow).
__global__ void sort_keys_kernel(const mgpu_match* recs, uint32_t n, uint64_t lo, uint64_t* keys, uint32_t* vals) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 h = reinterpret_cast<const uint4*>(recs)[2 * (size_t)i];  // offset (8 bytes), len, item_type | kind << 8 | ...
    const uint64_t off = ((uint64_t)h.y << 32) | h.x;
    keys[i] = ((off - lo) << 4) | (uint64_t)(h.w & 0xFu);
    vals[i] = i;
  }
}
__global__ void sort_gather_kernel(const mgpu_match* recs, uint32_t n, const uint32_t* vals, mgpu_match* out) {
  const uint4* src = reinterpret_cast<const uint4*>(recs);
  uint4* dst = reinterpret_cast<uint4*>(out);
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < 2 * (size_t)n; j += (size_t)gridDim.x * blockDim.x)
    dst[j] = src[2 * (size_t)vals[j >> 1] + (j & 1)];
}
// (h_nrec: the same snapshot in pinned host memory — the host reads it as soon as the piece's event has fired and starts the
//  copy of the piece's records while later pieces are still being scanned, see drain_records)
__global__ void piece_end_kernel(DevCounters* ctr, const ScanTotals* tot, uint32_t* h_nrec) {
  ctr->n_rec = tot->n_rec; ctr->n_ids = tot->n_ids;
  if (h_nrec) *h_nrec = tot->n_rec;
}

// Cut points of a resident buffer, all at once: out[k] = position just after the last '\n' in [at[k] - span, at[k]), or
// NONE64 when that window holds none.  One block per cut.
static const uint64_t NONE64 = ~0ULL;
__global__ void cuts_back_kernel(const uint8_t* buf, const uint64_t* at, uint64_t span, uint64_t* out) {
  __shared__ unsigned long long best;
  const uint64_t hi = at[blockIdx.x], lo = hi > span ? hi - span : 0;
  uint64_t end = hi;
  while (end > lo) {
    uint64_t beg = end - lo > 8192 ? end - 8192 : lo;
    if (threadIdx.x == 0) best = 0;
    __syncthreads();
    unsigned long long mine = 0;
    for (uint64_t i = beg + threadIdx.x; i < end; i += blockDim.x) if (buf[i] == '\n') mine = i + 1;
    if (mine) atomicMax(&best, mine);
    __syncthreads();
    unsigned long long b = best;
    __syncthreads();
    if (b) { if (threadIdx.x == 0) out[blockIdx.x] = b; return; }
    end = beg;
  }
  if (threadIdx.x == 0) out[blockIdx.x] = NONE64;
}

// position just after the first '\n' at or after `from` (or n): where a newline-aligned cut may be made
__global__ void cut_kernel(const uint8_t* buf, uint64_t n, const uint64_t* from, uint64_t* out, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint64_t p = from[i];
  while (p < n && buf[p] != '\n') p++;
  out[i] = p < n ? p + 1 : n;
}

// position just after the LAST '\n' in buf[lo, hi), or 0 if there is none (one block)
__global__ void rfind_nl_kernel(const uint8_t* buf, uint64_t lo, uint64_t hi, uint64_t* out) {
  __shared__ unsigned long long best;
  uint64_t end = hi;
  while (end > lo) {
    uint64_t beg = end - lo > 8192 ? end - 8192 : lo;
    if (threadIdx.x == 0) best = 0;
    __syncthreads();
    unsigned long long mine = 0;
    for (uint64_t i = beg + threadIdx.x; i < end; i += blockDim.x) if (buf[i] == '\n') mine = i + 1;
    if (mine) atomicMax(&best, mine);
    __syncthreads();
    unsigned long long b = best;
    __syncthreads();
    if (b) { if (threadIdx.x == 0) *out = b; return; }
    end = beg;
  }
  if (threadIdx.x == 0) *out = 0;
}

__global__ void fill_kernel(uint4* p, size_t n16, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(v, v, v, v);
}

// Debug (mgpu_set_option "verify_tokens"): audit a piece's IP token list after the token kernel.  The list is poisoned before
// the token kernel runs; every slot below n_ip must then be a token or a padding slot, the tokens must add up to the
// per-type counters, and their lookups (recomputed here, one thread per slot, plain atomics) must add up to the IP records
// the IP-trie kernel emits.  dbg: [0] slots still poisoned, [1] padding, [2] IPv4 tokens, [3] IPv6 tokens, [4] lookup hits,
// [5] slots of any other type, [6] slots audited.
static const uint32_t TOK_POISON = 0xEEEEEEEEu;
__global__ void verify_tokens_kernel(ScanArgs a, unsigned long long* dbg) {
  const uint32_t n = min(a.ctr->n_ip, a.cap_ip);
  unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    IpTok t = a.ip[i];
    uint32_t off = 0; uint8_t pl = 0;
    if (t.type == TOK_RAW4) {  // (unparsed numeric word of the fused scan kernel: count it as what iptrie_kernel will make of it)
      uint32_t addr = 0;
      if (parse_ipv4_hexdot(t.w, t.len, addr)) { t.type = MGPU_T_IPV4; t.w[0] = addr; } else t.type = TOK_INVALID;
    }
    if (t.type == TOK_POISON) c[0]++;
    else if (t.type == TOK_INVALID) c[1]++;
    else if (t.type == MGPU_T_IPV4) { c[2]++; if (a.db.has_ip && trie_lookup_v4(a.db, t.w[0], off, pl)) c[4]++; }
    else if (t.type == MGPU_T_IPV6) {
      c[3]++;
      uint16_t seg[8];
      for (int k = 0; k < 4; k++) { seg[2 * k] = (uint16_t)(t.w[k] >> 16); seg[2 * k + 1] = (uint16_t)t.w[k]; }
      if (a.db.has_ip && trie_lookup_v6(a.db, seg, off, pl)) c[4]++;
    } else c[5]++;
  }
  for (int k = 0; k < 6; k++) {
    unsigned long long v = c[k];
    for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&dbg[k], v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&dbg[6], (unsigned long long)n);
}

// single-query lookups through the same device functions (Database::lookup / lookup_ip)
__global__ void query_string_kernel(ScanArgs a, uint32_t len, uint32_t* out_n) {
  uint32_t n = 0;
  uint32_t pid, off;
  if (a.db.has_literal && lh_lookup(a.db, a.buf, len, pid) && lh_data_offset(a.db, pid, off)) { if (n < a.cap_ids) { a.ids[n].pattern_id = pid; a.ids[n].data_offset = off; } n++; }
  uint32_t g0 = n;
  AcAccel acc;
  acc.root_tab = ac_root_table(a.db); acc.gram2 = a.db.ac_gram2;  // global copies: one thread, no shared-memory staging
  if (a.db.has_glob) find_all_visit(a.db, a.buf, len, acc, [&](uint32_t p) { if (n < a.cap_ids) a.ids[n].pattern_id = p; n++; });
  uint32_t k = n < a.cap_ids ? n : a.cap_ids;
  for (uint32_t x = g0 + 1; x < k; x++) {
    uint32_t v = a.ids[x].pattern_id, y = x;
    while (y > g0 && a.ids[y - 1].pattern_id > v) { a.ids[y].pattern_id = a.ids[y - 1].pattern_id; y--; }
    a.ids[y].pattern_id = v;
  }
  uint32_t u = g0;
  for (uint32_t x = g0; x < k; x++) if (x == g0 || a.ids[x].pattern_id != a.ids[u - 1].pattern_id) a.ids[u++].pattern_id = a.ids[x].pattern_id;
  for (uint32_t x = g0; x < u; x++) { uint32_t o; a.ids[x].data_offset = glob_data_offset(a.db, a.ids[x].pattern_id, o) ? o : MGPU_NO_DATA; }
  *out_n = (n > a.cap_ids) ? NONE32 : u;
}
__global__ void query_ip_kernel(ScanArgs a, int is_v6, uint32_t* out3) {
  uint32_t off = 0; uint8_t pl = 0; bool hit = false;
  if (a.db.has_ip) {
    if (!is_v6) hit = trie_lookup_v4(a.db, ((uint32_t)a.buf[0] << 24) | ((uint32_t)a.buf[1] << 16) | ((uint32_t)a.buf[2] << 8) | a.buf[3], off, pl);
    else {
      uint16_t seg[8];
      for (int k = 0; k < 8; k++) seg[k] = (uint16_t)(((uint32_t)a.buf[2 * k] << 8) | a.buf[2 * k + 1]);
      hit = trie_lookup_v6(a.db, seg, off, pl);
    }
  }
  out3[0] = hit; out3[1] = off; out3[2] = pl;
}

}  // namespace mgpu

// =========================================================================================================
// Host side
// =========================================================================================================
using namespace mgpu;

static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
      return MGPU_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

// Growable array in pinned host memory (results land here by DMA; pageable std::vector storage would be staged by the
// driver at a fraction of the PCIe rate).  The storage is kept between scans.
template <typename T>
struct PinnedVec {
  T* p = nullptr; size_t n = 0, cap = 0;
  ~PinnedVec() { if (p) cudaFreeHost(p); }
  bool reserve(size_t want) {
    if (want <= cap) return true;
    size_t nc = std::max(want, cap * 2);
    if (nc < 4096) nc = 4096;
    T* q = nullptr;
    if (cudaMallocHost(&q, nc * sizeof(T)) != cudaSuccess) return false;
    if (n) memcpy(q, p, n * sizeof(T));
    if (p) cudaFreeHost(p);
    p = q; cap = nc;
    return true;
  }
  bool resize(size_t m) { if (!reserve(m)) return false; n = m; return true; }
  void clear() { n = 0; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T* data() { return p; }
  T* begin() { return p; }
  T* end() { return p + n; }
  T& operator[](size_t i) { return p[i]; }
};

struct mgpu_ctx {
  static const int MAX_BATCH = 32;  // pieces launched back to back before the host looks at their counters
  int device = 0;
  int sm_count = 148;
  size_t chunk_bytes = 0;
  cudaStream_t compute = nullptr, copy = nullptr;
  cudaStream_t lookup = nullptr;  // IP-trie and exact string lookups of piece i run here, beside the tokenizer of piece i+1
  cudaEvent_t ev_tokens[MAX_BATCH] = {}, ev_looked[MAX_BATCH] = {};  // token lists of the slot's piece written / consumed
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_k[MAX_BATCH][MGPU_K_COUNT + 1] = {};
  cudaEvent_t ev_l0[MAX_BATCH] = {};  // on the lookup stream, after it has waited for the piece's tokens: where the IP-trie kernel's time starts
  cudaEvent_t ev_scan[2] = {nullptr, nullptr};
  // log staging (double buffered) and pinned bounce buffers for pageable callers
  uint8_t* d_log[2] = {nullptr, nullptr};
  uint8_t* h_pin[2] = {nullptr, nullptr};
  // work buffers
  ScanArgs args;  // device pointers + capacities (buf/n/base/flags filled per chunk)
  DevCounters* h_ctr = nullptr;  // pinned, MAX_BATCH entries (args.ctr points at the device copy of entry 0)
  ScanTotals* d_tot = nullptr;
  uint64_t* d_cut = nullptr; uint64_t* h_cut = nullptr;
  uint8_t* d_small = nullptr; uint32_t* d_small_out = nullptr;
  void* d_flush = nullptr;
  // database
  bool db_loaded = false;
  std::vector<void*> db_allocs;
  mxy::Layout layout;
  mgpu_db_info info{};
  // PSL
  std::vector<uint8_t> psl_text;
  void* d_psl_keys = nullptr; void* d_psl_vals = nullptr; void* d_psl_pool = nullptr; void* d_psl_tld = nullptr;
  // results of the last scan
  PinnedVec<mgpu_match> recs;
  PinnedVec<mgpu_match> stage;  // a batch's records in pinned HOST memory, in the order the device appended them
  // How they get there.  Until the round's last session the lookup kernels stored them straight into `stage` over PCIe (no device-side
  // record buffer).  ncu then showed what that costs when a database matches often: iptrie_kernel on config 3 ran for 403 us per
  // 500 MB piece at pcie__write_bytes = 46 GB/s — the kernel's length WAS the 15 MB of records crossing the link, with its blocks
  // holding registers the next piece's tokenizer was waiting for.  Now the kernels append to d_recs in HBM, piece_end_kernel leaves
  // the running record count in pinned memory, and the host — idle while a batch runs — copies every piece's records with the
  // copy engine (stream d2h) as soon as the piece's lookups are done, beside the kernels of the pieces behind it.
  // MATCHY_B200_RECS_ZEROCOPY=1 restores the direct stores (d_recs stays null).
  mgpu_match* d_recs = nullptr;
  uint32_t* h_nrec = nullptr;   // pinned, MAX_BATCH entries
  cudaStream_t d2h = nullptr;
  cudaEvent_t ev_d2h = nullptr;
  // device-side result sort (see sort_keys_kernel): key / index double buffers, the sorted records, cub's scratch
  uint64_t* d_keys[2] = {nullptr, nullptr};
  uint32_t* d_vals[2] = {nullptr, nullptr};
  mgpu_match* d_sorted = nullptr;   // null: records leave the device in append order (MATCHY_B200_HOST_SORT=1) and the host sorts
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  // A piece is sorted on the device when it produced at least this many records (option "device_sort_min", MATCHY_B200_DEVICE_SORT_MIN).
  // Measured (scripts/gpu_r2am.sh, device-timed / wall-clock GB/s): config 3, 2 M records per piece: 696 / 206 with the host sort,
  // 572 / 484 with the device sort; config 5: 539 / 274 -> 514 / 422; config 2, 10 K records per piece: 928 / 758 -> 840 / 772 — the
  // sort's kernels take SM slots from the single-wave scan kernels, which a 1 ms host sort of 53 K records is not worth.
  uint32_t dev_sort_min = 131072;
  bool dev_sorting = false;     // a piece of the current scan reached dev_sort_min: the pieces behind it are sorted on the device too, whatever
                                // their size (a short last piece in append order would send the whole result through the host's sort)
  bool order_ok = true;         // current scan: every piece that produced records left the device sorted, and the pieces were gathered in
                                // offset order (nothing redone after later pieces): the result is sorted by construction.  finish_scan
                                // re-checks results of up to 2^20 records (or all, with MATCHY_B200_VERIFY_ORDER=1) and trusts it above that:
                                // reading 41 M records once more cost 29 ms of a 212 ms config-5 step
  bool verify_order = false;
  bool arrived_sorted = false;  // the last scan's records were in order when the host looked (no host sort ran)
  uint64_t piece_base[MAX_BATCH] = {}, piece_len[MAX_BATCH] = {};  // absolute offset of the slot's piece buffer and its length (the sort keys are relative to it)
  PinnedVec<mgpu_id_pair> ids;
  std::vector<mgpu_id_pair> ids_tmp;
  std::vector<mgpu_match> sort_tmp;   // sort_records' scratch
  uint32_t sub_block = 0;             // ScanArgs::sub_block of the two-kernel path (option "sub_block"; MATCHY_B200_SUB_BLOCK).  Measured, config 2 at 8 GB:
                                      // token_kernel 4.29 ms with 0, 5.61 / 4.89 / 4.63 ms with 8 / 16 / 32 KiB — the short groups at every sub-block edge cost more than the L2 hits bring
  int iptrie_minb = 4;                // iptrie_kernel instance: 4 = 64 registers, 5 = 48, 6 = 40 (MATCHY_B200_IPTRIE_MINB; config 3 at 8 GB: 588 / 577 / 526 GB/s — the spills cost more than the warps bring)
  uint64_t scan_lo = 0, scan_hi = ~0ull;  // absolute offsets the current scan can report (the range of sort_records' partition)
  uint64_t host_us[4] = {0, 0, 0, 0}; // microseconds of the last mgpu_scan_device on the host: whole call, sort, id re-pack, launches + gather
  std::unique_ptr<WorkerPool> pool;   // host threads of the result sort, started by the first scan that returns >= 4096 records
  mgpu_counters counters{};
  mgpu_timing timing{};
  bool keep_results = true;
  bool force_ac_walk = false;
  bool looked_pending = false; int looked_slot = 0;  // the lookup stream still owns the token lists (event ev_looked[looked_slot])
  bool force_generic = false;  // tests: run the generic string path (lithash + acglob kernels) even when the fast path applies
  std::vector<StrTok> x_str; std::vector<IpTok> x_ip;  // extraction-only capture
  bool capture_tokens = false;
  // test / debug switches (mgpu_set_option)
  bool fused = true;  // scan_kernel (one pass) instead of tokenize_kernel + the word passes of token_kernel
  bool serial = false;  // everything on one stream (MATCHY_B200_SERIAL=1): per-kernel times without scheduling waits
  int wide_warps = SK_WARPS_NOHOT;
  bool wide_scan = false;  // scan_kernel always as the 24-warp block that reads the hot filter through L1 (MATCHY_B200_WIDE_SCAN=1)
  // Working sets: the candidate queues and token lists of piece i are read by its lookups while piece i+1 is scanned, so
  // there are two of each (fused mode) and a piece uses set (slot & 1); set_slot[s] = batch slot of the piece that used set s last.
  struct BufSet { Cand* q_dotted; Cand* q_hash; uint32_t* q_at; uint32_t* q_c2; Cand* q_numeric; Cand* q_long; uint32_t* seg_cnt; StrTok* str; IpTok* ip; };
  BufSet sets[2] = {};
  int nsets = 1;
  bool set_pending[2] = {false, false}; int set_slot[2] = {0, 0};
  bool verify_tokens = false;
  unsigned long long* d_dbg = nullptr;  // 16 audit accumulators (verify_tokens_kernel: 0..6, iptrie_kernel: 8..10)
  uint32_t alloc_cap_str = 0, alloc_cap_ip = 0, alloc_cap_rec = 0, alloc_cap_ids = 0;  // what the buffers really hold
};

static int launch_grid(mgpu_ctx* c, int per_sm) { return c->sm_count * per_sm; }

extern "C" {

const char* mgpu_last_error(void) { return g_err.c_str(); }

static void free_db(mgpu_ctx* c) {
  for (void* p : c->db_allocs) cudaFree(p);
  c->db_allocs.clear();
  c->db_loaded = false;
}

void mgpu_destroy(mgpu_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_db(c);
  for (int s = 0; s < 2; s++) {
    if (c->d_log[s]) cudaFree(c->d_log[s]);
    if (c->h_pin[s]) cudaFreeHost(c->h_pin[s]);
    if (c->ev_copied[s]) cudaEventDestroy(c->ev_copied[s]);
    if (c->ev_free[s]) cudaEventDestroy(c->ev_free[s]);
  }
  for (auto& row : c->ev_k) for (auto& e : row) if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_l0) if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_scan) if (e) cudaEventDestroy(e);
  for (auto& b : c->sets) { void* q[] = {b.q_dotted, b.q_hash, b.q_at, b.q_c2, b.q_numeric, b.q_long, b.seg_cnt, b.str, b.ip}; for (void* p : q) if (p) cudaFree(p); }
  void* bufs[] = {c->args.defer, c->args.defer_tk, c->args.lh_res,
                  c->args.ids, c->args.ctr, c->d_tot, c->d_cut, c->d_small, c->d_small_out, c->d_dbg, c->d_flush, c->d_psl_keys, c->d_psl_vals, c->d_psl_pool, c->d_psl_tld};
  for (void* p : bufs) if (p) cudaFree(p);
  if (c->h_ctr) cudaFreeHost(c->h_ctr);
  if (c->h_nrec) cudaFreeHost(c->h_nrec);
  if (c->d_recs) cudaFree(c->d_recs);
  if (c->d2h) cudaStreamDestroy(c->d2h);
  { void* sb[] = {c->d_keys[0], c->d_keys[1], c->d_vals[0], c->d_vals[1], c->d_sorted, c->d_sort_tmp}; for (void* p : sb) if (p) cudaFree(p); }
  if (c->ev_d2h) cudaEventDestroy(c->ev_d2h);
  if (c->h_cut) cudaFreeHost(c->h_cut);
  if (c->compute) cudaStreamDestroy(c->compute);
  if (c->copy) cudaStreamDestroy(c->copy);
  if (c->lookup) cudaStreamDestroy(c->lookup);
  for (auto& e : c->ev_tokens) if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_looked) if (e) cudaEventDestroy(e);
  delete c;
}

static int create_impl(mgpu_ctx* c, int device, size_t chunk_bytes) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_err("no such CUDA device"); return MGPU_E_PARAM; }
  c->device = device;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { set_err("matchy_b200 needs an sm_100-class GPU"); return MGPU_E_CUDA; }
  c->sm_count = prop.multiProcessorCount;
  if (chunk_bytes == 0) chunk_bytes = (size_t)256 << 20;
  chunk_bytes = (chunk_bytes + TILE_BYTES - 1) / TILE_BYTES * TILE_BYTES;
  if (chunk_bytes > ((size_t)2 << 30)) chunk_bytes = (size_t)2 << 30;
  if (chunk_bytes < (size_t)64 << 10) chunk_bytes = (size_t)64 << 10;
  c->chunk_bytes = chunk_bytes;
  CK(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->lookup, cudaStreamNonBlocking));
  for (auto& e : c->ev_tokens) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : c->ev_looked) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (int s = 0; s < 2; s++) {
    CK(cudaEventCreateWithFlags(&c->ev_copied[s], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_free[s], cudaEventDisableTiming));
    CK(cudaMalloc(&c->d_log[s], chunk_bytes + TILE_BYTES));
    CK(cudaMallocHost(&c->h_pin[s], std::min(chunk_bytes, (size_t)64 << 20)));
  }
  for (auto& row : c->ev_k) for (auto& e : row) CK(cudaEventCreate(&e));
  for (auto& e : c->ev_l0) CK(cudaEventCreate(&e));
  for (auto& e : c->ev_scan) CK(cudaEventCreate(&e));
  ScanArgs& a = c->args;
  memset(&a, 0, sizeof a);
  auto cap32 = [](size_t v) { return (uint32_t)std::min<size_t>(v, 0x7FFFFFFFu); };
  // Which first stage: tokenize_kernel + token_kernel (two passes over the log, 32 warps per SM each — the faster pair on the
  // BASELINE configs, measured: DESIGN.md §4) or the single-pass scan_kernel (MATCHY_B200_FUSED=1; one pass, 1.2 DRAM bytes per
  // log byte, but 16..28 warps per SM).  The work buffers differ, so the choice is made when the context is created.
  c->fused = getenv("MATCHY_B200_FUSED") != nullptr && getenv("MATCHY_B200_UNFUSED") == nullptr;
  c->nsets = c->fused ? 2 : 1;
  c->wide_scan = getenv("MATCHY_B200_HOT_SMEM") == nullptr;  // (28-warp block reading the hot filter through L1: measured 15 % faster than 16 warps + shared-memory copy)
  c->wide_warps = 28;
  c->serial = getenv("MATCHY_B200_SERIAL") != nullptr;
  if (const char* mb = getenv("MATCHY_B200_IPTRIE_MINB")) c->iptrie_minb = atoi(mb);
  if (const char* vv = getenv("MATCHY_B200_VARIANT")) a.variant = (uint32_t)atoi(vv);  // (experiment switch, see ScanArgs::variant)
  if (const char* sb = getenv("MATCHY_B200_SUB_BLOCK")) c->sub_block = (uint32_t)atoi(sb);
  if (const char* ww = getenv("MATCHY_B200_WIDE_WARPS")) { int v = atoi(ww); if (v == 24 || v == 28 || v == 32) c->wide_warps = v; }
  // Candidate queue segments in HBM: one per scanning warp.  Capacities follow what a log can plausibly hold; a piece that
  // needs more sets its overflow flag and is split at newlines and redone (scan_piece), so exactness never depends on them.
  // No segment is smaller than the worst case of a single 1 KiB tile, so that splitting always ends overflows.
  if (c->fused) {
    // scan_kernel keeps dotted / numeric / hash words in shared memory; HBM sees the slow path and the rare anchors only
    a.nseg_max = (uint32_t)launch_grid(c, 1) * 32;
    const size_t share = chunk_bytes / a.nseg_max;
    a.seg_cap[Q_DOTTED] = cap32(std::max<size_t>(share / 64 + 64, 288));
    a.seg_cap[Q_HASH] = cap32(std::max<size_t>(share / 1024 + 32, 64));
    a.seg_cap[Q_AT] = cap32(std::max<size_t>(share / 64 + 64, 1024));
    a.seg_cap[Q_COLON2] = cap32(std::max<size_t>(share / 64 + 64, 512));
    a.seg_cap[Q_NUMERIC] = cap32(std::max<size_t>(share / 1024 + 32, 64));
    a.seg_cap[Q_LONG] = cap32(std::max<size_t>(share / 128 + 8, 40));
    a.cap_str = cap32(chunk_bytes / 32 + 65536);
    a.cap_ip = cap32(chunk_bytes / 32 + 65536);
    a.cap_rec = cap32(chunk_bytes / 64 + 65536);
    a.cap_ids = cap32(chunk_bytes / 64 + 65536);
  } else {
    a.nseg_max = (uint32_t)launch_grid(c, 2) * K1_WARPS;
    const size_t share = chunk_bytes / a.nseg_max;
    a.seg_cap[Q_DOTTED] = cap32(std::max<size_t>(share / 8 + 32, 256));
    a.seg_cap[Q_HASH] = cap32(std::max<size_t>(share / 33 + 8, 32));
    a.seg_cap[Q_AT] = cap32(std::max<size_t>(share / 16 + 32, 1024));
    a.seg_cap[Q_COLON2] = cap32(std::max<size_t>(share / 16 + 32, 512));
    a.seg_cap[Q_NUMERIC] = cap32(std::max<size_t>(share / 8 + 32, 256));
    a.seg_cap[Q_LONG] = cap32(std::max<size_t>(share / 27 + 8, 40));
    a.cap_str = cap32(chunk_bytes / 8 + 1024);
    a.cap_ip = cap32(chunk_bytes / 8 + 1024);
    a.cap_rec = cap32(chunk_bytes / 64 + 65536);
    a.cap_ids = cap32(chunk_bytes / 8 + 8192);
  }
  a.tok_unit = TOK_RESERVE;
  if (const char* tu = getenv("MATCHY_B200_TOK_RESERVE")) { const int v = atoi(tu); if (v >= 32 && v <= 65536) a.tok_unit = (uint32_t)v; }
  c->alloc_cap_str = a.cap_str; c->alloc_cap_ip = a.cap_ip; c->alloc_cap_rec = a.cap_rec; c->alloc_cap_ids = a.cap_ids;
  for (int k = 0; k < c->nsets; k++) {
    mgpu_ctx::BufSet& b = c->sets[k];
    CK(cudaMalloc(&b.q_dotted, (size_t)a.seg_cap[Q_DOTTED] * a.nseg_max * sizeof(Cand)));
    CK(cudaMalloc(&b.q_hash, (size_t)a.seg_cap[Q_HASH] * a.nseg_max * sizeof(Cand)));
    CK(cudaMalloc(&b.q_at, (size_t)a.seg_cap[Q_AT] * a.nseg_max * 4));
    CK(cudaMalloc(&b.q_c2, (size_t)a.seg_cap[Q_COLON2] * a.nseg_max * 4));
    CK(cudaMalloc(&b.q_numeric, (size_t)a.seg_cap[Q_NUMERIC] * a.nseg_max * sizeof(Cand)));
    CK(cudaMalloc(&b.q_long, (size_t)a.seg_cap[Q_LONG] * a.nseg_max * sizeof(Cand)));
    CK(cudaMalloc(&b.seg_cnt, (size_t)Q_COUNT * a.nseg_max * 4));
    CK(cudaMalloc(&b.str, (size_t)a.cap_str * sizeof(StrTok)));
    CK(cudaMalloc(&b.ip, (size_t)a.cap_ip * sizeof(IpTok)));
  }
  {
    const mgpu_ctx::BufSet& b = c->sets[0];
    a.q_dotted = b.q_dotted; a.q_hash = b.q_hash; a.q_at = b.q_at; a.q_c2 = b.q_c2; a.q_numeric = b.q_numeric; a.q_long = b.q_long; a.seg_cnt = b.seg_cnt;
    a.str = b.str; a.ip = b.ip;
  }
  CK(cudaMalloc(&a.defer, (size_t)launch_grid(c, 1) * TK_WARPS * DEFER_CAP * sizeof(StrTok)));
  CK(cudaMalloc(&a.defer_tk, (size_t)launch_grid(c, 1) * TK_WARPS * DEFER_CAP * sizeof(StrTok)));
  CK(cudaMalloc(&a.lh_res, (size_t)a.cap_str * 4));
  // match records go straight to pinned host memory (unified addressing: the host pointer is the device pointer)
  if (!c->stage.reserve(a.cap_rec) || !c->recs.reserve(a.cap_rec)) { set_err("out of pinned host memory for match records"); return MGPU_E_CUDA; }
  a.recs = c->stage.p;
  if (getenv("MATCHY_B200_RECS_ZEROCOPY") == nullptr) {
    CK(cudaMalloc(&c->d_recs, (size_t)a.cap_rec * sizeof(mgpu_match)));
    CK(cudaMallocHost(&c->h_nrec, sizeof(uint32_t) * mgpu_ctx::MAX_BATCH));
    // (highest priority: the sort's short kernels take the first SM slots a kernel boundary of the scan frees — at default
    //  priority they starved behind the single-wave scan kernels until the batch ended, and the copies with them)
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&c->d2h, cudaStreamNonBlocking, getenv("MATCHY_B200_D2H_PRIO0") ? prio_lo : prio_hi));
    CK(cudaEventCreateWithFlags(&c->ev_d2h, cudaEventDisableTiming));
    a.recs = c->d_recs;
    if (getenv("MATCHY_B200_HOST_SORT") == nullptr) {
      for (int k = 0; k < 2; k++) { CK(cudaMalloc(&c->d_keys[k], (size_t)a.cap_rec * 8)); CK(cudaMalloc(&c->d_vals[k], (size_t)a.cap_rec * 4)); }
      CK(cudaMalloc(&c->d_sorted, (size_t)a.cap_rec * sizeof(mgpu_match)));
      cub::DoubleBuffer<uint64_t> dk(c->d_keys[0], c->d_keys[1]);
      cub::DoubleBuffer<uint32_t> dv(c->d_vals[0], c->d_vals[1]);
      CK(cub::DeviceRadixSort::SortPairs(nullptr, c->sort_tmp_bytes, dk, dv, (int)a.cap_rec, 0, 64, c->d2h));
      c->sort_tmp_bytes += 4096;
      if (const char* m = getenv("MATCHY_B200_DEVICE_SORT_MIN")) c->dev_sort_min = (uint32_t)strtoul(m, nullptr, 10);
      c->verify_order = getenv("MATCHY_B200_VERIFY_ORDER") != nullptr;
      CK(cudaMalloc(&c->d_sort_tmp, c->sort_tmp_bytes));
    }
  }
  CK(cudaMalloc(&a.ids, (size_t)a.cap_ids * sizeof(mgpu_id_pair)));
  CK(cudaMalloc(&a.ctr, sizeof(DevCounters) * mgpu_ctx::MAX_BATCH));
  CK(cudaMalloc(&c->d_tot, sizeof(ScanTotals)));
  a.tot = c->d_tot;
  CK(cudaMallocHost(&c->h_ctr, sizeof(DevCounters) * mgpu_ctx::MAX_BATCH));
  CK(cudaMalloc(&c->d_cut, 4096 * sizeof(uint64_t) * 2));
  CK(cudaMallocHost(&c->h_cut, 4096 * sizeof(uint64_t) * 2));
  CK(cudaMalloc(&c->d_small, 65536 + TILE_BYTES));
  CK(cudaMalloc(&c->d_small_out, 64));
  CK(cudaMalloc(&c->d_dbg, 64 * sizeof(unsigned long long)));
  CK(cudaMemset(c->d_dbg, 0, 64 * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(acglob_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACGLOB_SMEM));
  CK(cudaFuncSetAttribute(token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TOKEN_SMEM));
  CK(cudaFuncSetAttribute(scan_kernel<0, true, SK_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(true, SK_WARPS)));
  CK(cudaFuncSetAttribute(scan_kernel<K1_DEFAULT_FLAGS, true, SK_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(true, SK_WARPS)));
  CK(cudaFuncSetAttribute(scan_kernel<0, false, 28>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(false, 28)));
  CK(cudaFuncSetAttribute(scan_kernel<K1_DEFAULT_FLAGS, false, 28>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(false, 28)));
  CK(cudaFuncSetAttribute(scan_kernel<0, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(false, 32)));
  CK(cudaFuncSetAttribute(scan_kernel<K1_DEFAULT_FLAGS, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(false, 32)));
  CK(cudaFuncSetAttribute(scan_kernel<0, false, SK_WARPS_NOHOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(false, SK_WARPS_NOHOT)));
  CK(cudaFuncSetAttribute(scan_kernel<K1_DEFAULT_FLAGS, false, SK_WARPS_NOHOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem(false, SK_WARPS_NOHOT)));
  CK(cudaFuncSetAttribute(tokenize_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 * 8));
  CK(cudaFuncSetAttribute(tokenize_kernel<K1_DEFAULT_FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 * 8));
  return MGPU_OK;
}

mgpu_ctx* mgpu_create(int device, size_t chunk_bytes) {
  mgpu_ctx* c = new mgpu_ctx();
  if (create_impl(c, device, chunk_bytes) != MGPU_OK) { std::string keep = g_err; mgpu_destroy(c); g_err = keep; return nullptr; }
  return c;
}

void mgpu_set_keep_results(mgpu_ctx* c, int keep) { c->keep_results = keep != 0; }
void mgpu_set_ac_mode(mgpu_ctx* c, int mode) { c->force_ac_walk = mode == 1; c->force_generic = mode == 1 || mode == 2; }

// Test / debug switches.  "tok_reserve": slots per reservation of the token lists; "cap_str" / "cap_ip" / "cap_rec" /
// "cap_ids": pretend the work buffers are this small (never above what was allocated; 0 restores the allocation), which
// drives the overflow -> split-and-redo path; "verify_tokens": audit every piece's IP token list (verify_tokens_kernel);
// "device_sort_min": records a piece must produce to be sorted on the device (1: always — the tests; see mgpu_ctx::dev_sort_min).
int mgpu_set_option(mgpu_ctx* c, const char* key, uint64_t value) {
  if (!c || !key) { set_err("null argument"); return MGPU_E_PARAM; }
  const std::string k(key);
  auto cap = [&](uint32_t& field, uint32_t alloc) { field = value == 0 ? alloc : (uint32_t)std::min<uint64_t>(value, alloc); };
  if (k == "tok_reserve") { if (value < 32 || value > 65536) { set_err("tok_reserve out of range"); return MGPU_E_PARAM; } c->args.tok_unit = (uint32_t)value; }
  else if (k == "cap_str") cap(c->args.cap_str, c->alloc_cap_str);
  else if (k == "cap_ip") cap(c->args.cap_ip, c->alloc_cap_ip);
  else if (k == "cap_rec") cap(c->args.cap_rec, c->alloc_cap_rec);
  else if (k == "cap_ids") cap(c->args.cap_ids, c->alloc_cap_ids);
  else if (k == "verify_tokens") {
    c->verify_tokens = value != 0;
    c->args.dbg = c->verify_tokens ? c->d_dbg : nullptr;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaMemset(c->d_dbg, 0, 64 * sizeof(unsigned long long)) != cudaSuccess) { set_err("cudaMemset failed"); return MGPU_E_CUDA; }
  }
  else if (k == "variant") c->args.variant = (uint32_t)value;
  else if (k == "device_sort_min") c->dev_sort_min = (uint32_t)std::min<uint64_t>(value, 0xFFFFFFFFu);
  else if (k == "sub_block") c->sub_block = (uint32_t)value;
  else if (k == "fused") {
    if ((value != 0) != (c->nsets == 2)) { set_err("the first-stage kernels are chosen when the context is created (MATCHY_B200_FUSED=1): the work buffers differ"); return MGPU_E_PARAM; }
  }
  else { set_err("unknown option: " + k); return MGPU_E_PARAM; }
  return MGPU_OK;
}
// accumulators of the token-list audit since "verify_tokens" was last set (see verify_tokens_kernel)
int mgpu_debug_get(mgpu_ctx* c, uint64_t out[64]) {
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, c->d_dbg, 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 4; k++) out[60 + k] = c->host_us[k];
  out[59] = c->arrived_sorted ? 1 : 0;  // the last scan's records came off the device in (offset, item_type, len) order
  // host time of the last resident scan: whole call, result sort, id re-pack, launch + gather
  return MGPU_OK;
}

// ---- PSL ------------------------------------------------------------------------------------------------
int mgpu_set_psl(mgpu_ctx* c, const uint8_t* text, size_t len) {
  CK(cudaSetDevice(c->device));
  PslTable t;
  std::string err;
  if (!build_psl(text, len, t, err)) { set_err(err); return MGPU_E_FORMAT; }
  if (c->d_psl_keys) { cudaFree(c->d_psl_keys); cudaFree(c->d_psl_vals); cudaFree(c->d_psl_pool); cudaFree(c->d_psl_tld); }
  CK(cudaMalloc(&c->d_psl_keys, t.keys.size() * 8));
  CK(cudaMalloc(&c->d_psl_vals, t.vals.size() * 4));
  CK(cudaMalloc(&c->d_psl_pool, t.pool.size() + 16));
  CK(cudaMemcpy(c->d_psl_keys, t.keys.data(), t.keys.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_psl_vals, t.vals.data(), t.vals.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_psl_pool, t.pool.data(), t.pool.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->d_psl_tld, t.tld.size() * 8));
  CK(cudaMemcpy(c->d_psl_tld, t.tld.data(), t.tld.size() * 8, cudaMemcpyHostToDevice));
  DbView& db = c->args.db;
  db.psl_keys = (const uint64_t*)c->d_psl_keys; db.psl_vals = (const uint32_t*)c->d_psl_vals; db.psl_pool = (const uint8_t*)c->d_psl_pool;
  db.psl_mask = t.mask; db.psl_max_len = t.max_len; db.psl_tld = (const uint64_t*)c->d_psl_tld;
  return MGPU_OK;
}

// ---- database upload ------------------------------------------------------------------------------------
static int dev_copy(mgpu_ctx* c, const void* src, size_t len, size_t align_off, size_t align, void** out_base) {
  // device copy of [src, src+len) placed so that (address % align) == align_off
  void* raw = nullptr;
  CK(cudaMalloc(&raw, len + align + 64));
  c->db_allocs.push_back(raw);
  uintptr_t p = (uintptr_t)raw;
  uintptr_t q = (p + align - 1) / align * align + align_off;
  if (len) CK(cudaMemcpy((void*)q, src, len, cudaMemcpyHostToDevice));
  CK(cudaMemset((uint8_t*)q + len, 0, 32));
  *out_base = (void*)q;
  return MGPU_OK;
}

int mgpu_db_upload(mgpu_ctx* c, const uint8_t* d, size_t n) {
  CK(cudaSetDevice(c->device));
  std::string err;
  PreparedDb P;
  if (!prepare_db(d, n, P, err)) { set_err(err); return MGPU_E_FORMAT; }  // (a bad file leaves the database in place: hot reload)
  CK(cudaDeviceSynchronize());  // nothing may still be reading the tables that are about to be freed
  free_db(c);
  const mxy::Layout& L = P.L;
  DbView db = P.view;  // scalar fields filled; pointers below
  db.psl_keys = c->args.db.psl_keys; db.psl_vals = c->args.db.psl_vals; db.psl_pool = c->args.db.psl_pool;
  db.psl_mask = c->args.db.psl_mask; db.psl_max_len = c->args.db.psl_max_len; db.psl_tld = c->args.db.psl_tld;
  void* p;
  int rc = dev_copy(c, d, (size_t)L.tree_size, 0, 256, &p);
  if (rc) return rc;
  db.tree = (const uint8_t*)p;
  if (!P.top16.empty()) {
    rc = dev_copy(c, P.top16.data(), P.top16.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.v4_top16 = (const uint32_t*)p;
    rc = dev_copy(c, P.top16_depth.data(), P.top16_depth.size(), 0, 256, &p);
    if (rc) return rc;
    db.v4_top16_depth = (const uint8_t*)p;
  }
  if (!P.v6_top16.empty()) {
    rc = dev_copy(c, P.v6_top16.data(), P.v6_top16.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.v6_top16 = (const uint32_t*)p;
    rc = dev_copy(c, P.v6_top16_depth.data(), P.v6_top16_depth.size(), 0, 256, &p);
    if (rc) return rc;
    db.v6_top16_depth = (const uint8_t*)p;
  }
  if (L.has_literal) {
    // the slot table starts at 4 (mod 16) inside the section: base = 12 (mod 16) makes every 16-byte entry aligned
    rc = dev_copy(c, d + L.lit_off, (size_t)L.lit_len, 12, 256, &p);
    if (rc) return rc;
    db.lh = (const uint8_t*)p;
    rc = dev_copy(c, P.lh_index.data(), P.lh_index.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.lh_data_index = (const uint32_t*)p;
    rc = dev_copy(c, P.lh_bloom.data(), P.lh_bloom.size() * 8, 0, 256, &p);
    if (rc) return rc;
    db.lh_bloom = (const uint64_t*)p;
  }
  if (L.has_glob) {
    rc = dev_copy(c, d + L.pg_off, (size_t)L.pg_len, 0, 256, &p);
    if (rc) return rc;
    db.pg = (const uint8_t*)p;
    rc = dev_copy(c, P.aclh.data(), P.aclh.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.aclh_index = (const uint32_t*)p;
    rc = dev_copy(c, P.gram2.data(), P.gram2.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.ac_gram2 = (const uint32_t*)p;
    rc = dev_copy(c, P.gram3.data(), P.gram3.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.ac_gram3 = (const uint32_t*)p;
    if (!P.pfx_keys.empty()) {
      rc = dev_copy(c, P.pfx_keys.data(), P.pfx_keys.size() * 8, 0, 256, &p);
      if (rc) return rc;
      db.ac_pfx_keys = (const uint64_t*)p;
      rc = dev_copy(c, P.pfx_vals.data(), P.pfx_vals.size() * 4, 0, 256, &p);
      if (rc) return rc;
      db.ac_pfx_vals = (const uint32_t*)p;
    }
    rc = dev_copy(c, d + L.map_off, (size_t)L.map_count * 4, 0, 256, &p);
    if (rc) return rc;
    db.glob_data = (const uint32_t*)p;
  }
  if (db.fast_ok && db.has_generic) {
    rc = dev_copy(c, P.gen2.data(), P.gen2.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.gen_gram2 = (const uint32_t*)p;
    rc = dev_copy(c, P.gen3.data(), P.gen3.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.gen_gram3 = (const uint32_t*)p;
  }
  if (db.fast_ok) {
    rc = dev_copy(c, P.hot.data(), P.hot.size() * 4, 0, 256, &p);
    if (rc) return rc;
    db.hot = (const uint32_t*)p;
    rc = dev_copy(c, P.cold.data(), P.cold.size() * 8, 0, 256, &p);
    if (rc) return rc;
    db.cold = (const uint64_t*)p;
  }
  c->args.db = db;
  c->layout = L;
  c->db_loaded = true;
  mgpu_db_info& I = c->info;
  I.node_count = L.node_count; I.record_bits = L.record_bits; I.ip_version = L.ip_version; I.match_mode = L.match_mode;
  I.has_ip = db.has_ip; I.has_literal = db.has_literal; I.has_glob = db.has_glob;
  I.literal_count = L.literal_count; I.glob_count = L.glob_count; I.ac_node_count = P.ac_node_count;
  I.tree_bytes = L.tree_size; I.literal_bytes = L.lit_len; I.paraglob_bytes = L.pg_len; I.file_bytes = n;
  return MGPU_OK;
}

int mgpu_db_info_get(mgpu_ctx* c, mgpu_db_info* out) {
  if (!c->db_loaded) { set_err("no database uploaded"); return MGPU_E_NODB; }
  *out = c->info;
  return MGPU_OK;
}

uint32_t mgpu_default_flags(mgpu_ctx* c) {
  uint32_t f = 0;
  if (!c->db_loaded) return 0;
  if (c->args.db.has_ip) f |= MGPU_X_IPV4 | MGPU_X_IPV6;
  if (c->args.db.has_literal || c->args.db.has_glob) f |= MGPU_X_DOMAINS | MGPU_X_EMAILS | MGPU_X_HASHES;
  return f;
}

// ---- scanning -------------------------------------------------------------------------------------------
// Launch the kernels of one piece into batch slot `slot` (counter block + event row).  No host synchronisation.
static int launch_piece(mgpu_ctx* c, int slot, const uint8_t* d_buf, uint64_t lo, uint64_t n, uint64_t base, uint32_t flags, bool lookups) {
  ScanArgs a = c->args;
  a.buf = d_buf; a.lo = lo; a.n = n; a.base = base; a.flags = flags;
  c->piece_base[slot] = base; c->piece_len[slot] = n;
  a.ctr = c->args.ctr + slot;
  if (c->force_ac_walk) a.db.ac_anchored = 0;
  const int set = c->nsets == 2 ? (slot & 1) : 0;
  {
    const mgpu_ctx::BufSet& b = c->sets[set];
    a.q_dotted = b.q_dotted; a.q_hash = b.q_hash; a.q_at = b.q_at; a.q_c2 = b.q_c2; a.q_numeric = b.q_numeric; a.q_long = b.q_long; a.seg_cnt = b.seg_cnt;
    a.str = b.str; a.ip = b.ip;
  }
  cudaStream_t st = c->compute, ls = c->serial ? c->compute : c->lookup;
  cudaEvent_t* ev = c->ev_k[slot];
  const bool fast = lookups && a.db.fast_ok && !c->force_generic && (a.db.has_literal || a.db.has_glob);
  a.fast = fast ? 1u : 0u;
  a.lookups = lookups ? 1u : 0u;
  a.sub_block = c->fused ? 0u : c->sub_block;  // (scan_kernel's slow-path entries are not sorted by position)
  a.ip_skip = lookups && a.db.ip_empty && !c->fused && !c->verify_tokens ? 1u : 0u;  // (scan_kernel leaves numeric words to iptrie_kernel's parser; the audit wants the list)
  const uint64_t tiles = (n + TILE_BYTES - 1) / TILE_BYTES;
  if (c->fused) {
    // compute stream: scan_kernel of piece i.  It fills working set (i & 1), which the lookups of piece i-2 may still read.
    if (c->set_pending[set]) { CK(cudaStreamWaitEvent(st, c->ev_looked[c->set_slot[set]], 0)); c->set_pending[set] = false; }
    if (c->verify_tokens) fill_kernel<<<launch_grid(c, 8), 256, 0, st>>>((uint4*)a.ip, (size_t)c->alloc_cap_ip * sizeof(IpTok) / 16, TOK_POISON);
    CK(cudaEventRecord(ev[0], st));
    const bool hot = fast && a.db.hot_tags != 0 && !c->wide_scan;
    const int warps = hot ? SK_WARPS : c->wide_warps;
    uint64_t want_blocks = (tiles + warps - 1) / warps;
    int grid = (int)std::min<uint64_t>(want_blocks, (uint64_t)launch_grid(c, 1));
    if (grid < 1) grid = 1;
    a.nseg = (uint32_t)grid * warps;
    if (hot) {
      if (flags == K1_DEFAULT_FLAGS) scan_kernel<K1_DEFAULT_FLAGS, true, SK_WARPS><<<grid, SK_WARPS * 32, scan_smem(true, SK_WARPS), st>>>(a);
      else scan_kernel<0, true, SK_WARPS><<<grid, SK_WARPS * 32, scan_smem(true, SK_WARPS), st>>>(a);
    } else if (warps == 28) {
      if (flags == K1_DEFAULT_FLAGS) scan_kernel<K1_DEFAULT_FLAGS, false, 28><<<grid, 28 * 32, scan_smem(false, 28), st>>>(a);
      else scan_kernel<0, false, 28><<<grid, 28 * 32, scan_smem(false, 28), st>>>(a);
    } else if (warps == 32) {
      if (flags == K1_DEFAULT_FLAGS) scan_kernel<K1_DEFAULT_FLAGS, false, 32><<<grid, 32 * 32, scan_smem(false, 32), st>>>(a);
      else scan_kernel<0, false, 32><<<grid, 32 * 32, scan_smem(false, 32), st>>>(a);
    } else {
      if (flags == K1_DEFAULT_FLAGS) scan_kernel<K1_DEFAULT_FLAGS, false, SK_WARPS_NOHOT><<<grid, SK_WARPS_NOHOT * 32, scan_smem(false, SK_WARPS_NOHOT), st>>>(a);
      else scan_kernel<0, false, SK_WARPS_NOHOT><<<grid, SK_WARPS_NOHOT * 32, scan_smem(false, SK_WARPS_NOHOT), st>>>(a);
    }
    CK(cudaEventRecord(ev[1], st));
    CK(cudaEventRecord(c->ev_tokens[slot], st));
    // lookup stream, beside the scan of piece i+1: the anchors and slow-path words (token_kernel), then the lookups
    CK(cudaStreamWaitEvent(ls, c->ev_tokens[slot], 0));
    CK(cudaEventRecord(c->ev_l0[slot], ls));
    token_kernel<<<launch_grid(c, 1), TK_THREADS, TOKEN_SMEM, ls>>>(a);
    if (flags & MGPU_X_CRYPTO) crypto_kernel<<<launch_grid(c, 8), 128, 0, ls>>>(a);
    if (c->verify_tokens) verify_tokens_kernel<<<launch_grid(c, 8), 256, 0, ls>>>(a, c->d_dbg);
    CK(cudaEventRecord(ev[2], ls));
  } else {
    CK(cudaEventRecord(ev[0], st));
    {
      uint64_t want_blocks = (tiles + K1_WARPS - 1) / K1_WARPS;
      int grid = (int)std::min<uint64_t>(want_blocks, (uint64_t)launch_grid(c, 2));
      if (grid < 1) grid = 1;
      a.nseg = (uint32_t)grid * K1_WARPS;
      size_t smem = 256 * 32 * 8;
      if (flags == K1_DEFAULT_FLAGS) tokenize_kernel<K1_DEFAULT_FLAGS><<<grid, K1_THREADS, smem, st>>>(a);
      else tokenize_kernel<0><<<grid, K1_THREADS, smem, st>>>(a);
    }
    CK(cudaEventRecord(ev[1], st));
    // The token kernel overwrites the token lists the previous piece's lookups read: wait for them.  The lookups themselves
    // (latency-bound, a fraction of the SMs busy) run on a second stream, beside the tokenizer of the next piece.
    if (c->set_pending[0]) { CK(cudaStreamWaitEvent(st, c->ev_looked[c->set_slot[0]], 0)); c->set_pending[0] = false; }
    if (c->verify_tokens) fill_kernel<<<launch_grid(c, 8), 256, 0, st>>>((uint4*)a.ip, (size_t)c->alloc_cap_ip * sizeof(IpTok) / 16, TOK_POISON);
    token_kernel<<<launch_grid(c, 1), TK_THREADS, TOKEN_SMEM, st>>>(a);
    if (flags & MGPU_X_CRYPTO) crypto_kernel<<<launch_grid(c, 8), 128, 0, st>>>(a);
    if (c->verify_tokens) verify_tokens_kernel<<<launch_grid(c, 8), 256, 0, st>>>(a, c->d_dbg);
    CK(cudaEventRecord(ev[2], st));
    CK(cudaEventRecord(c->ev_tokens[slot], st));
    CK(cudaStreamWaitEvent(ls, c->ev_tokens[slot], 0));
    CK(cudaEventRecord(c->ev_l0[slot], ls));
  }
  if (lookups) {
    if (!a.ip_skip) {
      if (c->iptrie_minb == 5) iptrie_kernel<5><<<launch_grid(c, 8), 256, 0, ls>>>(a);
      else if (c->iptrie_minb == 6) iptrie_kernel<6><<<launch_grid(c, 8), 256, 0, ls>>>(a);
      else iptrie_kernel<4><<<launch_grid(c, 8), 256, 0, ls>>>(a);
    }
    CK(cudaEventRecord(ev[3], ls));
    if (a.db.has_literal && !fast) lithash_kernel<<<launch_grid(c, 6), KT_THREADS, 0, ls>>>(a);
    CK(cudaEventRecord(ev[4], ls));
    if (fast) exact_kernel<<<launch_grid(c, 8), 256, 0, ls>>>(a);
    else if (a.db.has_literal || a.db.has_glob) acglob_kernel<<<launch_grid(c, 4), KT_THREADS, ACGLOB_SMEM, ls>>>(a);  // 4 blocks/SM: register- and shared-memory-limited
    CK(cudaEventRecord(ev[5], ls));
  } else {
    if (c->fused) iptrie_kernel<4><<<launch_grid(c, 8), 256, 0, ls>>>(a);  // (extraction only: parses the scan kernel's raw numeric words)
    for (int k = 3; k <= 5; k++) CK(cudaEventRecord(ev[k], ls));
  }
  piece_end_kernel<<<1, 1, 0, ls>>>(a.ctr, a.tot, c->d_recs ? c->h_nrec + slot : nullptr);
  CK(cudaEventRecord(c->ev_looked[slot], ls));
  c->set_pending[set] = true; c->set_slot[set] = slot;
  c->looked_pending = true; c->looked_slot = slot;
  CK(cudaGetLastError());
  c->timing.launches[MGPU_K_TOKENIZE]++; c->timing.launches[MGPU_K_TOKEN]++;
  if (lookups) {
    c->timing.launches[MGPU_K_IPTRIE]++;
    if (a.db.has_literal && !fast) c->timing.launches[MGPU_K_LITHASH]++;
    if (a.db.has_literal || a.db.has_glob) c->timing.launches[MGPU_K_STRINGS]++;
  }
  c->timing.aux_launches++;
  c->timing.chunks++;
  return MGPU_OK;
}

// Start a batch: zero the counter blocks and the running totals.
static int begin_batch(mgpu_ctx* c, int pieces) {
  CK(cudaMemsetAsync(c->args.ctr, 0, sizeof(DevCounters) * pieces, c->compute));
  CK(cudaMemsetAsync(c->d_tot, 0, sizeof(ScanTotals), c->compute));
  return MGPU_OK;
}
// Finish a batch: counters to the host, one synchronisation, kernel times.
static int end_batch(mgpu_ctx* c, int pieces) {
  if (c->looked_pending) { CK(cudaStreamWaitEvent(c->compute, c->ev_looked[c->looked_slot], 0)); c->looked_pending = false; }  // (one lookup stream: the latest event covers all)
  c->set_pending[0] = c->set_pending[1] = false;
  CK(cudaMemcpyAsync(c->h_ctr, c->args.ctr, sizeof(DevCounters) * pieces, cudaMemcpyDeviceToHost, c->compute));
  // The scan's device span (mgpu_timing.scan_ms) ends with the last device operation, not with the host's gathering of the
  // results: re-recorded after every batch and after every copy of id pairs, the last record stands.
  CK(cudaEventRecord(c->ev_scan[1], c->compute));
  CK(cudaStreamSynchronize(c->compute));
  for (int p = 0; p < pieces; p++) {
    for (int k = 0; k < MGPU_K_COUNT; k++) {
      float ms = 0;
      // (the lookup kernels run on their own stream: their span starts where that stream has finished waiting for the tokens)
      // (whichever kernel opens the lookup stream's part of the piece — token_kernel when scan_kernel did the first pass, else
      // iptrie_kernel — is timed from ev_l0, recorded after that stream's wait for the tokens)
      const int first_on_lookup = c->fused ? MGPU_K_TOKEN : MGPU_K_IPTRIE;
      CK(cudaEventElapsedTime(&ms, k == first_on_lookup ? c->ev_l0[p] : c->ev_k[p][k], c->ev_k[p][k + 1]));
      c->timing.kernel_ms[k] += ms;
    }
    float tot = 0;
    CK(cudaEventElapsedTime(&tot, c->ev_k[p][0], c->ev_k[p][MGPU_K_COUNT]));
    c->timing.total_ms += tot;
  }
  return MGPU_OK;
}
// The records of the batch's pieces 0..nb-1, device -> `stage`, piece by piece as their lookups finish (the kernels of later
// pieces keep running meanwhile; the ranges are disjoint: the record counter only grows).  The compute stream then waits for the
// last copy, so that end_batch's synchronisation — and the scan's device span — covers it.
static int drain_records(mgpu_ctx* c, int nb) {
  if (!c->d_recs) return MGPU_OK;
  uint32_t r_prev = 0;
  for (int k = 0; k < nb; k++) {
    CK(cudaEventSynchronize(c->ev_looked[k]));
    const uint32_t r_now = std::min(c->h_nrec[k], c->args.cap_rec);
    if (r_now > r_prev) {
      if (c->keep_results) {
        const uint32_t n = r_now - r_prev;
        const mgpu_match* src = c->d_recs + r_prev;
        if (n >= c->dev_sort_min) c->dev_sorting = true;
        bool sorted_here = n <= 1;
        if (c->d_sorted && n > 1 && c->dev_sorting) {
          cub::DoubleBuffer<uint64_t> dk(c->d_keys[0], c->d_keys[1]);
          cub::DoubleBuffer<uint32_t> dv(c->d_vals[0], c->d_vals[1]);
          int end_bit = 5;
          while (end_bit < 64 && (c->piece_len[k] >> (end_bit - 4)) != 0) end_bit++;
          size_t need = 0;
          CK(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, end_bit, c->d2h));
          if (need <= c->sort_tmp_bytes) {
            const int grid = (int)std::min<uint32_t>((n + 255) / 256, (uint32_t)launch_grid(c, 4));
            sort_keys_kernel<<<grid, 256, 0, c->d2h>>>(src, n, c->piece_base[k], c->d_keys[0], c->d_vals[0]);
            need = c->sort_tmp_bytes;
            CK(cub::DeviceRadixSort::SortPairs(c->d_sort_tmp, need, dk, dv, (int)n, 0, end_bit, c->d2h));
            sort_gather_kernel<<<(int)std::min<uint32_t>((2 * n + 255) / 256, (uint32_t)launch_grid(c, 8)), 256, 0, c->d2h>>>(src, n, dv.Current(), c->d_sorted + r_prev);
            src = c->d_sorted + r_prev;
            sorted_here = true;
            c->timing.aux_launches += 2;  // (sort_keys_kernel, sort_gather_kernel; cub's own kernels are library launches and not counted)
          }
        }
        if (!sorted_here) c->order_ok = false;
        CK(cudaMemcpyAsync(c->stage.p + r_prev, src, (size_t)n * sizeof(mgpu_match), cudaMemcpyDeviceToHost, c->d2h));
      }
      r_prev = r_now;
    }
  }
  CK(cudaEventRecord(c->ev_d2h, c->d2h));
  CK(cudaStreamWaitEvent(c->compute, c->ev_d2h, 0));
  return MGPU_OK;
}
// Put the records [r_lo, r_hi) of the batch's record buffer (pinned host memory the device wrote into) and the id pairs
// [i_lo, i_hi) of the device buffer behind the host vectors.  whole = these are all the records the batch produced and nothing
// is waiting to be redone: then the two pinned buffers simply change places.
static int fetch_results(mgpu_ctx* c, uint32_t r_lo, uint32_t r_hi, uint32_t i_lo, uint32_t i_hi, bool whole = false) {
  if (!c->keep_results || r_hi <= r_lo) return MGPU_OK;
  size_t r0 = c->recs.size(), i0 = c->ids.size();
  if (whole && r0 == 0 && r_lo == 0 && c->recs.cap == c->stage.cap) {
    std::swap(c->recs.p, c->stage.p);
    c->recs.n = r_hi;
    if (!c->d_recs) c->args.recs = c->stage.p;
  } else {
    if (!c->recs.resize(r0 + (r_hi - r_lo))) { set_err("out of pinned host memory for match records"); return MGPU_E_CUDA; }
    // (later batches of a long resident scan: tens of millions of records — copy with several threads)
    const size_t nrec = r_hi - r_lo;
    const unsigned nt = nrec < (1u << 20) ? 1u : std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
    mgpu_match* dst = c->recs.data() + r0;
    const mgpu_match* src = c->stage.p + r_lo;
    if (nt == 1) memcpy(dst, src, nrec * sizeof(mgpu_match));
    else {
      std::vector<std::thread> th;
      for (unsigned k = 0; k < nt; k++) th.emplace_back([=] { const size_t a = nrec * k / nt, b = nrec * (k + 1) / nt; memcpy(dst + a, src + a, (b - a) * sizeof(mgpu_match)); });
      for (auto& t : th) t.join();
    }
  }
  if (i_hi > i_lo) {
    if (!c->ids.resize(i0 + (i_hi - i_lo))) { set_err("out of pinned host memory for match ids"); return MGPU_E_CUDA; }
    CK(cudaMemcpyAsync(c->ids.data() + i0, c->args.ids + i_lo, (size_t)(i_hi - i_lo) * sizeof(mgpu_id_pair), cudaMemcpyDeviceToHost, c->compute));
    CK(cudaEventRecord(c->ev_scan[1], c->compute));
    CK(cudaStreamSynchronize(c->compute));
  }
  if ((size_t)i_lo != i0)  // (the first run of a scan lands at the same indices: nothing to rebase — 8 M records per step on config 3)
    for (size_t k = r0; k < c->recs.size(); k++) if (c->recs[k].kind == MGPU_KIND_PATTERN) c->recs[k].ids_index = (uint32_t)(c->recs[k].ids_index - i_lo + i0);
  return MGPU_OK;
}
static void add_counters(mgpu_ctx* c, const DevCounters& h, uint64_t bytes, uint32_t n_rec) {
  c->counters.lines += h.lines;
  c->counters.bytes += bytes;
  for (int k = 0; k < 12; k++) { c->counters.by_type[k] += h.by_type[k]; c->counters.candidates += h.by_type[k]; }
  c->counters.matches += n_rec;
}

// first newline-aligned cut point at or after each of `from` inside dev[0..n)
static int find_cuts(mgpu_ctx* c, const uint8_t* dev, uint64_t n, const std::vector<uint64_t>& from, std::vector<uint64_t>& out) {
  int cnt = (int)from.size();
  if (cnt > 64) { set_err("too many cuts"); return MGPU_E_PARAM; }
  for (int k = 0; k < cnt; k++) c->h_cut[k] = from[k];
  CK(cudaMemcpyAsync(c->d_cut, c->h_cut, cnt * 8, cudaMemcpyHostToDevice, c->compute));
  cut_kernel<<<1, 64, 0, c->compute>>>(dev, n, c->d_cut, c->d_cut + 64, cnt);
  c->timing.aux_launches++;
  CK(cudaMemcpyAsync(c->h_cut + 64, c->d_cut + 64, cnt * 8, cudaMemcpyDeviceToHost, c->compute));
  CK(cudaStreamSynchronize(c->compute));
  out.assign(c->h_cut + 64, c->h_cut + 64 + cnt);
  return MGPU_OK;
}

// Process the resident piece dev[pos, end) (dev 16-byte aligned; the piece may start anywhere).  On buffer
// exhaustion split it at newlines and retry the parts.  `base` = absolute log offset of dev[0].
static int scan_piece(mgpu_ctx* c, const uint8_t* dev, uint64_t pos, uint64_t end, uint64_t base, uint32_t flags, bool lookups, int depth) {
  if (end <= pos) return MGPU_OK;
  const uint64_t al = pos & ~(uint64_t)15;
  int rc = begin_batch(c, 1);
  if (rc) return rc;
  rc = launch_piece(c, 0, dev + al, pos - al, end - al, base + al, flags, lookups);
  if (rc) return rc;
  rc = drain_records(c, 1);
  if (rc) return rc;
  rc = end_batch(c, 1);
  if (rc) return rc;
  DevCounters& h = c->h_ctr[0];
  if (h.overflow) {
    if (depth >= 6 || end - pos < 4096) { set_err("result buffers exhausted (match density too high for the configured chunk size)"); return MGPU_E_OVERFLOW; }
    std::vector<uint64_t> from, cuts;
    for (int k = 1; k < 8; k++) from.push_back(pos + (end - pos) / 8 * k);
    rc = find_cuts(c, dev, end, from, cuts);
    if (rc) return rc;
    uint64_t prev = pos;
    cuts.push_back(end);
    for (uint64_t cut : cuts) {
      if (cut <= prev) continue;
      rc = scan_piece(c, dev, prev, cut, base, flags, lookups, depth + 1);
      if (rc) return rc;
      prev = cut;
    }
    return MGPU_OK;
  }
  // collect
  add_counters(c, h, end - pos, h.n_rec);
  if (c->capture_tokens) {
    size_t s0 = c->x_str.size(), i0 = c->x_ip.size();
    c->x_str.resize(s0 + h.n_str); c->x_ip.resize(i0 + h.n_ip);
    if (h.n_str) CK(cudaMemcpy(c->x_str.data() + s0, c->args.str, (size_t)h.n_str * sizeof(StrTok), cudaMemcpyDeviceToHost));
    if (h.n_ip) CK(cudaMemcpy(c->x_ip.data() + i0, c->args.ip, (size_t)h.n_ip * sizeof(IpTok), cudaMemcpyDeviceToHost));
    // drop the padding slots of the per-warp reservations, then make offsets absolute (extraction is limited to < 4 GiB inputs)
    c->x_str.erase(std::remove_if(c->x_str.begin() + s0, c->x_str.end(), [](const StrTok& t) { return t.type == TOK_INVALID; }), c->x_str.end());
    c->x_ip.erase(std::remove_if(c->x_ip.begin() + i0, c->x_ip.end(), [](const IpTok& t) { return t.type == TOK_INVALID; }), c->x_ip.end());
    for (size_t k = s0; k < c->x_str.size(); k++) c->x_str[k].start += (uint32_t)(base + al);
    for (size_t k = i0; k < c->x_ip.size(); k++) c->x_ip[k].start += (uint32_t)(base + al);
  }
  return fetch_results(c, 0, h.n_rec, 0, h.n_ids, /*whole=*/true);
}

static void begin_scan(mgpu_ctx* c) {
  c->recs.clear(); c->ids.clear();
  c->dev_sorting = false;
  c->order_ok = true;
  c->x_str.clear(); c->x_ip.clear();
  memset(&c->counters, 0, sizeof c->counters);
  memset(&c->timing, 0, sizeof c->timing);
  cudaEventRecord(c->ev_scan[0], c->compute);
  cudaEventRecord(c->ev_scan[1], c->compute);  // (stands only for an empty input: every batch records it again)
}

static void finish_scan(mgpu_ctx* c) {
  cudaEventSynchronize(c->ev_scan[1]);
  cudaEventElapsedTime(&c->timing.scan_ms, c->ev_scan[0], c->ev_scan[1]);
  // deterministic output: records by (offset, item_type, len), id pairs re-packed in record order (the device appends
  // both with atomics, so their raw order varies from run to run)
  const size_t nrec = c->recs.size();
  const auto t_sort = std::chrono::steady_clock::now();
  if (nrec >= 4096 && !c->pool) {
    const unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (nt >= 2) c->pool.reset(new WorkerPool(nt));
  }
  c->arrived_sorted = c->d_sorted && c->order_ok && ((nrec > ((size_t)1 << 20) && !c->verify_order) || records_sorted(c->recs.data(), nrec, c->pool.get()));
  if (!c->arrived_sorted) sort_records(c->recs.data(), nrec, c->scan_lo, c->scan_hi, c->sort_tmp, c->pool.get());
  const auto t_ids = std::chrono::steady_clock::now();
  if (!c->ids.empty()) {
    const size_t np = repack_ids(c->recs.data(), nrec, c->ids.data(), c->ids_tmp, c->pool.get());
    c->ids.n = np;
    if (np) memcpy(c->ids.data(), c->ids_tmp.data(), np * sizeof(mgpu_id_pair));
  }
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
  c->host_us[1] = us(t_sort, t_ids);
  c->host_us[2] = us(t_ids, std::chrono::steady_clock::now());
}

static int check_ready(mgpu_ctx* c, uint32_t flags) {
  if (!c) { set_err("null context"); return MGPU_E_PARAM; }
  if (!c->db_loaded) { set_err("no database uploaded"); return MGPU_E_NODB; }
  if (flags & ~MGPU_X_SUPPORTED) { set_err("unknown extractor flags"); return MGPU_E_PARAM; }
  if ((flags & (MGPU_X_DOMAINS | MGPU_X_EMAILS)) && !c->args.db.psl_keys) { set_err("Public Suffix List not set (mgpu_set_psl)"); return MGPU_E_PARAM; }
  return MGPU_OK;
}

// dev[0..len) resident in HBM: pieces of at most chunk_bytes, each ending just after a newline (any newline-aligned
// partition gives the same result; FileReader::next_batch, processing/mod.rs:206-251, is one such
// partition).  All cut points come from one kernel; the pieces of a batch are launched back to back with a counter
// block each and the host synchronises once per batch.  A piece that ran out of room is redone through scan_piece().
static int scan_device_impl(mgpu_ctx* c, const uint8_t* dev, size_t len, uint64_t base, uint32_t flags, bool lookups) {
  if (((uintptr_t)dev & 15) != 0) { set_err("device buffer must be 16-byte aligned"); return MGPU_E_PARAM; }
  if (len == 0) return MGPU_OK;
  const uint64_t slack = std::min<uint64_t>((uint64_t)1 << 20, c->chunk_bytes / 4);  // longest line the parallel cut search allows
  const uint64_t stride = c->chunk_bytes - 16 - slack;  // launch_piece realigns a piece's start downwards by up to 15 bytes
  std::vector<uint64_t> cuts;  // piece k = [cuts[k], cuts[k+1])
  cuts.push_back(0);
  const uint64_t ncut = (len - 1) / stride;  // nominal boundaries stride, 2*stride, ... < len
  bool sequential = c->capture_tokens || ncut > 4096;
  if (!sequential && ncut > 0) {
    for (uint64_t k = 0; k < ncut; k++) c->h_cut[k] = (k + 1) * stride;
    CK(cudaMemcpyAsync(c->d_cut, c->h_cut, ncut * 8, cudaMemcpyHostToDevice, c->compute));
    cuts_back_kernel<<<(unsigned)ncut, 256, 0, c->compute>>>(dev, c->d_cut, slack, c->d_cut + 4096);
    c->timing.aux_launches++;
    CK(cudaMemcpyAsync(c->h_cut + 4096, c->d_cut + 4096, ncut * 8, cudaMemcpyDeviceToHost, c->compute));
    CK(cudaStreamSynchronize(c->compute));
    for (uint64_t k = 0; k < ncut && !sequential; k++) {
      if (c->h_cut[4096 + k] == NONE64) sequential = true;  // a line longer than `slack`: take the sequential search below
      else if (c->h_cut[4096 + k] > cuts.back()) cuts.push_back(c->h_cut[4096 + k]);
    }
  }
  if (sequential) {
    uint64_t pos = 0;
    const uint64_t window = c->chunk_bytes - 16;
    while (pos < len) {
      uint64_t end = std::min<uint64_t>(len, pos + window);
      if (end < len) {
        rfind_nl_kernel<<<1, 256, 0, c->compute>>>(dev, pos, end, c->d_cut);
        c->timing.aux_launches++;
        CK(cudaMemcpyAsync(c->h_cut, c->d_cut, 8, cudaMemcpyDeviceToHost, c->compute));
        CK(cudaStreamSynchronize(c->compute));
        if (c->h_cut[0] == 0) { set_err("a single line is longer than the scan chunk; raise chunk_bytes"); return MGPU_E_OVERFLOW; }
        end = c->h_cut[0];
      }
      int rc = scan_piece(c, dev, pos, end, base, flags, lookups, 0);
      if (rc) return rc;
      pos = end;
    }
    return MGPU_OK;
  }
  cuts.push_back(len);
  const size_t npieces = cuts.size() - 1;
  for (size_t p0 = 0; p0 < npieces; p0 += mgpu_ctx::MAX_BATCH) {
    const int nb = (int)std::min<size_t>(mgpu_ctx::MAX_BATCH, npieces - p0);
    int rc = begin_batch(c, nb);
    if (rc) return rc;
    for (int k = 0; k < nb; k++) {
      const uint64_t pos = cuts[p0 + k], end = cuts[p0 + k + 1], al = pos & ~(uint64_t)15;
      rc = launch_piece(c, k, dev + al, pos - al, end - al, base + al, flags, lookups);
      if (rc) return rc;
    }
    rc = drain_records(c, nb);
    if (rc) return rc;
    rc = end_batch(c, nb);
    if (rc) return rc;
    // gather: runs of pieces without overflow share one copy; overflowed pieces are redone afterwards
    std::vector<DevCounters> h(c->h_ctr, c->h_ctr + nb);  // (scan_piece below reuses the pinned block)
    std::vector<int> redo;
    bool any_overflow = false;
    for (int k = 0; k < nb; k++) any_overflow |= h[k].overflow != 0;
    uint32_t r_prev = 0, i_prev = 0, run_r = 0, run_i = 0;
    bool in_run = false;
    for (int k = 0; k <= nb; k++) {
      const bool ok = k < nb && !h[k].overflow;
      if (ok) {
        if (!in_run) { run_r = r_prev; run_i = i_prev; in_run = true; }
        add_counters(c, h[k], cuts[p0 + k + 1] - cuts[p0 + k], h[k].n_rec - r_prev);
      } else {
        if (in_run) { rc = fetch_results(c, run_r, r_prev, run_i, i_prev, !any_overflow); if (rc) return rc; in_run = false; }
        if (k < nb) redo.push_back(k);
      }
      if (k < nb) { r_prev = std::min(h[k].n_rec, c->args.cap_rec); i_prev = std::min(h[k].n_ids, c->args.cap_ids); }
    }
    if (!redo.empty()) c->order_ok = false;  // (their records follow those of later pieces)
    for (int k : redo) {
      rc = scan_piece(c, dev, cuts[p0 + k], cuts[p0 + k + 1], base, flags, lookups, 0);
      if (rc) return rc;
    }
  }
  return MGPU_OK;
}

int mgpu_scan_device(mgpu_ctx* c, const uint8_t* dev, size_t len, uint64_t base, uint32_t flags) {
  int rc = check_ready(c, flags);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  const auto t0 = std::chrono::steady_clock::now();
  begin_scan(c);
  c->scan_lo = base; c->scan_hi = base + len;
  rc = scan_device_impl(c, dev, len, base, flags, true);
  if (rc) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  finish_scan(c);
  c->host_us[3] = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
  c->host_us[0] = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
  return MGPU_OK;
}

static int scan_host_impl(mgpu_ctx* c, const uint8_t* host, size_t len, uint64_t base, uint32_t flags, bool lookups) {
  // FileReader::next_batch semantics (processing/mod.rs:206-251): pieces end just after a '\n', the tail goes last.
  cudaPointerAttributes attr;
  bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  size_t pos = 0;
  struct Piece { size_t pos, len; int slot; };
  auto stage = [&](size_t p, size_t l, int s) -> int {
    // H2D of piece [p, p+l) into d_log[s] on the copy stream; waits until the kernels that last read d_log[s] are done
    CK(cudaStreamWaitEvent(c->copy, c->ev_free[s], 0));
    if (pinned) CK(cudaMemcpyAsync(c->d_log[s], host + p, l, cudaMemcpyHostToDevice, c->copy));
    else {
      size_t bounce = std::min(c->chunk_bytes, (size_t)64 << 20);
      for (size_t o = 0; o < l; o += bounce) {
        size_t m = std::min(bounce, l - o);
        CK(cudaStreamSynchronize(c->copy));  // the bounce buffer is reused
        memcpy(c->h_pin[s], host + p + o, m);
        CK(cudaMemcpyAsync(c->d_log[s] + o, c->h_pin[s], m, cudaMemcpyHostToDevice, c->copy));
      }
    }
    CK(cudaEventRecord(c->ev_copied[s], c->copy));
    return MGPU_OK;
  };
  auto next_piece = [&](size_t p, size_t& l) -> int {
    size_t want = std::min(len - p, c->chunk_bytes);
    if (p + want < len) {
      const uint8_t* q = (const uint8_t*)memrchr(host + p, '\n', want);
      if (!q) { set_err("a single line is longer than the scan chunk; raise chunk_bytes"); return MGPU_E_OVERFLOW; }
      want = (size_t)(q - (host + p)) + 1;
    }
    l = want;
    return MGPU_OK;
  };
  if (len == 0) return MGPU_OK;
  Piece cur{0, 0, 0};
  int rc = next_piece(0, cur.len);
  if (rc) return rc;
  rc = stage(cur.pos, cur.len, cur.slot);
  if (rc) return rc;
  pos = cur.len;
  while (true) {
    Piece nxt{pos, 0, cur.slot ^ 1};
    bool have_next = pos < len;
    if (have_next) {
      rc = next_piece(pos, nxt.len);
      if (rc) return rc;
      rc = stage(nxt.pos, nxt.len, nxt.slot);  // overlaps with the kernels of `cur`
      if (rc) return rc;
      pos += nxt.len;
    }
    CK(cudaStreamWaitEvent(c->compute, c->ev_copied[cur.slot], 0));
    rc = scan_piece(c, c->d_log[cur.slot], 0, cur.len, base + cur.pos, flags, lookups, 0);
    if (rc) return rc;
    CK(cudaEventRecord(c->ev_free[cur.slot], c->compute));
    if (!have_next) break;
    cur = nxt;
  }
  return MGPU_OK;
}

int mgpu_scan(mgpu_ctx* c, const uint8_t* host, size_t len, uint64_t base, uint32_t flags) {
  int rc = check_ready(c, flags);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  const bool trace = getenv("MGPU_TRACE") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  begin_scan(c);
  c->scan_lo = base; c->scan_hi = base + len;
  rc = scan_host_impl(c, host, len, base, flags, true);
  if (rc) return rc;
  auto t1 = std::chrono::steady_clock::now();
  finish_scan(c);
  if (trace) {
    auto t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "mgpu_scan: %zu bytes, pieces %u, copy+kernels %.2f ms, finish (sort, repack) %.2f ms\n", len, c->timing.chunks,
            std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
  }
  return MGPU_OK;
}

int mgpu_results(mgpu_ctx* c, const mgpu_match** recs, size_t* n_recs, const mgpu_id_pair** ids, size_t* n_ids) {
  *recs = c->recs.data(); *n_recs = c->recs.size(); *ids = c->ids.data(); *n_ids = c->ids.size();
  return MGPU_OK;
}
int mgpu_counters_get(mgpu_ctx* c, mgpu_counters* out) { *out = c->counters; return MGPU_OK; }
int mgpu_timing_get(mgpu_ctx* c, mgpu_timing* out) { *out = c->timing; return MGPU_OK; }

int64_t mgpu_extract(mgpu_ctx* c, const uint8_t* host, size_t len, uint32_t flags, uint64_t* out, size_t cap) {
  if (!c) { set_err("null context"); return MGPU_E_PARAM; }
  if (flags & ~MGPU_X_SUPPORTED) { set_err("unsupported extractor flags"); return MGPU_E_PARAM; }
  if ((flags & (MGPU_X_DOMAINS | MGPU_X_EMAILS)) && !c->args.db.psl_keys) { set_err("Public Suffix List not set"); return MGPU_E_PARAM; }
  if (len >= ((size_t)1 << 32)) { set_err("mgpu_extract is limited to inputs < 4 GiB"); return MGPU_E_PARAM; }
  if (cudaSetDevice(c->device) != cudaSuccess) { set_err("cudaSetDevice failed"); return MGPU_E_CUDA; }
  begin_scan(c);
  c->capture_tokens = true;
  int rc = scan_host_impl(c, host, len, 0, flags, false);
  c->capture_tokens = false;
  if (rc) return rc;
  std::vector<std::array<uint64_t, 3>> items;
  for (auto& t : c->x_str) items.push_back({t.type, t.start, (uint64_t)t.start + t.len});
  for (auto& t : c->x_ip) items.push_back({t.type, t.start, (uint64_t)t.start + t.len});
  std::sort(items.begin(), items.end(), [](auto& x, auto& y) { return x[1] != y[1] ? x[1] < y[1] : x[0] < y[0]; });
  for (size_t k = 0; k < items.size() && k < cap; k++) { out[3 * k] = items[k][0]; out[3 * k + 1] = items[k][1]; out[3 * k + 2] = items[k][2]; }
  return (int64_t)items.size();
}

int mgpu_lookup_string(mgpu_ctx* c, const uint8_t* q, size_t len, mgpu_id_pair* out, size_t cap) {
  if (!c || !c->db_loaded) { set_err("no database uploaded"); return MGPU_E_NODB; }
  if (len > 65536) { set_err("query too long"); return MGPU_E_PARAM; }
  CK(cudaSetDevice(c->device));
  ScanArgs a = c->args;
  CK(cudaMemcpyAsync(c->d_small, q, len, cudaMemcpyHostToDevice, c->compute));
  a.buf = c->d_small; a.n = len;
  query_string_kernel<<<1, 1, 0, c->compute>>>(a, (uint32_t)len, c->d_small_out);
  uint32_t n = 0;
  CK(cudaMemcpyAsync(&n, c->d_small_out, 4, cudaMemcpyDeviceToHost, c->compute));
  CK(cudaStreamSynchronize(c->compute));
  if (n == NONE32) { set_err("too many matching patterns"); return MGPU_E_OVERFLOW; }
  std::vector<mgpu_id_pair> tmp(n);
  if (n) CK(cudaMemcpy(tmp.data(), a.ids, (size_t)n * sizeof(mgpu_id_pair), cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < n && k < cap; k++) out[k] = tmp[k];
  return (int)n;
}

int mgpu_lookup_ip(mgpu_ctx* c, const uint8_t ip16[16], int is_v6, uint32_t* data_offset, uint8_t* prefix_len) {
  if (!c || !c->db_loaded) { set_err("no database uploaded"); return MGPU_E_NODB; }
  CK(cudaSetDevice(c->device));
  ScanArgs a = c->args;
  CK(cudaMemcpyAsync(c->d_small, ip16, 16, cudaMemcpyHostToDevice, c->compute));
  a.buf = c->d_small; a.n = 16;
  query_ip_kernel<<<1, 1, 0, c->compute>>>(a, is_v6, c->d_small_out);
  uint32_t r[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(r, c->d_small_out, 12, cudaMemcpyDeviceToHost, c->compute));
  CK(cudaStreamSynchronize(c->compute));
  *data_offset = r[1]; *prefix_len = (uint8_t)r[2];
  return (int)r[0];
}

// ---- memory helpers ---------------------------------------------------------------------------------------
void* mgpu_dev_alloc(mgpu_ctx* c, size_t bytes) {
  if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes + 2 * TILE_BYTES) != cudaSuccess) { set_err("cudaMalloc failed"); return nullptr; }
  return p;
}
void mgpu_dev_free(mgpu_ctx* c, void* p) { cudaSetDevice(c->device); cudaFree(p); }
int mgpu_dev_upload(mgpu_ctx* c, void* dst, const void* src, size_t bytes) {
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return MGPU_OK;
}
int mgpu_dev_download(mgpu_ctx* c, void* dst, const void* src, size_t bytes) {
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return MGPU_OK;
}
void* mgpu_host_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { set_err("cudaMallocHost failed"); return nullptr; }
  return p;
}
void mgpu_host_free_pinned(void* p) { cudaFreeHost(p); }
int mgpu_flush_l2(mgpu_ctx* c) {
  // overwrite a buffer larger than the 126 MB L2
  CK(cudaSetDevice(c->device));
  const size_t bytes = (size_t)256 << 20;
  if (!c->d_flush) CK(cudaMalloc(&c->d_flush, bytes));
  fill_kernel<<<c->sm_count * 4, 256, 0, c->compute>>>((uint4*)c->d_flush, bytes / 16, 0x0A0A0A0Au);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->compute));
  return MGPU_OK;
}

}  // extern "C"
