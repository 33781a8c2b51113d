// crypto_addr.cuh — device side of the crypto-address extractors (matchy-extractor/src/lib.rs:1269-1409, 1799-1920):
// Bitcoin Base58Check (bs58 decode + double SHA-256), Bitcoin bech32/bech32m (hrp "bc"), Ethereum "0x" + 40 hex with the
// EIP-55 mixed-case checksum (Keccak-256), Monero (plain bs58 decode + Keccak-256, exactly as the reference does it).
// Plain functions of byte pointers, MGPU_HD like device_fns.cuh so that the host emulation can check them against the oracle.
// The candidates are rare (long boundary-delimited words with the right first bytes), so these run one thread per candidate
// in a kernel of their own (crypto_kernel) and favour small code over speed.
#pragma once
#include "../../include/matchy_b200.h"
#include "device_fns.cuh"

namespace mgpu {

MGPU_HD uint32_t ca_rotr(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
// SHA-256 of msg[0..len), len <= 119 (at most two blocks)
MGPU_HDN void ca_sha256(const uint8_t* msg, uint32_t len, uint8_t out[32]) {
  const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74,
      0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d,
      0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e,
      0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5,
      0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  const uint32_t blocks = len <= 55 ? 1u : 2u, total = blocks * 64;
  for (uint32_t off = 0; off < total; off += 64) {
    uint32_t w[64];
    for (uint32_t i = 0; i < 16; i++) {
      uint32_t v = 0;
      for (uint32_t k = 0; k < 4; k++) {
        const uint32_t pos = off + 4 * i + k;
        uint8_t b = pos < len ? msg[pos] : (pos == len ? 0x80 : 0);
        if (pos >= total - 8) { const uint64_t bits = (uint64_t)len * 8; b = (uint8_t)(bits >> (8 * (total - 1 - pos))); }
        v = (v << 8) | b;
      }
      w[i] = v;
    }
    for (uint32_t i = 16; i < 64; i++) {
      const uint32_t s0 = ca_rotr(w[i - 15], 7) ^ ca_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = ca_rotr(w[i - 2], 17) ^ ca_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (uint32_t i = 0; i < 64; i++) {
      const uint32_t t1 = hh + (ca_rotr(e, 6) ^ ca_rotr(e, 11) ^ ca_rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      const uint32_t t2 = (ca_rotr(a, 2) ^ ca_rotr(a, 13) ^ ca_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  for (uint32_t i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}

MGPU_HD uint64_t ca_rotl64(uint64_t x, uint32_t r) { return r ? (x << r) | (x >> (64 - r)) : x; }
// Keccak-256 (tiny-keccak Keccak::v256: rate 136, padding 0x01 … 0x80) of msg[0..len), len <= 135 (one block); first 4 digest bytes
// are all the callers need besides the EIP-55 nibbles, so the whole 32-byte digest is returned.
MGPU_HDN void ca_keccak256(const uint8_t* msg, uint32_t len, uint8_t out[32]) {
  const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
                           0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                           0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
                           0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  const uint8_t ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};  // [x + 5y]
  uint64_t st[25];
  for (uint32_t i = 0; i < 25; i++) st[i] = 0;
  for (uint32_t i = 0; i < 136; i++) {
    uint8_t b = i < len ? msg[i] : (i == len ? 0x01 : 0);
    if (i == 135) b |= 0x80;
    st[i >> 3] ^= (uint64_t)b << (8 * (i & 7));
  }
  for (uint32_t round = 0; round < 24; round++) {
    uint64_t C[5], B[25];
    for (uint32_t x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (uint32_t i = 0; i < 25; i++) { const uint32_t x = i % 5; st[i] ^= C[(x + 4) % 5] ^ ca_rotl64(C[(x + 1) % 5], 1); }
    for (uint32_t x = 0; x < 5; x++)
      for (uint32_t y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = ca_rotl64(st[x + 5 * y], ROT[x + 5 * y]);
    for (uint32_t y = 0; y < 5; y++)
      for (uint32_t x = 0; x < 5; x++) st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[round];
  }
  for (uint32_t i = 0; i < 32; i++) out[i] = (uint8_t)(st[i >> 3] >> (8 * (i & 7)));
}

// bs58 0.5.1 decode_into, Bitcoin alphabet.  s[0..n), n <= 110 -> out (big-endian, leading '1's as zero bytes), length in out_n
// (<= 96).  false on a byte outside the alphabet.
MGPU_HD int ca_b58_value(uint8_t c) {
  if (c >= '1' && c <= '9') return c - '1';
  if (c >= 'A' && c <= 'H') return c - 'A' + 9;
  if (c >= 'J' && c <= 'N') return c - 'J' + 17;
  if (c >= 'P' && c <= 'Z') return c - 'P' + 22;
  if (c >= 'a' && c <= 'k') return c - 'a' + 33;
  if (c >= 'm' && c <= 'z') return c - 'm' + 44;
  return -1;
}
MGPU_HDN bool ca_bs58_decode(const uint8_t* s, uint32_t n, uint8_t out[96], uint32_t& out_n) {
  uint8_t le[96];
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; i++) {
    const int v = ca_b58_value(s[i]);
    if (v < 0) return false;
    uint32_t val = (uint32_t)v;
    for (uint32_t k = 0; k < m; k++) { val += (uint32_t)le[k] * 58; le[k] = (uint8_t)val; val >>= 8; }
    while (val > 0) { if (m >= 96) return false; le[m++] = (uint8_t)val; val >>= 8; }
  }
  for (uint32_t i = 0; i < n && s[i] == '1'; i++) { if (m >= 96) return false; le[m++] = 0; }
  for (uint32_t k = 0; k < m; k++) out[k] = le[m - 1 - k];
  out_n = m;
  return true;
}
MGPU_HDN bool ca_bitcoin_base58(const uint8_t* s, uint32_t n) {  // validate_bitcoin_base58
  uint8_t d[96], h1[32], h2[32];
  uint32_t dn;
  if (!ca_bs58_decode(s, n, d, dn) || dn < 5) return false;
  ca_sha256(d, dn - 4, h1);
  ca_sha256(h1, 32, h2);
  return h2[0] == d[dn - 4] && h2[1] == d[dn - 3] && h2[2] == d[dn - 2] && h2[3] == d[dn - 1];
}
MGPU_HDN bool ca_monero(const uint8_t* s, uint32_t n) {  // validate_monero_address
  uint8_t d[96], h[32];
  uint32_t dn;
  if (!ca_bs58_decode(s, n, d, dn) || dn < 5) return false;
  ca_keccak256(d, dn - 4, h);
  return h[0] == d[dn - 4] && h[1] == d[dn - 3] && h[2] == d[dn - 2] && h[3] == d[dn - 1];
}
// bech32 0.11.1 `decode` succeeds with hrp == "bc": characters after the LAST '1' are bech32 characters (either case), no
// mixed case in the whole string, the separator is the third character, at least 6 data characters, the checksum residue
// is the Bech32 (1) or the Bech32m (0x2bc830a3) constant.
MGPU_HD int ca_bech32_value(uint8_t c) {
  const char* CH = "qpzry9x8gf2tvdw0s3jn54khce6mua7l";
  if (c >= 'A' && c <= 'Z') c = (uint8_t)(c + 32);
  for (int i = 0; i < 32; i++) if ((uint8_t)CH[i] == c) return i;
  return -1;
}
MGPU_HDN bool ca_bitcoin_bech32(const uint8_t* s, uint32_t n) {
  bool upper = false, lower = false, seen_sep = false;
  uint32_t sep = 0;
  for (uint32_t i = n; i-- > 0;) {
    const uint8_t ch = s[i];
    if (ch > 127) return false;
    if (ch == '1' && !seen_sep) { seen_sep = true; sep = i; }
    else if (!seen_sep && ca_bech32_value(ch) < 0) return false;
    if (ch >= 'A' && ch <= 'Z') upper = true; else if (ch >= 'a' && ch <= 'z') lower = true;
  }
  if ((upper && lower) || !seen_sep || sep != 2) return false;
  if (!((s[0] | 32) == 'b' && (s[1] | 32) == 'c')) return false;
  if (n - 3 < 6) return false;
  const uint32_t GEN[5] = {0x3b6a57b2, 0x26508e6d, 0x1ea119fa, 0x3d4233dd, 0x2a1462b3};
  uint32_t chk = 1;
  const uint32_t pre[5] = {'b' >> 5, 'c' >> 5, 0, 'b' & 31, 'c' & 31};
  for (uint32_t i = 0; i < 5 + (n - 3); i++) {
    const uint32_t v = i < 5 ? pre[i] : (uint32_t)ca_bech32_value(s[3 + (i - 5)]);
    const uint32_t b = chk >> 25;
    chk = ((chk & 0x1ffffff) << 5) ^ v;
    for (uint32_t k = 0; k < 5; k++) if ((b >> k) & 1u) chk ^= GEN[k];
  }
  return chk == 1u || chk == 0x2bc830a3u;
}
// "0x" + 40 hex digits (both established by the caller): validate_ethereum_checksum
MGPU_HDN bool ca_ethereum(const uint8_t* s) {
  bool lower = false, upper = false;
  uint8_t lc[40], h[32];
  for (uint32_t i = 0; i < 40; i++) {
    const uint8_t c = s[2 + i];
    if (c >= 'a' && c <= 'f') lower = true;
    if (c >= 'A' && c <= 'F') upper = true;
    lc[i] = (c >= 'A' && c <= 'F') ? (uint8_t)(c + 32) : c;
  }
  if (!(lower && upper)) return true;
  ca_keccak256(lc, 40, h);
  for (uint32_t i = 0; i < 40; i++) {
    const uint8_t c = s[2 + i];
    if (!((c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'))) continue;
    const uint32_t nib = (i & 1u) ? (h[i >> 1] & 15u) : (h[i >> 1] >> 4);
    if ((c <= 'F') != (nib >= 8)) return false;
  }
  return true;
}

// A boundary-delimited word w[0..n) of 26..62 or 90..110 bytes (the tokenizer's crypto queue): which address type is it, if any?
// Returns MGPU_T_BITCOIN / MGPU_T_ETHEREUM / MGPU_T_MONERO or NONE32.  `flags` = enabled extractors.
MGPU_HDN uint32_t crypto_word_type(const uint8_t* w, uint32_t n, uint32_t flags) {
  if ((flags & MGPU_X_ETHEREUM) && n == 42 && w[0] == '0' && w[1] == 'x') {
    bool hex = true;
    for (uint32_t i = 2; i < 42 && hex; i++) hex = is_hex(w[i]);
    if (hex && ca_ethereum(w)) return MGPU_T_ETHEREUM;
  }
  if ((flags & MGPU_X_BITCOIN) && n >= 26 && n <= 62) {
    // (std::str::from_utf8 must succeed first: a validator that accepted would have rejected any byte >= 0x80 anyway)
    if (w[0] == 'b' && w[1] == 'c' && w[2] == '1') { if (ca_bitcoin_bech32(w, n)) return MGPU_T_BITCOIN; }
    else if (w[0] == '1' || w[0] == '3') { if (ca_bitcoin_base58(w, n)) return MGPU_T_BITCOIN; }
  }
  if ((flags & MGPU_X_MONERO) && n >= 90 && n <= 110 && (w[0] == '4' || w[0] == '8') && ca_monero(w, n)) return MGPU_T_MONERO;
  return NONE32;
}

}  // namespace mgpu
