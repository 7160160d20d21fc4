// Streaming Blake2b-512 and Keccak-256 for the Fiat-Shamir replay (one state per proof, in registers /
// local memory).  Byte rules restated from the reference transcript/mod.rs:
//   Blake2bRead::init   :118-134   Blake2b, 64-byte digest, personal "Halo2-Transcript" (blake2b_simd 1.x)
//   Keccak256Read::init :136-151   sha3 0.9.1 Keccak256 (original 0x01 padding), pre-loaded with "Halo2-Transcript"
//   squeeze_challenge   :209-214 / :239-254   finalise a CLONE, keep the running state
#pragma once
#include "field.cuh"

namespace h2v {

H2V_HD u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
H2V_HD u64 rotl64(u64 x, int n) { return (x << n) | (x >> (64 - n)); }

H2V_HD constexpr u64 blake2b_iv(int i) {
  constexpr u64 v[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                        0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  return v[i];
}
H2V_HD constexpr u8 blake2b_sigma(int r, int i) {
  constexpr u8 s[12][16] = {
      {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
      {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
      {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
      {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
      {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
      {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
  return s[r][i];
}

struct Blake2b {
  u64 h[8];
  u64 t;       // bytes compressed so far (proof transcripts are far below 2^64)
  u64 m[16];   // 128-byte block buffer as little-endian words
  u32 buflen;  // bytes currently in m

  H2V_HDN void init_halo2() {
    // parameter block: digest_length 64, fanout 1, depth 1, personal = "Halo2-Transcript"
    for (int i = 0; i < 8; i++) h[i] = blake2b_iv(i);
    h[0] ^= 0x01010040ull;
    h[6] ^= 0x72542d326f6c6148ull;  // "Halo2-Tr" little-endian
    h[7] ^= 0x7470697263736e61ull;  // "anscript"
    t = 0;
    buflen = 0;
    for (int i = 0; i < 16; i++) m[i] = 0;
  }

  H2V_HDN void compress(bool last) {
    u64 v[16];
    for (int i = 0; i < 8; i++) {
      v[i] = h[i];
      v[i + 8] = blake2b_iv(i);
    }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define H2V_B2G(a, b, c, d, x, y)      \
  v[a] = v[a] + v[b] + (x);            \
  v[d] = rotr64(v[d] ^ v[a], 32);      \
  v[c] = v[c] + v[d];                  \
  v[b] = rotr64(v[b] ^ v[c], 24);      \
  v[a] = v[a] + v[b] + (y);            \
  v[d] = rotr64(v[d] ^ v[a], 16);      \
  v[c] = v[c] + v[d];                  \
  v[b] = rotr64(v[b] ^ v[c], 63);
#pragma unroll
    for (int r = 0; r < 12; r++) {
      H2V_B2G(0, 4, 8, 12, m[blake2b_sigma(r, 0)], m[blake2b_sigma(r, 1)]);
      H2V_B2G(1, 5, 9, 13, m[blake2b_sigma(r, 2)], m[blake2b_sigma(r, 3)]);
      H2V_B2G(2, 6, 10, 14, m[blake2b_sigma(r, 4)], m[blake2b_sigma(r, 5)]);
      H2V_B2G(3, 7, 11, 15, m[blake2b_sigma(r, 6)], m[blake2b_sigma(r, 7)]);
      H2V_B2G(0, 5, 10, 15, m[blake2b_sigma(r, 8)], m[blake2b_sigma(r, 9)]);
      H2V_B2G(1, 6, 11, 12, m[blake2b_sigma(r, 10)], m[blake2b_sigma(r, 11)]);
      H2V_B2G(2, 7, 8, 13, m[blake2b_sigma(r, 12)], m[blake2b_sigma(r, 13)]);
      H2V_B2G(3, 4, 9, 14, m[blake2b_sigma(r, 14)], m[blake2b_sigma(r, 15)]);
    }
#undef H2V_B2G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
  }

  H2V_HD void flush_block() {  // only called when more input follows (the last block is finalised in digest)
    t += 128;
    compress(false);
    buflen = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = 0;
  }
  H2V_HD void update_byte(u8 b) {
    if (buflen == 128) flush_block();
    m[buflen >> 3] |= (u64)b << ((buflen & 7) * 8);
    buflen++;
  }
  // absorb 4 bytes (little-endian word) at any byte alignment: at most two read-modify-writes of the block buffer
  H2V_HD void update_u32(u32 w) {
    if (buflen == 128) flush_block();
    const u32 sh = (buflen & 7) * 8, wi = buflen >> 3;
    m[wi] |= (u64)w << sh;
    if (sh <= 32) {
      buflen += 4;
      return;
    }
    const u64 hi = (u64)w >> (64 - sh);  // the bytes that did not fit into word wi
    if (wi + 1 < 16) {
      m[wi + 1] |= hi;
      buflen += 4;
    } else {
      buflen = 128;
      flush_block();
      m[0] = hi;
      buflen = (sh - 32) / 8;
    }
  }
  H2V_HDN void update(const u8* data, u32 len) {
    for (u32 i = 0; i < len; i++) update_byte(data[i]);
  }
  // absorb a 256-bit little-endian value given as 8 u32 limbs
  H2V_HDN void update_limbs(const u32* l) {
#pragma unroll
    for (int i = 0; i < 8; i++) update_u32(l[i]);
  }
  // prefix byte + 256-bit value (the transcript's common_scalar / one coordinate of common_point)
  H2V_HDN void update_prefixed(u8 prefix, const u32* l) {
    update_byte(prefix);
#pragma unroll
    for (int i = 0; i < 8; i++) update_u32(l[i]);
  }
  // digest of a clone as 16 little-endian words; `this` keeps running
  H2V_HDN void digest_words(u32* out16) const {
    Blake2b c = *this;
    c.t += c.buflen;
    c.compress(true);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      out16[2 * i] = (u32)c.h[i];
      out16[2 * i + 1] = (u32)(c.h[i] >> 32);
    }
  }
  // digest of a clone; `this` keeps running
  H2V_HDN void digest(u8* out64) const {
    Blake2b c = *this;
    c.t += c.buflen;
    c.compress(true);
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 8; j++) out64[8 * i + j] = (u8)(c.h[i] >> (8 * j));
  }
};

H2V_HD constexpr u64 keccak_rc(int i) {
  constexpr u64 rc[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
                          0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
                          0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
                          0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
                          0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                          0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  return rc[i];
}

struct Keccak256 {
  u64 s[25];
  u32 pos;  // bytes absorbed into the current 136-byte block

  H2V_HDN void init_halo2() {
    for (int i = 0; i < 25; i++) s[i] = 0;
    pos = 0;
    const char* p = "Halo2-Transcript";
    for (int i = 0; i < 16; i++) update_byte((u8)p[i]);
  }
  H2V_HDN void permute() {
    constexpr int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    constexpr int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
      u64 bc[5];
#pragma unroll
      for (int i = 0; i < 5; i++) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
#pragma unroll
      for (int i = 0; i < 5; i++) {
        u64 t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
#pragma unroll
        for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
      }
      u64 t = s[1];
#pragma unroll
      for (int i = 0; i < 24; i++) {
        int j = piln[i];
        u64 b = s[j];
        s[j] = rotl64(t, rotc[i]);
        t = b;
      }
#pragma unroll
      for (int j = 0; j < 25; j += 5) {
#pragma unroll
        for (int i = 0; i < 5; i++) bc[i] = s[j + i];
#pragma unroll
        for (int i = 0; i < 5; i++) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
      }
      s[0] ^= keccak_rc(round);
    }
  }
  H2V_HD void update_byte(u8 b) {
    s[pos >> 3] ^= (u64)b << ((pos & 7) * 8);
    pos++;
    if (pos == 136) {
      permute();
      pos = 0;
    }
  }
  H2V_HD void update_u32(u32 w) {  // 4 bytes at any alignment inside the 136-byte rate
    if (pos + 4 > 136) {
      update_byte((u8)w);
      update_byte((u8)(w >> 8));
      update_byte((u8)(w >> 16));
      update_byte((u8)(w >> 24));
      return;
    }
    const u32 sh = (pos & 7) * 8, wi = pos >> 3;
    s[wi] ^= (u64)w << sh;
    if (sh > 32) s[wi + 1] ^= (u64)w >> (64 - sh);
    pos += 4;
    if (pos == 136) {
      permute();
      pos = 0;
    }
  }
  H2V_HDN void update(const u8* data, u32 len) {
    for (u32 i = 0; i < len; i++) update_byte(data[i]);
  }
  H2V_HDN void update_limbs(const u32* l) {
#pragma unroll
    for (int i = 0; i < 8; i++) update_u32(l[i]);
  }
  H2V_HDN void update_prefixed(u8 prefix, const u32* l) {
    update_byte(prefix);
#pragma unroll
    for (int i = 0; i < 8; i++) update_u32(l[i]);
  }
  H2V_HDN void digest_words_with_suffix(u8 suffix, u32* out8) const {
    Keccak256 c = *this;
    c.update_byte(suffix);
    c.s[c.pos >> 3] ^= (u64)0x01 << ((c.pos & 7) * 8);
    c.s[16] ^= 0x8000000000000000ull;  // last byte of the 136-byte rate
    c.permute();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      out8[2 * i] = (u32)c.s[i];
      out8[2 * i + 1] = (u32)(c.s[i] >> 32);
    }
  }
  // digest of a clone after absorbing one extra byte (the lo / hi challenge prefixes 10 / 11)
  H2V_HDN void digest_with_suffix(u8 suffix, u8* out32) const {
    Keccak256 c = *this;
    c.update_byte(suffix);
    c.s[c.pos >> 3] ^= (u64)0x01 << ((c.pos & 7) * 8);
    c.s[16] ^= 0x8000000000000000ull;  // last byte of the 136-byte rate
    c.permute();
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 8; j++) out32[8 * i + j] = (u8)(c.s[i] >> (8 * j));
  }
};

}  // namespace h2v
