/* TEST INFRASTRUCTURE ONLY (oracle): BN254 field / curve / pairing arithmetic in plain C.
 *
 * CPU restatement used as the checker of the CUDA path and as the CPU baseline of bench.py; never
 * linked into or called by the product (halo2-verifier_b200/).  The reference has none of this
 * in-tree: it calls halo2curves (git ChainSafe/halo2curves, branch no-std, un-pinned) at
 * transcript/mod.rs:161-172,220-228,502, arithmetic.rs:42-92, poly/kzg/msm.rs:78-86,186-202.
 * Representation: 4 x 64-bit limbs, Montgomery form, CIOS with unsigned __int128 (the portable
 * path halo2curves takes with default-features = false).  Independent of csrc/ (8 x 32-bit limbs).
 * PARITY: pinned only through the SRS fixture KATs and agreement with the Python oracle (tests/test_c_oracle.py);
 * at the level of challenges / accumulators / verdicts parity with the Rust binary is unpinned.
 */
#ifndef H2VO_FIELD_H
#define H2VO_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct {
  uint64_t l[4];
} fe;

typedef struct {
  uint64_t m[4];
  uint64_t inv; /* -m^-1 mod 2^64 */
  fe one, r2, r3;
} field_t;

extern field_t FQ, FR;
void h2vo_fields_init(void);

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
  return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int raw_geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static inline uint64_t raw_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
static inline uint64_t raw_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - b[i] - br;
    r[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}
static inline void fe_add(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[4];
  raw_add(t, a->l, b->l);
  if (raw_geq(t, F->m)) raw_sub(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_sub(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[4];
  if (raw_sub(t, a->l, b->l)) raw_add(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_neg(fe* r, const fe* a, const field_t* F) {
  if (fe_is_zero(a)) {
    *r = *a;
    return;
  }
  uint64_t t[4];
  raw_sub(t, F->m, a->l);
  memcpy(r->l, t, 32);
}
static inline void fe_dbl(fe* r, const fe* a, const field_t* F) { fe_add(r, a, a, F); }

/* Montgomery product a*b/2^256 mod m; b < m, a < 2^256 */
static inline void fe_mul(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    const uint64_t bi = b->l[i];
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * bi + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    const uint64_t mm = t[0] * F->inv;
    c = (u128)mm * F->m[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)mm * F->m[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  if (t[4] || raw_geq(t, F->m)) raw_sub(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_sqr(fe* r, const fe* a, const field_t* F) { fe_mul(r, a, a, F); }

void fe_pow(fe* r, const fe* a, const uint64_t e[4], const field_t* F);
void fe_inv(fe* r, const fe* a, const field_t* F); /* inv(0) = 0 */
void fe_from_canon(fe* r, const uint64_t c[4], const field_t* F);
void fe_to_canon(uint64_t c[4], const fe* a, const field_t* F);
void fe_from_u64(fe* r, uint64_t v, const field_t* F);
/* 32 little-endian bytes <-> canonical limbs */
static inline void le_load(uint64_t c[4], const uint8_t* b) {
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | b[8 * i + k];
    c[i] = v;
  }
}
static inline void le_store(uint8_t* b, const uint64_t c[4]) {
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 8; k++) b[8 * i + k] = (uint8_t)(c[i] >> (8 * k));
}
/* canonical 32-byte encoding -> field element; returns 0 if >= modulus (from_repr) */
int fe_from_repr(fe* r, const uint8_t* b, const field_t* F);
void fe_to_repr(uint8_t* b, const fe* a, const field_t* F);
/* ff::FromUniformBytes<64>: 512-bit little-endian integer mod r (transcript/mod.rs:502) */
void fr_from_uniform(fe* r, const uint8_t* b64);

/* ---- G1: y^2 = x^3 + 3 */
typedef struct {
  fe x, y;
  int inf;
} g1a;
typedef struct {
  fe X, Y, Z; /* Z = 0: identity */
} g1j;
void g1j_identity(g1j* r);
void g1j_from_affine(g1j* r, const g1a* a);
void g1j_double(g1j* r, const g1j* p);
void g1j_add_affine(g1j* r, const g1j* p, const g1a* q);
void g1j_add(g1j* r, const g1j* p, const g1j* q);
void g1j_to_affine(g1a* r, const g1j* p);
void g1_mul(g1j* r, const g1a* p, const uint64_t k[4]); /* canonical scalar */
int g1_on_curve(const g1a* p);
/* halo2curves compressed form (transcript/mod.rs:161-162): returns 0 on invalid encoding or identity */
int g1_decompress(g1a* r, const uint8_t* b32);
/* RawBytes: x | y Montgomery limbs little-endian (helpers.rs:40-98); checked = on-curve test */
int g1_read_raw(g1a* r, const uint8_t* b64, int checked);
/* affine canonical x | y (64 bytes LE), all-zero = identity */
void g1_to_bytes64(uint8_t* out, const g1a* p);
/* best_multiexp / multiexp_serial of the reference (arithmetic.rs:7-108) */
void g1_multiexp_serial(g1j* acc, const fe* scalars_mont, const g1a* bases, size_t n);

/* ---- tower and pairing */
typedef struct {
  fe c0, c1;
} fq2;
typedef struct {
  fq2 c0, c1, c2;
} fq6;
typedef struct {
  fq6 c0, c1;
} fq12;
typedef struct {
  fq2 x, y;
} g2a;
typedef struct {
  fq2 nlam, c; /* line: yP + (nlam * xP) w + c w^3 */
} g2line;
#define H2VO_ATE_LINES 102
typedef struct {
  g2line l[H2VO_ATE_LINES];
} g2prep;
int g2_read(g2a* r, const uint8_t* b, int fmt); /* fmt 0: compressed 64 B, else raw 128 B (1 = checked) */
void g2_neg(g2a* r, const g2a* a);
int g2_on_curve(const g2a* q);
void g2_prepare(g2prep* out, const g2a* q);
/* e(p0, Q0) e(p1, Q1) == 1 with prepared lines (DualMSM::check, msm.rs:185-203) */
int pairing_check2(const g1a* p0, const g2prep* q0, const g1a* p1, const g2prep* q1);
int h2vo_selftest(void); /* 0 = ok */
#endif
