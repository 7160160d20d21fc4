/* TEST INFRASTRUCTURE ONLY (oracle): `verify_proof` of ChainSafe/halo2-verifier restated in C.
 *
 * Follows (paths relative to /root/reference/halo2_verifier/src), in the order of this file:
 *   transcript/mod.rs:16-39,118-272,484-515     Blake2b / Keccak256 transcripts, Challenge255
 *   helpers.rs:7-166, poly/kzg/commitment.rs:155-207, plonk/vk.rs:76-115,274-365,514-546   byte formats
 *   poly/domain.rs:34-73,172-212                omega, 1/n, rotate_omega, l_i_range
 *   plonk/vk.rs:396-455,478-512,579-586         blinding_factors, query indices, expression evaluation
 *   lib.rs:33-425                               the driver
 *   plonk/permutation.rs:189-325, lookup.rs:159-271, shuffle.rs:148-225, vanishing.rs:92-136
 *   poly/kzg/multiopen/shplonk.rs:58-267, gwc.rs:54-163, arithmetic.rs:137-206
 *   poly/kzg/msm.rs:57-95,185-203, strategy.rs:125-176
 * One circuit instance per proof (what the C ABI of the product accepts).  Used by tests/ as the
 * full-size checker and by bench.py as the CPU baseline (threads over proofs).  Never called by the
 * product.  PARITY with the Rust binary: unpinned (no toolchain / vectors), see DESIGN.md section 2.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "h2vo_field.h"

/* ============================================================================ hashes */
typedef struct {
  uint64_t h[8], t;
  uint8_t buf[128];
  size_t len;
} blake2b_t;
static const uint64_t B2_IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                  0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
static const uint8_t B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static void b2_compress(blake2b_t* s, const uint8_t* block, int last) {
  uint64_t m[16], v[16];
  for (int i = 0; i < 16; i++) {
    uint64_t w = 0;
    for (int k = 7; k >= 0; k--) w = (w << 8) | block[8 * i + k];
    m[i] = w;
  }
  for (int i = 0; i < 8; i++) v[i] = s->h[i], v[8 + i] = B2_IV[i];
  v[12] ^= s->t;
  if (last) v[14] = ~v[14];
#define B2G(a, b, c, d, x, y)          \
  v[a] = v[a] + v[b] + (x);            \
  v[d] = rotr64(v[d] ^ v[a], 32);      \
  v[c] = v[c] + v[d];                  \
  v[b] = rotr64(v[b] ^ v[c], 24);      \
  v[a] = v[a] + v[b] + (y);            \
  v[d] = rotr64(v[d] ^ v[a], 16);      \
  v[c] = v[c] + v[d];                  \
  v[b] = rotr64(v[b] ^ v[c], 63);
  for (int r = 0; r < 12; r++) {
    const uint8_t* g = B2_SIGMA[r];
    B2G(0, 4, 8, 12, m[g[0]], m[g[1]]) B2G(1, 5, 9, 13, m[g[2]], m[g[3]]) B2G(2, 6, 10, 14, m[g[4]], m[g[5]])
    B2G(3, 7, 11, 15, m[g[6]], m[g[7]]) B2G(0, 5, 10, 15, m[g[8]], m[g[9]]) B2G(1, 6, 11, 12, m[g[10]], m[g[11]])
    B2G(2, 7, 8, 13, m[g[12]], m[g[13]]) B2G(3, 4, 9, 14, m[g[14]], m[g[15]])
  }
  for (int i = 0; i < 8; i++) s->h[i] ^= v[i] ^ v[8 + i];
}
static void b2_init_halo2(blake2b_t* s) { /* Blake2b-512, personal "Halo2-Transcript" (transcript/mod.rs:118-134) */
  memcpy(s->h, B2_IV, 64);
  s->h[0] ^= 0x01010040ull;
  uint64_t p0 = 0, p1 = 0;
  const char* pers = "Halo2-Transcript";
  for (int k = 7; k >= 0; k--) p0 = (p0 << 8) | (uint8_t)pers[k], p1 = (p1 << 8) | (uint8_t)pers[8 + k];
  s->h[6] ^= p0;
  s->h[7] ^= p1;
  s->t = 0;
  s->len = 0;
}
static void b2_update(blake2b_t* s, const uint8_t* d, size_t n) {
  while (n) {
    if (s->len == 128) { /* a full buffer is compressed only when more input follows */
      s->t += 128;
      b2_compress(s, s->buf, 0);
      s->len = 0;
    }
    size_t k = 128 - s->len;
    if (k > n) k = n;
    memcpy(s->buf + s->len, d, k);
    s->len += k;
    d += k;
    n -= k;
  }
}
static void b2_digest(const blake2b_t* s0, uint8_t out[64]) {
  blake2b_t s = *s0;
  s.t += s.len;
  memset(s.buf + s.len, 0, 128 - s.len);
  b2_compress(&s, s.buf, 1);
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(s.h[i] >> (8 * k));
}

typedef struct {
  uint64_t a[25];
  uint8_t buf[136];
  size_t len;
} keccak_t;
static const uint64_t K_RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
                                  0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
                                  0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
                                  0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                                  0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int K_ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
static inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t a[25]) {
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(a[x + 5 * y], K_ROT[x][y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= K_RC[rnd];
  }
}
static void k_absorb(keccak_t* s, const uint8_t* block) {
  for (int i = 0; i < 17; i++) {
    uint64_t w = 0;
    for (int k = 7; k >= 0; k--) w = (w << 8) | block[8 * i + k];
    s->a[i] ^= w;
  }
  keccak_f(s->a);
}
static void k_update(keccak_t* s, const uint8_t* d, size_t n) {
  while (n) {
    size_t k = 136 - s->len;
    if (k > n) k = n;
    memcpy(s->buf + s->len, d, k);
    s->len += k;
    d += k;
    n -= k;
    if (s->len == 136) {
      k_absorb(s, s->buf);
      s->len = 0;
    }
  }
}
static void k_init_halo2(keccak_t* s) { /* sha3 0.9.1 Keccak256 pre-loaded with "Halo2-Transcript" (:136-151) */
  memset(s, 0, sizeof(*s));
  k_update(s, (const uint8_t*)"Halo2-Transcript", 16);
}
static void k_digest(const keccak_t* s0, uint8_t out[32]) { /* original Keccak padding 0x01 .. 0x80 */
  keccak_t s = *s0;
  memset(s.buf + s.len, 0, 136 - s.len);
  s.buf[s.len] ^= 0x01;
  s.buf[135] ^= 0x80;
  k_absorb(&s, s.buf);
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(s.a[i] >> (8 * k));
}

/* ============================================================================ transcript */
enum { ST_OK = 0, ST_INVALID_INSTANCES = 1, ST_TRANSCRIPT = 2, ST_OPENING = 3, ST_CSF = 4, ST_PANIC = 5 };
typedef struct {
  int keccak;
  blake2b_t b2;
  keccak_t kc;
  const uint8_t* proof;
  size_t len, pos;
  int failed; /* io::Error raised */
  fe* chal;   /* squeeze log */
  uint32_t n_chal, cap_chal;
} transcript_t;
static void tr_update(transcript_t* t, const uint8_t* d, size_t n) {
  if (t->keccak) k_update(&t->kc, d, n);
  else b2_update(&t->b2, d, n);
}
static void tr_init(transcript_t* t, int keccak, const uint8_t* proof, size_t len) {
  memset(t, 0, sizeof(*t));
  t->keccak = keccak;
  if (keccak) k_init_halo2(&t->kc);
  else b2_init_halo2(&t->b2);
  t->proof = proof;
  t->len = len;
}
static fe tr_squeeze(transcript_t* t) {
  uint8_t z = 0, d[64];
  if (t->failed) { /* the reference has already returned Err: nothing is squeezed any more */
    fe zero;
    memset(&zero, 0, sizeof(zero));
    return zero;
  }
  tr_update(t, &z, 1);
  if (!t->keccak) {
    b2_digest(&t->b2, d);
  } else {
    keccak_t lo = t->kc, hi = t->kc;
    uint8_t a = 10, b = 11;
    k_update(&lo, &a, 1);
    k_update(&hi, &b, 1);
    k_digest(&lo, d);
    k_digest(&hi, d + 32);
  }
  fe c;
  fr_from_uniform(&c, d);
  if (t->n_chal == t->cap_chal) {
    t->cap_chal = t->cap_chal ? 2 * t->cap_chal : 16;
    t->chal = (fe*)realloc(t->chal, sizeof(fe) * t->cap_chal);
  }
  t->chal[t->n_chal++] = c;
  return c;
}
static void tr_common_scalar(transcript_t* t, const fe* s) {
  uint8_t p = 2, b[32];
  tr_update(t, &p, 1);
  fe_to_repr(b, s, &FR);
  tr_update(t, b, 32);
}
static void tr_common_point(transcript_t* t, const g1a* pt) {
  uint8_t p = 1, b[64];
  tr_update(t, &p, 1);
  g1_to_bytes64(b, pt);
  tr_update(t, b, 64);
}
static int tr_read_point(transcript_t* t, g1a* out) { /* transcript/mod.rs:153-166 (+ identity refused :218-219) */
  memset(out, 0, sizeof(*out));
  out->inf = 1;
  if (t->failed) return 0;
  if (t->pos + 32 > t->len || !g1_decompress(out, t->proof + t->pos)) {
    t->failed = 1;
    out->inf = 1;
    return 0;
  }
  t->pos += 32;
  tr_common_point(t, out);
  return 1;
}
static int tr_read_scalar(transcript_t* t, fe* out) {
  memset(out, 0, sizeof(*out));
  if (t->failed) return 0;
  if (t->pos + 32 > t->len || !fe_from_repr(out, t->proof + t->pos, &FR)) {
    t->failed = 1;
    return 0;
  }
  t->pos += 32;
  tr_common_scalar(t, out);
  return 1;
}

/* ============================================================================ formats */
typedef struct {
  uint32_t coeff, var_begin, var_end;
} term_t;
typedef struct {
  uint32_t term_begin, term_end;
} poly_t;
typedef struct {
  uint32_t n, *in_polys, *tab_polys;
} arg_t; /* lookup / shuffle: input and table expression lists */
typedef struct h2vo_vk {
  /* params */
  uint32_t pk;
  g1a g;
  g2a g2, s_g2;
  g2prep prep_s, prep_n;
  /* vk */
  uint32_t k, n_fixed_commit, cs_degree;
  g1a* fixed_commit;
  uint32_t n_fixed, n_advice, n_instance, n_selectors, n_challenges, n_gates, n_lookups, n_shuffles, n_coeff;
  uint8_t *advice_phase, *challenge_phase;
  uint32_t* n_advice_q;
  uint32_t n_aq;
  struct {
    uint32_t col;
    uint8_t phase;
    int32_t rot;
  }* aq;
  struct {
    uint32_t col;
    int32_t rot;
  } *iq, *fq;
  uint32_t n_perm;
  struct {
    uint32_t idx;
    uint8_t typ;
  }* perm_cols;
  poly_t* polys;
  uint32_t n_polys;
  term_t* terms;
  uint32_t n_terms;
  struct {
    uint32_t var, pow;
  }* vars;
  uint32_t n_vars;
  uint32_t* gate_polys;
  arg_t *lookups, *shuffles;
  fe* coeffs;
  g1a* perm_commit;
  fe transcript_repr;
  fe omega, omega_inv, inv_n, delta;
  char err[128];
} h2vo_vk;

typedef struct {
  const uint8_t* d;
  size_t n, p;
  int bad;
} rd_t;
static const uint8_t* rd_take(rd_t* r, size_t k) {
  static const uint8_t zeros[128] = {0};
  if (r->bad || r->p + k > r->n) {
    r->bad = 1;
    return zeros;
  }
  const uint8_t* q = r->d + r->p;
  r->p += k;
  return q;
}
static uint32_t rd_u32(rd_t* r) { /* big-endian (helpers.rs:120-164) */
  const uint8_t* b = rd_take(r, 4);
  return (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
}
static uint32_t rd_u16(rd_t* r) {
  const uint8_t* b = rd_take(r, 2);
  return (uint32_t)b[0] << 8 | b[1];
}
static uint8_t rd_u8(rd_t* r) { return rd_take(r, 1)[0]; }
static void rd_g1(rd_t* r, int fmt, g1a* out) {
  if (fmt == 0) {
    const uint8_t* b = rd_take(r, 32);
    int allz = 1;
    for (int i = 0; i < 32; i++) allz &= b[i] == 0;
    if (allz) {
      memset(out, 0, sizeof(*out));
      out->inf = 1;
    } else if (!g1_decompress(out, b)) {
      r->bad = 1;
    }
  } else if (!g1_read_raw(out, rd_take(r, 64), fmt == 1)) {
    r->bad = 1;
  }
}
static void rd_fr(rd_t* r, int fmt, fe* out) {
  const uint8_t* b = rd_take(r, 32);
  if (fmt == 0) {
    if (!fe_from_repr(out, b, &FR)) r->bad = 1;
  } else {
    le_load(out->l, b);
    if (raw_geq(out->l, FR.m)) r->bad = 1;
  }
}
static uint32_t vk_read_poly(h2vo_vk* v, rd_t* r) { /* plonk/vk.rs:514-546; returns the poly id */
  (void)rd_u32(r); /* num_vars */
  const uint32_t nt = rd_u32(r);
  if (r->bad || nt > (1u << 20)) {
    r->bad = 1;
    return 0;
  }
  v->polys = (poly_t*)realloc(v->polys, sizeof(poly_t) * (v->n_polys + 1));
  poly_t* p = &v->polys[v->n_polys];
  p->term_begin = v->n_terms;
  for (uint32_t t = 0; t < nt && !r->bad; t++) {
    v->terms = (term_t*)realloc(v->terms, sizeof(term_t) * (v->n_terms + 1));
    term_t* tm = &v->terms[v->n_terms++];
    tm->coeff = rd_u16(r);
    const uint32_t nv = rd_u32(r);
    if (r->bad || nv > (1u << 16)) {
      r->bad = 1;
      break;
    }
    tm->var_begin = v->n_vars;
    v->vars = realloc(v->vars, sizeof(*v->vars) * (v->n_vars + nv + 1));
    for (uint32_t i = 0; i < nv; i++) {
      v->vars[v->n_vars].var = rd_u32(r);
      v->vars[v->n_vars].pow = rd_u32(r);
      v->n_vars++;
    }
    tm->var_end = v->n_vars;
  }
  p->term_end = v->n_terms;
  return v->n_polys++;
}
static void vk_read_args(h2vo_vk* v, rd_t* r, arg_t** dst, uint32_t count) { /* lookup.rs:51-68 / shuffle.rs:85-102: interleaved pairs */
  *dst = (arg_t*)calloc(count ? count : 1, sizeof(arg_t));
  for (uint32_t i = 0; i < count && !r->bad; i++) {
    const uint32_t m = rd_u32(r);
    if (r->bad || m > (1u << 16)) {
      r->bad = 1;
      return;
    }
    (*dst)[i].n = m;
    (*dst)[i].in_polys = (uint32_t*)calloc(m ? m : 1, 4);
    (*dst)[i].tab_polys = (uint32_t*)calloc(m ? m : 1, 4);
    for (uint32_t j = 0; j < m && !r->bad; j++) {
      (*dst)[i].in_polys[j] = vk_read_poly(v, r);
      (*dst)[i].tab_polys[j] = vk_read_poly(v, r);
    }
  }
}

void h2vo_free(h2vo_vk* v) {
  if (!v) return;
  for (uint32_t i = 0; v->lookups && i < v->n_lookups; i++) free(v->lookups[i].in_polys), free(v->lookups[i].tab_polys);
  for (uint32_t i = 0; v->shuffles && i < v->n_shuffles; i++) free(v->shuffles[i].in_polys), free(v->shuffles[i].tab_polys);
  free(v->lookups), free(v->shuffles), free(v->fixed_commit), free(v->advice_phase), free(v->challenge_phase), free(v->n_advice_q);
  free(v->aq), free(v->iq), free(v->fq), free(v->perm_cols), free(v->polys), free(v->terms), free(v->vars), free(v->gate_polys);
  free(v->coeffs), free(v->perm_commit);
  free(v);
}

/* ParamsKZG::read_custom (commitment.rs:155-207) + VerifyingKey::read (vk.rs:76-115) + the domain constants */
int h2vo_load(const uint8_t* params, size_t plen, int pfmt, const uint8_t* vkb, size_t vlen, int vfmt, h2vo_vk** out) {
  h2vo_fields_init();
  h2vo_vk* v = (h2vo_vk*)calloc(1, sizeof(h2vo_vk));
  *out = v;
  rd_t r = {params, plen, 0, 0};
  const uint8_t* kb = rd_take(&r, 4);
  v->pk = (uint32_t)kb[0] | (uint32_t)kb[1] << 8 | (uint32_t)kb[2] << 16 | (uint32_t)kb[3] << 24; /* k is little-endian here */
  rd_g1(&r, pfmt, &v->g);
  const size_t g2sz = pfmt == 0 ? 64 : 128;
  if (!g2_read(&v->g2, rd_take(&r, g2sz), pfmt) || !g2_read(&v->s_g2, rd_take(&r, g2sz), pfmt)) r.bad = 1;
  if (r.bad) {
    snprintf(v->err, sizeof(v->err), "malformed params");
    return -1;
  }
  g2a ng2;
  g2_neg(&ng2, &v->g2);
  g2_prepare(&v->prep_s, &v->s_g2);
  g2_prepare(&v->prep_n, &ng2);

  r = (rd_t){vkb, vlen, 0, 0};
  v->k = rd_u32(&r);
  v->n_fixed_commit = rd_u32(&r);
  if (r.bad || v->k > 28 || v->n_fixed_commit > (1u << 20)) goto bad;
  v->fixed_commit = (g1a*)calloc(v->n_fixed_commit + 1, sizeof(g1a));
  for (uint32_t i = 0; i < v->n_fixed_commit; i++) rd_g1(&r, vfmt, &v->fixed_commit[i]);
  v->cs_degree = rd_u32(&r);
  v->n_fixed = rd_u32(&r);
  v->n_advice = rd_u32(&r);
  v->n_instance = rd_u32(&r);
  v->n_selectors = rd_u32(&r);
  v->n_challenges = rd_u32(&r);
  v->n_gates = rd_u32(&r);
  v->n_lookups = rd_u32(&r);
  v->n_shuffles = rd_u32(&r);
  v->n_coeff = rd_u32(&r);
  if (r.bad || (v->n_fixed | v->n_advice | v->n_instance | v->n_selectors | v->n_challenges | v->n_gates | v->n_lookups | v->n_shuffles | v->n_coeff) > (1u << 20))
    goto bad;
  v->advice_phase = (uint8_t*)calloc(v->n_advice + 1, 1);
  v->challenge_phase = (uint8_t*)calloc(v->n_challenges + 1, 1);
  for (uint32_t i = 0; i < v->n_advice; i++) v->advice_phase[i] = rd_u8(&r);
  for (uint32_t i = 0; i < v->n_challenges; i++) v->challenge_phase[i] = rd_u8(&r);
  v->n_advice_q = (uint32_t*)calloc(v->n_advice + 1, 4);
  for (uint32_t i = 0; i < v->n_advice; i++) {
    v->n_advice_q[i] = rd_u32(&r);
    if (v->n_advice_q[i] > (1u << 20)) r.bad = 1;
    v->n_aq += v->n_advice_q[i];
  }
  if (r.bad) goto bad;
  v->aq = calloc(v->n_aq + 1, sizeof(*v->aq));
  for (uint32_t i = 0; i < v->n_aq; i++) {
    v->aq[i].col = rd_u32(&r);
    v->aq[i].phase = rd_u8(&r);
    v->aq[i].rot = (int32_t)rd_u32(&r);
  }
  v->iq = calloc(v->n_instance + 1, sizeof(*v->iq));
  for (uint32_t i = 0; i < v->n_instance; i++) v->iq[i].col = rd_u32(&r), v->iq[i].rot = (int32_t)rd_u32(&r);
  v->fq = calloc(v->n_fixed + 1, sizeof(*v->fq));
  for (uint32_t i = 0; i < v->n_fixed; i++) v->fq[i].col = rd_u32(&r), v->fq[i].rot = (int32_t)rd_u32(&r);
  v->n_perm = rd_u32(&r);
  if (r.bad || v->n_perm > (1u << 20)) goto bad;
  v->perm_cols = calloc(v->n_perm + 1, sizeof(*v->perm_cols));
  for (uint32_t i = 0; i < v->n_perm; i++) {
    v->perm_cols[i].idx = rd_u32(&r);
    v->perm_cols[i].typ = rd_u8(&r);
    if (!(v->perm_cols[i].typ == 255 || v->perm_cols[i].typ == 254 || v->perm_cols[i].typ <= 2)) r.bad = 1; /* circuit.rs:53-65 */
  }
  v->gate_polys = (uint32_t*)calloc(v->n_gates + 1, 4);
  for (uint32_t i = 0; i < v->n_gates && !r.bad; i++) v->gate_polys[i] = vk_read_poly(v, &r);
  vk_read_args(v, &r, &v->lookups, v->n_lookups);
  vk_read_args(v, &r, &v->shuffles, v->n_shuffles);
  v->coeffs = (fe*)calloc(v->n_coeff + 1, sizeof(fe));
  for (uint32_t i = 0; i < v->n_coeff; i++) {
    rd_fr(&r, vfmt, &v->coeffs[i]);
  }
  v->perm_commit = (g1a*)calloc(v->n_perm + 1, sizeof(g1a));
  for (uint32_t i = 0; i < v->n_perm; i++) rd_g1(&r, vfmt, &v->perm_commit[i]);
  (void)rd_take(&r, (size_t)v->n_selectors * ((((size_t)1 << v->k) + 7) / 8));
  rd_fr(&r, vfmt, &v->transcript_repr);
  if (r.bad) goto bad;
  /* EvaluationDomain::new (domain.rs:34-73): omega = ROOT_OF_UNITY^(2^(S-k)), S = 28, ROOT = 7^((r-1)/2^28) */
  {
    fe seven;
    fe_from_u64(&seven, 7, &FR);
    uint64_t e[4], one[4] = {1, 0, 0, 0};
    raw_sub(e, FR.m, one);
    for (int s = 0; s < 28; s++)
      for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);
    fe root;
    fe_pow(&root, &seven, e, &FR);
    v->omega = root;
    for (uint32_t i = v->k; i < 28; i++) fe_sqr(&v->omega, &v->omega, &FR);
    fe_inv(&v->omega_inv, &v->omega, &FR);
    fe nn;
    fe_from_u64(&nn, (uint64_t)1 << v->k, &FR);
    fe_inv(&v->inv_n, &nn, &FR);
    v->delta = seven; /* DELTA = 7^(2^28) (permutation.rs:268) */
    for (int i = 0; i < 28; i++) fe_sqr(&v->delta, &v->delta, &FR);
  }
  return 0;
bad:
  snprintf(v->err, sizeof(v->err), "malformed verifying key");
  return -1;
}
const char* h2vo_error(const h2vo_vk* v) { return v ? v->err : "null"; }

/* ============================================================================ small containers */
typedef struct {
  fe s;
  g1a base;
} mterm;
typedef struct {
  mterm* t;
  size_t n, cap;
} msm_t;
static void msm_push(msm_t* m, const fe* s, const g1a* b) { /* msm.rs:57-60 append_term */
  if (m->n == m->cap) {
    m->cap = m->cap ? 2 * m->cap : 32;
    m->t = (mterm*)realloc(m->t, sizeof(mterm) * m->cap);
  }
  m->t[m->n].s = *s;
  m->t[m->n].base = *b;
  m->n++;
}
static void msm_scale(msm_t* m, const fe* f) { /* msm.rs:67-75 */
  for (size_t i = 0; i < m->n; i++) fe_mul(&m->t[i].s, &m->t[i].s, f, &FR);
}
static void msm_extend(msm_t* d, const msm_t* s) { /* msm.rs:62-65 add_msm */
  for (size_t i = 0; i < s->n; i++) msm_push(d, &s->t[i].s, &s->t[i].base);
}
static void msm_eval(g1a* out, const msm_t* m) { /* msm.rs:81-86: batch_normalize (bases are affine already) + best_multiexp */
  fe* sc = (fe*)malloc(sizeof(fe) * (m->n ? m->n : 1));
  g1a* bs = (g1a*)malloc(sizeof(g1a) * (m->n ? m->n : 1));
  for (size_t i = 0; i < m->n; i++) sc[i] = m->t[i].s, bs[i] = m->t[i].base;
  g1j acc;
  g1j_identity(&acc);
  g1_multiexp_serial(&acc, sc, bs, m->n);
  g1j_to_affine(out, &acc);
  free(sc);
  free(bs);
}

#define ID_FIXED 1000000
#define ID_SIGMA 2000000
#define ID_HMSM 3000000
typedef struct {
  int ident; /* proof point slot, ID_FIXED + i, ID_SIGMA + i, ID_HMSM: commitment identity (query.rs:63-74 compares pointers) */
  fe point, eval;
  const g1a* pt; /* NULL for the h MSM */
} query_t;

static void batch_invert_skip_zero(fe* v, size_t n) { /* ff::BatchInvert: zeros are skipped and stay zero */
  fe* pre = (fe*)malloc(sizeof(fe) * (n + 1));
  fe acc = FR.one;
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    if (!fe_is_zero(&v[i])) fe_mul(&acc, &acc, &v[i], &FR);
  }
  fe_inv(&acc, &acc, &FR);
  for (size_t i = n; i-- > 0;) {
    if (fe_is_zero(&v[i])) continue;
    fe t;
    fe_mul(&t, &acc, &pre[i], &FR);
    fe_mul(&acc, &acc, &v[i], &FR);
    v[i] = t;
  }
  free(pre);
}
static fe rotate_omega(const h2vo_vk* v, const fe* val, int32_t rot) { /* domain.rs:172-182 */
  fe w = rot >= 0 ? v->omega : v->omega_inv, acc = *val;
  uint32_t e = (uint32_t)(rot >= 0 ? rot : -rot);
  fe base = w;
  while (e) {
    if (e & 1) fe_mul(&acc, &acc, &base, &FR);
    fe_sqr(&base, &base, &FR);
    e >>= 1;
  }
  return acc;
}
/* l_i(x) for rot in [lo, lo + cnt)   (domain.rs:187-212) */
static fe* l_i_range(const h2vo_vk* v, const fe* x, const fe* xn, int32_t lo, uint32_t cnt) {
  fe* res = (fe*)malloc(sizeof(fe) * (cnt + 1));
  for (uint32_t i = 0; i < cnt; i++) {
    fe w = rotate_omega(v, &FR.one, lo + (int32_t)i);
    fe_sub(&res[i], x, &w, &FR);
  }
  batch_invert_skip_zero(res, cnt);
  fe common;
  fe_sub(&common, xn, &FR.one, &FR);
  fe_mul(&common, &common, &v->inv_n, &FR);
  for (uint32_t i = 0; i < cnt; i++) {
    fe t;
    fe_mul(&t, &res[i], &common, &FR);
    res[i] = rotate_omega(v, &t, lo + (int32_t)i);
  }
  return res;
}
static void fe_pow_u32(fe* r, const fe* a, uint32_t e) { /* pow_vartime (vk.rs:583) */
  fe acc = FR.one, base = *a;
  while (e) {
    if (e & 1) fe_mul(&acc, &acc, &base, &FR);
    fe_sqr(&base, &base, &FR);
    e >>= 1;
  }
  *r = acc;
}
/* IndexedExpressionPoly::evaluate (vk.rs:478-512,579-586); returns 0 when the reference would panic (no terms) */
static int eval_poly(const h2vo_vk* v, uint32_t pid, const fe* vars, uint32_t nvars, fe* out) {
  const poly_t* p = &v->polys[pid];
  if (p->term_begin == p->term_end) return 0;
  fe acc;
  memset(&acc, 0, sizeof(acc));
  for (uint32_t t = p->term_begin; t < p->term_end; t++) {
    const term_t* tm = &v->terms[t];
    fe prod = FR.one;
    for (uint32_t i = tm->var_begin; i < tm->var_end; i++) {
      if (v->vars[i].var >= nvars) return 0;
      fe pw;
      fe_pow_u32(&pw, &vars[v->vars[i].var], v->vars[i].pow);
      fe_mul(&prod, &prod, &pw, &FR);
    }
    if (tm->coeff >= v->n_coeff) return 0;
    fe_mul(&prod, &v->coeffs[tm->coeff], &prod, &FR);
    fe_add(&acc, &acc, &prod, &FR);
  }
  *out = acc;
  return 1;
}

typedef struct {
  int status;
  uint32_t n_chal;
  fe chal[64];
  g1a L, R;
  int have_lr;
} result_t;

/* ============================================================================ multiopen */
static int fe_in(const fe* set, size_t n, const fe* x) {
  for (size_t i = 0; i < n; i++)
    if (fe_eq(&set[i], x)) return 1;
  return 0;
}
static void add_commit(msm_t* dst, const query_t* q, const msm_t* hmsm, const fe* scale) {
  if (q->pt) {
    msm_push(dst, scale, q->pt);
  } else {
    for (size_t i = 0; i < hmsm->n; i++) {
      fe s;
      fe_mul(&s, &hmsm->t[i].s, scale, &FR);
      msm_push(dst, &s, &hmsm->t[i].base);
    }
  }
}
/* shplonk.rs:58-149 + 175-267 */
static int shplonk_verify(const h2vo_vk* v, transcript_t* tr, const query_t* q, size_t nq, const msm_t* hmsm, msm_t* left, msm_t* right) {
  /* commitment map in first-appearance order */
  int* idents = (int*)malloc(sizeof(int) * nq);
  size_t nc = 0;
  fe* sup = (fe*)malloc(sizeof(fe) * nq);
  size_t nsup = 0;
  for (size_t i = 0; i < nq; i++) {
    if (!fe_in(sup, nsup, &q[i].point)) sup[nsup++] = q[i].point;
    size_t c = 0;
    for (; c < nc; c++)
      if (idents[c] == q[i].ident) break;
    if (c == nc) idents[nc++] = q[i].ident;
  }
  /* point set of each commitment, then groups of commitments with equal sets */
  fe** cpts = (fe**)malloc(sizeof(fe*) * nc);
  size_t* cn = (size_t*)calloc(nc, sizeof(size_t));
  for (size_t c = 0; c < nc; c++) {
    cpts[c] = (fe*)malloc(sizeof(fe) * nq);
    for (size_t i = 0; i < nq; i++)
      if (q[i].ident == idents[c] && !fe_in(cpts[c], cn[c], &q[i].point)) cpts[c][cn[c]++] = q[i].point;
  }
  int* set_of = (int*)malloc(sizeof(int) * nc);
  size_t* set_rep = (size_t*)malloc(sizeof(size_t) * nc);
  size_t nsets = 0;
  for (size_t c = 0; c < nc; c++) {
    size_t s = 0;
    for (; s < nsets; s++) {
      const size_t r = set_rep[s];
      if (cn[r] != cn[c]) continue;
      int same = 1;
      for (size_t k = 0; k < cn[c]; k++) same &= fe_in(cpts[r], cn[r], &cpts[c][k]);
      if (same) break;
    }
    if (s == nsets) set_rep[nsets++] = c;
    set_of[c] = (int)s;
  }
  fe y = tr_squeeze(tr), vv = tr_squeeze(tr), u;
  g1a h1, h2;
  tr_read_point(tr, &h1);
  u = tr_squeeze(tr);
  tr_read_point(tr, &h2);
  int rc = ST_OK;
  if (tr->failed) {
    rc = ST_OPENING;
    goto done;
  }
  {
    fe z0, z0_diff_inv, r_outer, pow_v = FR.one;
    memset(&z0, 0, sizeof(z0));
    memset(&z0_diff_inv, 0, sizeof(z0_diff_inv));
    memset(&r_outer, 0, sizeof(r_outer));
    msm_t outer = {0, 0, 0};
    for (size_t s = 0; s < nsets; s++) {
      const fe* pts = cpts[set_rep[s]];
      const size_t m = cn[set_rep[s]];
      fe zdiff = FR.one, t;
      for (size_t i = 0; i < nsup; i++)
        if (!fe_in(pts, m, &sup[i])) {
          fe_sub(&t, &u, &sup[i], &FR);
          fe_mul(&zdiff, &t, &zdiff, &FR);
        }
      if (s == 0) {
        z0 = FR.one;
        for (size_t i = 0; i < m; i++) {
          fe_sub(&t, &u, &pts[i], &FR);
          fe_mul(&z0, &t, &z0, &FR);
        }
        if (fe_is_zero(&zdiff)) { /* shplonk.rs:215 unwrap */
          rc = ST_PANIC;
          free(outer.t);
          goto done;
        }
        fe_inv(&z0_diff_inv, &zdiff, &FR);
        zdiff = FR.one;
      } else {
        fe_mul(&zdiff, &zdiff, &z0_diff_inv, &FR);
      }
      msm_t inner = {0, 0, 0};
      fe r_inner, pow_y = FR.one;
      memset(&r_inner, 0, sizeof(r_inner));
      for (size_t c = 0; c < nc; c++) {
        if (set_of[c] != (int)s) continue;
        /* evals of this commitment at the set's points, in the points' order */
        fe evals[16], denom[16], coef[17], fin[17];
        const query_t* qc = NULL;
        for (size_t k = 0; k < m; k++)
          for (size_t i = 0; i < nq; i++)
            if (q[i].ident == idents[c] && fe_eq(&q[i].point, &pts[k])) {
              evals[k] = q[i].eval;
              qc = &q[i];
              break;
            }
        /* lagrange_interpolate (arithmetic.rs:149-202), coefficient form */
        memset(fin, 0, sizeof(fin));
        if (m == 1) {
          fin[0] = evals[0];
        } else {
          for (size_t j = 0; j < m; j++) {
            size_t nd = 0;
            for (size_t k = 0; k < m; k++)
              if (k != j) fe_sub(&denom[nd++], &pts[j], &pts[k], &FR);
            batch_invert_skip_zero(denom, nd);
            size_t len = 1;
            coef[0] = FR.one;
            nd = 0;
            for (size_t k = 0; k < m; k++) {
              if (k == j) continue;
              const fe d = denom[nd++];
              fe ndx, prod[17];
              fe_mul(&ndx, &d, &pts[k], &FR);
              fe_neg(&ndx, &ndx, &FR);
              for (size_t i = 0; i <= len; i++) {
                fe a, b;
                memset(&a, 0, sizeof(a));
                memset(&b, 0, sizeof(b));
                if (i < len) fe_mul(&a, &coef[i], &ndx, &FR);
                if (i > 0) fe_mul(&b, &coef[i - 1], &d, &FR);
                fe_add(&prod[i], &a, &b, &FR);
              }
              len++;
              memcpy(coef, prod, sizeof(fe) * len);
            }
            for (size_t i = 0; i < len; i++) {
              fe_mul(&t, &coef[i], &evals[j], &FR);
              fe_add(&fin[i], &fin[i], &t, &FR);
            }
          }
        }
        fe r_eval;
        memset(&r_eval, 0, sizeof(r_eval));
        for (size_t i = m; i-- > 0;) { /* eval_polynomial (arithmetic.rs:137-144) */
          fe_mul(&r_eval, &r_eval, &u, &FR);
          fe_add(&r_eval, &r_eval, &fin[i], &FR);
        }
        fe_mul(&r_eval, &r_eval, &pow_y, &FR);
        add_commit(&inner, qc, hmsm, &pow_y);
        fe_add(&r_inner, &r_inner, &r_eval, &FR);
        fe_mul(&pow_y, &pow_y, &y, &FR);
      }
      fe sc;
      fe_mul(&sc, &pow_v, &zdiff, &FR);
      msm_scale(&inner, &sc);
      msm_extend(&outer, &inner);
      free(inner.t);
      fe_mul(&t, &pow_v, &r_inner, &FR);
      fe_mul(&t, &t, &zdiff, &FR);
      fe_add(&r_outer, &r_outer, &t, &FR);
      fe_mul(&pow_v, &pow_v, &vv, &FR);
    }
    fe nr, nz;
    fe_neg(&nr, &r_outer, &FR);
    fe_neg(&nz, &z0, &FR);
    msm_push(&outer, &nr, &v->g);
    msm_push(&outer, &nz, &h1);
    msm_push(&outer, &u, &h2);
    msm_push(left, &FR.one, &h2);
    msm_extend(right, &outer);
    free(outer.t);
  }
done:
  for (size_t c = 0; c < nc; c++) free(cpts[c]);
  free(cpts), free(cn), free(set_of), free(set_rep), free(idents), free(sup);
  return rc;
}
/* gwc.rs:54-163 */
static int gwc_verify(const h2vo_vk* v, transcript_t* tr, const query_t* q, size_t nq, const msm_t* hmsm, msm_t* left, msm_t* right) {
  fe vv = tr_squeeze(tr);
  fe* pts = (fe*)malloc(sizeof(fe) * nq);
  size_t np = 0;
  for (size_t i = 0; i < nq; i++)
    if (!fe_in(pts, np, &q[i].point)) pts[np++] = q[i].point;
  g1a* w = (g1a*)malloc(sizeof(g1a) * (np + 1));
  for (size_t i = 0; i < np; i++) tr_read_point(tr, &w[i]);
  fe u = tr_squeeze(tr);
  int rc = ST_OK;
  if (tr->failed) {
    rc = ST_OPENING;
  } else {
    msm_t cm = {0, 0, 0}, wit = {0, 0, 0}, wita = {0, 0, 0};
    fe eval_multi, pow_u = FR.one, t;
    memset(&eval_multi, 0, sizeof(eval_multi));
    for (size_t p = 0; p < np; p++) {
      msm_t batch = {0, 0, 0};
      fe eval_batch, pow_v = FR.one;
      memset(&eval_batch, 0, sizeof(eval_batch));
      for (size_t i = 0; i < nq; i++) {
        if (!fe_eq(&q[i].point, &pts[p])) continue;
        add_commit(&batch, &q[i], hmsm, &pow_v);
        fe_mul(&t, &pow_v, &q[i].eval, &FR);
        fe_add(&eval_batch, &eval_batch, &t, &FR);
        fe_mul(&pow_v, &pow_v, &vv, &FR);
      }
      msm_scale(&batch, &pow_u);
      msm_extend(&cm, &batch);
      free(batch.t);
      fe_mul(&t, &pow_u, &eval_batch, &FR);
      fe_add(&eval_multi, &eval_multi, &t, &FR);
      fe_mul(&t, &pow_u, &pts[p], &FR);
      msm_push(&wita, &t, &w[p]);
      msm_push(&wit, &pow_u, &w[p]);
      fe_mul(&pow_u, &pow_u, &u, &FR);
    }
    msm_extend(left, &wit);
    msm_extend(right, &wita);
    msm_extend(right, &cm);
    g1a ng = v->g;
    fe_neg(&ng.y, &ng.y, &FQ);
    msm_push(right, &eval_multi, &ng);
    free(cm.t), free(wit.t), free(wita.t);
  }
  free(pts), free(w);
  return rc;
}

/* ============================================================================ the driver (lib.rs:33-425) */
static int find_query(const h2vo_vk* v, uint32_t idx, uint8_t typ) { /* get_any_query_index (vk.rs:413-455), rotation 0 */
  if (typ == 255) {
    for (uint32_t i = 0; i < v->n_fixed; i++)
      if (v->fq[i].col == idx && v->fq[i].rot == 0) return (int)i;
  } else if (typ == 254) {
    for (uint32_t i = 0; i < v->n_instance; i++)
      if (v->iq[i].col == idx && v->iq[i].rot == 0) return (int)i;
  } else {
    for (uint32_t i = 0; i < v->n_aq; i++)
      if (v->aq[i].col == idx && v->aq[i].phase == typ && v->aq[i].rot == 0) return (int)i;
  }
  return -1;
}

/* m = circuit instances carried by the proof (`instances.len()`, lib.rs:63,92,117,134); the instance columns are laid out
 * instance-major (ncols = m x instance columns of the VK), everything per instance repeats in the reference's interleaving */
static void verify_one(const h2vo_vk* v, const uint8_t* proof, size_t plen, const uint8_t* inst, const uint32_t* col_len, uint32_t ncols,
                       uint32_t m, int multiopen, int keccak, int check_pairing, result_t* res) {
  memset(res, 0, sizeof(*res));
  res->L.inf = res->R.inf = 1;
  if (m == 0 || ncols != v->n_instance * m) { /* lib.rs:51-55 */
    res->status = ST_INVALID_INSTANCES;
    return;
  }
  uint32_t n_inst = 0, max_len = 0;
  for (uint32_t c = 0; c < ncols; c++) {
    n_inst += col_len[c];
    if (col_len[c] > max_len) max_len = col_len[c];
  }
  fe* ivals = (fe*)malloc(sizeof(fe) * (n_inst + 1));
  for (uint32_t i = 0; i < n_inst; i++)
    if (!fe_from_repr(&ivals[i], inst + 32 * (size_t)i, &FR)) { /* not an Fr: cannot even be passed to the reference */
      res->status = ST_INVALID_INSTANCES;
      free(ivals);
      return;
    }
  transcript_t tr;
  tr_init(&tr, keccak, proof, plen);
  tr_common_scalar(&tr, &v->transcript_repr); /* lib.rs:66 */
  for (uint32_t i = 0; i < n_inst; i++) tr_common_scalar(&tr, &ivals[i]); /* lib.rs:76-82 */

  const uint32_t chunk = v->cs_degree - 2; /* permutation.rs:72 */
  const uint32_t n_sets = v->n_perm ? (v->n_perm + chunk - 1) / chunk : 0;
  const uint32_t n_h = v->cs_degree - 1;
  const uint32_t n_pts_max = m * (v->n_advice + 3 * v->n_lookups + n_sets + v->n_shuffles) + 1 + n_h + 8;
  g1a* P = (g1a*)calloc(n_pts_max, sizeof(g1a));
  uint32_t np = 0;
#define READ_POINT() (tr_read_point(&tr, &P[np]), (int)np++)
  int* adv_slot = (int*)malloc(sizeof(int) * ((size_t)m * v->n_advice + 1)); /* [pi][column] */
  fe* user_chal = (fe*)calloc(v->n_challenges + 1, sizeof(fe));
  uint8_t max_phase = 0;
  for (uint32_t i = 0; i < v->n_advice; i++)
    if (v->advice_phase[i] > max_phase) max_phase = v->advice_phase[i];
  for (uint32_t ph = 0; ph <= max_phase; ph++) { /* lib.rs:91-109 */
    for (uint32_t pi = 0; pi < m; pi++)
      for (uint32_t c = 0; c < v->n_advice; c++)
        if (v->advice_phase[c] == ph) adv_slot[(size_t)pi * v->n_advice + c] = READ_POINT();
    for (uint32_t c = 0; c < v->n_challenges; c++)
      if (v->challenge_phase[c] == ph) user_chal[c] = tr_squeeze(&tr);
  }
  const fe theta = tr_squeeze(&tr);
  int* lk_in = (int*)malloc(sizeof(int) * ((size_t)m * v->n_lookups + 1)); /* all of these: [pi][index] */
  int* lk_tab = (int*)malloc(sizeof(int) * ((size_t)m * v->n_lookups + 1));
  int* lk_prod = (int*)malloc(sizeof(int) * ((size_t)m * v->n_lookups + 1));
  int* sh_prod = (int*)malloc(sizeof(int) * ((size_t)m * v->n_shuffles + 1));
  int* pm_slot = (int*)malloc(sizeof(int) * ((size_t)m * n_sets + 1));
  int* h_slot = (int*)malloc(sizeof(int) * (n_h + 1));
  for (uint32_t i = 0; i < m * v->n_lookups; i++) lk_in[i] = READ_POINT(), lk_tab[i] = READ_POINT();
  const fe beta = tr_squeeze(&tr), gamma = tr_squeeze(&tr);
  for (uint32_t i = 0; i < m * n_sets; i++) pm_slot[i] = READ_POINT();
  for (uint32_t i = 0; i < m * v->n_lookups; i++) lk_prod[i] = READ_POINT();
  for (uint32_t i = 0; i < m * v->n_shuffles; i++) sh_prod[i] = READ_POINT();
  const int random_slot = READ_POINT(); /* vanishing.rs:49-57 */
  const fe y = tr_squeeze(&tr);
  for (uint32_t i = 0; i < n_h; i++) h_slot[i] = READ_POINT(); /* vanishing.rs:61-73 */
  const fe x = tr_squeeze(&tr);

  /* instance evals (lib.rs:180-217) */
  fe xn = x;
  for (uint32_t i = 0; i < v->pk; i++) fe_sqr(&xn, &xn, &FR);
  int32_t min_rot = 0, max_rot = 0;
  for (uint32_t i = 0; i < v->n_instance; i++) {
    if (v->iq[i].rot < min_rot) min_rot = v->iq[i].rot;
    else if (v->iq[i].rot > max_rot) max_rot = v->iq[i].rot;
  }
  fe* lis = l_i_range(v, &x, &xn, -max_rot, (uint32_t)max_rot + max_len + (uint32_t)(-min_rot));
  const uint32_t nvars = v->n_aq + v->n_fixed + v->n_instance + v->n_challenges;
  /* per instance: advice | fixed | instance | challenges (vk.rs:490-500); instance pi at vars + pi * nvars */
  fe* vars = (fe*)calloc((size_t)m * nvars + 1, sizeof(fe));
#define ADV_EV(pi) (vars + (size_t)(pi) * nvars)
#define FIX_EV(pi) (ADV_EV(pi) + v->n_aq)
#define INS_EV(pi) (FIX_EV(pi) + v->n_fixed)
  for (uint32_t pi = 0; pi < m; pi++) {
    memcpy(INS_EV(pi) + v->n_instance, user_chal, sizeof(fe) * v->n_challenges);
    for (uint32_t qi = 0; qi < v->n_instance; qi++) {
      const uint32_t col = pi * v->n_instance + v->iq[qi].col;
      uint32_t cbeg = 0;
      for (uint32_t c = 0; c < col && c < ncols; c++) cbeg += col_len[c];
      const uint32_t clen = (v->iq[qi].col < v->n_instance && col < ncols) ? col_len[col] : 0, off = (uint32_t)(max_rot - v->iq[qi].rot);
      fe acc, t;
      memset(&acc, 0, sizeof(acc));
      for (uint32_t i = 0; i < clen; i++) {
        fe_mul(&t, &ivals[cbeg + i], &lis[off + i], &FR);
        fe_add(&acc, &acc, &t, &FR);
      }
      INS_EV(pi)[qi] = acc;
    }
  }
  free(lis);
  for (uint32_t pi = 0; pi < m; pi++)
    for (uint32_t i = 0; i < v->n_aq; i++) tr_read_scalar(&tr, &ADV_EV(pi)[i]);
  for (uint32_t i = 0; i < v->n_fixed; i++) tr_read_scalar(&tr, &FIX_EV(0)[i]);
  for (uint32_t pi = 1; pi < m; pi++) memcpy(FIX_EV(pi), FIX_EV(0), sizeof(fe) * v->n_fixed);
  fe* fix_ev = FIX_EV(0);
  fe random_eval;
  tr_read_scalar(&tr, &random_eval);
  fe* sigma_ev = (fe*)calloc(v->n_perm + 1, sizeof(fe));
  for (uint32_t i = 0; i < v->n_perm; i++) tr_read_scalar(&tr, &sigma_ev[i]);
  fe(*pm_ev_all)[3] = calloc((size_t)m * n_sets + 1, sizeof(*pm_ev_all)); /* [pi][set]: eval, next, last (permutation.rs:105-131) */
  for (uint32_t i = 0; i < m * n_sets; i++) {
    tr_read_scalar(&tr, &pm_ev_all[i][0]);
    tr_read_scalar(&tr, &pm_ev_all[i][1]);
    if (i % n_sets != n_sets - 1) tr_read_scalar(&tr, &pm_ev_all[i][2]);
  }
  fe(*lk_ev_all)[5] = calloc((size_t)m * v->n_lookups + 1, sizeof(*lk_ev_all)); /* product, product_next, input, input_inv, table */
  for (uint32_t i = 0; i < m * v->n_lookups; i++)
    for (int k = 0; k < 5; k++) tr_read_scalar(&tr, &lk_ev_all[i][k]);
  fe(*sh_ev_all)[2] = calloc((size_t)m * v->n_shuffles + 1, sizeof(*sh_ev_all));
  for (uint32_t i = 0; i < m * v->n_shuffles; i++)
    for (int k = 0; k < 2; k++) tr_read_scalar(&tr, &sh_ev_all[i][k]);
  msm_t hmsm = {0, 0, 0}, left = {0, 0, 0}, right = {0, 0, 0};
  query_t* q = NULL;
  if (tr.failed) {
    res->status = ST_TRANSCRIPT;
    goto out;
  }
  {
    /* vanishing argument (lib.rs:257-347) */
    uint32_t factors = 1;
    if (v->n_advice) {
      factors = 0;
      for (uint32_t i = 0; i < v->n_advice; i++)
        if (v->n_advice_q[i] > factors) factors = v->n_advice_q[i];
    }
    const uint32_t bf = (factors > 3 ? factors : 3) + 2; /* vk.rs:396-401 */
    fe* le = l_i_range(v, &x, &xn, -(int32_t)(bf + 1), bf + 2);
    fe l_last = le[0], l_0 = le[bf + 1], l_blind, active, t, one = FR.one;
    memset(&l_blind, 0, sizeof(l_blind));
    for (uint32_t i = 1; i <= bf; i++) fe_add(&l_blind, &l_blind, &le[i], &FR);
    free(le);
    fe_add(&t, &l_last, &l_blind, &FR);
    fe_sub(&active, &one, &t, &FR);
    fe h;
    memset(&h, 0, sizeof(h));
    int panic = 0;
#define FOLD(e)                    \
  do {                             \
    fe_mul(&h, &h, &y, &FR);       \
    fe_add(&h, &h, &(e), &FR);     \
  } while (0)
    for (uint32_t pi = 0; pi < m && !panic; pi++) { /* lib.rs:273-344: all expressions of instance pi, then the next instance */
    const fe *vars_pi = ADV_EV(pi), *adv_ev = ADV_EV(pi), *ins_ev = INS_EV(pi);
    fe(*pm_ev)[3] = pm_ev_all + (size_t)pi * n_sets;
    fe(*lk_ev)[5] = lk_ev_all + (size_t)pi * v->n_lookups;
    fe(*sh_ev)[2] = sh_ev_all + (size_t)pi * v->n_shuffles;
    for (uint32_t g = 0; g < v->n_gates && !panic; g++) {
      fe e;
      if (!eval_poly(v, v->gate_polys[g], vars_pi, nvars, &e)) panic = 1;
      else FOLD(e);
    }
    if (n_sets && !panic) { /* permutation.rs:189-288 */
      fe e;
      fe_sub(&t, &one, &pm_ev[0][0], &FR);
      fe_mul(&e, &l_0, &t, &FR);
      FOLD(e);
      fe_sqr(&t, &pm_ev[n_sets - 1][0], &FR);
      fe_sub(&t, &t, &pm_ev[n_sets - 1][0], &FR);
      fe_mul(&e, &t, &l_last, &FR);
      FOLD(e);
      for (uint32_t s = 1; s < n_sets; s++) {
        fe_sub(&t, &pm_ev[s][0], &pm_ev[s - 1][2], &FR);
        fe_mul(&e, &t, &l_0, &FR);
        FOLD(e);
      }
      for (uint32_t s = 0; s < n_sets && !panic; s++) {
        const uint32_t c0 = s * chunk, c1 = (s + 1) * chunk < v->n_perm ? (s + 1) * chunk : v->n_perm;
        fe lft = pm_ev[s][1], rgt = pm_ev[s][0], cur, dpow;
        fe_pow_u32(&dpow, &v->delta, c0);
        fe_mul(&cur, &beta, &x, &FR);
        fe_mul(&cur, &cur, &dpow, &FR);
        for (uint32_t c = c0; c < c1; c++) {
          const int qi = find_query(v, v->perm_cols[c].idx, v->perm_cols[c].typ);
          if (qi < 0) {
            panic = 1;
            break;
          }
          const fe* ce = v->perm_cols[c].typ == 255 ? &fix_ev[qi] : (v->perm_cols[c].typ == 254 ? &ins_ev[qi] : &adv_ev[qi]);
          fe a;
          fe_mul(&a, &beta, &sigma_ev[c], &FR);
          fe_add(&a, &a, ce, &FR);
          fe_add(&a, &a, &gamma, &FR);
          fe_mul(&lft, &lft, &a, &FR);
          fe_add(&a, ce, &cur, &FR);
          fe_add(&a, &a, &gamma, &FR);
          fe_mul(&rgt, &rgt, &a, &FR);
          fe_mul(&cur, &cur, &v->delta, &FR);
        }
        fe_sub(&t, &lft, &rgt, &FR);
        fe_mul(&e, &t, &active, &FR);
        FOLD(e);
      }
    }
    for (uint32_t li = 0; li < v->n_lookups + v->n_shuffles && !panic; li++) { /* lookup.rs:159-230, shuffle.rs:148-203 */
      const int is_lk = li < v->n_lookups;
      const arg_t* a = is_lk ? &v->lookups[li] : &v->shuffles[li - v->n_lookups];
      fe cin, ctab, e, pe, pne;
      memset(&cin, 0, sizeof(cin));
      memset(&ctab, 0, sizeof(ctab));
      for (uint32_t j = 0; j < a->n && !panic; j++) {
        fe ev;
        if (!eval_poly(v, a->in_polys[j], vars_pi, nvars, &ev)) panic = 1;
        fe_mul(&cin, &cin, &theta, &FR);
        fe_add(&cin, &cin, &ev, &FR);
        if (!eval_poly(v, a->tab_polys[j], vars_pi, nvars, &ev)) panic = 1;
        fe_mul(&ctab, &ctab, &theta, &FR);
        fe_add(&ctab, &ctab, &ev, &FR);
      }
      if (panic) break;
      if (is_lk) {
        pe = lk_ev[li][0], pne = lk_ev[li][1];
        const fe pie = lk_ev[li][2], piie = lk_ev[li][3], pte = lk_ev[li][4];
        fe lft, rgt, a1, a2;
        fe_add(&a1, &pie, &beta, &FR);
        fe_add(&a2, &pte, &gamma, &FR);
        fe_mul(&lft, &pne, &a1, &FR);
        fe_mul(&lft, &lft, &a2, &FR);
        fe_add(&a1, &cin, &beta, &FR);
        fe_add(&a2, &ctab, &gamma, &FR);
        fe_mul(&rgt, &pe, &a1, &FR);
        fe_mul(&rgt, &rgt, &a2, &FR);
        fe_sub(&t, &one, &pe, &FR);
        fe_mul(&e, &l_0, &t, &FR);
        FOLD(e);
        fe_sqr(&t, &pe, &FR);
        fe_sub(&t, &t, &pe, &FR);
        fe_mul(&e, &l_last, &t, &FR);
        FOLD(e);
        fe_sub(&t, &lft, &rgt, &FR);
        fe_mul(&e, &t, &active, &FR);
        FOLD(e);
        fe_sub(&t, &pie, &pte, &FR);
        fe_mul(&e, &l_0, &t, &FR);
        FOLD(e);
        fe_sub(&a1, &pie, &piie, &FR);
        fe_mul(&e, &t, &a1, &FR);
        fe_mul(&e, &e, &active, &FR);
        FOLD(e);
      } else {
        const uint32_t si = li - v->n_lookups;
        pe = sh_ev[si][0], pne = sh_ev[si][1];
        fe lft, rgt, a1;
        fe_add(&a1, &ctab, &gamma, &FR);
        fe_mul(&lft, &pne, &a1, &FR);
        fe_add(&a1, &cin, &gamma, &FR);
        fe_mul(&rgt, &pe, &a1, &FR);
        fe_sub(&t, &one, &pe, &FR);
        fe_mul(&e, &l_0, &t, &FR);
        FOLD(e);
        fe_sqr(&t, &pe, &FR);
        fe_sub(&t, &t, &pe, &FR);
        fe_mul(&e, &l_last, &t, &FR);
        FOLD(e);
        fe_sub(&t, &lft, &rgt, &FR);
        fe_mul(&e, &t, &active, &FR);
        FOLD(e);
      }
    }
    } /* pi */
    fe xn_m1;
    fe_sub(&xn_m1, &xn, &one, &FR);
    if (panic || fe_is_zero(&xn_m1)) { /* vanishing.rs:100 unwrap */
      res->status = ST_PANIC;
      goto out;
    }
    fe_inv(&t, &xn_m1, &FR);
    fe_mul(&h, &h, &t, &FR);
    for (uint32_t i = n_h; i-- > 0;) { /* vanishing.rs:102-112 */
      msm_scale(&hmsm, &xn);
      msm_push(&hmsm, &one, &P[h_slot[i]]);
    }
    /* queries (lib.rs:349-414) */
    const size_t qcap = (size_t)m * (v->n_aq + 3 * n_sets + 5 * v->n_lookups + 2 * v->n_shuffles) + v->n_fixed + v->n_perm + 4;
    q = (query_t*)calloc(qcap, sizeof(query_t));
    size_t nq = 0;
#define ADDQ(id, ptr, rot, ev)                      \
  do {                                              \
    q[nq].ident = (id);                             \
    q[nq].pt = (ptr);                               \
    q[nq].point = rotate_omega(v, &x, (rot));       \
    q[nq].eval = (ev);                              \
    nq++;                                           \
  } while (0)
    for (uint32_t pi = 0; pi < m; pi++) {
    const fe* adv_ev = ADV_EV(pi);
    const int *adv_slot_pi = adv_slot + (size_t)pi * v->n_advice, *pm_slot_pi = pm_slot + (size_t)pi * n_sets;
    const int *lk_in_pi = lk_in + (size_t)pi * v->n_lookups, *lk_tab_pi = lk_tab + (size_t)pi * v->n_lookups;
    const int *lk_prod_pi = lk_prod + (size_t)pi * v->n_lookups, *sh_prod_pi = sh_prod + (size_t)pi * v->n_shuffles;
    fe(*pm_ev)[3] = pm_ev_all + (size_t)pi * n_sets;
    fe(*lk_ev)[5] = lk_ev_all + (size_t)pi * v->n_lookups;
    fe(*sh_ev)[2] = sh_ev_all + (size_t)pi * v->n_shuffles;
    for (uint32_t i = 0; i < v->n_aq; i++) {
      if (v->aq[i].col >= v->n_advice) {
        res->status = ST_PANIC;
        goto out;
      }
      ADDQ(adv_slot_pi[v->aq[i].col], &P[adv_slot_pi[v->aq[i].col]], v->aq[i].rot, adv_ev[i]);
    }
    for (uint32_t s = 0; s < n_sets; s++) { /* permutation.rs:290-325 */
      ADDQ(pm_slot_pi[s], &P[pm_slot_pi[s]], 0, pm_ev[s][0]);
      ADDQ(pm_slot_pi[s], &P[pm_slot_pi[s]], 1, pm_ev[s][1]);
    }
    for (uint32_t s = n_sets > 0 ? n_sets - 1 : 0; s-- > 0;) ADDQ(pm_slot_pi[s], &P[pm_slot_pi[s]], -(int32_t)(bf + 1), pm_ev[s][2]);
    for (uint32_t i = 0; i < v->n_lookups; i++) { /* lookup.rs:232-271 */
      ADDQ(lk_prod_pi[i], &P[lk_prod_pi[i]], 0, lk_ev[i][0]);
      ADDQ(lk_in_pi[i], &P[lk_in_pi[i]], 0, lk_ev[i][2]);
      ADDQ(lk_tab_pi[i], &P[lk_tab_pi[i]], 0, lk_ev[i][4]);
      ADDQ(lk_in_pi[i], &P[lk_in_pi[i]], -1, lk_ev[i][3]);
      ADDQ(lk_prod_pi[i], &P[lk_prod_pi[i]], 1, lk_ev[i][1]);
    }
    for (uint32_t i = 0; i < v->n_shuffles; i++) { /* shuffle.rs:205-225 */
      ADDQ(sh_prod_pi[i], &P[sh_prod_pi[i]], 0, sh_ev[i][0]);
      ADDQ(sh_prod_pi[i], &P[sh_prod_pi[i]], 1, sh_ev[i][1]);
    }
    } /* pi */
    for (uint32_t i = 0; i < v->n_fixed; i++) {
      if (v->fq[i].col >= v->n_fixed_commit) {
        res->status = ST_PANIC;
        goto out;
      }
      ADDQ(ID_FIXED + (int)v->fq[i].col, &v->fixed_commit[v->fq[i].col], v->fq[i].rot, fix_ev[i]);
    }
    for (uint32_t i = 0; i < v->n_perm; i++) ADDQ(ID_SIGMA + (int)i, &v->perm_commit[i], 0, sigma_ev[i]);
    ADDQ(ID_HMSM, NULL, 0, h); /* vanishing.rs:124-136 */
    ADDQ(random_slot, &P[random_slot], 0, random_eval);
    int rc = multiopen == 0 ? shplonk_verify(v, &tr, q, nq, &hmsm, &left, &right) : gwc_verify(v, &tr, q, nq, &hmsm, &left, &right);
    if (rc != ST_OK) {
      res->status = rc;
      goto out;
    }
    msm_eval(&res->L, &left); /* DualMSM::check (msm.rs:185-203), SingleStrategy (strategy.rs:164-176) */
    msm_eval(&res->R, &right);
    res->have_lr = 1;
    if (check_pairing && !pairing_check2(&res->L, &v->prep_s, &res->R, &v->prep_n)) res->status = ST_CSF;
  }
out:
  res->n_chal = tr.n_chal < 64 ? tr.n_chal : 64;
  memcpy(res->chal, tr.chal, sizeof(fe) * res->n_chal);
  free(tr.chal), free(ivals), free(P), free(adv_slot), free(user_chal), free(lk_in), free(lk_tab), free(lk_prod), free(sh_prod), free(pm_slot), free(h_slot);
  free(vars), free(sigma_ev), free(pm_ev_all), free(lk_ev_all), free(sh_ev_all), free(hmsm.t), free(left.t), free(right.t), free(q);
}

/* ============================================================================ C API (ctypes) */
int h2vo_verify_multi(const h2vo_vk* v, const uint8_t* proof, size_t plen, const uint8_t* inst, const uint32_t* col_len, uint32_t ncols, uint32_t m,
                      int multiopen, int hash, int check_pairing, uint8_t* challenges, uint32_t* n_challenges, uint8_t* LR);
/* one proof; col_len[ncols] scalars per instance column.  challenges: up to 64 x 32 B canonical; LR: 128 B affine L | R */
int h2vo_verify(const h2vo_vk* v, const uint8_t* proof, size_t plen, const uint8_t* inst, const uint32_t* col_len, uint32_t ncols, int multiopen,
                int hash, int check_pairing, uint8_t* challenges, uint32_t* n_challenges, uint8_t* LR) {
  return h2vo_verify_multi(v, proof, plen, inst, col_len, ncols, 1, multiopen, hash, check_pairing, challenges, n_challenges, LR);
}
/* the same for a proof that carries m circuit instances: col_len[ncols], ncols = m x instance columns, instance-major */
int h2vo_verify_multi(const h2vo_vk* v, const uint8_t* proof, size_t plen, const uint8_t* inst, const uint32_t* col_len, uint32_t ncols, uint32_t m,
                      int multiopen, int hash, int check_pairing, uint8_t* challenges, uint32_t* n_challenges, uint8_t* LR) {
  result_t r;
  verify_one(v, proof, plen, inst, col_len, ncols, m, multiopen, hash, check_pairing, &r);
  if (challenges)
    for (uint32_t i = 0; i < r.n_chal; i++) fe_to_repr(challenges + 32 * i, &r.chal[i], &FR);
  if (n_challenges) *n_challenges = r.n_chal;
  if (LR) {
    g1_to_bytes64(LR, &r.L);
    g1_to_bytes64(LR + 64, &r.R);
  }
  return r.status;
}

typedef struct {
  const h2vo_vk* v;
  uint32_t n, tid, nthreads, ncols, n_chal_cap;
  const uint8_t *proofs, *inst;
  const uint64_t *poff, *ioff;
  int multiopen, hash, check_pairing;
  uint8_t *status, *LR, *chal;
} job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  uint32_t* cl = (uint32_t*)malloc(4 * (j->ncols + 1));
  for (uint32_t i = j->tid; i < j->n; i += j->nthreads) {
    const uint64_t tot = j->ioff[i + 1] - j->ioff[i];
    result_t r;
    if (j->ncols == 0 ? tot != 0 : tot % j->ncols != 0) {
      memset(&r, 0, sizeof(r));
      r.status = ST_INVALID_INSTANCES;
      r.L.inf = r.R.inf = 1;
    } else {
      for (uint32_t c = 0; c < j->ncols; c++) cl[c] = (uint32_t)(tot / j->ncols);
      verify_one(j->v, j->proofs + j->poff[i], (size_t)(j->poff[i + 1] - j->poff[i]), j->inst + 32 * j->ioff[i], cl, j->ncols, 1, j->multiopen, j->hash,
                 j->check_pairing, &r);
    }
    j->status[i] = (uint8_t)r.status;
    if (j->LR) {
      g1_to_bytes64(j->LR + 128 * (size_t)i, &r.L);
      g1_to_bytes64(j->LR + 128 * (size_t)i + 64, &r.R);
    }
    if (j->chal) {
      memset(j->chal + 32 * (size_t)i * j->n_chal_cap, 0, 32 * (size_t)j->n_chal_cap);
      for (uint32_t c = 0; c < r.n_chal && c < j->n_chal_cap; c++) fe_to_repr(j->chal + 32 * ((size_t)i * j->n_chal_cap + c), &r.chal[c], &FR);
    }
  }
  free(cl);
  return NULL;
}
/* n proofs, `threads` host threads over proofs, verify_proof with SingleStrategy each (one pairing per proof when
 * check_pairing).  Instances: ioff in scalars, equal split over the VK's instance columns.  Optional outputs:
 * LR n x 128 B, chal n x chal_cap x 32 B.  Returns wall seconds of the verification in *seconds. */
int h2vo_verify_many(const h2vo_vk* v, uint32_t n, const uint8_t* proofs, const uint64_t* poff, const uint8_t* inst, const uint64_t* ioff, int multiopen,
                     int hash, int check_pairing, int threads, uint8_t* status, uint8_t* LR, uint8_t* chal, uint32_t chal_cap, double* seconds) {
  if (threads < 1) threads = 1;
  if ((uint32_t)threads > n) threads = (int)(n ? n : 1);
  job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job_t){v, n, (uint32_t)t, (uint32_t)threads, v->n_instance, chal_cap, proofs, inst, poff, ioff, multiopen, hash, check_pairing, status, LR, chal};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  free(jobs), free(th);
  return 0;
}
/* AccumulatorStrategy (strategy.rs:125-140): fold the per-proof accumulators with c_j = prod_{i>j} r_i over the
 * proofs with include[j] != 0 (those whose accumulators exist), then DualMSM::check.  out_LR: 128 B. */
int h2vo_fold(const h2vo_vk* v, uint32_t n, const uint8_t* LR, const uint8_t* rs, const uint8_t* include, uint8_t* out_LR, int* verdict) {
  g1j accL, accR;
  g1j_identity(&accL);
  g1j_identity(&accR);
  fe c = FR.one;
  for (uint32_t j = n; j-- > 0;) {
    if (include[j]) {
      uint64_t k[4];
      fe_to_canon(k, &c, &FR);
      for (int side = 0; side < 2; side++) {
        const uint8_t* b = LR + 128 * (size_t)j + 64 * side;
        int allz = 1;
        for (int i = 0; i < 64; i++) allz &= b[i] == 0;
        if (allz) continue;
        g1a p;
        if (!fe_from_repr(&p.x, b, &FQ) || !fe_from_repr(&p.y, b + 32, &FQ)) return -1;
        p.inf = 0;
        g1j t;
        g1_mul(&t, &p, k);
        g1j_add(side ? &accR : &accL, side ? &accR : &accL, &t);
      }
    }
    fe r;
    uint64_t rc[4];
    le_load(rc, rs + 32 * (size_t)j);
    fe_from_canon(&r, rc, &FR); /* r_j may exceed the modulus: reduced like the device path */
    fe_mul(&c, &c, &r, &FR);
  }
  g1a L, R;
  g1j_to_affine(&L, &accL);
  g1j_to_affine(&R, &accR);
  if (out_LR) {
    g1_to_bytes64(out_LR, &L);
    g1_to_bytes64(out_LR + 64, &R);
  }
  if (verdict) *verdict = pairing_check2(&L, &v->prep_s, &R, &v->prep_n);
  return 0;
}
