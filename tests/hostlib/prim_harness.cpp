// Host build of the HD device headers, for CPU-side unit tests (tests only; never shipped).
#include "tower.cuh"
#include "pairing_cta.cuh"
#include "hash.cuh"
#include "glv.cuh"
#include <string.h>
using namespace h2v;
extern "C" {
// op: 0 mul, 1 add, 2 sub, 3 inv, 4 to_canonical, 5 from_canonical, 6 sqrt-candidate, 7 inv by the binary Euclid ; field: 0 Fq, 1 Fr
void t_field(int field, int op, const u32* a, const u32* b, u32* out) {
  if (field == 0) {
    Fq x, y, r; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.inv(); break; case 7: r = x.inv_bin(); break;
      case 4: r = x.to_canonical(); break; case 5: r = Fq::from_canonical(x); break; default: r = fq_sqrt_candidate(x); }
    memcpy(out, r.l, 32);
  } else {
    Fr x, y, r; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.inv(); break; case 7: r = x.inv_bin(); break;
      case 4: r = x.to_canonical(); break; default: r = Fr::from_canonical(x); }
    memcpy(out, r.l, 32);
  }
}
void t_from_uniform(const u8* b64, u32* out) { Fr r = Fr::from_uniform(b64).to_canonical(); memcpy(out, r.l, 32); }
int t_decompress(const u8* b, u32* xy) {
  G1Affine p; if (!g1_decompress(b, p)) return 0;
  Fq x = p.x.to_canonical(), y = p.y.to_canonical(); memcpy(xy, x.l, 32); memcpy(xy + 8, y.l, 32); return 1;
}
// k canonical scalar; p affine canonical; out affine canonical (zeros for identity)
void t_g1_mul(const u32* pxy, const u32* k, u32* out) {
  G1Affine p; Fq t; memcpy(t.l, pxy, 32); p.x = Fq::from_canonical(t); memcpy(t.l, pxy + 8, 32); p.y = Fq::from_canonical(t);
  G1Jac r = g1_mul_canonical(p, k); G1Affine a; g1_to_affine(r, a);
  Fq x = a.x.to_canonical(), y = a.y.to_canonical(); memcpy(out, x.l, 32); memcpy(out + 8, y.l, 32);
}
// GLV (glv.cuh): the two halves of a canonical scalar (5 limbs + sign each), and [k]P as the sum of the two half-length parts
void t_glv_decompose(const u32* k, u32* out12) {
  GlvHalf a, b; glv_decompose(k, a, b);
  memcpy(out12, a.l, 20); out12[5] = a.neg; memcpy(out12 + 6, b.l, 20); out12[11] = b.neg;
}
void t_g1_mul_glv(const u32* pxy, const u32* k, u32* out) {
  G1Affine p; Fq t; memcpy(t.l, pxy, 32); p.x = Fq::from_canonical(t); memcpy(t.l, pxy + 8, 32); p.y = Fq::from_canonical(t);
  GlvHalf a, b; glv_decompose(k, a, b);
  G1Jac r = g1_add(g1_add(g1_mul_glv_part(p, a, false, 64, 65, 64), g1_mul_glv_part(p, a, false, 0, 64, 0)),
                   g1_add(g1_mul_glv_part(p, b, true, 64, 65, 64), g1_mul_glv_part(p, b, true, 0, 64, 0)));
  G1Affine q; const bool ok = g1_to_affine(r, q);
  Fq x = q.x.to_canonical(), y = q.y.to_canonical(); memcpy(out, x.l, 32); memcpy(out + 8, y.l, 32); (void)ok;
}
void t_g1_add(const u32* pxy, const u32* qxy, int neg, u32* out) {
  G1Affine p, q; Fq t; memcpy(t.l, pxy, 32); p.x = Fq::from_canonical(t); memcpy(t.l, pxy + 8, 32); p.y = Fq::from_canonical(t);
  memcpy(t.l, qxy, 32); q.x = Fq::from_canonical(t); memcpy(t.l, qxy + 8, 32); q.y = Fq::from_canonical(t);
  G1Jac r = g1_add_mixed(g1_double(g1_double(G1Jac::from_affine(p))), q, neg != 0);  // 4p +- q
  r = g1_add(r, g1_neg(g1_double(G1Jac::from_affine(p))));                             // 2p +- q
  G1Affine a; g1_to_affine(r, a);
  Fq x = a.x.to_canonical(), y = a.y.to_canonical(); memcpy(out, x.l, 32); memcpy(out + 8, y.l, 32);
}
static G2Affine load_g2(const u32* c) {
  G2Affine q; Fq t[4]; for (int i = 0; i < 4; i++) { Fq r; memcpy(r.l, c + 8 * i, 32); t[i] = Fq::from_canonical(r); }
  q.x = {t[0], t[1]}; q.y = {t[2], t[3]}; return q;
}
// pairing product check e(L,Q0) e(R,Q1) == 1; also returns final GT value (12 canonical Fq, a[i].c0,a[i].c1 order)
int t_pairing_check(const u32* Lxy, const u32* Rxy, int l_inf, int r_inf, const u32* q0, const u32* q1, u32* gt) {
  G2Affine Q0 = load_g2(q0), Q1 = load_g2(q1);
  if (!g2_on_curve(Q0) || !g2_on_curve(Q1)) return -1;
  static G2Line l0[H2V_ATE_LINES], l1[H2V_ATE_LINES];
  g2_prepare(Q0, l0); g2_prepare(Q1, l1);
  G1Affine p[2]; Fq t;
  memcpy(t.l, Lxy, 32); p[0].x = Fq::from_canonical(t); memcpy(t.l, Lxy + 8, 32); p[0].y = Fq::from_canonical(t);
  memcpy(t.l, Rxy, 32); p[1].x = Fq::from_canonical(t); memcpy(t.l, Rxy + 8, 32); p[1].y = Fq::from_canonical(t);
  bool skip[2] = {l_inf != 0, r_inf != 0};
  const G2Line* lines[2] = {l0, l1};
  Fq12 f = final_exponentiation(miller_loop2(p, skip, lines));
  for (int i = 0; i < 6; i++) { Fq a = f.a[i].c0.to_canonical(), b = f.a[i].c1.to_canonical(); memcpy(gt + 16 * i, a.l, 32); memcpy(gt + 16 * i + 8, b.l, 32); }
  return f.is_one() ? 1 : 0;
}
// ---- host emulation of the cooperative Fq12 engine (pairing_cta.cuh): lanes run as loops
static const LinTables h_lin = H2V_LIN_TABLES_INIT;
struct HostEngine {
  Fq scr[2 * E12_N];
  void mul(E12* dst, const E12* a, const E12* b) {
    for (int l = 0; l < E12_N; l++) e12_mul_p1(scr, a, b, l);
    for (int l = 0; l < E12_N; l++) e12_mul_p2(dst, scr, &h_lin, l);
  }
  void expand(E12* dst) { for (int l = 0; l < E12_N; l++) e12_expand_p(dst, scr, &h_lin, l); }
  void conj(E12* dst, const E12* a) { for (int b = 0; b < E12_NB; b++) scr[b] = e12_conj_base(a, b); expand(dst); }
  void frob(E12* dst, const E12* a) { for (int b = 0; b < E12_NB; b++) scr[b] = e12_frob_base(a, b); expand(dst); }
  void frob2(E12* dst, const E12* a) { for (int b = 0; b < E12_NB; b++) scr[b] = e12_frob2_base(a, b); expand(dst); }
  void from_fq12(E12* dst, const Fq12& x) { for (int b = 0; b < E12_NB; b++) scr[b] = fq12_base_coord(x, b); expand(dst); }
  void pow_u(E12* dst, const E12* x) {
    *dst = *x;
    for (int i = 61; i >= 0; i--) { mul(dst, dst, dst); if ((H2V_BN_U >> i) & 1) mul(dst, dst, x); }
  }
  Fq12 to_fq12(const E12* x) {
    Fq12 r;
    for (int b = 0; b < E12_NB; b++) { const int hh = b / 6, j = (b % 6) / 2; Fq2& c = r.a[2 * j + hh]; ((b & 1) ? c.c1 : c.c0) = x->e[e12_base_slot(b)]; }
    return r;
  }
};
static Fq12 load_fq12(const u32* c) {  // 12 canonical Fq: a[i].c0, a[i].c1
  Fq12 r; for (int i = 0; i < 6; i++) { Fq t; memcpy(t.l, c + 16 * i, 32); r.a[i].c0 = Fq::from_canonical(t); memcpy(t.l, c + 16 * i + 8, 32); r.a[i].c1 = Fq::from_canonical(t); }
  return r;
}
static void store_fq12(const Fq12& f, u32* out) {
  for (int i = 0; i < 6; i++) { Fq a = f.a[i].c0.to_canonical(), b = f.a[i].c1.to_canonical(); memcpy(out + 16 * i, a.l, 32); memcpy(out + 16 * i + 8, b.l, 32); }
}
// the 128-thread split of the linear map (gf_mul of pairing_cta.cuh), threads emulated: per-thread partial column
// accumulators, the row's lead thread adds its followers' and reduces
static const u32 h_lin_fast[H2V_LIN_FAST_THREADS * (FAST_W + 1)] = H2V_LIN_FAST_INIT;
static void fast_mul(HostEngine& he, E12* dst, const E12* a, const E12* b) {
  for (int l = 0; l < E12_N; l++) e12_mul_p1(he.scr, a, b, l);
  static u64 part[H2V_LIN_FAST_THREADS][8];
  FastTerms ft[H2V_LIN_FAST_THREADS];
  for (int t = 0; t < H2V_LIN_FAST_THREADS; t++) {
    for (int k = 0; k < FAST_W; k++) ft[t].w[k] = h_lin_fast[t * (FAST_W + 1) + k];
    ft[t].meta = h_lin_fast[t * (FAST_W + 1) + FAST_W];
    fast_partial(part[t], ft[t], he.scr);
  }
  int rows_done = 0;
  for (int t = 0; t < H2V_LIN_FAST_THREADS; t++) {
    if (!ft[t].lead()) continue;
    u64 sum[8];
    for (int i = 0; i < 8; i++) sum[i] = part[t][i];
    // followers sit in the same warp, right behind the lead (the device combines them with __shfl_down_sync)
    if (ft[t].followers() > FAST_MAX_FOLLOWERS || (t % 32) + ft[t].followers() > 31) { memset(dst, 0xff, sizeof(E12)); return; }
    for (int f = 1; f <= ft[t].followers(); f++) {
      if (ft[t + f].lead() || ft[t + f].row() != ft[t].row()) { memset(dst, 0xff, sizeof(E12)); return; }
      for (int i = 0; i < 8; i++) sum[i] += part[t + f][i];
    }
    dst->e[ft[t].row()] = lin_reduce(sum);
    rows_done++;
  }
  if (rows_done != E12_N) memset(dst, 0xff, sizeof(E12));
}
// op: 0 mul, 1 conj, 2 frob, 3 frob2, 4 pow_u, 5 mul through the 128-thread split map
void t_e12_op(int op, const u32* a, const u32* b, u32* out) {
  HostEngine he; E12 x, y, z;
  he.from_fq12(&x, load_fq12(a)); he.from_fq12(&y, load_fq12(b));
  switch (op) { case 0: he.mul(&z, &x, &y); break; case 5: fast_mul(he, &z, &x, &y); break; case 1: he.conj(&z, &x); break; case 2: he.frob(&z, &x); break; case 3: he.frob2(&z, &x); break; default: he.pow_u(&z, &x); }
  // the expanded form must be self-consistent: re-expanding the base coordinates reproduces it
  E12 chk; he.from_fq12(&chk, he.to_fq12(&z));
  for (int l = 0; l < E12_N; l++) if (chk.e[l] != z.e[l]) { memset(out, 0xff, 12 * 32); return; }
  store_fq12(he.to_fq12(&z), out);
}
// decomposed pairing-product check: prod_k e(S_k, Q_k) == 1 with S_k Jacobian (X, Y, Z canonical, 24 words
// each; Z = 0: identity) and Q_k affine G2 (32 words each).  Mirrors k_lines + k_pairing_check.
int t_pairing_windows(int n, const u32* jac, const u32* g2s) {
  G2Affine* q = new G2Affine[n];
  G2Line* lines = new G2Line[(size_t)n * H2V_ATE_LINES];
  for (int k = 0; k < n; k++) { q[k] = load_g2(g2s + 32 * k); if (!g2_on_curve(q[k])) return -1; }
  g2_prepare_many(q, n, lines);
  HostEngine he;
  static E12 M[H2V_ATE_ITERS];
  for (int it = 0; it < H2V_ATE_ITERS; it++) {
    const int n0 = it >= 64 ? H2V_ATE_LINES - 2 : ate_line_index(it), ns = ate_lines_in_iteration(it);
    bool first = true; E12 acc, tmp;
    for (int s = 0; s < ns; s++) for (int k = 0; k < n; k++) {
      Fq X, Y, Z, t; memcpy(t.l, jac + 24 * k, 32); X = Fq::from_canonical(t); memcpy(t.l, jac + 24 * k + 8, 32); Y = Fq::from_canonical(t);
      memcpy(t.l, jac + 24 * k + 16, 32); Z = Fq::from_canonical(t);
      const Fq XZ = X * Z, Z3 = Z * Z * Z;
      for (int b = 0; b < E12_NB; b++) he.scr[b] = e12_line_base(lines[(size_t)k * H2V_ATE_LINES + n0 + s], Y, XZ, Z3, Z.is_zero(), b);
      he.expand(first ? &acc : &tmp);
      if (!first) he.mul(&acc, &acc, &tmp);
      first = false;
    }
    M[it] = acc;
  }
  E12 f = M[0], tt, fu, fu2, fu3, a, b, y0, T0, T1, N;
  for (int it = 1; it < H2V_ATE_ITERS; it++) { if (it < 64) he.mul(&f, &f, &f); he.mul(&f, &f, &M[it]); }
  he.frob2(&a, &f); he.mul(&tt, &a, &f);
  he.pow_u(&fu, &tt); he.pow_u(&fu2, &fu); he.pow_u(&fu3, &fu2);
  he.frob(&a, &tt); he.frob2(&b, &tt); he.mul(&y0, &a, &b); he.frob(&a, &b); he.mul(&y0, &y0, &a);
  he.frob2(&a, &fu2); he.mul(&a, &a, &a); he.mul(&b, &a, &a); he.mul(&b, &b, &a); he.mul(&N, &y0, &b);
  he.frob(&a, &fu3); he.mul(&a, &fu3, &a); he.mul(&T0, &a, &a);
  he.frob(&a, &fu2); he.mul(&a, &fu, &a); he.mul(&T0, &T0, &a); he.mul(&T0, &T0, &fu2);
  he.frob(&a, &fu); he.mul(&T1, &a, &fu2); he.mul(&T1, &T1, &T0);
  he.mul(&T1, &T1, &T1); he.mul(&T1, &T1, &T0); he.mul(&T1, &T1, &T1);
  he.mul(&T0, &T1, &tt); he.mul(&T0, &T0, &T0); he.mul(&T0, &T0, &T1);
  he.conj(&a, &N); he.mul(&a, &a, &T0); he.conj(&b, &T0); he.mul(&b, &N, &b);
  bool ok = true;
  for (int c = 0; c < E12_NB; c++) ok = ok && a.e[e12_base_slot(c)] == b.e[e12_base_slot(c)];
  delete[] q; delete[] lines;
  return ok ? 1 : 0;
}
void t_blake2b(const u8* data, u32 len, u8* out) { Blake2b b; b.init_halo2(); b.update(data, len); b.digest(out); }
void t_keccak(const u8* data, u32 len, u8 suffix, u8* out) { Keccak256 k; k.init_halo2(); k.update(data, len); k.digest_with_suffix(suffix, out); }
}
