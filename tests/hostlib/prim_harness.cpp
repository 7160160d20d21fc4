// Host build of the HD device headers, for CPU-side unit tests (tests only; never shipped).
#include "tower.cuh"
#include "hash.cuh"
#include <string.h>
using namespace h2v;
extern "C" {
// op: 0 mul, 1 add, 2 sub, 3 inv, 4 to_canonical, 5 from_canonical, 6 sqrt-candidate ; field: 0 Fq, 1 Fr
void t_field(int field, int op, const u32* a, const u32* b, u32* out) {
  if (field == 0) {
    Fq x, y, r; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.inv(); break;
      case 4: r = x.to_canonical(); break; case 5: r = Fq::from_canonical(x); break; default: r = fq_sqrt_candidate(x); }
    memcpy(out, r.l, 32);
  } else {
    Fr x, y, r; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.inv(); break;
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
void t_blake2b(const u8* data, u32 len, u8* out) { Blake2b b; b.init_halo2(); b.update(data, len); b.digest(out); }
void t_keccak(const u8* data, u32 len, u8 suffix, u8* out) { Keccak256 k; k.init_halo2(); k.update(data, len); k.digest_with_suffix(suffix, out); }
}
