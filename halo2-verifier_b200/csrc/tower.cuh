// BN254 extension tower and optimal-ate pairing check.
//
// Replaces the halo2curves calls of DualMSM::check (reference poly/kzg/msm.rs:185-203):
// `G2Prepared::from`, `multi_miller_loop`, `final_exponentiation`, `is_identity`.
// Tower: Fq2 = Fq[u]/(u^2+1), Fq12 = Fq2[w]/(w^6 - xi), xi = 9 + u; an Fq12 element is stored as its
// six Fq2 coefficients a[i] of w^i (the Fq6 halves of the usual 2-3-2 tower are the even / odd
// coefficients).  Only `is_identity` of the final value is observable by the reference.
//
// G2 arguments of the check are verifier parameters ([s]G2 and -G2), so their Miller-loop lines
// are prepared once per context (g2_prepare, host side at h2v_ctx_create) exactly like
// `G2Prepared`; the device Miller loop only evaluates prepared lines at the two G1 accumulators.
#pragma once
#include "curve.cuh"
#include <stddef.h>

#include "tower_consts.inc"

namespace h2v {

// ------------------------------------------------------------------------------------------ Fq2
struct Fq2 {
  Fq c0, c1;
  static H2V_HD Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static H2V_HD Fq2 one() { return {Fq::one(), Fq::zero()}; }
  H2V_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  H2V_HD bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  friend H2V_HD Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  friend H2V_HD Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  H2V_HD Fq2 neg() const { return {c0.neg(), c1.neg()}; }
  H2V_HD Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  H2V_HD Fq2 conj() const { return {c0, c1.neg()}; }
  friend H2V_HDN Fq2 operator*(const Fq2& a, const Fq2& b) {  // Karatsuba, 3 MM
    Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
    Fq t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
    return {t0 - t1, t2 - t0 - t1};
  }
  H2V_HDN Fq2 sqr() const {  // 2 MM
    Fq t = c0 * c1;
    return {(c0 + c1) * (c0 - c1), t.dbl()};
  }
  H2V_HD Fq2 mul_fq(const Fq& s) const { return {c0 * s, c1 * s}; }
  H2V_HD Fq2 mul_xi() const {  // (9 + u)(c0 + c1 u) = (9 c0 - c1) + (9 c1 + c0) u
    Fq t0 = c0.dbl().dbl().dbl() + c0;
    Fq t1 = c1.dbl().dbl().dbl() + c1;
    return {t0 - c1, t1 + c0};
  }
  H2V_HDN Fq2 inv() const {
    Fq d = (c0.sqr() + c1.sqr()).inv();
    return {c0 * d, (c1 * d).neg()};
  }
};

H2V_HD Fq fq_from_limbs(const u32* v) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = v[i];
  return r;
}
H2V_HDN inline Fq2 tower_gamma1(int i) {
  constexpr u32 v[6][2][8] = H2V_GAMMA1_INIT;
  return {fq_from_limbs(v[i][0]), fq_from_limbs(v[i][1])};
}
H2V_HDN inline Fq tower_gamma2(int i) {
  constexpr u32 v[6][8] = H2V_GAMMA2_INIT;
  return fq_from_limbs(v[i]);
}
H2V_HDN inline Fq2 twist_b() {
  constexpr u32 v[2][8] = H2V_TWIST_B_INIT;
  return {fq_from_limbs(v[0]), fq_from_limbs(v[1])};
}

// ------------------------------------------------------------------------------------------ Fq6 (helper on Fq2 triples)
struct Fq6 {
  Fq2 b0, b1, b2;  // b0 + b1 v + b2 v^2, v^3 = xi
  friend H2V_HD Fq6 operator+(const Fq6& a, const Fq6& b) { return {a.b0 + b.b0, a.b1 + b.b1, a.b2 + b.b2}; }
  friend H2V_HD Fq6 operator-(const Fq6& a, const Fq6& b) { return {a.b0 - b.b0, a.b1 - b.b1, a.b2 - b.b2}; }
  H2V_HD Fq6 neg() const { return {b0.neg(), b1.neg(), b2.neg()}; }
  H2V_HD Fq6 mul_v() const { return {b2.mul_xi(), b0, b1}; }
};

H2V_HDN inline Fq6 fq6_mul(const Fq6& a, const Fq6& b) {  // 6 Fq2 mul
  Fq2 t0 = a.b0 * b.b0, t1 = a.b1 * b.b1, t2 = a.b2 * b.b2;
  Fq6 r;
  r.b0 = t0 + ((a.b1 + a.b2) * (b.b1 + b.b2) - t1 - t2).mul_xi();
  r.b1 = (a.b0 + a.b1) * (b.b0 + b.b1) - t0 - t1 + t2.mul_xi();
  r.b2 = (a.b0 + a.b2) * (b.b0 + b.b2) - t0 - t2 + t1;
  return r;
}
// a * (c0 + c1 v): 5 Fq2 mul
H2V_HDN inline Fq6 fq6_mul_by_01(const Fq6& a, const Fq2& c0, const Fq2& c1) {
  Fq2 t0 = a.b0 * c0, t1 = a.b1 * c1;
  Fq6 r;
  r.b0 = t0 + ((a.b1 + a.b2) * c1 - t1).mul_xi();
  r.b1 = (a.b0 + a.b1) * (c0 + c1) - t0 - t1;
  r.b2 = (a.b0 + a.b2) * c0 - t0 + t1;
  return r;
}
H2V_HDN inline Fq6 fq6_inv(const Fq6& a) {
  Fq2 t0 = a.b0.sqr() - (a.b1 * a.b2).mul_xi();
  Fq2 t1 = a.b2.sqr().mul_xi() - a.b0 * a.b1;
  Fq2 t2 = a.b1.sqr() - a.b0 * a.b2;
  Fq2 d = a.b0 * t0 + (a.b2 * t1 + a.b1 * t2).mul_xi();
  Fq2 di = d.inv();
  return {t0 * di, t1 * di, t2 * di};
}

// ------------------------------------------------------------------------------------------ Fq12
struct Fq12 {
  Fq2 a[6];  // sum a[i] w^i, w^6 = xi
  static H2V_HD Fq12 one() {
    Fq12 r;
    r.a[0] = Fq2::one();
#pragma unroll
    for (int i = 1; i < 6; i++) r.a[i] = Fq2::zero();
    return r;
  }
  H2V_HD Fq6 even() const { return {a[0], a[2], a[4]}; }
  H2V_HD Fq6 odd() const { return {a[1], a[3], a[5]}; }
  static H2V_HD Fq12 from_halves(const Fq6& e, const Fq6& o) {
    Fq12 r;
    r.a[0] = e.b0; r.a[2] = e.b1; r.a[4] = e.b2;
    r.a[1] = o.b0; r.a[3] = o.b1; r.a[5] = o.b2;
    return r;
  }
  H2V_HD bool is_one() const {
    bool ok = a[0] == Fq2::one();
#pragma unroll
    for (int i = 1; i < 6; i++) ok = ok && a[i].is_zero();
    return ok;
  }
  H2V_HD Fq12 conj() const {  // ^(p^6): w -> -w
    Fq12 r = *this;
    r.a[1] = a[1].neg(); r.a[3] = a[3].neg(); r.a[5] = a[5].neg();
    return r;
  }
};

H2V_HDN inline Fq12 fq12_mul(const Fq12& x, const Fq12& y) {  // 18 Fq2 mul
  Fq6 x0 = x.even(), x1 = x.odd(), y0 = y.even(), y1 = y.odd();
  Fq6 t0 = fq6_mul(x0, y0), t1 = fq6_mul(x1, y1);
  Fq6 t2 = fq6_mul(x0 + x1, y0 + y1);
  return Fq12::from_halves(t0 + t1.mul_v(), t2 - t0 - t1);
}
H2V_HDN inline Fq12 fq12_sqr(const Fq12& x) {  // complex squaring, 12 Fq2 mul
  Fq6 x0 = x.even(), x1 = x.odd();
  Fq6 t = fq6_mul(x0, x1);
  Fq6 s = fq6_mul(x0 + x1, x0 + x1.mul_v());
  return Fq12::from_halves(s - t - t.mul_v(), t + t);
}
H2V_HDN inline Fq12 fq12_inv(const Fq12& x) {
  Fq6 x0 = x.even(), x1 = x.odd();
  Fq6 d = fq6_mul(x0, x0) - fq6_mul(x1, x1).mul_v();
  Fq6 di = fq6_inv(d);
  return Fq12::from_halves(fq6_mul(x0, di), fq6_mul(x1, di).neg());
}
H2V_HDN inline Fq12 fq12_frob(const Fq12& x) {  // ^p
  Fq12 r;
  r.a[0] = x.a[0].conj();
  for (int i = 1; i < 6; i++) r.a[i] = x.a[i].conj() * tower_gamma1(i);
  return r;
}
H2V_HDN inline Fq12 fq12_frob2(const Fq12& x) {  // ^(p^2)
  Fq12 r;
  r.a[0] = x.a[0];
  for (int i = 1; i < 6; i++) r.a[i] = x.a[i].mul_fq(tower_gamma2(i));
  return r;
}
// f * (A + B w + C w^3) with A in Fq: 10 Fq2 mul + 6 Fq mul
H2V_HDN inline Fq12 fq12_mul_by_line(const Fq12& f, const Fq& A, const Fq2& B, const Fq2& C) {
  Fq6 f0 = f.even(), f1 = f.odd();
  Fq6 t0 = {f0.b0.mul_fq(A), f0.b1.mul_fq(A), f0.b2.mul_fq(A)};  // f0 * (A,0,0)
  Fq6 t1 = fq6_mul_by_01(f1, B, C);                               // f1 * (B,C,0)
  Fq2 AB = B;
  AB.c0 = AB.c0 + A;
  Fq6 t2 = fq6_mul_by_01(f0 + f1, AB, C);
  return Fq12::from_halves(t0 + t1.mul_v(), t2 - t0 - t1);
}

// ------------------------------------------------------------------------------------------ G2 line preparation
struct G2Affine {
  Fq2 x, y;
};
struct G2Line {
  Fq2 nlam;  // -lambda'           (coefficient of xP * w)
  Fq2 c;     // lambda' xT - yT    (coefficient of w^3)
};

// 6u+2 = 0x1_9d797039_be763ba8 (65 bits); the loop runs over bits 63..0.
static constexpr u64 H2V_ATE_LOOP_LOW = 0x9d797039be763ba8ull;
static constexpr int H2V_ATE_LINES = 64 + 36 + 2;  // doublings + additions (popcount of low 64 bits) + 2 Frobenius
static constexpr u64 H2V_BN_U = 0x44e992b44a6909f1ull;

#if !defined(__CUDA_ARCH__)
// Host only (context creation): affine line schedule for Q.  The order matches miller_loop().
inline void g2_line_step(G2Affine& t, const G2Affine& q, bool is_double, G2Line& out) {
  Fq2 lam;
  if (is_double) {
    Fq2 x2 = t.x.sqr();
    lam = (x2.dbl() + x2) * t.y.dbl().inv();
  } else {
    lam = (q.y - t.y) * (q.x - t.x).inv();
  }
  Fq2 x3 = lam.sqr() - t.x - (is_double ? t.x : q.x);
  Fq2 y3 = lam * (t.x - x3) - t.y;
  out.nlam = lam.neg();
  out.c = lam * t.x - t.y;
  t.x = x3;
  t.y = y3;
}
inline G2Affine g2_frob(const G2Affine& q) {
  // (x', y') -> (conj(x') xi^((p-1)/3), conj(y') xi^((p-1)/2))
  return {q.x.conj() * tower_gamma1(2), q.y.conj() * tower_gamma1(3)};
}
inline void g2_prepare(const G2Affine& q, G2Line* lines /* H2V_ATE_LINES */) {
  G2Affine t = q;
  int n = 0;
  for (int i = 63; i >= 0; i--) {
    g2_line_step(t, t, true, lines[n++]);
    if ((H2V_ATE_LOOP_LOW >> i) & 1) g2_line_step(t, q, false, lines[n++]);
  }
  G2Affine q1 = g2_frob(q);
  G2Affine q2 = g2_frob(q1);
  q2.y = q2.y.neg();
  g2_line_step(t, q1, false, lines[n++]);
  g2_line_step(t, q2, false, lines[n++]);
}
inline bool g2_on_curve(const G2Affine& q) { return q.y.sqr() == q.x.sqr() * q.x + twist_b(); }
inline G2Affine g2_double_affine(const G2Affine& q) {
  G2Affine t = q;
  G2Line unused;
  g2_line_step(t, t, true, unused);
  return t;
}
// Line tables of `count` points at once (lines[k * H2V_ATE_LINES ..] for q[k]); the slope
// denominators of one step are inverted together (Montgomery's trick), one Fq2 inversion per step.
inline void g2_prepare_many(const G2Affine* q, int count, G2Line* lines) {
  if (count <= 0) return;
  G2Affine* t = new G2Affine[count];
  G2Affine* other = new G2Affine[count];
  Fq2* den = new Fq2[count];
  Fq2* pre = new Fq2[count];
  for (int k = 0; k < count; k++) t[k] = q[k];
  int n = 0;
  auto step = [&](bool is_double, const G2Affine* rhs) {
    Fq2 acc = Fq2::one();
    for (int k = 0; k < count; k++) {
      den[k] = is_double ? t[k].y.dbl() : rhs[k].x - t[k].x;
      pre[k] = acc;
      acc = acc * den[k];
    }
    Fq2 inv = acc.inv();
    for (int k = count; k-- > 0;) {
      const Fq2 di = inv * pre[k];
      inv = inv * den[k];
      Fq2 lam;
      if (is_double) {
        Fq2 x2 = t[k].x.sqr();
        lam = (x2.dbl() + x2) * di;
      } else {
        lam = (rhs[k].y - t[k].y) * di;
      }
      const Fq2 x3 = lam.sqr() - t[k].x - (is_double ? t[k].x : rhs[k].x);
      const Fq2 y3 = lam * (t[k].x - x3) - t[k].y;
      G2Line& out = lines[(size_t)k * H2V_ATE_LINES + n];
      out.nlam = lam.neg();
      out.c = lam * t[k].x - t[k].y;
      t[k].x = x3;
      t[k].y = y3;
    }
    n++;
  };
  for (int i = 63; i >= 0; i--) {
    step(true, nullptr);
    if ((H2V_ATE_LOOP_LOW >> i) & 1) step(false, q);
  }
  for (int k = 0; k < count; k++) other[k] = g2_frob(q[k]);
  step(false, other);
  for (int k = 0; k < count; k++) {
    other[k] = g2_frob(other[k]);
    other[k].y = other[k].y.neg();
  }
  step(false, other);
  delete[] t;
  delete[] other;
  delete[] den;
  delete[] pre;
}
#endif

// ------------------------------------------------------------------------------------------ pairing check
// prod_k e(P_k, Q_k) over the prepared line tables; a pair whose G1 point is the identity is
// skipped (halo2curves' multi_miller_loop does the same).  lines[k] has H2V_ATE_LINES entries.
H2V_HDN inline Fq12 miller_loop2(const G1Affine* p, const bool* skip, const G2Line* const* lines) {
  Fq12 f = Fq12::one();
  int n = 0;
  for (int i = 63; i >= 0; i--) {
    f = fq12_sqr(f);
    for (int k = 0; k < 2; k++)
      if (!skip[k]) f = fq12_mul_by_line(f, p[k].y, lines[k][n].nlam.mul_fq(p[k].x), lines[k][n].c);
    n++;
    if ((H2V_ATE_LOOP_LOW >> i) & 1) {
      for (int k = 0; k < 2; k++)
        if (!skip[k]) f = fq12_mul_by_line(f, p[k].y, lines[k][n].nlam.mul_fq(p[k].x), lines[k][n].c);
      n++;
    }
  }
  for (int e = 0; e < 2; e++) {
    for (int k = 0; k < 2; k++)
      if (!skip[k]) f = fq12_mul_by_line(f, p[k].y, lines[k][n].nlam.mul_fq(p[k].x), lines[k][n].c);
    n++;
  }
  return f;
}

H2V_HDN inline Fq12 fq12_pow_u(const Fq12& x) {
  Fq12 acc = x;
  for (int i = 61; i >= 0; i--) {  // u has 63 bits, top bit consumed by acc = x
    acc = fq12_sqr(acc);
    if ((H2V_BN_U >> i) & 1) acc = fq12_mul(acc, x);
  }
  return acc;
}

// f^((p^12-1)/r): easy part, then the Devegili-Scott-Dahab / Scott et al. y0..y6 chain, which
// equals the exact exponent (p^4-p^2+1)/r (checked against a plain power in the oracle tests).
H2V_HDN inline Fq12 final_exponentiation(const Fq12& f) {
  Fq12 t1 = fq12_mul(f.conj(), fq12_inv(f));
  t1 = fq12_mul(fq12_frob2(t1), t1);
  Fq12 fp = fq12_frob(t1), fp2 = fq12_frob2(t1), fp3 = fq12_frob(fp2);
  Fq12 fu = fq12_pow_u(t1), fu2 = fq12_pow_u(fu), fu3 = fq12_pow_u(fu2);
  Fq12 y3 = fq12_frob(fu).conj();
  Fq12 fu2p = fq12_frob(fu2), fu3p = fq12_frob(fu3);
  Fq12 y2 = fq12_frob2(fu2);
  Fq12 y0 = fq12_mul(fq12_mul(fp, fp2), fp3);
  Fq12 y1 = t1.conj();
  Fq12 y5 = fu2.conj();
  Fq12 y4 = fq12_mul(fu, fu2p).conj();
  Fq12 y6 = fq12_mul(fu3, fu3p).conj();
  Fq12 t0 = fq12_mul(fq12_mul(fq12_sqr(y6), y4), y5);
  Fq12 T1 = fq12_mul(fq12_mul(y3, y5), t0);
  t0 = fq12_mul(t0, y2);
  T1 = fq12_sqr(fq12_mul(fq12_sqr(T1), t0));
  t0 = fq12_mul(T1, y1);
  T1 = fq12_mul(T1, y0);
  return fq12_mul(fq12_sqr(t0), T1);
}

// e(L, Q0) * e(R, Q1) == 1 ?   (DualMSM::check with Q0 = [s]G2, Q1 = -G2)
H2V_HDN inline bool pairing_check2(const G1Jac& L, const G1Jac& R, const G2Line* lines0, const G2Line* lines1) {
  G1Affine p[2];
  bool skip[2];
  skip[0] = !g1_to_affine(L, p[0]);
  skip[1] = !g1_to_affine(R, p[1]);
  const G2Line* lines[2] = {lines0, lines1};
  Fq12 f = miller_loop2(p, skip, lines);
  return final_exponentiation(f).is_one();
}

}  // namespace h2v
