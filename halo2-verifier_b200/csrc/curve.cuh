// BN254 G1 (y^2 = x^3 + 3) group law on Montgomery-form Fq, Jacobian accumulators.
//
// Replaces the halo2curves calls the reference makes at arithmetic.rs:42,56-59,89-92 (double / add /
// mixed add inside multiexp_serial), msm.rs:78-86 (batch_normalize + best_multiexp) and the point
// decompression behind transcript/mod.rs:161-162 (`C::from_bytes`).
#pragma once
#include "field.cuh"

namespace h2v {

struct G1Affine {
  Fq x, y;  // Montgomery form; the identity never appears as an affine proof point
};

struct G1Jac {
  Fq X, Y, Z;  // Z == 0 <=> identity
  static H2V_HD G1Jac identity() {
    G1Jac r;
    r.X = Fq::one();
    r.Y = Fq::one();
    r.Z = Fq::zero();
    return r;
  }
  H2V_HD bool is_identity() const { return Z.is_zero(); }
  static H2V_HD G1Jac from_affine(const G1Affine& a) {
    G1Jac r;
    r.X = a.x;
    r.Y = a.y;
    r.Z = Fq::one();
    return r;
  }
};

H2V_HD Fq fq_b3() { return Fq::from_u32(3); }

H2V_HD bool g1_on_curve(const G1Affine& p) {
  Fq rhs = p.x.sqr() * p.x + fq_b3();
  return p.y.sqr() == rhs;
}

// dbl-2009-l (a = 0): 2M + 5S.  Products through the out-of-line multiplication (Fq::mul_c / sqr_c): g1_double and g1_add serve
// the latency-bound kernels (bucket reduction, window combination, attribution), where a few warps per SM walk the code once
// per point: with every product inlined the pair was ~4.6 k instructions and 27 % of k_msm_chunk_reduce's warp stalls were
// instruction fetches (ncu, profiles/r2b_narrow_summary.txt).  The multiplier-bound kernels use g1_add_mixed (inlined).
H2V_HDN inline G1Jac g1_double(const G1Jac& p) {
  if (p.is_identity()) return p;
  Fq A = Fq::sqr_c(p.X);
  Fq B = Fq::sqr_c(p.Y);
  Fq C = Fq::sqr_c(B);
  Fq t = p.X + B;
  Fq D = (Fq::sqr_c(t) - A - C).dbl();
  Fq E = A.dbl() + A;
  Fq F = Fq::sqr_c(E);
  G1Jac r;
  r.X = F - D.dbl();
  Fq C8 = C.dbl().dbl().dbl();
  r.Y = Fq::mul_c(E, D - r.X) - C8;
  r.Z = Fq::mul_c(p.Y, p.Z).dbl();
  return r;
}

// Jacobian += affine (madd-2007-bl style, 7M + 4S) with the exceptional cases handled exactly:
// equal points -> doubling, opposite points -> identity.  `negate` adds -q.
H2V_HDN inline G1Jac g1_add_mixed(const G1Jac& p, const G1Affine& q, bool negate = false) {
  Fq qy = negate ? q.y.neg() : q.y;
  if (p.is_identity()) {
    G1Jac r;
    r.X = q.x;
    r.Y = qy;
    r.Z = Fq::one();
    return r;
  }
  Fq Z1Z1 = p.Z.sqr();
  Fq U2 = q.x * Z1Z1;
  Fq S2 = qy * p.Z * Z1Z1;
  Fq H = U2 - p.X;
  Fq rr = S2 - p.Y;
  if (H.is_zero()) {
    if (rr.is_zero()) return g1_double(p);
    return G1Jac::identity();
  }
  Fq HH = H.sqr();
  Fq HHH = HH * H;
  Fq V = p.X * HH;
  G1Jac r;
  r.X = rr.sqr() - HHH - V.dbl();
  r.Y = rr * (V - r.X) - p.Y * HHH;
  r.Z = p.Z * H;
  return r;
}

// Jacobian + Jacobian (11M + 5S), exceptional cases handled exactly.
H2V_HDN inline G1Jac g1_add(const G1Jac& p, const G1Jac& q) {
  if (p.is_identity()) return q;
  if (q.is_identity()) return p;
  Fq Z1Z1 = Fq::sqr_c(p.Z);
  Fq Z2Z2 = Fq::sqr_c(q.Z);
  Fq U1 = Fq::mul_c(p.X, Z2Z2);
  Fq U2 = Fq::mul_c(q.X, Z1Z1);
  Fq S1 = Fq::mul_c(Fq::mul_c(p.Y, q.Z), Z2Z2);
  Fq S2 = Fq::mul_c(Fq::mul_c(q.Y, p.Z), Z1Z1);
  Fq H = U2 - U1;
  Fq rr = S2 - S1;
  if (H.is_zero()) {
    if (rr.is_zero()) return g1_double(p);
    return G1Jac::identity();
  }
  Fq HH = Fq::sqr_c(H);
  Fq HHH = Fq::mul_c(HH, H);
  Fq V = Fq::mul_c(U1, HH);
  G1Jac r;
  r.X = Fq::sqr_c(rr) - HHH - V.dbl();
  r.Y = Fq::mul_c(rr, V - r.X) - Fq::mul_c(S1, HHH);
  r.Z = Fq::mul_c(Fq::mul_c(p.Z, q.Z), H);
  return r;
}

H2V_HD G1Jac g1_neg(const G1Jac& p) {
  G1Jac r = p;
  r.Y = p.Y.neg();
  return r;
}

// Jacobian -> affine; returns false for the identity (x = y = 0 then).
H2V_HDN inline bool g1_to_affine(const G1Jac& p, G1Affine& out) {
  if (p.is_identity()) {
    out.x = Fq::zero();
    out.y = Fq::zero();
    return false;
  }
  Fq zi = p.Z.inv();
  Fq zi2 = zi.sqr();
  out.x = p.X * zi2;
  out.y = p.Y * zi2 * zi;
  return true;
}

// [k]P for a canonical (non-Montgomery) 256-bit scalar, plain left-to-right double-and-add.
H2V_HDN inline G1Jac g1_mul_canonical(const G1Affine& p, const u32* k) {
  G1Jac acc = G1Jac::identity();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) acc = g1_double(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) {
      acc = g1_add_mixed(acc, p);
      started = true;
    }
  }
  return acc;
}

// Point decompression, the halo2curves `from_bytes` convention (kept in ONE place):
// 32 bytes little-endian x; bit 7 of byte 31 = parity of canonical y; bit 6 must be clear.
// Returns false for every encoding the reference's read_point rejects (invalid encoding, or the
// identity, which common_point refuses: transcript/mod.rs:161-163,218-219).
// canon (optional): the CANONICAL coordinates x | y as the transcript absorbs them (transcript/mod.rs:216-224), which fall
// out of the decompression anyway (x is the input, y's canonical form decides the sign).
H2V_HDN inline bool g1_decompress(const u8* b, G1Affine& out, Fq* canon = nullptr) {
  Fq x = Fq::load_le(b);
  const bool sign = (x.l[7] >> 31) & 1;
  const bool inf_bit = (x.l[7] >> 30) & 1;
  x.l[7] &= 0x3FFFFFFFu;
  if (inf_bit || x.geq_mod()) return false;
  Fq xm = Fq::from_canonical(x);
  Fq rhs = xm.sqr() * xm + fq_b3();
  Fq y = fq_sqrt_candidate(rhs);
  if (y.sqr() != rhs) return false;  // also rejects the all-zero string: 3 is a non-residue
  Fq yc = y.to_canonical();
  if ((bool)(yc.l[0] & 1) != sign) {
    y = y.neg();
    if (canon) {  // y != 0 here (0 is even: a set sign bit on y = 0 cannot occur since 3 is a non-residue, x^3 + 3 != 0)
      u32 m[8];
#pragma unroll
      for (int i = 0; i < 8; i++) m[i] = FqP::mod(i);
      Fq::sub_raw(yc.l, m, yc.l);
    }
  }
  if (canon) {
    canon[0] = x;
    canon[1] = yc;
  }
  out.x = xm;
  out.y = y;
  return true;
}

#if defined(__CUDACC__)
// Fully inlined variants for the serial chains (window combination, conversion to affine): device only.
__device__ __forceinline__ Fq fq_inv_inl(const Fq& a) {  // a^(p-2), plain square-and-multiply kept in registers
  Fq acc = Fq::one();
#pragma unroll 1
  for (int i = 253; i >= 0; i--) {
    acc = Fq::mul(acc, acc);
    u32 limb;  // bit i of p - 2
    switch (i >> 5) {
      case 0: limb = FqP::mod(0) - 2; break;
      case 1: limb = FqP::mod(1); break;
      case 2: limb = FqP::mod(2); break;
      case 3: limb = FqP::mod(3); break;
      case 4: limb = FqP::mod(4); break;
      case 5: limb = FqP::mod(5); break;
      case 6: limb = FqP::mod(6); break;
      default: limb = FqP::mod(7); break;
    }
    if ((limb >> (i & 31)) & 1) acc = Fq::mul(acc, a);
  }
  return acc;
}
// Jacobian doubling with everything inlined (serial window combination: ~255 dependent doublings)
__device__ __forceinline__ void g1_double_inl(G1Jac& p) {
  if (p.Z.is_zero()) return;
  Fq A = Fq::mul(p.X, p.X);
  Fq B = Fq::mul(p.Y, p.Y);
  Fq C = Fq::mul(B, B);
  Fq t = p.X + B;
  Fq D = (Fq::mul(t, t) - A - C).dbl();
  Fq E = A.dbl() + A;
  Fq F = Fq::mul(E, E);
  Fq Z3 = Fq::mul(p.Y, p.Z).dbl();
  p.X = F - D.dbl();
  p.Y = Fq::mul(E, D - p.X) - C.dbl().dbl().dbl();
  p.Z = Z3;
}
__device__ __forceinline__ bool g1_to_affine_inl(const G1Jac& p, G1Affine& out) {
  if (p.Z.is_zero()) {
    out.x = Fq::zero();
    out.y = Fq::zero();
    return false;
  }
  Fq zi = fq_inv_inl(p.Z);
  Fq zi2 = Fq::mul(zi, zi);
  out.x = Fq::mul(p.X, zi2);
  out.y = Fq::mul(Fq::mul(p.Y, zi2), zi);
  return true;
}
#endif

}  // namespace h2v
