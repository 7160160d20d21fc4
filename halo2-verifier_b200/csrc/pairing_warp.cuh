// Warp-cooperative BN254 pairing check (device only).
//
// The batch verdict is ONE 2-pair pairing (reference poly/kzg/msm.rs:185-203) and sits on the
// critical path of every batch, so its latency matters more than its work.  A single thread walks
// ~22k dependent Montgomery multiplications; here one warp shares each Fq12 operation instead:
// an Fq12 product is three Fq6 products = 18 independent Fq2 products, one per lane, with operands
// and results staged in shared memory.  Values are bit-identical to the single-thread tower in
// tower.cuh (tests compare both with the oracle).
//
// Storage: W12 keeps the two Fq6 halves contiguously, h[0][j] = a[2j], h[1][j] = a[2j+1] of Fq12::a.
#pragma once
#include "tower.cuh"

namespace h2v {

struct W12 {
  Fq2 h[2][3];
};
struct WScratch {
  Fq2 T[3][3];  // Fq6 products
  Fq2 S[2][3];  // operand sums
  Fq2 pr[18];   // Fq2 products
  W12 line;     // line element being multiplied in
};

__device__ __forceinline__ Fq2 fq2_mul_inl(const Fq2& a, const Fq2& b) {
  Fq t0 = Fq::mul(a.c0, b.c0), t1 = Fq::mul(a.c1, b.c1);
  Fq t2 = Fq::mul(a.c0 + a.c1, b.c0 + b.c1);
  return {t0 - t1, t2 - t0 - t1};
}
__device__ __forceinline__ Fq fq_inv_inl(const Fq& a) {  // a^(p-2), plain square-and-multiply kept in registers
  Fq acc = Fq::one();
#pragma unroll 1
  for (int i = 253; i >= 0; i--) {
    acc = Fq::mul(acc, acc);
    u32 e = FqP::mod(0);
    // bit i of p - 2
    u32 limb;
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
    (void)e;
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

template <class T>
__device__ __forceinline__ T* sel3(int g, T* a, T* b, T* c) {
  return g == 0 ? a : (g == 1 ? b : c);
}

// up to 3 independent Fq6 products; lanes [6g, 6g+6) multiply, lanes [3g, 3g+3) combine
__device__ __noinline__ void w_fq6_mul(int nmul, Fq2* D0, Fq2* D1, Fq2* D2, const Fq2* A0, const Fq2* A1, const Fq2* A2,
                                          const Fq2* B0, const Fq2* B1, const Fq2* B2, Fq2* pr, int lane) {
  if (lane < 6 * nmul) {
    const int g = lane / 6, i = lane % 6;
    const Fq2* a = sel3(g, A0, A1, A2);
    const Fq2* b = sel3(g, B0, B1, B2);
    Fq2 oa, ob;
    if (i < 3) {
      oa = a[i];
      ob = b[i];
    } else {
      const int u = i == 3 ? 1 : 0, v = i == 4 ? 1 : 2;  // (1,2) (0,1) (0,2)
      oa = a[u] + a[v];
      ob = b[u] + b[v];
    }
    pr[lane] = fq2_mul_inl(oa, ob);
  }
  __syncwarp();
  if (lane < 3 * nmul) {
    const int g = lane / 3, j = lane % 3;
    const Fq2* p = pr + 6 * g;
    Fq2 c;
    if (j == 0) c = p[0] + (p[3] - p[1] - p[2]).mul_xi();
    else if (j == 1) c = p[4] - p[0] - p[1] + p[2].mul_xi();
    else c = p[5] - p[0] - p[2] + p[1];
    sel3(g, D0, D1, D2)[j] = c;
  }
  __syncwarp();
}

// dst = x * y (dst may alias x or y)
__device__ __noinline__ void w_fq12_mul(W12* dst, const W12* x, const W12* y, WScratch* ws, int lane) {
  if (lane < 3) ws->S[0][lane] = x->h[0][lane] + x->h[1][lane];
  else if (lane < 6) ws->S[1][lane - 3] = y->h[0][lane - 3] + y->h[1][lane - 3];
  __syncwarp();
  w_fq6_mul(3, ws->T[0], ws->T[1], ws->T[2], x->h[0], x->h[1], ws->S[0], y->h[0], y->h[1], ws->S[1], ws->pr, lane);
  if (lane < 3) {
    const int j = lane;
    Fq2 vt = j == 0 ? ws->T[1][2].mul_xi() : ws->T[1][j - 1];
    dst->h[0][j] = ws->T[0][j] + vt;
  } else if (lane < 6) {
    const int j = lane - 3;
    dst->h[1][j] = ws->T[2][j] - ws->T[0][j] - ws->T[1][j];
  }
  __syncwarp();
}

// dst = x^2 (complex squaring; dst may alias x)
__device__ __noinline__ void w_fq12_sqr(W12* dst, const W12* x, WScratch* ws, int lane) {
  if (lane < 3) {
    ws->S[0][lane] = x->h[0][lane] + x->h[1][lane];
  } else if (lane < 6) {
    const int j = lane - 3;
    Fq2 vx = j == 0 ? x->h[1][2].mul_xi() : x->h[1][j - 1];
    ws->S[1][j] = x->h[0][j] + vx;
  }
  __syncwarp();
  w_fq6_mul(2, ws->T[0], ws->T[1], ws->T[2], x->h[0], ws->S[0], nullptr, x->h[1], ws->S[1], nullptr, ws->pr, lane);
  if (lane < 3) {
    const int j = lane;
    Fq2 vt = j == 0 ? ws->T[0][2].mul_xi() : ws->T[0][j - 1];
    dst->h[0][j] = ws->T[1][j] - ws->T[0][j] - vt;
  } else if (lane < 6) {
    dst->h[1][lane - 3] = ws->T[0][lane - 3].dbl();
  }
  __syncwarp();
}

__device__ __forceinline__ void w_copy(W12* dst, const W12* x, int lane) {
  if (lane < 6) dst->h[lane & 1][lane >> 1] = x->h[lane & 1][lane >> 1];
  __syncwarp();
}
__device__ __forceinline__ void w_conj(W12* dst, const W12* x, int lane) {  // ^(p^6)
  if (lane < 6) {
    const int hh = lane & 1, j = lane >> 1;
    dst->h[hh][j] = hh ? x->h[hh][j].neg() : x->h[hh][j];
  }
  __syncwarp();
}
__device__ __noinline__ void w_frob(W12* dst, const W12* x, int lane) {  // ^p
  if (lane < 6) {
    const int hh = lane & 1, j = lane >> 1;  // w-power index i = lane
    Fq2 c = x->h[hh][j].conj();
    dst->h[hh][j] = lane == 0 ? c : fq2_mul_inl(c, tower_gamma1(lane));
  }
  __syncwarp();
}
__device__ __noinline__ void w_frob2(W12* dst, const W12* x, int lane) {  // ^(p^2)
  if (lane < 6) {
    const int hh = lane & 1, j = lane >> 1;
    Fq2 c = x->h[hh][j];
    if (lane) {
      Fq g = tower_gamma2(lane);
      c = {Fq::mul(c.c0, g), Fq::mul(c.c1, g)};
    }
    dst->h[hh][j] = c;
  }
  __syncwarp();
}
__device__ __forceinline__ void w_set_one(W12* dst, int lane) {
  if (lane < 6) dst->h[lane & 1][lane >> 1] = lane == 0 ? Fq2::one() : Fq2::zero();
  __syncwarp();
}

// dst = x^-1; the single Fq inversion runs on lane 0
__device__ __noinline__ void w_fq12_inv(W12* dst, const W12* x, WScratch* ws, int lane) {
  // d = x0^2 - v x1^2
  w_fq6_mul(2, ws->T[0], ws->T[1], ws->T[2], x->h[0], x->h[1], nullptr, x->h[0], x->h[1], nullptr, ws->pr, lane);
  if (lane < 3) {
    Fq2 vt = lane == 0 ? ws->T[1][2].mul_xi() : ws->T[1][lane - 1];
    ws->S[0][lane] = ws->T[0][lane] - vt;
  }
  __syncwarp();
  if (lane == 0) {  // Fq6 inverse of S[0] -> S[1]
    const Fq2 c0 = ws->S[0][0], c1 = ws->S[0][1], c2 = ws->S[0][2];
    Fq2 t0 = fq2_mul_inl(c0, c0) - fq2_mul_inl(c1, c2).mul_xi();
    Fq2 t1 = fq2_mul_inl(c2, c2).mul_xi() - fq2_mul_inl(c0, c1);
    Fq2 t2 = fq2_mul_inl(c1, c1) - fq2_mul_inl(c0, c2);
    Fq2 d = fq2_mul_inl(c0, t0) + (fq2_mul_inl(c2, t1) + fq2_mul_inl(c1, t2)).mul_xi();
    Fq nrm = fq_inv_inl(Fq::mul(d.c0, d.c0) + Fq::mul(d.c1, d.c1));
    Fq2 di = {Fq::mul(d.c0, nrm), Fq::mul(d.c1, nrm).neg()};
    ws->S[1][0] = fq2_mul_inl(t0, di);
    ws->S[1][1] = fq2_mul_inl(t1, di);
    ws->S[1][2] = fq2_mul_inl(t2, di);
  }
  __syncwarp();
  w_fq6_mul(2, ws->T[0], ws->T[1], ws->T[2], x->h[0], x->h[1], nullptr, ws->S[1], ws->S[1], nullptr, ws->pr, lane);
  if (lane < 3) dst->h[0][lane] = ws->T[0][lane];
  else if (lane < 6) dst->h[1][lane - 3] = ws->T[1][lane - 3].neg();
  __syncwarp();
}

__device__ __noinline__ void w_pow_u(W12* dst, const W12* x, WScratch* ws, int lane) {  // dst != x
  w_copy(dst, x, lane);
#pragma unroll 1
  for (int i = 61; i >= 0; i--) {
    w_fq12_sqr(dst, dst, ws, lane);
    if ((H2V_BN_U >> i) & 1) w_fq12_mul(dst, dst, x, ws, lane);
  }
}

// f *= line_k(P) for one prepared line
__device__ __noinline__ void w_mul_line(W12* f, const G2Line& ln, const G1Affine& p, WScratch* ws, int lane) {
  if (lane < 6) {
    Fq2 v = Fq2::zero();
    if (lane == 0) v = {p.y, Fq::zero()};                                                        // h[0][0] = yP
    else if (lane == 1) v = {Fq::mul(ln.nlam.c0, p.x), Fq::mul(ln.nlam.c1, p.x)};               // h[1][0] = -lambda xP
    else if (lane == 3) v = ln.c;                                                               // h[1][1] = lambda xT - yT
    ws->line.h[lane & 1][lane >> 1] = v;
  }
  __syncwarp();
  w_fq12_mul(f, f, &ws->line, ws, lane);
}

// e(P0, Q0) e(P1, Q1) == 1 with prepared lines; pool: 10 W12 values of shared memory.
__device__ __noinline__ bool w_pairing_check2(const G1Affine* p, const bool* skip, const G2Line* lines0, const G2Line* lines1,
                                                 W12* pool, WScratch* ws, int lane) {
  W12* f = &pool[0];
  w_set_one(f, lane);
  int n = 0;
#pragma unroll 1
  for (int i = 63; i >= 0; i--) {
    w_fq12_sqr(f, f, ws, lane);
    if (!skip[0]) w_mul_line(f, lines0[n], p[0], ws, lane);
    if (!skip[1]) w_mul_line(f, lines1[n], p[1], ws, lane);
    n++;
    if ((H2V_ATE_LOOP_LOW >> i) & 1) {
      if (!skip[0]) w_mul_line(f, lines0[n], p[0], ws, lane);
      if (!skip[1]) w_mul_line(f, lines1[n], p[1], ws, lane);
      n++;
    }
  }
  for (int e = 0; e < 2; e++) {
    if (!skip[0]) w_mul_line(f, lines0[n], p[0], ws, lane);
    if (!skip[1]) w_mul_line(f, lines1[n], p[1], ws, lane);
    n++;
  }
  // final exponentiation (same chain as final_exponentiation() in tower.cuh)
  W12 *t1 = &pool[1], *a = &pool[2], *b = &pool[3], *fu = &pool[4], *fu2 = &pool[5], *fu3 = &pool[6], *y0 = &pool[7],
      *t0 = &pool[8], *T1 = &pool[9];
  w_fq12_inv(a, f, ws, lane);
  w_conj(b, f, lane);
  w_fq12_mul(t1, b, a, ws, lane);   // f^(p^6-1)
  w_frob2(a, t1, lane);
  w_fq12_mul(t1, a, t1, ws, lane);  // ^(p^2+1)
  w_pow_u(fu, t1, ws, lane);
  w_pow_u(fu2, fu, ws, lane);
  w_pow_u(fu3, fu2, ws, lane);
  // y0 = frob(t1) frob2(t1) frob3(t1)
  w_frob(a, t1, lane);
  w_frob2(b, t1, lane);
  w_fq12_mul(y0, a, b, ws, lane);
  w_frob(a, b, lane);
  w_fq12_mul(y0, y0, a, ws, lane);
  // t0 = y6^2 y4 y5 with y6 = conj(fu3 frob(fu3)), y4 = conj(fu frob(fu2)), y5 = conj(fu2)
  w_frob(a, fu3, lane);
  w_fq12_mul(a, fu3, a, ws, lane);
  w_conj(a, a, lane);
  w_fq12_sqr(t0, a, ws, lane);
  w_frob(a, fu2, lane);
  w_fq12_mul(a, fu, a, ws, lane);
  w_conj(a, a, lane);
  w_fq12_mul(t0, t0, a, ws, lane);
  w_conj(b, fu2, lane);  // y5
  w_fq12_mul(t0, t0, b, ws, lane);
  // T1 = y3 y5 t0 with y3 = conj(frob(fu))
  w_frob(a, fu, lane);
  w_conj(a, a, lane);
  w_fq12_mul(T1, a, b, ws, lane);
  w_fq12_mul(T1, T1, t0, ws, lane);
  // t0 = t0 y2, y2 = frob2(fu2)
  w_frob2(a, fu2, lane);
  w_fq12_mul(t0, t0, a, ws, lane);
  // T1 = (T1^2 t0)^2
  w_fq12_sqr(T1, T1, ws, lane);
  w_fq12_mul(T1, T1, t0, ws, lane);
  w_fq12_sqr(T1, T1, ws, lane);
  // t0 = T1 y1 (y1 = conj(t1)); T1 = T1 y0; result = t0^2 T1
  w_conj(a, t1, lane);
  w_fq12_mul(t0, T1, a, ws, lane);
  w_fq12_mul(T1, T1, y0, ws, lane);
  w_fq12_sqr(t0, t0, ws, lane);
  w_fq12_mul(t0, t0, T1, ws, lane);
  bool ok = true;
  if (lane < 6) {
    const Fq2 v = t0->h[lane & 1][lane >> 1];
    ok = lane == 0 ? (v == Fq2::one()) : v.is_zero();
  }
  return __all_sync(0xFFFFFFFFu, ok);
}

static constexpr int H2V_WPOOL = 10;

}  // namespace h2v
