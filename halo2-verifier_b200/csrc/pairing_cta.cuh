// Cooperative Fq12 engine and the batch pairing check built on it.
//
// The batch verdict is ONE pairing-product check (reference poly/kzg/msm.rs:185-203,
// `DualMSM::check`) and sits on the critical path of every batch, so it is organised for LATENCY:
//
//  * No window combination and no inversion.  The folded MSM leaves one Jacobian sum S_w per
//    (channel, window); instead of the serial Horner chain  R = sum_w 2^(c w) S_w  (c*W dependent
//    doublings) the check uses bilinearity,  e(R, Q) = prod_w e(S_w, [2^(c w)] Q):  the G2 multiples
//    are verifier parameters and their Miller lines are prepared once per window geometry.  Lines
//    are evaluated at PROJECTIVE G1 points (scaled by Z^3, an Fq factor the final exponentiation
//    kills), and the final exponentiation is checked without dividing:  f^((p^6-1) M) = 1  <=>
//    conj(F) = F  with F = f^M = N / D,  i.e.  conj(N) D = N conj(D)  (N, D: the positive- and
//    negative-exponent halves of the hard-part addition chain).
//  * k_lines: every Miller-loop iteration's line product  prod_pairs l(S_pair)  is independent of
//    the running value, so all 65 of them are computed by 65 thread blocks in parallel.
//  * k_pairing_check: one 64-thread group then walks  f = f^2 * M_i  and the final exponentiation;
//    every Fq12 product is 54 independent Montgomery multiplications (one per lane) plus one
//    table-driven linear map (pairing_lin.inc), with the operands kept in Karatsuba-expanded form.
//
// Values are identical to the single-thread tower in tower.cuh (tests/hostlib runs the per-lane
// phase functions on the host against fq12_mul; GPU tests compare verdicts with the oracle).
#pragma once
#include "pairing_lin.inc"
#include "tower.cuh"

namespace h2v {

static constexpr int E12_N = 54;   // expanded coordinates
static constexpr int E12_NB = 12;  // base coordinates b = 2 * (3 * hh + j) + part  <->  Fq12::a[2 j + hh].c{part}

struct E12 {
  Fq e[E12_N];
};

struct alignas(16) LinTables {  // term lists 8-byte aligned: four terms are fetched with one 64-bit load
  alignas(8) uint16_t full_terms[H2V_LIN_FULL_NTERMS];
  alignas(8) uint16_t exp_terms[H2V_LIN_EXP_NTERMS];
  uint16_t full_start[E12_N + 1];
  uint16_t exp_start[E12_N + 1];
};
#define H2V_LIN_TABLES_INIT {H2V_LIN_FULL_TERMS_INIT, H2V_LIN_EXP_TERMS_INIT, H2V_LIN_FULL_START_INIT, H2V_LIN_EXP_START_INIT}

H2V_HD int e12_base_slot(int b) {  // position of base coordinate b inside the expanded form
  const int hh = b / 6, j = (b % 6) / 2, part = b & 1;
  return 3 * (6 * hh + j) + part;
}

// accumulates terms [k0, k1) of a row (k0, k1 multiples of 4: rows are padded) into 64-bit column accumulators
H2V_HD void lin_row_acc(u64* acc, const uint16_t* terms, int k0, int k1, const Fq* src) {
  for (int k = k0; k < k1; k += 4) {
    const u64 tt = *(const u64*)(terms + k);  // 4 packed (index | coefficient << 8) terms
    const u32 tx = (u32)tt, ty = (u32)(tt >> 32);
    const Fq v0 = src[tx & 0xFF], v1 = src[(tx >> 16) & 0xFF], v2 = src[ty & 0xFF], v3 = src[(ty >> 16) & 0xFF];
    const u32 c0 = (tx >> 8) & 0xFF, c1 = tx >> 24, c2 = (ty >> 8) & 0xFF, c3 = ty >> 24;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] += (u64)c0 * v0.l[i] + (u64)c1 * v1.l[i];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] += (u64)c2 * v2.l[i] + (u64)c3 * v3.l[i];
  }
}
// column accumulators (value < 2^9 p) -> reduced field element
H2V_HD Fq lin_reduce(const u64* acc) {
  u32 V[9];
  u64 cy = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    cy += acc[i];
    V[i] = (u32)cy;
    cy >>= 32;
  }
  V[8] = (u32)cy;
  // quotient estimate from the top bits: q <= floor(V / p) <= q + 2   (1354 = floor(2^32 / ((p >> 232) + 1)))
  const u32 q = (u32)(((u64)((V[8] << 24) | (V[7] >> 8)) * 1354u) >> 32);
  Fq r;
  u64 mc = 0;
  u32 br = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    mc += (u64)q * FqP::mod(i);
    const u32 m = (u32)mc;
    mc >>= 32;
    const u64 d = (u64)V[i] - m - br;
    r.l[i] = (u32)d;
    br = (u32)(d >> 32) & 1u;
  }
  r.cond_sub_mod();  // remainder < 3 p < 2^256
  r.cond_sub_mod();
  return r;
}
// sum_k coef_k * src[idx_k] mod p for one table row: 64-bit column accumulators, one reduction.
// Every coefficient is positive (negative terms address the negated copy of the source) and every
// source value is <= p, so the sum is < 2^9 p.
H2V_HD Fq lin_row(const uint16_t* start, const uint16_t* terms, int row, const Fq* src) {
  u64 acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0;
  lin_row_acc(acc, terms, start[row], start[row + 1], src);
  return lin_reduce(acc);
}

// ---- the latency-critical variant of the linear map (k_pairing_check: one 128-thread group walks ~435 dependent Fq12
// products): every row of FULL is split over 2-3 consecutive threads of one warp, and a thread's terms (H2V_LIN_FAST_CAP
// packed 16-bit terms + a meta word, pairing_lin.inc) stay in REGISTERS for the whole kernel, so that a product's map is a
// fully unrolled run of independent shared-memory loads and multiply-adds instead of a table walk.
static constexpr int FAST_W = H2V_LIN_FAST_CAP / 2;  // term words per thread
struct FastTerms {
  u32 w[FAST_W];
  u32 meta;  // row | lead << 8 | followers << 16; row 255 = idle thread
  H2V_HD int row() const { return (int)(meta & 0xFF); }
  H2V_HD bool lead() const { return ((meta >> 8) & 1) != 0; }
  H2V_HD int followers() const { return (int)(meta >> 16); }
};
static constexpr int FAST_MAX_FOLLOWERS = H2V_LIN_FAST_MAX_FOLLOWERS;
// partial column accumulators of one thread's terms
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void lds_fq(u32 (&v)[8], u32 saddr) {  // explicit shared-memory loads: through a generic pointer they became LD
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%8];\n\tld.shared.v4.u32 {%4, %5, %6, %7}, [%8 + 16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(saddr));
}
#endif
H2V_HD void fast_partial(u64* acc, const FastTerms& ft, const Fq* src) {
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0;
#if defined(__CUDA_ARCH__)
  const u32 sbase = (u32)__cvta_generic_to_shared(src);
#endif
#pragma unroll
  for (int k = 0; k < FAST_W; k++) {
    const u32 tt = ft.w[k];
    const u32 c0 = (tt >> 8) & 0xFF, c1 = tt >> 24;
#if defined(__CUDA_ARCH__)
    u32 v0[8], v1[8];
    lds_fq(v0, sbase + (tt & 0xFF) * 32u);
    lds_fq(v1, sbase + ((tt >> 16) & 0xFF) * 32u);
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] += (u64)c0 * v0[i] + (u64)c1 * v1[i];
#else
    const Fq v0 = src[tt & 0xFF], v1 = src[(tt >> 16) & 0xFF];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] += (u64)c0 * v0.l[i] + (u64)c1 * v1.l[i];
#endif
  }
}

// ---- per-lane phases (host/device; the device wrappers below add the barriers)
// products: pr[lane] = a[lane] * b[lane], pr[54 + lane] = p - pr[lane]
H2V_HD void e12_mul_p1(Fq* pr, const E12* a, const E12* b, int lane) {
  const Fq t = Fq::mul(a->e[lane], b->e[lane]);
  pr[lane] = t;
  Fq n;
  u32 m[8];
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = FqP::mod(i);
  Fq::sub_raw(n.l, m, t.l);
  pr[E12_N + lane] = n;
}
H2V_HD void e12_mul_p2(E12* dst, const Fq* pr, const LinTables* lt, int lane) {
  dst->e[lane] = lin_row(lt->full_start, lt->full_terms, lane, pr);
}
H2V_HD void e12_expand_p(E12* dst, const Fq* base12, const LinTables* lt, int lane) {
  dst->e[lane] = lin_row(lt->exp_start, lt->exp_terms, lane, base12);
}
// base coordinate b of conj(x) = x^(p^6), of x^p and of x^(p^2)   (lane = b < 12)
H2V_HD Fq e12_conj_base(const E12* x, int b) {
  const Fq v = x->e[e12_base_slot(b)];
  return b >= 6 ? v.neg() : v;
}
H2V_HD Fq e12_frob_base(const E12* x, int b) {
  const int hh = b / 6, j = (b % 6) / 2, part = b & 1, i = 2 * j + hh;
  const Fq re = x->e[e12_base_slot(b & ~1)], im = x->e[e12_base_slot(b | 1)];
  if (i == 0) return part ? im.neg() : re;
  const Fq2 g = tower_gamma1(i);  // conj(c) * g = (re g0 + im g1) + (re g1 - im g0) u
  return part ? Fq::mul(re, g.c1) - Fq::mul(im, g.c0) : Fq::mul(re, g.c0) + Fq::mul(im, g.c1);
}
H2V_HD Fq e12_frob2_base(const E12* x, int b) {
  const int hh = b / 6, j = (b % 6) / 2, i = 2 * j + hh;
  const Fq v = x->e[e12_base_slot(b)];
  return i == 0 ? v : Fq::mul(v, tower_gamma2(i));
}
// base coordinate b of the line value  Y + (nlam * XZ) w + (c * Z3) w^3  at the projective point
// (X, Y, Z) of G1, given XZ = X Z and Z3 = Z^3 (affine line scaled by Z^3; identity point: line = 1)
H2V_HD Fq e12_line_base(const G2Line& ln, const Fq& Y, const Fq& XZ, const Fq& Z3, bool identity, int b) {
  if (identity) return b == 0 ? Fq::one() : Fq::zero();
  switch (b) {
    case 0: return Y;                         // a[0].c0
    case 6: return Fq::mul(ln.nlam.c0, XZ);   // a[1].c0   (hh = 1, j = 0)
    case 7: return Fq::mul(ln.nlam.c1, XZ);
    case 8: return Fq::mul(ln.c.c0, Z3);      // a[3].c0   (hh = 1, j = 1)
    case 9: return Fq::mul(ln.c.c1, Z3);
    default: return Fq::zero();
  }
}
// Fq12 (six Fq2 coefficients of w^i) <-> base coordinates, for tests and for loading constants
H2V_HD Fq fq12_base_coord(const Fq12& x, int b) {
  const int hh = b / 6, j = (b % 6) / 2, part = b & 1;
  const Fq2& c = x.a[2 * j + hh];
  return part ? c.c1 : c.c0;
}

// index of the first prepared line of Miller-loop iteration `it` (0..63; 64 = the two Frobenius lines)
H2V_HD int ate_line_index(int it) {
  int n = 0;
  for (int k = 0; k < it && k < 64; k++) n += 1 + (int)((H2V_ATE_LOOP_LOW >> (63 - k)) & 1);
  return n;
}
H2V_HD int ate_lines_in_iteration(int it) { return it >= 64 ? 2 : 1 + (int)((H2V_ATE_LOOP_LOW >> (63 - it)) & 1); }
static constexpr int H2V_ATE_ITERS = 65;

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// device: a "group" is 64 threads (2 warps) sharing one named barrier
__device__ const LinTables g_lin_tables = H2V_LIN_TABLES_INIT;

struct Grp {
  int lane;  // 0..nthr-1
  int bar;   // named barrier id (1..15)
  const LinTables* lt;
  Fq* scr;   // 2 * E12_N values of shared scratch
  int nthr;  // 64, or 128: every row of the linear map is split over two lanes (latency-critical single group)
  u64* part; // nthr == 128: E12_N x 8 partial column accumulators
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory"); }
};

__device__ __forceinline__ void g_mul(const Grp& g, E12* dst, const E12* a, const E12* b) {  // dst may alias a, b
  if (g.lane < E12_N) e12_mul_p1(g.scr, a, b, g.lane);
  g.sync();
  if (g.nthr == 64) {
    if (g.lane < E12_N) e12_mul_p2(dst, g.scr, g.lt, g.lane);
    g.sync();
    return;
  }
  const int half = g.lane >> 6, row = g.lane & 63;
  u64 acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0;
  if (row < E12_N) {
    const int k0 = g.lt->full_start[row], k1 = g.lt->full_start[row + 1];
    const int mid = k0 + (((k1 - k0) / 4 + 1) / 2) * 4;
    lin_row_acc(acc, g.lt->full_terms, half ? mid : k0, half ? k1 : mid, g.scr);
    if (half) {
#pragma unroll
      for (int i = 0; i < 8; i++) g.part[row * 8 + i] = acc[i];
    }
  }
  g.sync();
  if (!half && row < E12_N) {
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] += g.part[row * 8 + i];
    dst->e[row] = lin_reduce(acc);
  }
  g.sync();
}
// fast group (128 threads): products by lanes 0..53, then the register-resident split map with shuffle combination
__device__ const u32 g_lin_fast[H2V_LIN_FAST_THREADS * (FAST_W + 1)] = H2V_LIN_FAST_INIT;
__device__ __forceinline__ FastTerms fast_terms_of(int t) {
  FastTerms ft;
#pragma unroll
  for (int k = 0; k < FAST_W; k++) ft.w[k] = g_lin_fast[t * (FAST_W + 1) + k];
  ft.meta = g_lin_fast[t * (FAST_W + 1) + FAST_W];
  return ft;
}
__device__ __forceinline__ void gf_mul(const Grp& g, const FastTerms& ft, E12* dst, const E12* a, const E12* b) {  // dst may alias a, b
  if (g.lane < E12_N) e12_mul_p1(g.scr, a, b, g.lane);
  g.sync();
  u64 acc[8], sum[8];
  fast_partial(acc, ft, g.scr);
#pragma unroll
  for (int i = 0; i < 8; i++) sum[i] = acc[i];
  const int fol = ft.followers();
#pragma unroll
  for (int f = 1; f <= FAST_MAX_FOLLOWERS; f++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const u64 o = __shfl_down_sync(0xFFFFFFFFu, acc[i], f);
      if (f <= fol) sum[i] += o;
    }
  }
  if (ft.lead()) dst->e[ft.row()] = lin_reduce(sum);
  g.sync();
}

__device__ __forceinline__ void g_expand(const Grp& g, E12* dst) {  // from base coordinates in scr[0..11]
  g.sync();
  if (g.lane < E12_N) e12_expand_p(dst, g.scr, g.lt, g.lane);
  g.sync();
}
__device__ __forceinline__ void g_conj(const Grp& g, E12* dst, const E12* a) {
  if (g.lane < E12_NB) g.scr[g.lane] = e12_conj_base(a, g.lane);
  g_expand(g, dst);
}
__device__ __noinline__ void g_frob(const Grp& g, E12* dst, const E12* a) {
  if (g.lane < E12_NB) g.scr[g.lane] = e12_frob_base(a, g.lane);
  g_expand(g, dst);
}
__device__ __noinline__ void g_frob2(const Grp& g, E12* dst, const E12* a) {
  if (g.lane < E12_NB) g.scr[g.lane] = e12_frob2_base(a, g.lane);
  g_expand(g, dst);
}
__device__ __forceinline__ void g_one(const Grp& g, E12* dst) {
  if (g.lane < E12_NB) g.scr[g.lane] = g.lane == 0 ? Fq::one() : Fq::zero();
  g_expand(g, dst);
}
__device__ __forceinline__ void g_copy(const Grp& g, E12* dst, const E12* a) {
  if (g.lane < E12_N) dst->e[g.lane] = a->e[g.lane];
  g.sync();
}
__device__ __noinline__ void g_pow_u(const Grp& g, E12* dst, const E12* x) {  // dst != x
  g_copy(g, dst, x);
#pragma unroll 1
  for (int i = 61; i >= 0; i--) {
    g_mul(g, dst, dst, dst);
    if ((H2V_BN_U >> i) & 1) g_mul(g, dst, dst, x);
  }
}

// (inlined: as a separate function its FastTerms argument would live in local memory and the scratch pointer would be generic)
__device__ __forceinline__ void gf_pow_u(const Grp& g, const FastTerms& ft, E12* dst, const E12* x) {  // dst != x
  g_copy(g, dst, x);
#pragma unroll 1
  for (int i = 61; i >= 0; i--) {
    gf_mul(g, ft, dst, dst, dst);
    if ((H2V_BN_U >> i) & 1) gf_mul(g, ft, dst, dst, x);
  }
}

// ---- k_lines: M[it] = prod over pairs and over the lines of iteration `it` of line(S_pair)
struct LinesArgs {
  u32 n_pairs;  // (channel, window) pairs = window sums; pair p uses lines[p * H2V_ATE_LINES ..]
};
// Which of the launched check groups are live (attribution runs many small checks and most of them are not needed):
// group index i = base + block index; skipped when `count` is given and i >= *count, or when flags[i / div] != 0.
struct PairSkip {
  const u32* flags;
  const u32* count;
  u32 div, base;
  __device__ __forceinline__ bool skip(u32 blk) const {
    const u32 i = base + blk;
    return (count && i >= *count) || (flags && flags[i / div] != 0);
  }
};
struct LinesSmem {  // dynamic shared memory layout of k_lines<GROUPS>
  LinTables lt;
  Fq Y[128], XZ[128], Z3[128];
  u32 ident[128];
};
template <int GROUPS>
__global__ void __launch_bounds__(64 * GROUPS) k_lines(LinesArgs la, const G1Jac* __restrict__ wsums, const G2Line* __restrict__ lines,
                                                       E12* __restrict__ M, PairSkip sk) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // programmatic dependent launch, see pdl_prologue() in kernels.cu
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (sk.skip(blockIdx.y)) return;
  ::TlScope tl_(9, M);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LinesSmem* sm = (LinesSmem*)smem_raw;
  E12* accs = (E12*)(sm + 1);            // [GROUPS]
  E12* tmps = accs + GROUPS;             // [GROUPS]
  Fq* scrs = (Fq*)(tmps + GROUPS);       // [GROUPS][2 * E12_N]
  const int t = threadIdx.x, gi = t >> 6;
  wsums += (size_t)blockIdx.y * la.n_pairs;  // fold group blockIdx.y: its window sums, its iteration products
  M += (size_t)blockIdx.y * H2V_ATE_ITERS;
  for (int i = t; i < (int)(sizeof(LinTables) / 2); i += blockDim.x) ((uint16_t*)&sm->lt)[i] = ((const uint16_t*)&g_lin_tables)[i];
  if (t < (int)la.n_pairs) {
    const G1Jac s = wsums[t];
    const bool id = s.Z.is_zero();
    sm->ident[t] = id;
    sm->Y[t] = s.Y;
    sm->XZ[t] = Fq::mul(s.X, s.Z);
    sm->Z3[t] = Fq::mul(Fq::mul(s.Z, s.Z), s.Z);
  }
  __syncthreads();
  Grp g{t & 63, gi + 1, &sm->lt, scrs + gi * 2 * E12_N, 64, nullptr};
  E12 *acc = accs + gi, *tmp = tmps + gi;
  const int it = blockIdx.x;
  const int n0 = it >= 64 ? H2V_ATE_LINES - 2 : ate_line_index(it), ns = ate_lines_in_iteration(it);
  const int items = ns * (int)la.n_pairs;
  bool first = true;
  for (int item = gi; item < items; item += GROUPS) {
    const int p = item % (int)la.n_pairs, s = item / (int)la.n_pairs;
    if (g.lane < E12_NB)
      g.scr[g.lane] = e12_line_base(lines[(size_t)p * H2V_ATE_LINES + n0 + s], sm->Y[p], sm->XZ[p], sm->Z3[p], sm->ident[p] != 0, g.lane);
    g_expand(g, first ? acc : tmp);
    if (!first) g_mul(g, acc, acc, tmp);
    first = false;
  }
  if (first) g_one(g, acc);
  __syncthreads();
  for (int s = GROUPS / 2; s >= 1; s >>= 1) {
    if (gi < s) g_mul(g, acc, acc, accs + gi + s);
    __syncthreads();
  }
  if (gi == 0 && g.lane < E12_N) M[it].e[g.lane] = acc->e[g.lane];
}
template <int GROUPS>
constexpr size_t k_lines_smem() {
  return sizeof(LinesSmem) + GROUPS * (2 * sizeof(E12) + 2 * E12_N * sizeof(Fq));
}

// ---- Miller accumulation over the prepared iteration products + final check
// The accumulation  f = M_64 * prod_{it < 64} M_it^(2^(63 - it))  is a chain of 127 dependent products when one block walks
// f = f^2 * M_it.  k_miller_segments walks FOUR segments of the iterations on four blocks (four SMs) in parallel and squares
// every partial product into place: segments [0,5) [5,15) [15,35) [35,64] cost 67 products each; k_pairing_check<true> then
// multiplies the four partial products: 70 instead of 127 products on the critical path.  (Four engines inside ONE block were
// measured first: they share the SM's shared-memory bandwidth and the kernel got slower, 0.72 -> 0.82 ms.)  Used when the
// launch has few groups (the batch path); the many small checks of the attribution keep the single-block walk.
static constexpr int MILLER_SEGS = 4;
__device__ __forceinline__ int miller_seg_begin(int e) { return e == 0 ? 0 : e == 1 ? 5 : e == 2 ? 15 : 35; }
__device__ __forceinline__ int miller_seg_end(int e) { return e == 0 ? 5 : e == 1 ? 15 : e == 2 ? 35 : 64; }

// f = M_a; f = f^2 * M_it for it in (a, b); then (64 - b) squarings; the last segment also takes the Frobenius lines M_64
__device__ __forceinline__ void miller_walk(const Grp& g, const FastTerms& ft, E12* f, E12* m, const E12* __restrict__ M, int a, int b, bool last) {
  g_copy(g, f, &M[a]);
#pragma unroll 1
  for (int it = a + 1; it < b; it++) {
    g_copy(g, m, &M[it]);
    gf_mul(g, ft, f, f, f);
    gf_mul(g, ft, f, f, m);
  }
#pragma unroll 1
  for (int i = b; i < 64; i++) gf_mul(g, ft, f, f, f);
  if (last) {
    g_copy(g, m, &M[64]);
    gf_mul(g, ft, f, f, m);
  }
}

__global__ void __launch_bounds__(128, 1) k_miller_segments(const E12* __restrict__ M, E12* __restrict__ seg) {  // grid (MILLER_SEGS, groups)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  ::TlScope tl_(10, M);
  __shared__ LinTables lt;
  __shared__ E12 slot[2];
  __shared__ Fq scr[2 * E12_N];
  const int t = threadIdx.x, e = blockIdx.x;
  M += (size_t)blockIdx.y * H2V_ATE_ITERS;
  seg += (size_t)blockIdx.y * MILLER_SEGS + e;
  for (int i = t; i < (int)(sizeof(LinTables) / 2); i += 128) ((uint16_t*)&lt)[i] = ((const uint16_t*)&g_lin_tables)[i];
  __syncthreads();
  Grp g{t, 1, &lt, scr, 128, nullptr};
  const FastTerms ft = fast_terms_of(t);
  miller_walk(g, ft, &slot[0], &slot[1], M, miller_seg_begin(e), miller_seg_end(e), e == MILLER_SEGS - 1);
  if (t < E12_N) seg->e[t] = slot[0].e[t];
}

// SEG: M holds the MILLER_SEGS partial products of every group (k_miller_segments) instead of the 65 iteration products
template <bool SEG>
__global__ void __launch_bounds__(128, 1) k_pairing_check(const E12* __restrict__ M, u32* verdict, PairSkip sk) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // programmatic dependent launch, see pdl_prologue() in kernels.cu
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (sk.skip(blockIdx.x)) return;
  ::TlScope tl_(10, M);
  __shared__ LinTables lt;
  __shared__ E12 slot[12];
  __shared__ Fq scr[2 * E12_N];
  const int t = threadIdx.x;
  M += (size_t)blockIdx.x * (SEG ? MILLER_SEGS : H2V_ATE_ITERS);  // one block per fold group
  verdict += blockIdx.x;
  for (int i = t; i < (int)(sizeof(LinTables) / 2); i += 128) ((uint16_t*)&lt)[i] = ((const uint16_t*)&g_lin_tables)[i];
  __syncthreads();
  Grp g{t, 1, &lt, scr, 128, nullptr};
  const FastTerms ft = fast_terms_of(t);
  E12 *f = &slot[0], *tt = &slot[1], *fu = &slot[2], *fu2 = &slot[3], *fu3 = &slot[4], *a = &slot[5], *b = &slot[6], *y0 = &slot[7],
      *T0 = &slot[8], *T1 = &slot[9], *N = &slot[10], *m = &slot[11];
  if (SEG) {
    g_copy(g, f, &M[0]);
#pragma unroll 1
    for (int e = 1; e < MILLER_SEGS; e++) {
      g_copy(g, m, &M[e]);
      gf_mul(g, ft, f, f, m);
    }
  } else {
    miller_walk(g, ft, f, m, M, 0, 64, true);
  }
  // t = f^(p^2 + 1); the easy factor p^6 - 1 is replaced by the conjugation test at the end
  g_frob2(g, a, f);
  gf_mul(g, ft, tt, a, f);
  gf_pow_u(g, ft, fu, tt);
  gf_pow_u(g, ft, fu2, fu);
  gf_pow_u(g, ft, fu3, fu2);
  // numerator: y0 = t^p t^(p^2) t^(p^3), y2 = (t^(u^2))^(p^2);  N = y0 y2^6
  g_frob(g, a, tt);
  g_frob2(g, b, tt);
  gf_mul(g, ft, y0, a, b);
  g_frob(g, a, b);
  gf_mul(g, ft, y0, y0, a);
  g_frob2(g, a, fu2);
  gf_mul(g, ft, a, a, a);        // y2^2
  gf_mul(g, ft, b, a, a);        // y2^4
  gf_mul(g, ft, b, b, a);        // y2^6
  gf_mul(g, ft, N, y0, b);
  // denominator (positive powers of the inverted terms): Y1 = t, Y3 = (t^u)^p, Y4 = t^u (t^(u^2))^p, Y5 = t^(u^2),
  // Y6 = t^(u^3) (t^(u^3))^p;  D = Y1^2 Y3^12 Y4^18 Y5^30 Y6^36 by the Scott et al. vector chain
  g_frob(g, a, fu3);
  gf_mul(g, ft, a, fu3, a);      // Y6
  gf_mul(g, ft, T0, a, a);       // Y6^2
  g_frob(g, a, fu2);
  gf_mul(g, ft, a, fu, a);       // Y4
  gf_mul(g, ft, T0, T0, a);
  gf_mul(g, ft, T0, T0, fu2);    // T0 = Y6^2 Y4 Y5
  g_frob(g, a, fu);         // Y3
  gf_mul(g, ft, T1, a, fu2);
  gf_mul(g, ft, T1, T1, T0);     // T1 = Y3 Y5 T0
  gf_mul(g, ft, T1, T1, T1);
  gf_mul(g, ft, T1, T1, T0);
  gf_mul(g, ft, T1, T1, T1);     // T1 = (T1^2 T0)^2
  gf_mul(g, ft, T0, T1, tt);     // T0 = T1 Y1
  gf_mul(g, ft, T0, T0, T0);
  gf_mul(g, ft, T0, T0, T1);     // D = T0^2 T1
  // accept  <=>  conj(N) D == N conj(D)
  g_conj(g, a, N);
  gf_mul(g, ft, a, a, T0);
  g_conj(g, b, T0);
  gf_mul(g, ft, b, N, b);
  bool ok = true;
  if (t < E12_NB) ok = a->e[e12_base_slot(t)] == b->e[e12_base_slot(t)];
  ok = __syncthreads_and(ok);
  if (t == 0) *verdict = ok ? 1u : 0u;
}
#endif  // __CUDACC__

}  // namespace h2v
