// Fiat-Shamir replay with FOUR lanes per proof (device only; Blake2b transcripts).
//
// The thread-per-proof replay (stages.cuh: transcript_stage, kept for Keccak transcripts and for the host build) spends its
// time in ~23 dependent Blake2b compressions per proof (VM shape; 136 for the k = 18 shape) on a handful of warps.  Here a
// QUAD of lanes shares one compression: lane q holds column q of the 4 x 4 state matrix, the four column G functions and
// the four diagonal G functions of a round run in parallel, the diagonalisation is a rotation of the rows between the lanes
// (warp shuffles).  The 128-byte block buffer of a proof lives in shared memory; the lanes write 8 bytes each of every
// 32-byte item.  The byte stream is the reference's (transcript/mod.rs:16-39, 118-134, 205-232): personalised Blake2b-512,
// prefix 0x02 | scalar, 0x01 | x | y of the affine point, 0x00 then a digest of a CLONE for every challenge.
//
// Inputs prepared by wider kernels: the canonical point coordinates come out of the decompression (k_decompress);
// the Montgomery conversion of the proof scalars (one multiplication each) is spread over the lanes after the hashing.
#pragma once
#include "stages.cuh"

namespace h2v {
#if defined(__CUDACC__)

static constexpr int TQ_PROOFS_PER_BLOCK = 32;  // 128 threads
static constexpr int TQ_BUF_STRIDE = 144;       // bytes between the 128-byte block buffers of neighbouring proofs: the eight quads of a warp then hit
                                                // disjoint banks (a 128-byte stride put all of them on the same banks: 88 M conflict cycles per launch set, ncu)
__constant__ u64 c_tq_iv[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                               0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
__constant__ u8 c_tq_sigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

struct TqState {
  u64 a, b;      // h[q], h[q + 4]
  u64 t;         // bytes compressed so far
  u32 pos;       // bytes in the block buffer
  u32 qmask;     // the four lanes of this quad
  u32 q;         // lane inside the quad
  u32 lane0;     // first lane of the quad inside the warp
  u8* buf;       // 128-byte block buffer of this proof (shared memory)
  u32 sig[6];    // message word indices of this lane: 12 rounds x 4 nibbles (column x, y; diagonal x, y)
};

// lane / quad bookkeeping and this lane's message-word schedule (4 nibbles per round: column x, y; diagonal x, y)
__device__ __forceinline__ void tq_setup(TqState& st, u8* buf) {
  const u32 lane = threadIdx.x & 31;
  st.q = lane & 3;
  st.lane0 = lane & ~3u;
  st.qmask = 0xFu << st.lane0;
  st.buf = buf;
#pragma unroll
  for (int r = 0; r < 12; r++) {
    const u32 s4 = (u32)c_tq_sigma[r][2 * st.q] | ((u32)c_tq_sigma[r][2 * st.q + 1] << 4) | ((u32)c_tq_sigma[r][8 + 2 * st.q] << 8) |
                   ((u32)c_tq_sigma[r][9 + 2 * st.q] << 12);
    if (r & 1) st.sig[r >> 1] |= s4 << 16;
    else st.sig[r >> 1] = s4;
  }
}

__device__ __forceinline__ u64 tq_shfl(const TqState& st, u64 v, u32 src_q) { return __shfl_sync(st.qmask, v, (int)(st.lane0 + src_q)); }

__device__ __forceinline__ void tq_g(u64& a, u64& b, u64& c, u64& d, u64 x, u64 y) {
  a = a + b + x;
  d = rotr64(d ^ a, 32);
  c = c + d;
  b = rotr64(b ^ c, 24);
  a = a + b + y;
  d = rotr64(d ^ a, 16);
  c = c + d;
  b = rotr64(b ^ c, 63);
}

// one compression of the block buffer into (a, b); `last`: final block of a digest.
// NOT inlined: the replay reaches it from five places (byte / word absorption, flush, digest) and an inlined copy is ~1.2 k
// instructions; the kernel was 48 k instructions, ncu (profiles/r2b_narrow_summary.txt) showed 7 k of them hot and 19.5 % of the
// warp-stall samples on instruction fetch.  Arguments by value (registers), the buffer as a shared-memory address.
__device__ __noinline__ ulonglong2 tq_compress_ni(u64 ha, u64 hb, u64 t, u32 last, u32 sbuf, u32 s0, u32 s1, u32 s2, u32 s3, u32 s4_, u32 s5) {
  const u32 lane = threadIdx.x & 31, q = lane & 3, lane0 = lane & ~3u, qmask = 0xFu << lane0;
  const u32 sig[6] = {s0, s1, s2, s3, s4_, s5};
  u64 a = ha, b = hb, c = c_tq_iv[q], d = c_tq_iv[q + 4];
  // v[12] ^= t (low word), v[13] ^= 0 (high word of the counter), v[14] inverted on the last block
  d ^= q == 0 ? t : 0ull;
  if (last && q == 2) d = ~d;
  auto ldm = [sbuf](u32 w) {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(sbuf + 8 * w));
    return v;
  };
#pragma unroll
  for (int r = 0; r < 12; r++) {
    const u32 s4 = (sig[r >> 1] >> ((r & 1) * 16)) & 0xFFFFu;
    tq_g(a, b, c, d, ldm(s4 & 15), ldm((s4 >> 4) & 15));
    // diagonalise: lane q takes b from q + 1, c from q + 2, d from q + 3
    b = __shfl_sync(qmask, b, (int)(lane0 + ((q + 1) & 3)));
    c = __shfl_sync(qmask, c, (int)(lane0 + ((q + 2) & 3)));
    d = __shfl_sync(qmask, d, (int)(lane0 + ((q + 3) & 3)));
    tq_g(a, b, c, d, ldm((s4 >> 8) & 15), ldm((s4 >> 12) & 15));
    b = __shfl_sync(qmask, b, (int)(lane0 + ((q + 3) & 3)));
    c = __shfl_sync(qmask, c, (int)(lane0 + ((q + 2) & 3)));
    d = __shfl_sync(qmask, d, (int)(lane0 + ((q + 1) & 3)));
  }
  return make_ulonglong2(ha ^ a ^ c, hb ^ b ^ d);
}
__device__ __forceinline__ void tq_compress(const TqState& st, u64& ha, u64& hb, u64 t, bool last) {
  const ulonglong2 r = tq_compress_ni(ha, hb, t, last ? 1u : 0u, (u32)__cvta_generic_to_shared(st.buf), st.sig[0], st.sig[1], st.sig[2], st.sig[3], st.sig[4], st.sig[5]);
  ha = r.x;
  hb = r.y;
}

__device__ __forceinline__ void tq_zero_buf(const TqState& st) {
  uint4* p = (uint4*)(st.buf + 32 * st.q);
  p[0] = make_uint4(0, 0, 0, 0);
  p[1] = make_uint4(0, 0, 0, 0);
  __syncwarp(st.qmask);
}
// the buffer is full and more input follows: compress it, start a new block
__device__ __forceinline__ void tq_flush(TqState& st) {
  __syncwarp(st.qmask);
  st.t += 128;
  tq_compress(st, st.a, st.b, st.t, false);
  __syncwarp(st.qmask);
  tq_zero_buf(st);
  st.pos = 0;
}
__device__ __forceinline__ void tq_byte(TqState& st, u8 v) {
  if (st.pos == 128) tq_flush(st);
  if (st.q == 0) st.buf[st.pos] = v;
  st.pos++;
}
// 32 bytes, lane q supplies bytes [8 q, 8 q + 8) as a little-endian u64
__device__ __forceinline__ void tq_word32(TqState& st, u64 w) {
  if (st.pos == 128) tq_flush(st);
  const u32 pos = st.pos, first = 128 - pos;  // bytes that still fit
  const u32 o = 8 * st.q;
#pragma unroll
  for (u32 i = 0; i < 8; i++)
    if (o + i < first) st.buf[pos + o + i] = (u8)(w >> (8 * i));
  if (first >= 32) {
    st.pos = pos + 32;
    return;
  }
  st.pos = 128;
  tq_flush(st);
#pragma unroll
  for (u32 i = 0; i < 8; i++)
    if (o + i >= first) st.buf[o + i - first] = (u8)(w >> (8 * i));
  st.pos = 32 - first;
}
// digest of a clone (the running state continues): lane q returns h[q] in lo, h[q + 4] in hi
__device__ __forceinline__ void tq_digest(const TqState& st, u64& lo, u64& hi) {
  __syncwarp(st.qmask);
  lo = st.a;
  hi = st.b;
  tq_compress(st, lo, hi, st.t + st.pos, true);
}

// 8 bytes at any alignment
__device__ __forceinline__ u64 tq_load8(const u8* p) {
  if (((size_t)p & 7) == 0) return *(const u64*)p;
  u64 v = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) v |= (u64)p[i] << (8 * i);
  return v;
}
// is the 256-bit little-endian value whose word q this lane holds >= the Fr modulus?  (quad-uniform result)
__device__ __forceinline__ bool tq_geq_r(const TqState& st, u64 w) {
  const u64 mq = (u64)FrP::mod(2 * (int)st.q) | ((u64)FrP::mod(2 * (int)st.q + 1) << 32);
  const int c = w > mq ? 1 : (w < mq ? -1 : 0);
  int r = 0;  // decided by the most significant differing word
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int cs = __shfl_sync(st.qmask, c, (int)(st.lane0 + s));
    if (cs != 0) r = cs;
  }
  return r >= 0;
}

// The whole replay of proof j by its quad.  Same contract as transcript_stage: Montgomery-form proof scalars and
// challenges into the value table, first failing item (merged with the decompression's) returned, inst_bad set.
__device__ __forceinline__ u32 transcript_quad(const PlanView& pv, const u8* proof, u32 len, const u8* inst, u32 inst_total, const u32* ptsc, Fr* vals, u32 j,
                                               u32 n, u32 bad_item, bool& inst_bad, TqState& st) {
  const PlanHeader& hd = pv.h();
  const TranscriptOp* ops = pv.sec<TranscriptOp>(hd.off_tops);
  // init (Blake2bRead::init, transcript/mod.rs:118-134): parameter block digest_length 64, fanout 1, depth 1, personal "Halo2-Transcript"
  st.a = c_tq_iv[st.q] ^ (st.q == 0 ? 0x01010040ull : 0ull);
  st.b = c_tq_iv[st.q + 4] ^ (st.q == 2 ? 0x72542d326f6c6148ull : st.q == 3 ? 0x7470697263736e61ull : 0ull);
  st.t = 0;
  st.pos = 0;
  tq_zero_buf(st);
  u32 item = 0, pslot = 0, cidx = 0;
  inst_bad = false;
  // ONE absorption site for every kind of item (prefix byte + one or two 32-byte words): with a site per kind the inlined
  // buffer handling was most of the kernel's then 7 k instructions and 28 % of its warp stalls were instruction fetches (ncu); 6.4 k now
  for (u32 o = 0; o < hd.n_tops; o++) {
    const u32 kind = ops[o].kind, count = ops[o].count;
    if (kind == T_SQUEEZE) {
      for (u32 i = 0; i < count; i++, cidx++) {
        tq_byte(st, 0);
        u64 lo, hi;
        tq_digest(st, lo, hi);
        // the 64-byte digest h[0..8) -> lane cidx & 3, which reduces it mod r (Challenge255::new, transcript/mod.rs:500-509)
        u32 d[16];
#pragma unroll
        for (int s = 0; s < 4; s++) {
          const u64 l = tq_shfl(st, lo, s), h = tq_shfl(st, hi, s);
          d[2 * s] = (u32)l;
          d[2 * s + 1] = (u32)(l >> 32);
          d[8 + 2 * s] = (u32)h;
          d[8 + 2 * s + 1] = (u32)(h >> 32);
        }
        if (st.q == (cidx & 3)) vals[(size_t)(hd.v_chal + cidx) * n + j] = Fr::from_uniform_words(d);
      }
      continue;
    }
    const u32 reps = kind == T_ABS_VK ? 1u : kind == T_ABS_INST ? inst_total : count;
    for (u32 i = 0; i < reps; i++) {
      u64 w = 0, w2 = 0;
      u32 prefix = 2, words = 1;
      if (kind == T_ABS_VK) {
        const Fr c = pv.cst(hd.c_vk_repr).to_canonical();
        w = (u64)c.l[2 * st.q] | ((u64)c.l[2 * st.q + 1] << 32);
      } else if (kind == T_ABS_INST) {
        w = tq_load8(inst + 32 * (size_t)i + 8 * st.q);
        if (tq_geq_r(st, w)) inst_bad = true;
      } else if (kind == T_POINTS) {
        const u64* c = (const u64*)(ptsc + 16 * ((size_t)pslot * n + j));  // canonical x | y (zeros when the point was rejected)
        prefix = 1;
        words = 2;
        w = c[st.q];
        w2 = c[4 + st.q];
        item++;
        pslot++;
      } else {  // T_SCALARS
        if ((item + 1) * 32 <= len) {
          w = tq_load8(proof + item * 32 + 8 * st.q);
          if (tq_geq_r(st, w)) {
            if (item < bad_item) bad_item = item;
            w = 0;
          }
        } else if (item < bad_item) {
          bad_item = item;
        }
        item++;
      }
      tq_byte(st, (u8)prefix);
      for (u32 k = 0; k < words; k++) tq_word32(st, k ? w2 : w);
    }
  }
  // Montgomery form of the proof scalars, spread over the lanes (values the hashing above replaced by zero stay zero)
  const u32* sc_item = pv.sec<u32>(hd.off_sc_item);
  for (u32 sslot = st.q; sslot < hd.n_scalars; sslot += 4) {
    const u32 it = sc_item[sslot];
    Fr c = Fr::zero();
    if ((it + 1) * 32 <= len) {
      c = Fr::load_le_fast(proof + it * 32);
      if (c.geq_mod()) c = Fr::zero();
    }
    vals[(size_t)sslot * n + j] = Fr::from_canonical(c);
  }
  return bad_item;
}

#endif
}  // namespace h2v
