// Per-block timeline (diagnostics and bench timing; off unless h2v_debug_timeline_start was called): thread 0 of
// every block of an instrumented kernel appends {kernel id, block, SM, context tag, start, end} on the global
// nanosecond timer.  Kernel ids: 1 decompress, 2 transcript, 3 scalar, 4 digits, 5 scatter, 6 bucket_sum,
// 7 chunk_reduce, 8 window_reduce, 9 lines, 10 pairing_check, 11 window_combine, 12 pp_mul_list, 13 pp_reduce_list, 14 rlc_scan,
// 15 shared_reduce, 16 bucket_order.
#pragma once
#include "field.cuh"

#if defined(__CUDACC__)
using h2v::u32;
using h2v::u64;
struct TlRec {
  u32 kid, block, smid, tag;
  u64 t0, t1;
};
__device__ TlRec* g_tl_buf = nullptr;
__device__ u32 g_tl_cap = 0;
__device__ u32 g_tl_count = 0;
struct TlScope {
  u64 t0;
  u32 kid, tag;
  bool on;
  __device__ __forceinline__ TlScope(u32 kid_, const void* tagp) {
    on = threadIdx.x == 0 && g_tl_buf != nullptr;
    if (on) {
      kid = kid_;
      tag = (u32)((size_t)tagp >> 8);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    }
  }
  __device__ __forceinline__ ~TlScope() {
    if (on) {
      u64 t1;
      u32 sm;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      const u32 i = atomicAdd(&g_tl_count, 1u);
      if (i < g_tl_cap) g_tl_buf[i] = TlRec{kid, blockIdx.x, sm, tag, t0, t1};
    }
  }
};

#endif
