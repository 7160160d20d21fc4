// Micro-benchmarks that calibrate the integer roofline of the field kernels on the box:
//   * issue rate of IMAD (32-bit multiply-add) and IMAD.WIDE.U32 (32x32+64) per SM
//   * throughput / dependent-chain latency of the two Montgomery multipliers in field.cuh
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I halo2-verifier_b200/csrc tools/microbench.cu -o tools/microbench.x
#include <cuda_runtime.h>
#include <stdio.h>
#include "field.cuh"
#include "pairing_cta.cuh"
using namespace h2v;

__global__ void k_imad_lo(u32 iters, u32* out) {
  u32 a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  const u32 m = blockIdx.x * 2654435761u + 12345u, k = threadIdx.x | 1u;
  for (u32 i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] = a[j] * m + k;
  }
  u32 x = 0;
  for (int i = 0; i < 8; i++) x ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_imad_wide(u32 iters, u32* out) {
  u64 a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  const u32 m = blockIdx.x * 2654435761u + 12345u;
  for (u32 i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int j = 0; j < 8; j++) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[j]) : "r"((u32)a[j]), "r"(m));
  }
  u64 x = 0;
  for (int i = 0; i < 8; i++) x ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)x ^ (u32)(x >> 32);
}
// chain of dependent Montgomery products; CH independent chains per thread
template <int MODE, int CH>
__global__ void k_mm(u32 iters, Fq* io) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x[CH], y = io[t];
  for (int c = 0; c < CH; c++) { x[c] = y; x[c].l[0] ^= c; x[c].l[7] &= 0x0FFFFFFF; }
  y.l[7] &= 0x0FFFFFFF;
  for (u32 i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = MODE == 0 ? Fq::mul_any(x[c], y) : MODE == 2 ? x[c].sqr() : Fq::mul(x[c], y);
  }
  Fq r = x[0];
  for (int c = 1; c < CH; c++) r = r + x[c];
  io[t] = r;
}
__global__ void k_check(u32 count, u32* bad) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  u64 s = 0x9E3779B97F4A7C15ull * (t + 1);
  Fq a, b;
  for (int i = 0; i < 8; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull; a.l[i] = (u32)(s >> 32);
    s = s * 6364136223846793005ull + 1442695040888963407ull; b.l[i] = (u32)(s >> 32);
  }
  a.l[7] &= 0x1FFFFFFF; b.l[7] &= 0x1FFFFFFF;
  if (t % 5 == 0) { a = Fq::zero() - Fq::one(); }
  if (t % 7 == 0) { b = Fq::zero() - Fq::one(); }
  if (t % 11 == 0) { for (int i = 0; i < 8; i++) a.l[i] = FqP::mod(i); a.l[0] -= 1; }
  if (t % 13 == 0) { for (int i = 0; i < 8; i++) b.l[i] = FqP::mod(i); b.l[0] -= 1; }
  Fq x = Fq::mul(a, b), y = Fq::mul_portable(a, b), z = Fq::mul_any(a, b);
  Fr fa, fb;
  for (int i = 0; i < 8; i++) { fa.l[i] = a.l[i]; fb.l[i] = b.l[i]; }
  fa.l[7] &= 0x0FFFFFFF; fb.l[7] &= 0x0FFFFFFF;
  Fr fx = Fr::mul(fa, fb), fy = Fr::mul_portable(fa, fb);
  if (x != y || z != y || fx != fy) atomicAdd(bad, 1u);
  if (a.sqr() != Fq::mul_portable(a, a) || b.sqr() != Fq::mul_portable(b, b) || fa.sqr() != Fr::mul_portable(fa, fa)) atomicAdd(bad, 1u);
}

// latency of the cooperative Fq12 product: one 64-thread group, `iters` dependent g_mul
__global__ void __launch_bounds__(128) k_e12_chain(u32 iters, Fq* io, int mode) {
  __shared__ LinTables lt;
  __shared__ E12 x, y;
  __shared__ Fq scr[2 * E12_N];
  __shared__ u64 part[E12_N * 8];
  const int t = threadIdx.x;
  for (int i = t; i < (int)(sizeof(LinTables) / 2); i += blockDim.x) ((uint16_t*)&lt)[i] = ((const uint16_t*)&g_lin_tables)[i];
  if (t < E12_NB) { Fq v = io[t]; v.l[7] &= 0x0FFFFFFF; scr[t] = v; }
  __syncthreads();
  Grp g{t, 1, &lt, scr, (int)blockDim.x, part};
  g_expand(g, &x);
  g_copy(g, &y, &x);
  for (u32 i = 0; i < iters; i++) {
    if (mode == 0) g_mul(g, &x, &x, &y);
    else if (mode == 1) { if (t < E12_N) e12_mul_p1(scr, &x, &y, t); g.sync(); if (t < E12_N) x.e[t] = scr[t]; g.sync(); }
    else { if (t < E12_N) e12_mul_p2(&x, scr, &lt, t); g.sync(); }
  }
  if (t < E12_N) io[64 + t] = x.e[t];
}

template <class F>
static double time_ms(F launch, int reps = 3) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  u32* d; cudaMalloc(&d, 4u << 22);
  u32* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  k_check<<<4096, 256>>>(1u << 20, bad);
  u32 hb = 0; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("mul / sqr vs mul_portable vs mul_any mismatches over 2^20 inputs (Fq and Fr): %u  [%s]\n", hb, cudaGetErrorString(cudaGetLastError()));
  {
    const u32 blocks = sms * 8, threads = 256, iters = 2048;
    double ms = time_ms([&] { k_imad_lo<<<blocks, threads>>>(iters, d); });
    printf("IMAD           : %.3f T/s  (%.1f per clk per SM at 1.965 GHz)\n", (double)blocks * threads * iters * 64 / ms / 1e9, (double)blocks * threads * iters * 64 / (ms * 1e-3) / sms / 1.965e9);
    ms = time_ms([&] { k_imad_wide<<<blocks, threads>>>(iters, d); });
    printf("IMAD.WIDE.U32  : %.3f T/s  (%.1f per clk per SM at 1.965 GHz)\n", (double)blocks * threads * iters * 64 / ms / 1e9, (double)blocks * threads * iters * 64 / (ms * 1e-3) / sms / 1.965e9);
  }
  Fq* io; cudaMalloc(&io, sizeof(Fq) << 20);
  cudaMemset(io, 0x5a, sizeof(Fq) << 20);
  const u32 iters = 2048;
  struct Cfg { int blocks_per_sm, threads; };
  const Cfg cfgs[] = {{1, 32}, {1, 128}, {2, 128}, {4, 128}, {4, 256}, {8, 256}};
  for (auto c : cfgs) {
    const u32 blocks = sms * c.blocks_per_sm;
    double m0 = time_ms([&] { k_mm<0, 1><<<blocks, c.threads>>>(iters, io); });
    double m1 = time_ms([&] { k_mm<1, 1><<<blocks, c.threads>>>(iters, io); });
    double m2 = time_ms([&] { k_mm<1, 2><<<blocks, c.threads>>>(iters, io); });
    double m3 = time_ms([&] { k_mm<1, 4><<<blocks, c.threads>>>(iters, io); });
    double m4 = time_ms([&] { k_mm<2, 1><<<blocks, c.threads>>>(iters, io); });
    const double n = (double)blocks * c.threads * iters;
    printf("warps/SM %2d: lo/hi CIOS %.2f G MM/s (chain %.0f ns/MM) | wide %.2f G MM/s (chain %.0f ns/MM) | wide x2 chains %.2f | wide x4 chains %.2f G MM/s | sqr %.2f G/s (chain %.0f ns)\n",
           c.blocks_per_sm * c.threads / 32, n / m0 / 1e6, m0 * 1e6 / iters, n / m1 / 1e6, m1 * 1e6 / iters, 2 * n / m2 / 1e6, 4 * n / m3 / 1e6, n / m4 / 1e6, m4 * 1e6 / iters);
  }
  for (int nthr = 64; nthr <= 128; nthr += 64)
    for (int mode = 0; mode < 3; mode++) {
      if (nthr == 128 && mode == 2) continue;
      double ms = time_ms([&] { k_e12_chain<<<1, nthr>>>(1000, io, mode); });
      printf("cooperative Fq12 engine, %d threads, %s: %.2f us per op\n", nthr, mode == 0 ? "g_mul" : mode == 1 ? "product phase only" : "linear phase only", ms);
    }
  printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
