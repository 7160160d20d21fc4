// Microbenchmark: hand-over latency between consecutive kernels of a stream when N streams are busy.
// Every stream runs a chain of K kernels (B blocks x T threads, each block busy for D microseconds either spinning on
// the timer or running dependent IMADs); block 0 logs start / end on the global timer.  Reports the mean gap between the end of
// kernel k and the start of kernel k+1 of the same stream, and the chain throughput, for plain launches, graph
// replay and programmatic dependent launch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/stream_gap.cu -o tools/stream_gap.x
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
__device__ __forceinline__ u64 gtime() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int WORK>  // 0 = spin on the timer, 1 = IMAD chain
__global__ void k_busy(u64 ns, u64* log, unsigned* sink) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const u64 t0 = gtime();
  if (WORK == 0) {
    while (gtime() - t0 < ns) {}
  } else {
    unsigned a = threadIdx.x, b = blockIdx.x | 1;
    while (gtime() - t0 < ns) {
#pragma unroll
      for (int i = 0; i < 256; i++) a = a * b + i;
    }
    if (a == 0x12345) *sink = a;
  }
  if (threadIdx.x == 0) {
    atomicMin(&log[0], t0);
    atomicMax(&log[1], gtime());
  }
}
struct Cfg { int streams, K, blocks, threads; double us; int work, mode; };  // mode 0 plain, 1 graph, 2 graph + PDL, 3 plain + PDL
static void launch(const Cfg& c, cudaStream_t s, u64* log, unsigned* sink, bool pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = c.blocks; cfg.blockDim = c.threads; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  u64 ns = (u64)(c.us * 1000);
  if (c.work) cudaLaunchKernelEx(&cfg, k_busy<1>, ns, log, sink); else cudaLaunchKernelEx(&cfg, k_busy<0>, ns, log, sink);
}
static void run(const Cfg& c) {
  const int reps = 6;
  std::vector<cudaStream_t> st(c.streams);
  for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  u64* log; unsigned* sink;
  const size_t nlog = (size_t)c.streams * c.K * reps * 2;
  cudaMalloc(&log, 8 * nlog); cudaMalloc(&sink, 4);
  std::vector<u64> h(nlog);
  for (size_t i = 0; i < nlog; i += 2) { h[i] = ~0ull; h[i + 1] = 0; }
  cudaMemcpy(log, h.data(), 8 * nlog, cudaMemcpyHostToDevice);
  std::vector<cudaGraphExec_t> ge(c.streams * reps, nullptr);
  const bool graph = c.mode == 1 || c.mode == 2, pdl = c.mode >= 2;
  if (graph)
    for (int s = 0; s < c.streams; s++)
      for (int r = 0; r < reps; r++) {
        cudaGraph_t g;
        cudaStreamBeginCapture(st[s], cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < c.K; k++) launch(c, st[s], log + 2 * (((size_t)s * reps + r) * c.K + k), sink, pdl);
        cudaStreamEndCapture(st[s], &g);
        cudaGraphInstantiate(&ge[s * reps + r], g, 0);
        cudaGraphDestroy(g);
      }
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int r = 0; r < reps; r++)
    for (int s = 0; s < c.streams; s++) {
      if (graph) cudaGraphLaunch(ge[s * reps + r], st[s]);
      else for (int k = 0; k < c.K; k++) launch(c, st[s], log + 2 * (((size_t)s * reps + r) * c.K + k), sink, pdl);
    }
  cudaDeviceSynchronize();
  cudaEventRecord(e1, 0); cudaEventSynchronize(e1);
  cudaMemcpy(h.data(), log, 8 * nlog, cudaMemcpyDeviceToHost);
  std::vector<double> gaps, durs;
  u64 tmin = ~0ull, tmax = 0;
  for (int s = 0; s < c.streams; s++)
    for (int r = 1; r < reps; r++)  // first rep = warm-up
      for (int k = 0; k < c.K; k++) {
        const size_t i = 2 * (((size_t)s * reps + r) * c.K + k);
        durs.push_back((h[i + 1] - h[i]) * 1e-3);
        tmin = std::min(tmin, h[i]); tmax = std::max(tmax, h[i + 1]);
        if (k) gaps.push_back(((double)h[i] - (double)h[i - 1]) * 1e-3);
      }
  std::sort(gaps.begin(), gaps.end());
  double gm = 0; for (double g : gaps) gm += g; gm /= gaps.size();
  double dm = 0; for (double d : durs) dm += d; dm /= durs.size();
  const double span = (tmax - tmin) * 1e-3, kernels = (double)c.streams * (reps - 1) * c.K;
  printf("streams %2d K %2d blocks %4d x %3d  D %5.0f us %s %-10s | kernel dur %7.1f us | gap mean %7.1f p50 %7.1f p90 %7.1f us | %7.1f us per kernel overall (ideal if packed: %.1f)\n",
         c.streams, c.K, c.blocks, c.threads, c.us, c.work ? "imad" : "spin", c.mode == 0 ? "plain" : c.mode == 1 ? "graph" : c.mode == 2 ? "graph+pdl" : "plain+pdl",
         dm, gm, gaps[gaps.size() / 2], gaps[gaps.size() * 9 / 10], span / kernels, c.us * std::max(1.0, (double)c.blocks * c.threads / (148.0 * 2048)) / 1.0);
  for (auto& g : ge) if (g) cudaGraphExecDestroy(g);
  for (auto& s : st) cudaStreamDestroy(s);
  cudaFree(log); cudaFree(sink);
}
int main(int argc, char** argv) {
  const int Ns[] = {1, 4, 16, 32};
  for (int work = 0; work < 2; work++)
    for (double us : {20.0, 300.0})
      for (int blocks : {64, 384})
        for (int n : Ns)
          for (int mode : {0, 1, 2}) run(Cfg{n, 12, blocks, 128, us, work, mode});
  printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
