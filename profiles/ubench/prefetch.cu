// Micro-benchmark: does prefetch.global.L1 (CCTL.E.PF1) / .L2 (PF2) shorten a later ld.global.nc on sm_100a?
// One warp per SM; each lane touches a random 128-byte line of a 1 GiB buffer; latency in SM clocks.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const unsigned short* buf, size_t nlines, int mode, int wait_clk, long long* out, unsigned* sink) {
  unsigned s = (blockIdx.x * 32 + threadIdx.x) * 2654435761u + 12345u + 7919u * (unsigned)(mode * 16 + wait_clk);
  long long tot = 0; unsigned acc = 0;
  const int iters = 64;
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const size_t line = ((size_t)s * 2654435761ull >> 7) % nlines;
    const unsigned short* p = buf + line * 64;
    if (mode == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    if (mode == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    if (mode == 3) { unsigned short v; asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p)); acc += v; }  // real load first (L1 allocate)
    long long t0 = clock64();
    while (clock64() - t0 < wait_clk) {}
    __syncwarp();
    long long a = clock64();
    unsigned short v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p + 1));
    acc += v;
    asm volatile("" ::"r"(acc));
    long long b = clock64();
    tot += b - a;
  }
  if (threadIdx.x == 0) out[blockIdx.x] = tot / iters;
  if (acc == 0x7fffffff) *sink = acc;
}
int main() {
  const size_t bytes = 8ull << 30; unsigned short* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes); cudaDeviceSynchronize();
  // warm the TLB-independent part: nothing; every (mode, wait) pair uses its own random lines, so loads come from DRAM unless a prefetch worked
  long long* out; cudaMallocManaged(&out, 148 * 8); unsigned* sink; cudaMalloc(&sink, 4);
  const char* names[] = {"no prefetch", "prefetch.global.L1", "prefetch.global.L2", "real load first"};
  for (int wait : {0, 1000, 3000}) for (int mode = 0; mode < 4; ++mode) {
    k<<<148, 32>>>(buf, bytes / 128, mode, wait, out, sink); cudaDeviceSynchronize();
    long long s = 0; for (int i = 0; i < 148; ++i) s += out[i];
    printf("wait %4d clk  %-20s load latency %lld clk\n", wait, names[mode], s / 148);
  }
  return 0;
}
