// Micro-benchmark: issue rate and dependent latency of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a,
// alone and mixed with ALU-pipe instructions (FMNMX), to decide whether K1's projection can be packed two points per lane.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE, int CH>
__global__ void k(float* out, float s, int iters) {
  float acc[CH]; u64 acc2[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i] = threadIdx.x * 1e-3f + i; acc2[i] = pk(acc[i], acc[i] + 1.f); }
  const u64 s2 = pk(s, s);
  float mn = 1e30f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[i]) : "f"(s));
        if (MODE == 1) acc2[i] = fma2(acc2[i], s2, s2);
        if (MODE == 2) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[i]) : "f"(s)); asm volatile("min.f32 %0, %0, %1;" : "+f"(mn) : "f"(acc[i])); }
        if (MODE == 3) { acc2[i] = fma2(acc2[i], s2, s2); float a, b; upk(acc2[i], a, b); asm volatile("min.f32 %0, %0, %1;" : "+f"(mn) : "f"(a)); }
      }
    }
  }
  float t = mn;
#pragma unroll
  for (int i = 0; i < CH; ++i) { float a, b; upk(acc2[i], a, b); t += acc[i] + a + b; }
  if (t == 123.456f) out[0] = t;
}
template <int MODE, int CH>
void run(const char* name, int blocks, int threads, int iters) {
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, CH><<<blocks, threads>>>(d, 0.999f, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE, CH><<<blocks, threads>>>(d, 0.999f, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double inst = (double)blocks * (threads / 32) * iters * 8.0 * CH * ((MODE >= 2) ? 2 : 1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-34s blocks=%d thr=%d  %.3f ms  warp-inst/clk/SM=%.3f (at %d MHz nominal)\n", name, blocks, threads, ms, inst / cyc / 148.0, clk / 1000);
  cudaFree(d);
}
int main() {
  const int it = 4096;
  run<0, 8>("FFMA  x8 chains, 32 warps/SM", 148 * 4, 256, it);
  run<1, 8>("FFMA2 x8 chains, 32 warps/SM", 148 * 4, 256, it);
  run<0, 1>("FFMA  x1 chain,  32 warps/SM", 148 * 4, 256, it);
  run<1, 1>("FFMA2 x1 chain,  32 warps/SM", 148 * 4, 256, it);
  run<0, 2>("FFMA  x2 chains, 32 warps/SM", 148 * 4, 256, it);
  run<1, 2>("FFMA2 x2 chains, 32 warps/SM", 148 * 4, 256, it);
  run<2, 8>("FFMA+FMNMX x8,   32 warps/SM", 148 * 4, 256, it);
  run<3, 8>("FFMA2+FMNMX x8,  32 warps/SM", 148 * 4, 256, it);
  run<0, 1>("FFMA  x1 chain, 1 warp/SM (latency)", 148, 32, it);
  run<1, 1>("FFMA2 x1 chain, 1 warp/SM (latency)", 148, 32, it);
  return 0;
}
