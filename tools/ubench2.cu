// Low-occupancy issue behaviour of FFMA vs FFMA2 (2 warps per scheduler, like the tf32x3 rollout):
// dependent chains of length ILP per warp.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr int ITERS = 4096;
template <int ILP> __global__ void k_ffma(float* out, float seed) {
  float a[ILP]; float b = seed, c = seed * 0.5f;
  for (int i = 0; i < ILP; ++i) a[i] = seed + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 12345.678f) out[threadIdx.x] = s;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int ILP> __global__ void k_ffma2(float* out, float seed) {
  uint64_t a[ILP]; float2 bb = make_float2(seed, seed * 1.1f), cc = make_float2(seed * 0.5f, seed * 0.3f);
  uint64_t b = *reinterpret_cast<uint64_t*>(&bb), c = *reinterpret_cast<uint64_t*>(&cc);
  for (int i = 0; i < ILP; ++i) { float2 t = make_float2(seed + i, seed - i); a[i] = *reinterpret_cast<uint64_t*>(&t); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma2(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&a[i]); s += t.x + t.y; }
  if (s == 12345.678f) out[threadIdx.x] = s;
}
template <typename F> void run(const char* name, F k, int threads, int ilp, int flops_per, int sms, double hz) {
  float* out; cudaMalloc(&out, 4096); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); k<<<sms, threads>>>(out, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double)sms * threads * ITERS * ilp;
  printf("%-8s warps/SM %2d ILP %2d: %7.1f lane-instr/clk/SM  %7.1f lane-FMA/clk/SM\n", name, threads / 32, ilp, inst / (ms * 1e-3) / sms / hz,
         inst * flops_per / (ms * 1e-3) / sms / hz);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); double hz = khz * 1e3; int sms = p.multiProcessorCount;
  for (int threads : {128, 256, 448, 1024}) {
    run("FFMA", k_ffma<4>, threads, 4, 1, sms, hz); run("FFMA", k_ffma<8>, threads, 8, 1, sms, hz); run("FFMA", k_ffma<16>, threads, 16, 1, sms, hz);
    run("FFMA2", k_ffma2<4>, threads, 4, 2, sms, hz); run("FFMA2", k_ffma2<8>, threads, 8, 2, sms, hz); run("FFMA2", k_ffma2<16>, threads, 16, 2, sms, hz);
  }
  return 0;
}
