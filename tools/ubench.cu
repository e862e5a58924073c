// SIMT pipe micro-benchmarks on B200: the rollout kernel is bound by issue slots of the FMA / ALU / MUFU / LSU
// pipes (DESIGN.md "SIMT roof"), so the instruction-level roofs are measured rather than assumed.
// Prints lane-operations per clock per SM for each instruction class.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 2048;
constexpr int ILP = 16;

#define BENCH_KERNEL(name, decl, body, sink)                                          \
  __global__ void __launch_bounds__(1024) name(float* out, float seed) {            \
    decl;                                                                             \
    for (int it = 0; it < ITERS; ++it) {                                              \
      _Pragma("unroll") for (int i = 0; i < ILP; ++i) { body; }                       \
    }                                                                                 \
    float s = 0.f;                                                                    \
    _Pragma("unroll") for (int i = 0; i < ILP; ++i) s += sink;                        \
    if (s == 12345.678f) out[threadIdx.x] = s;                                        \
  }

BENCH_KERNEL(k_ffma, float a[ILP]; float b = seed; float c = seed * 0.5f; for (int i = 0; i < ILP; ++i) a[i] = seed + i,
             a[i] = fmaf(a[i], b, c), a[i])
BENCH_KERNEL(k_fmul_fadd, float a[ILP]; float b = seed; float c = seed * 0.5f; for (int i = 0; i < ILP; ++i) a[i] = seed + i,
             a[i] = (a[i] - c) * b, a[i])

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__global__ void __launch_bounds__(1024) k_ffma2(float* out, float seed) {
  uint64_t a[ILP];
  float2 bb = make_float2(seed, seed * 1.1f), cc = make_float2(seed * 0.5f, seed * 0.3f);
  uint64_t b = *reinterpret_cast<uint64_t*>(&bb), c = *reinterpret_cast<uint64_t*>(&cc);
  for (int i = 0; i < ILP; ++i) {
    float2 t = make_float2(seed + i, seed - i);
    a[i] = *reinterpret_cast<uint64_t*>(&t);
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma2(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float2 t = *reinterpret_cast<float2*>(&a[i]);
    s += t.x + t.y;
  }
  if (s == 12345.678f) out[threadIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_ex2(float* out, float seed) {
  float a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = seed + i * 0.01f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 12345.678f) out[threadIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_imad_hi(float* out, float seed) {
  uint32_t a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = (uint32_t)(seed * 1000.f) + i * 77u;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = __umulhi(a[i], 0xD2511F53u) ^ (a[i] * 0xCD9E8D57u);
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 0x12345u) out[threadIdx.x] = (float)s;
}

__global__ void __launch_bounds__(1024) k_lds128(float* out, float seed) {
  __shared__ float4 buf[256];
  if (threadIdx.x < 256) buf[threadIdx.x] = make_float4(seed, seed, seed, seed);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = (int)seed & 255;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      float4 v;  // warp-uniform address: broadcast
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"((unsigned)__cvta_generic_to_shared(&buf[(idx + i) & 255])));
      acc.x += v.x;
      idx = (idx + 1) & 255;
    }
  }
  if (acc.x == 12345.678f) out[threadIdx.x] = acc.x + acc.y;
}

template <typename F>
void run(const char* name, F kernel, double lane_ops_per_thread_iter, int sms, double clock_hz) {
  float* out;
  cudaMalloc(&out, 4096);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    kernel<<<sms, 1024>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * 1024 * ITERS * ILP * lane_ops_per_thread_iter;
  printf("%-12s %8.3f ms  %7.1f lane-instr/clk/SM (at %.0f MHz)  err=%s\n", name, ms,
         ops / (ms * 1e-3) / sms / clock_hz, clock_hz / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double hz = khz * 1e3;
  printf("%s: %d SMs, clock attr %.0f MHz\n", p.name, sms, hz / 1e6);
  run("FFMA", k_ffma, 1, sms, hz);
  run("FSUB+FMUL", k_fmul_fadd, 2, sms, hz);
  run("FFMA2", k_ffma2, 1, sms, hz);
  run("MUFU.EX2", k_ex2, 1, sms, hz);
  run("IMAD.HI+LO", k_imad_hi, 3, sms, hz);
  run("LDS.128 bc", k_lds128, 1, sms, hz);
  return 0;
}
