// Probe of the tcgen05 conventions the tensor-core rollout relies on, checked against a host reference:
//   D[128 x N] (TMEM, fp32) = A[128 x K] (TMEM, written with tcgen05.st, row r <-> lane r) * B[N x K]^T (smem,
//   K-major, no swizzle, core matrix = 8 rows x 16 B, LBO = N*16 B between K chunks, SBO = 128 B between row groups)
// for kind::tf32 (K step 8) and kind::f16/bf16 (K step 16).  Exits non-zero on mismatch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../sde_sampler_lrds_b200/csrc/lrds_tc_ptx.cuh"

using namespace lrds::ptx;

constexpr int A_COL = 0, D_COL = 128, TMEM_COLS = 256;

template <bool BF16>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Bw,
                                                    float* __restrict__ D, int K, int N) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  constexpr int E = BF16 ? 8 : 4;  // elements per 16-byte K chunk
  for (int idx = tid; idx < N * K; idx += 128) {
    const int n = idx / K, k = idx % K, kc = k / E, e = k % E;
    uint8_t* dst = smem + (size_t)kc * N * 16 + n * 16;
    if (BF16) reinterpret_cast<__nv_bfloat16*>(dst)[e] = __float2bfloat16(Bw[idx]);
    else reinterpret_cast<float*>(dst)[e] = Bw[idx];
  }
  fence_proxy_async();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const float* arow = A + (size_t)tid * K;
  if (BF16) {
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 v = __floats2bfloat162_rn(arow[2 * (c0 + i)], arow[2 * (c0 + i) + 1]);
        r[i] = *reinterpret_cast<const uint32_t*>(&v);
      }
      tmem_st8(tmem + lane_base + A_COL + c0, r);
    }
  } else {
    for (int c0 = 0; c0 < K; c0 += 8) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(arow[c0 + i]);
      tmem_st8(tmem + lane_base + A_COL + c0, r);
    }
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t sbase = smem_u32(smem);
    const int ksteps = K / (BF16 ? 16 : 8);
    const uint32_t idesc = BF16 ? make_idesc_bf16(128, N) : make_idesc_tf32(128, N);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t bdesc = make_smem_desc(sbase + (uint32_t)ks * 2u * N * 16u, (uint32_t)N * 16u, 128u);
      const uint32_t a_addr = tmem + A_COL + ks * 8;
      if (BF16) mma_bf16_ts(tmem + D_COL, a_addr, bdesc, idesc, ks > 0);
      else mma_tf32_ts(tmem + D_COL, a_addr, bdesc, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tmem + lane_base + D_COL + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// MN-major B operand: the SAME shared-memory image (matrix Img[N][K], K-major, as above) read as the B operand of
//   D2[128 x K] = A2[128 x N] * Img          (contraction over the image's rows)
// i.e. B2[j][n] = Img[n][j] with the "MN" index j contiguous: instruction-descriptor bit 16 (b_major) = 1 and, per
// the canonical no-swizzle MN-major layout ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), SBO = byte distance between groups of
// 8 MN elements (= N*16, the image's K-chunk stride) and LBO = byte distance between groups of 8 K rows (= 128).
__global__ void __launch_bounds__(128) probe_mn_kernel(const float* __restrict__ A2, const float* __restrict__ Img,
                                                       float* __restrict__ D, int K, int N) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  for (int idx = tid; idx < N * K; idx += 128) {
    const int n = idx / K, k = idx % K, kc = k / 8, e = k % 8;
    reinterpret_cast<__nv_bfloat16*>(smem + (size_t)kc * N * 16 + n * 16)[e] = __float2bfloat16(Img[idx]);
  }
  fence_proxy_async();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const float* arow = A2 + (size_t)tid * N;
  for (int c0 = 0; c0 < N / 2; c0 += 8) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat162 v = __floats2bfloat162_rn(arow[2 * (c0 + i)], arow[2 * (c0 + i) + 1]);
      r[i] = *reinterpret_cast<const uint32_t*>(&v);
    }
    tmem_st8(tmem + lane_base + A_COL + c0, r);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t sbase = smem_u32(smem);
    const uint32_t idesc = make_idesc_bf16(128, K) | (1u << 16);
    for (int ks = 0; ks < N / 16; ++ks) {  // 16 image rows per MMA: two groups of 8, 128 bytes apart
      const uint64_t bdesc = make_smem_desc(sbase + (uint32_t)ks * 256u, 128u, (uint32_t)N * 16u);
      mma_bf16_ts(tmem + D_COL, tmem + A_COL + ks * 8, bdesc, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < K; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tmem + lane_base + D_COL + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) D[(size_t)tid * K + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}


// Both operands from shared memory, both MN-major (the weight-gradient GEMM of lrds_ctrl_grad.cuh):
//   D[m][n] = sum_rows G[row][m] * E[row][n],   rows = the contraction index (128 per call, 16 per MMA)
// operands stored [groups of 8 features][128 rows][16 bytes]: offset(row, f) = (f / 8) * 2048 + row * 16 + (f % 8) * 2,
// descriptor start = base + 256 * kstep, LBO = 128 (next 8 rows), SBO = 2048 (next 8 features).  G carries MF <= 128
// features (the rest of the 128 M rows reads whatever follows), E carries N.
__global__ void __launch_bounds__(128) probe_ss_kernel(const float* __restrict__ G, const float* __restrict__ E,
                                                       float* __restrict__ D, int MF, int N, int swap) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint8_t* gs = smem;                 // 16 groups x 2048 (zero where MF ends)
  uint8_t* es = smem + 16 * 2048;     // N / 8 groups x 2048
  for (int g = 0; g < 16; ++g) {      // thread = row
    __half v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2half(g * 8 + e < MF ? G[(size_t)tid * MF + g * 8 + e] : 0.f);
    *reinterpret_cast<uint4*>(gs + g * 2048 + tid * 16) = *reinterpret_cast<const uint4*>(v);
  }
  for (int g = 0; g < N / 8; ++g) {
    __half v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2half(E[(size_t)tid * N + g * 8 + e]);
    *reinterpret_cast<uint4*>(es + g * 2048 + tid * 16) = *reinterpret_cast<const uint4*>(v);
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, N) | IDESC_A_MN | IDESC_B_MN;
    const uint32_t lbo = swap ? 2048u : 128u, sbo = swap ? 128u : 2048u;
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t ad = make_smem_desc(smem_u32(gs) + ks * 256u, lbo, sbo);
      const uint64_t bd = make_smem_desc(smem_u32(es) + ks * 256u, lbo, sbo);
      mma_f16_ss(tmem + D_COL, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tmem + lane_base + D_COL + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

static float round_tf32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}
static float round_bf16(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  memcpy(&v, &u, 4);
  return v;
}

template <bool BF16>
int run_case(int K, int N) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N, -1.f);
  srand(K * 131 + N);
  for (auto& v : A) v = BF16 ? round_bf16((rand() % 2001 - 1000) / 500.f) : round_tf32((rand() % 2001 - 1000) / 500.f);
  for (auto& v : B) v = BF16 ? round_bf16((rand() % 2001 - 1000) / 700.f) : round_tf32((rand() % 2001 - 1000) / 700.f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)N * K * (BF16 ? 2 : 4);
  cudaFuncSetAttribute(probe_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<BF16><<<1, 128, smem>>>(dA, dB, dD, K, N);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s K=%d N=%d: CUDA error %s\n", BF16 ? "bf16" : "tf32", K, N, cudaGetErrorString(e));
    return 1;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
      worst = fmax(worst, fabs(ref - D[m * N + n]));
    }
  printf("%s K=%3d N=%3d: max abs err %.3e  %s\n", BF16 ? "bf16" : "tf32", K, N, worst, worst < 1e-4 ? "OK" : "MISMATCH");
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return worst < 1e-4 ? 0 : 1;
}

int run_mn_case(int K, int N) {  // image [N][K]; contraction over N
  std::vector<float> A(128 * N), B(N * K), D(128 * K, -1.f);
  srand(K * 17 + N);
  for (auto& v : A) v = round_bf16((rand() % 2001 - 1000) / 500.f);
  for (auto& v : B) v = round_bf16((rand() % 2001 - 1000) / 700.f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)N * K * 2;
  cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_mn_kernel<<<1, 128, smem>>>(dA, dB, dD, K, N);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mn-major K=%d N=%d: CUDA error %s\n", K, N, cudaGetErrorString(e));
    return 1;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < 128; ++m)
    for (int j = 0; j < K; ++j) {
      double ref = 0;
      for (int n = 0; n < N; ++n) ref += (double)A[m * N + n] * (double)B[n * K + j];
      worst = fmax(worst, fabs(ref - D[m * K + j]));
    }
  printf("bf16 MN-major B: image [%3d][%3d], contraction over rows: max abs err %.3e  %s\n", N, K, worst, worst < 1e-3 ? "OK" : "MISMATCH");
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return worst < 1e-3 ? 0 : 1;
}


static float round_f16(float v) { return __half2float(__float2half(v)); }

int run_ss_case(int MF, int N, int swap) {
  std::vector<float> G(128 * MF), E(128 * N), D(128 * N, -1.f);
  srand(MF * 7 + N);
  for (auto& v : G) v = round_f16((rand() % 2001 - 1000) / 512.f);
  for (auto& v : E) v = round_f16((rand() % 2001 - 1000) / 1024.f);
  float *dG, *dE, *dD;
  cudaMalloc(&dG, G.size() * 4);
  cudaMalloc(&dE, E.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 16 * 2048 + (size_t)(N / 8) * 2048;
  cudaFuncSetAttribute(probe_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_ss_kernel<<<1, 128, smem>>>(dG, dE, dD, MF, N, swap);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("ss MF=%d N=%d: CUDA error %s\n", MF, N, cudaGetErrorString(e));
    return 1;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      if (m < MF)
        for (int r = 0; r < 128; ++r) ref += (double)G[r * MF + m] * (double)E[r * N + n];
      worst = fmax(worst, fabs(ref - D[m * N + n]));
    }
  printf("f16 SS, A and B MN-major (swap=%d): M feats %3d N %3d: max abs err %.3e  %s\n", swap, MF, N, worst,
         worst < 1e-3 ? "OK" : "MISMATCH");
  cudaFree(dG);
  cudaFree(dE);
  cudaFree(dD);
  return worst < 1e-3 ? 0 : 1;
}

int main() {
  int bad_ss = 0;
  bad_ss += run_ss_case(72, 64, 0);
  bad_ss += run_ss_case(64, 56, 0);
  bad_ss += run_ss_case(128, 64, 0);
  if (bad_ss) run_ss_case(72, 64, 1);

  int bad = bad_ss;
  bad += run_mn_case(64, 176);
  bad += run_mn_case(48, 96);
  bad += run_mn_case(16, 16);
  bad += run_case<false>(8, 16);
  bad += run_case<false>(56, 64);
  bad += run_case<false>(64, 64);
  bad += run_case<false>(64, 112);
  bad += run_case<false>(104, 64);
  bad += run_case<true>(16, 16);
  bad += run_case<true>(64, 64);
  bad += run_case<true>(64, 112);
  bad += run_case<true>(112, 64);
  printf(bad ? "tc_probe: FAILED (%d cases)\n" : "tc_probe: all cases OK\n", bad);
  return bad ? 1 : 0;
}
