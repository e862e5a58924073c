// tcgen05 tensor-core drift network (LRDS_PRECISION_TF32 / TF32X3 / BF16) and the kernel wrapper that runs the
// shared rollout body (lrds_rollout_simt.cuh) with it.
//
// Mapping.  A CTA holds W <= 16 warps = up to four 128-particle tiles; warp w integrates particles
// [32 w, 32 w + 32) of the CTA and owns TMEM lanes 32 (w % 4) .. +31 of tile w / 4, i.e. thread <-> particle <->
// TMEM lane <-> row of every GEMM.  Per grid step and tile the FourierMLP (models/mlp.py:135-143) is four GEMMs
//     [128 x Kin] . W_in^T -> GELU -> [128 x 64] . W_h^T -> GELU -> ... -> [128 x 64] . W_out^T
// issued by the tile's first thread as tcgen05.mma with the A operand in TMEM (written by the particles' own
// threads with tcgen05.st), B = the weight image in shared memory (K-major, no swizzle; staged ONCE per CTA by a
// bulk TMA copy) and the fp32 accumulator in TMEM (read back with tcgen05.ld for bias + GELU).  Activations never
// touch shared or global memory.  Tiles of one CTA run out of phase, so one tile's MMA latency is covered by the
// other tiles' SIMT work (scores, noise, integrator).
//
// Precisions: TF32X3 splits both operands into tf32 (hi, lo) and accumulates lo*hi + hi*lo + hi*hi (fp32-grade
// products, the parity mode); F16X3 does the same with fp16 (hi, lo) pairs of power-of-two scaled operands (weights
// scaled per layer so that max |W| lands in [2^14, 2^15), hidden activations by 64; the accumulator is scaled back by
// the bias FFMA of the epilogue): the same 22 mantissa bits per operand at half the shared-memory image, half the
// TMEM columns (=> twice the resident tiles) and twice the tensor rate.  TF32 and BF16 are single-pass
// reduced-precision modes reported separately.
#pragma once
#include <cuda_fp16.h>

#include "lrds_rollout_simt.cuh"
#include "lrds_tc_ptx.cuh"

namespace lrds {

constexpr int TC_MAX_WARPS = 16;
// The tf32 3-pass split needs 192 TMEM columns per tile, i.e. at most two tiles = 8 warps per CTA: compile it for 256
// threads so that ptxas may use 255 registers and keep more operand loads in flight per warp.  The fp16 split needs
// 128 columns per tile; its kernels are compiled for three tiles = 12 warps (168 registers).
__host__ __device__ constexpr int tc_max_warps(int prec) {
  return prec == LRDS_PRECISION_TF32X3 ? 8 : prec == LRDS_PRECISION_F16X3 ? 12 : TC_MAX_WARPS;
}
constexpr int TC_TAIL_BYTES = 96;  // mbarriers (5 x 8 B) + TMEM slot (at 48) + the tiles' hand-off counters (4 x 4 B at 64) after the image
constexpr float TC_ACT_SCALE = 64.f;  // F16X3: hidden activations enter the tensor core as 64 * GELU(h)
__host__ __device__ constexpr bool tc_is_x3(int prec) { return prec == LRDS_PRECISION_TF32X3 || prec == LRDS_PRECISION_F16X3; }
__host__ __device__ constexpr bool tc_is_half(int prec) { return prec == LRDS_PRECISION_BF16 || prec == LRDS_PRECISION_F16X3; }

struct TcLayout {
  int prec, parts, es, kstep;
  int Kin, Nout, nh;
  uint32_t part_bytes, off_in, off_hid, off_out;  // one part = all layers of one (hi | lo) image
  uint32_t off_bhid, off_bout, off_scale, bytes;  // fp32 biases + 16 scale floats behind the parts; total bytes (multiple of 16)
  int a_cols, d_cols, tile_cols;                  // TMEM columns: one A part, the accumulator, one tile
};

__host__ __device__ inline TcLayout tc_layout(int d, int nh, int prec) {
  TcLayout L{};
  L.prec = prec;
  L.parts = tc_is_x3(prec) ? 2 : 1;
  L.es = tc_is_half(prec) ? 2 : 4;
  L.kstep = tc_is_half(prec) ? 16 : 8;
  L.Kin = (d + L.kstep - 1) / L.kstep * L.kstep;
  L.Nout = (d + 15) / 16 * 16;
  L.nh = nh;
  L.off_in = 0;
  L.off_hid = (uint32_t)(C * L.Kin * L.es);
  L.off_out = L.off_hid + (uint32_t)(nh * C * C * L.es);
  L.part_bytes = L.off_out + (uint32_t)(L.Nout * C * L.es);
  L.off_bhid = L.parts * L.part_bytes;
  L.off_bout = L.off_bhid + (uint32_t)((nh > 0 ? nh : 1) * C * 4);
  L.off_scale = L.off_bout + (uint32_t)(L.Nout * 4);  // [0..8) weight multipliers, [8..16) accumulator un-scales per layer
  L.bytes = L.off_scale + 64u;
  L.a_cols = (L.Kin > C ? L.Kin : C) * L.es / 4;
  L.d_cols = L.Nout > C ? L.Nout : C;
  L.tile_cols = L.parts * L.a_cols + L.d_cols;
  return L;
}

// ---- weight image ------------------------------------------------------------------------------------------------
// F16X3 scales (one block): layer l = 0 (input), 1..nh (hidden), nh+1 (output): multiplier 2^k with max |W| 2^k in
// [2^14, 2^15) and the accumulator un-scale 2^-k (/ TC_ACT_SCALE for the layers fed by hidden activations).
static __global__ void tc_scales_kernel(const lrds_mlp w, const TcLayout L, uint8_t* __restrict__ img) {
  __shared__ float red[32];
  float* sc = reinterpret_cast<float*>(img + L.off_scale);
  for (int l = 0; l < L.nh + 2; ++l) {
    const float* p = l == 0 ? w.w_in_t : (l <= L.nh ? w.w_hid_t + (int64_t)(l - 1) * C * C : w.w_out_t);
    const int n = l == 0 ? w.d * C : (l <= L.nh ? C * C : C * w.d_pad);
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(p[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
      int k = 0;
      if (m > 0.f && m < INFINITY) k = 14 - ilogbf(m);
      k = k > 100 ? 100 : (k < -100 ? -100 : k);
      sc[l] = ldexpf(1.0f, k);
      sc[8 + l] = ldexpf(1.0f, -k) / (l == 0 ? 1.0f : TC_ACT_SCALE);
    }
  }
}

// B operand of layer (N x K, K-major, no swizzle): 16-byte K chunk kc of row n at  kc * N * 16 + n * 16.
static __global__ void pack_tc_image_kernel(const lrds_mlp w, const TcLayout L, uint8_t* __restrict__ img) {
  const int E = 16 / L.es;
  const int n_in = C * L.Kin, n_hid = L.nh * C * C, n_out = L.Nout * C;
  const int total = n_in + n_hid + n_out;
  const float* sc = reinterpret_cast<const float*>(img + L.off_scale);  // written by tc_scales_kernel (F16X3)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    float v;
    uint32_t off;
    int n, k, N, layer;
    if (idx < n_in) {
      n = idx / L.Kin; k = idx % L.Kin; N = C; off = L.off_in; layer = 0;
      v = k < w.d ? w.w_in_t[(int64_t)k * C + n] : 0.f;
    } else if (idx < n_in + n_hid) {
      const int r = idx - n_in, l = r / (C * C);
      n = (r / C) % C; k = r % C; N = C; off = L.off_hid + (uint32_t)(l * C * C * L.es); layer = 1 + l;
      v = w.w_hid_t[(int64_t)l * C * C + (int64_t)k * C + n];
    } else {
      const int r = idx - n_in - n_hid;
      n = r / C; k = r % C; N = L.Nout; off = L.off_out; layer = L.nh + 1;
      v = n < w.d_pad ? w.w_out_t[(int64_t)k * w.d_pad + n] : 0.f;
    }
    const uint32_t byte = off + (uint32_t)(k / E) * N * 16u + (uint32_t)n * 16u + (uint32_t)(k % E) * L.es;
    if (L.prec == LRDS_PRECISION_BF16) {
      *reinterpret_cast<uint16_t*>(img + byte) = (uint16_t)(ptx::pack_bf16x2(v, 0.f) & 0xFFFFu);
    } else if (L.prec == LRDS_PRECISION_F16X3) {
      const float vs = v * sc[layer];  // exact (power of two)
      const __half hi = __float2half_rn(vs);
      *reinterpret_cast<__half*>(img + byte) = hi;
      *reinterpret_cast<__half*>(img + L.part_bytes + byte) = __float2half_rn(vs - __half2float(hi));
    } else {
      const float hi = ptx::to_tf32(v);
      *reinterpret_cast<float*>(img + byte) = hi;
      if (L.parts == 2) *reinterpret_cast<float*>(img + L.part_bytes + byte) = ptx::to_tf32(v - hi);
    }
  }
  float* bh = reinterpret_cast<float*>(img + L.off_bhid);
  float* bo = reinterpret_cast<float*>(img + L.off_bout);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (L.nh > 0 ? L.nh : 1) * C; i += gridDim.x * blockDim.x)
    bh[i] = i < L.nh * C ? w.b_hid[i] : 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.Nout; i += gridDim.x * blockDim.x)
    bo[i] = i < w.d_pad ? w.b_out[i] : 0.f;
}

// ---- the policy --------------------------------------------------------------------------------------------------
template <int PREC>
struct TcMlp {
  static constexpr bool kX3 = tc_is_x3(PREC);
  static constexpr bool kHalf = tc_is_half(PREC);
  static constexpr bool kF16 = PREC == LRDS_PRECISION_F16X3;
  static constexpr bool kPipe = tc_max_warps(PREC) <= 8;  // 255-register kernels prefetch their mixture operands
  TcLayout L;
  const uint8_t* img;  // weight image in shared memory
  uint32_t img_s;      // its shared-window address
  uint32_t tm_tile;    // TMEM address of the tile's column 0 at lane 0 (MMA operands)
  uint32_t tm_lane;    // the same at this warp's first lane (tcgen05.ld / st)
  uint64_t* bar;       // the tile's MMA-completion mbarrier
  uint32_t phase;
  int bar_id, bar_threads;  // (named barrier of the tile: only the kernels' own barriers use it now)
  bool issuer;              // this WARP issues the tile's batches (warp-uniform)
  uint32_t* hand_cnt;       // the tile's hand-off counter: one increment per warp and hand-off
  uint32_t hand_target = 0; // its value once every warp of the tile has arrived for the current hand-off
  uint32_t hand_warps;      // warps that take part in a hand-off
  int dp;              // columns of x held by the body (d rounded up to 8, zero padded)
  float vmax = 0.f, xmax = 0.f;  // F16X3: largest hidden pre-activation / |coordinate| this thread has converted to fp16
  __device__ __forceinline__ bool saturated() const { return kF16 && (vmax > F16X3_MAX_PREACT || xmax > F16X3_MAX_COORD); }

  // (hi, lo) operand bits of an activation.  The 3-pass splits truncate (one LOP3): hi = the top 11 significand bits,
  // lo = v - hi is exact in fp32; the tf32 tensor core drops lo's bits below tf32 itself, the fp16 conversion rounds
  // them (error 2^-22 of v either way).  The one-pass tf32 mode rounds.
  static __device__ __forceinline__ uint32_t split_hi(float v) {
    if (kX3) return __float_as_uint(v) & 0xFFFFE000u;
    return __float_as_uint(ptx::to_tf32(v));
  }
  __device__ __forceinline__ uint32_t a_col(int part) const { return (uint32_t)(part * L.a_cols); }
  __device__ __forceinline__ uint32_t d_col() const { return (uint32_t)(L.parts * L.a_cols); }
  __device__ __forceinline__ float unscale(int layer) const {  // F16X3: accumulator -> pre-activation
    return reinterpret_cast<const float*>(img + L.off_scale)[8 + layer];
  }

  // This warp's operand rows are stored / its accumulator rows read: increment the tile's counter (release, no round
  // trip) and go on; returns true in the tile's issuer warp once every warp of the tile has arrived (it polls the
  // counter): the caller then issues the batch under elect.sync and commits it to the tile's mbarrier.  Nobody else
  // blocks here - the warps wait for the batch on the mbarrier when they need its result (the named barrier that
  // round 1 had in front of every batch cost 6 - 11 % of the warps' time in the two-threads-per-particle kernels).
  __device__ __forceinline__ bool handoff() {
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncwarp();
    hand_target += hand_warps;
    if ((threadIdx.x & 31) == 0) ptx::red_add_release(hand_cnt, 1u);
    if (!issuer) return false;
    while ((int32_t)(ptx::ld_acquire(hand_cnt) - hand_target) < 0) {
    }
    ptx::tc_fence_after();
    return true;
  }

  // one layer's MMAs, issued by the tile's issuer warp and committed to the tile's mbarrier
  __device__ __forceinline__ void issue(uint32_t b_off, int K, int N) {
    const bool mine = handoff();
    if (mine && ptx::elect_one()) {
      const uint32_t idesc = PREC == LRDS_PRECISION_BF16 ? ptx::make_idesc_bf16(128, N)
                             : kF16                      ? ptx::make_idesc_f16(128, N)
                                                         : ptx::make_idesc_tf32(128, N);
      const uint32_t dcol = tm_tile + d_col();
      const int ksteps = K / L.kstep;
      const uint32_t kbytes = 2u * (uint32_t)N * 16u;  // one MMA consumes two 16-byte K chunks
      uint32_t acc = 0;
      auto pass = [&](int a_part, int b_part) {
        const uint32_t abase = tm_tile + a_col(a_part);
        const uint32_t bbase = img_s + (uint32_t)b_part * L.part_bytes + b_off;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t bdesc = ptx::make_smem_desc(bbase + (uint32_t)ks * kbytes, (uint32_t)N * 16u, 128u);
          if (kHalf) ptx::mma_bf16_ts(dcol, abase + ks * 8, bdesc, idesc, acc);  // kind::f16 (bf16 or fp16 per idesc)
          else ptx::mma_tf32_ts(dcol, abase + ks * 8, bdesc, idesc, acc);
          acc = 1;
        }
      };
      if (kX3) {  // small terms first
        pass(1, 0);
        pass(0, 1);
      }
      pass(0, 0);
      ptx::mma_commit(bar);
    }
    if (mine) __syncwarp();
  }
  __device__ __forceinline__ void wait() {
    ptx::mbar_wait(bar, phase);
    phase ^= 1u;
    ptx::tc_fence_after();
  }

  // A operand <- 8 consecutive fp32 values (columns c0 .. c0+7 of the K axis), tf32 kinds
  __device__ __forceinline__ void store8(int c0, const float (&v)[8]) {
    uint32_t hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hi[i] = split_hi(v[i]);
    ptx::tmem_st8(tm_lane + a_col(0) + c0, hi);
    if (kX3) {
      uint32_t lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
      ptx::tmem_st8(tm_lane + a_col(1) + c0, lo);
    }
  }
  // A operand <- 16 consecutive fp32 values as 8 packed columns c0 .. c0+7, 16-bit kinds
  __device__ __forceinline__ void store16_half(int c0, const float (&v)[16]) {
    uint32_t r[8];
    if (kF16) {
      uint32_t rl[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        u64 hi, lo;
        xmax = max3abs(xmax, v[2 * i], v[2 * i + 1]);
        f2::split(f2::pack(v[2 * i], v[2 * i + 1]), hi, lo);
        float h0, h1, l0, l1;
        f2::unpack(hi, h0, h1);
        f2::unpack(lo, l0, l1);
        r[i] = ptx::pack_f16x2(h0, h1);
        rl[i] = ptx::pack_f16x2(l0, l1);
      }
      ptx::tmem_st8(tm_lane + a_col(0) + c0, r);
      ptx::tmem_st8(tm_lane + a_col(1) + c0, rl);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = ptx::pack_bf16x2(v[2 * i], v[2 * i + 1]);
      ptx::tmem_st8(tm_lane + a_col(0) + c0, r);
    }
  }

  __device__ __forceinline__ void store_x(const Col4& x) {
    if (kHalf) {
      for (int c0 = 0; c0 < L.Kin / 2; c0 += 8) {  // 16 dims -> 8 packed columns
        float v[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j0 = 2 * c0 + 8 * h;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (j0 < dp) {
            a = x.ld4(j0 >> 2);
            b = x.ld4((j0 >> 2) + 1);
          }
          v[8 * h + 0] = a.x; v[8 * h + 1] = a.y; v[8 * h + 2] = a.z; v[8 * h + 3] = a.w;
          v[8 * h + 4] = b.x; v[8 * h + 5] = b.y; v[8 * h + 6] = b.z; v[8 * h + 7] = b.w;
        }
        store16_half(c0, v);
      }
    } else {
      for (int c0 = 0; c0 < L.Kin; c0 += 8) {  // Kin == dp for the tf32 kinds
        const float4 a = x.ld4(c0 >> 2), b = x.ld4((c0 >> 2) + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store8(c0, v);
      }
    }
  }

  // F16X3: the whole epilogue on packed fp32x2 arithmetic
  template <bool GLOBAL_BIAS>
  __device__ __forceinline__ void epilogue_f16(const float* __restrict__ bias, int layer, int c_begin = 0, int c_end = C) {
    const u64 us2 = f2::pk(unscale(layer));
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(tm_lane + d_col() + c0, r);
      ptx::tmem_wait_ld();
      const float4* b4 = reinterpret_cast<const float4*>(bias + c0);  // warp-uniform, 16-byte aligned
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = GLOBAL_BIAS ? __ldg(b4 + i) : b4[i];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const u64 acc = f2::pack(__uint_as_float(r[4 * i + 2 * e]), __uint_as_float(r[4 * i + 2 * e + 1]));
          const u64 v = f2::fma(acc, us2, e ? f2::pack(b.z, b.w) : f2::pack(b.x, b.y));
          {
            float va, vb;
            f2::unpack(v, va, vb);
            vmax = max3(vmax, va, vb);
          }
          const u64 g = gelu_pair(v, 5.0f, TC_ACT_SCALE);  // TC_ACT_SCALE / 2 = 2^5
          u64 hi, lo;
          f2::split(g, hi, lo);
          float h0, h1, l0, l1;
          f2::unpack(hi, h0, h1);
          f2::unpack(lo, l0, l1);
          ph[2 * i + e] = ptx::pack_f16x2(h0, h1);
          pl[2 * i + e] = ptx::pack_f16x2(l0, l1);
        }
      }
      ptx::tmem_st16(tm_lane + a_col(0) + c0 / 2, ph);
      ptx::tmem_st16(tm_lane + a_col(1) + c0 / 2, pl);
    }
  }

  // the same for 16 accumulator columns [c0, c0 + 16) (four threads per particle, lrds_rollout_mix_small.cuh)
  template <bool GLOBAL_BIAS>
  __device__ __forceinline__ void epilogue_f16_16(const float* __restrict__ bias, int layer, int c0) {
    const u64 us2 = f2::pk(unscale(layer));
    uint32_t r[16];
    ptx::tmem_ld16(tm_lane + d_col() + c0, r);
    ptx::tmem_wait_ld();
    const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
    uint32_t ph[8], pl[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = GLOBAL_BIAS ? __ldg(b4 + i) : b4[i];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const u64 acc = f2::pack(__uint_as_float(r[4 * i + 2 * e]), __uint_as_float(r[4 * i + 2 * e + 1]));
        const u64 v = f2::fma(acc, us2, e ? f2::pack(b.z, b.w) : f2::pack(b.x, b.y));
        {
          float va, vb;
          f2::unpack(v, va, vb);
          vmax = max3(vmax, va, vb);
        }
        const u64 g = gelu_pair(v, 5.0f, TC_ACT_SCALE);
        u64 hi, lo;
        f2::split(g, hi, lo);
        float h0, h1, l0, l1;
        f2::unpack(hi, h0, h1);
        f2::unpack(lo, l0, l1);
        ph[2 * i + e] = ptx::pack_f16x2(h0, h1);
        pl[2 * i + e] = ptx::pack_f16x2(l0, l1);
      }
    }
    ptx::tmem_st8(tm_lane + a_col(0) + c0 / 2, ph);
    ptx::tmem_st8(tm_lane + a_col(1) + c0 / 2, pl);
  }

  // accumulator (64 columns) [* un-scale] + bias -> GELU -> A operand of the next layer
  template <bool GLOBAL_BIAS>
  __device__ __forceinline__ void epilogue(const float* __restrict__ bias, int layer) {
    if constexpr (kF16) {
      epilogue_f16<GLOBAL_BIAS>(bias, layer);
      return;
    }
    const float us = kF16 ? unscale(layer) : 1.0f;
    const float gs = kF16 ? 0.5f * TC_ACT_SCALE : 0.5f;
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(tm_lane + d_col() + c0, r);
      ptx::tmem_wait_ld();
      float g[32];
      const float4* b4 = reinterpret_cast<const float4*>(bias + c0);  // warp-uniform, 16-byte aligned
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = GLOBAL_BIAS ? __ldg(b4 + i) : b4[i];
        if (kF16) {
          gelu_exact2(fmaf(__uint_as_float(r[4 * i + 0]), us, b.x), fmaf(__uint_as_float(r[4 * i + 1]), us, b.y), g[4 * i + 0], g[4 * i + 1], gs);
          gelu_exact2(fmaf(__uint_as_float(r[4 * i + 2]), us, b.z), fmaf(__uint_as_float(r[4 * i + 3]), us, b.w), g[4 * i + 2], g[4 * i + 3], gs);
        } else {
          gelu_exact2(__uint_as_float(r[4 * i + 0]) + b.x, __uint_as_float(r[4 * i + 1]) + b.y, g[4 * i + 0], g[4 * i + 1]);
          gelu_exact2(__uint_as_float(r[4 * i + 2]) + b.z, __uint_as_float(r[4 * i + 3]) + b.w, g[4 * i + 2], g[4 * i + 3]);
        }
      }
      if (kHalf) {
        float lo16[16], hi16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { lo16[i] = g[i]; hi16[i] = g[16 + i]; }
        store16_half(c0 / 2, lo16);
        store16_half(c0 / 2 + 8, hi16);
      } else {
        uint32_t hi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) hi[i] = split_hi(g[i]);
        ptx::tmem_st32(tm_lane + a_col(0) + c0, hi);
        if (kX3) {
#pragma unroll
          for (int i = 0; i < 32; ++i) hi[i] = __float_as_uint(g[i] - __uint_as_float(hi[i]));
          ptx::tmem_st32(tm_lane + a_col(1) + c0, hi);
        }
      }
    }
  }

  // BIAS_SH: the table row holding bias1 has been staged in shared memory
  template <bool BIAS_SH>
  __device__ __forceinline__ void hidden(const float* __restrict__ bias1, const Col4& x) {
    store_x(x);
    issue(L.off_in, L.Kin, C);
    const float* bh = reinterpret_cast<const float*>(img + L.off_bhid);
    for (int l = 0; l < L.nh; ++l) {
      wait();
      if (l == 0) epilogue<!BIAS_SH>(bias1, 0);
      else epilogue<false>(bh + (l - 1) * C, l);
      issue(L.off_hid + (uint32_t)(l * C * C * L.es), C, C);
    }
    wait();
    if (L.nh == 0) epilogue<!BIAS_SH>(bias1, 0);
    else epilogue<false>(bh + (L.nh - 1) * C, L.nh);
    issue(L.off_out, C, L.Nout);
    wait();
  }

  // the same as four (dim 2q, dim 2q+1) pairs (F16X3)
  __device__ __forceinline__ void out_chunk2(int j0, u64 (&out)[JC / 2]) {
    uint32_t r[8];
    ptx::tmem_ld8(tm_lane + d_col() + j0, r);
    ptx::tmem_wait_ld();
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(img + L.off_bout + j0 * 4);
    const ulonglong2 a = b2[0], b = b2[1];
    const u64 us2 = f2::pk(unscale(L.nh + 1));
    out[0] = f2::fma(f2::pack(__uint_as_float(r[0]), __uint_as_float(r[1])), us2, a.x);
    out[1] = f2::fma(f2::pack(__uint_as_float(r[2]), __uint_as_float(r[3])), us2, a.y);
    out[2] = f2::fma(f2::pack(__uint_as_float(r[4]), __uint_as_float(r[5])), us2, b.x);
    out[3] = f2::fma(f2::pack(__uint_as_float(r[6]), __uint_as_float(r[7])), us2, b.y);
  }

  __device__ __forceinline__ void out_chunk(int j0, float (&out)[JC]) {
    uint32_t r[8];
    ptx::tmem_ld8(tm_lane + d_col() + j0, r);
    ptx::tmem_wait_ld();
    const float4* b4 = reinterpret_cast<const float4*>(img + L.off_bout + j0 * 4);
    const float4 a = b4[0], b = b4[1];
    const float bs[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (kF16) {
      const float us = unscale(L.nh + 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) out[c] = fmaf(__uint_as_float(r[c]), us, bs[c]);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) out[c] = __uint_as_float(r[c]) + bs[c];
    }
  }
};

__host__ __device__ inline uint32_t tc_stage_bytes(const lrds_spec& s, int level) {
  return level > 0 ? (stage_layout(s, level).total + 15u) & ~15u : 0u;
}

// ---- kernel --------------------------------------------------------------------------------------------------------
// shared memory: [weight image | mbarriers + TMEM slot | operand stage (STAGED) | particle columns]
template <int KIND, int PREC, int STAGE, class TR>
__global__ void __launch_bounds__(tc_max_warps(PREC) * 32, 1)
rollout_tc_kernel(const RolloutArgs a, const uint8_t* __restrict__ image, const uint32_t tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const TcLayout TL = tc_layout(a.s.d, a.s.mlp.num_hidden, PREC);
  const int tid = threadIdx.x, warp = tid >> 5, nwarps = blockDim.x >> 5;
  uint8_t* img = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + TL.bytes);  // [0] image, [1 + t] tile t
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + TL.bytes + 48);
  uint8_t* stage = smem_raw + TL.bytes + TC_TAIL_BYTES;
  float* cols = reinterpret_cast<float*>(stage + tc_stage_bytes(a.s, STAGE));
  if (warp == 0) ptx::tmem_alloc(slot, tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) ptx::mbar_init(bars + i, 1);
    for (int i = 0; i < 4; ++i) reinterpret_cast<uint32_t*>(smem_raw + TL.bytes + 64)[i] = 0u;
    ptx::fence_mbar_init();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (tid == 0) {  // drift weights staged once per CTA by the TMA engine
    ptx::mbar_expect_tx(bars, TL.bytes);
    ptx::bulk_g2s(img, image, TL.bytes, bars);
  }
  ptx::mbar_wait(bars, 0);
  const uint32_t tmem = *slot;
  const int tile = warp >> 2;
  const int tile_warps = min(4, nwarps - 4 * tile);
  TcMlp<PREC> mlp;
  mlp.L = TL;
  mlp.img = img;
  mlp.img_s = ptx::smem_u32(img);
  mlp.tm_tile = tmem + (uint32_t)(tile * TL.tile_cols);
  mlp.tm_lane = mlp.tm_tile + ((uint32_t)((warp & 3) * 32) << 16);
  mlp.bar = bars + 1 + tile;
  mlp.phase = 0;
  mlp.bar_id = 1 + tile;
  mlp.bar_threads = tile_warps * 32;
  mlp.issuer = (warp & 3) == tile_warps - 1;  // the tile's last warp: its sub-partition carries the fewest warps
  mlp.hand_cnt = reinterpret_cast<uint32_t*>(smem_raw + TL.bytes + 64) + tile;
  mlp.hand_warps = (uint32_t)tile_warps;
  mlp.dp = a.s.mlp.d_pad;
  rollout_body<KIND, STAGE, TR>(a, cols, stage, mlp);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, tmem_cols);
}

// ---- host side -------------------------------------------------------------------------------------------------------
struct TcPlan {
  int warps, grid, staged;
  uint32_t tmem_cols;
  size_t smem;
};

// Warps per CTA: as few waves as possible over the SMs (one CTA per SM), then as few idle lanes as possible.
// Operand staging (LINEAR kind) is used when its shared-memory cost does not reduce the warps per CTA.
inline int plan_rollout_tc(const lrds_spec& s, int smem_cap, int sms, TcPlan* out, const char** why) {
  const TcLayout TL = tc_layout(s.d, s.mlp.num_hidden, s.precision);
  const ColLayout CL = col_layout(s);
  const size_t fixed = (size_t)TL.bytes + TC_TAIL_BYTES;
  const size_t per_warp = (size_t)CL.total * 32 * sizeof(float);
  if (TL.tile_cols > 512) { *why = "TMEM columns of one tile exceed 512"; return LRDS_ERR_RESOURCES; }
  if (fixed + per_warp > (size_t)smem_cap) { *why = "weight image + one warp of particle state exceed shared memory"; return LRDS_ERR_RESOURCES; }
  const int tmax = 512 / TL.tile_cols;
  const int need = (s.B + 31) / 32;
  auto pick = [&](size_t extra, int* wmax_out) {
    if (fixed + extra + per_warp > (size_t)smem_cap) { *wmax_out = 0; return 0; }
    int wmax = (int)(((size_t)smem_cap - fixed - extra) / per_warp);
    wmax = wmax < tc_max_warps(s.precision) ? wmax : tc_max_warps(s.precision);
    wmax = wmax < 4 * tmax ? wmax : 4 * tmax;
    *wmax_out = wmax;
    const int waves = (need + sms * wmax - 1) / (sms * wmax);
    int w = (need + sms * waves - 1) / (sms * waves);
    return w < 1 ? 1 : (w > wmax ? wmax : w);
  };
  int wm = 0;
  const int w_plain = pick(0, &wm);
  int level = 0, w = w_plain;
  size_t stage = 0;
  if (s.kind == LRDS_ROLLOUT_LINEAR) {
    for (int lv = 2; lv >= 1; --lv) {  // the deepest staging level that does not cost warps
      const size_t bytes = tc_stage_bytes(s, lv);
      const int wl = pick(bytes, &wm);
      if (wl >= w_plain) { level = lv; w = wl; stage = bytes; break; }
    }
  }
  const bool staged = level > 0;
  const int tiles = (w + 3) / 4;
  uint32_t cols = 32;
  while ((int)cols < tiles * TL.tile_cols) cols <<= 1;
  out->warps = w;
  out->grid = (need + w - 1) / w;
  out->staged = level;
  out->tmem_cols = cols;
  out->smem = fixed + (staged ? stage : 0) + per_warp * w;
  return LRDS_OK;
}

}  // namespace lrds
