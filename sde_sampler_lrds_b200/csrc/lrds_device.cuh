// Device-side building blocks shared by the rollout kernels (fp32 SIMT path and the epilogues of the
// tcgen05 path): per-particle column views of shared memory, exact GELU, Philox4x32-10 normals, the
// distribution scores / log-densities and the FourierMLP layers in thread-per-particle form.
//
// Arithmetic follows the reference (file:line relative to /root/reference/sde_sampler/):
//   GELU (exact erf)          conf/model/base/fouriermlp.yaml:5-6, models/mlp.py:141-143
//   diag-GMM logits / score   distr/gauss.py:67-73, 97-107, 124-126, 202-221
//   PhiFour U / grad_U        distr/phi_four.py:45-96
//   logistic regression       distr/logistic_regression.py:41-61 (+ autograd score, distr/base.py:146-154)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lrds_b200.h"

namespace lrds {

constexpr int C = LRDS_CHANNELS;  // hidden width
constexpr int JC = 8;             // dimension chunk processed per inner iteration

// A particle's private vector living in shared memory as a column: element j of thread t is at
// base[j * stride + t]; with stride = blockDim.x (a multiple of 32) a warp touches 32 consecutive banks.
struct Col {
  float* p;
  int stride;
  __device__ __forceinline__ float& operator()(int j) const { return p[j * stride]; }
};

__device__ __forceinline__ float clipf(float v, float c) { return c > 0.f ? fminf(fmaxf(v, -c), c) : v; }

__device__ __forceinline__ float gelu_exact(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller; spec in oracle/philox_ref.py (the numpy statement the tests compare with)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  const float u1 = (float)((ra >> 8) + 1u) * 5.9604644775390625e-08f;  // (0, 1]
  const float u2 = (float)(rb >> 8) * 5.9604644775390625e-08f;         // [0, 1)
  const float rad = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(fmaf(u2, 6.283185307179586f, -3.141592653589793f), &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

// four normals for dims [4*blk, 4*blk+4) of particle `pidx` at step `step`
__device__ __forceinline__ void normals4(uint64_t seed, uint32_t pidx, uint32_t step, uint32_t blk,
                                         uint32_t stream_id, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(pidx, step, blk, stream_id, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  box_muller(r[0], r[1], z[0], z[1]);
  box_muller(r[2], r[3], z[2], z[3]);
}

// ---------------------------------------------------------------------------------------------------
// diagonal Gaussian mixture
// ---------------------------------------------------------------------------------------------------
struct GmmView {
  int M;
  const float* logc;
  const float* mu;
  const float* ivar;
};

__device__ __forceinline__ GmmView gmm_at(const lrds_gmm& g, int step) {
  GmmView v;
  v.M = g.M;
  v.logc = g.logc + (int64_t)step * g.step_stride_logc;
  v.mu = g.mu + (int64_t)step * g.step_stride_param;
  v.ivar = g.ivar + (int64_t)step * g.step_stride_param;
  return v;
}

// Pass 1: responsibilities r(m) = softmax_m(logc_m - q_m / 2), q_m = sum_j (x_j - mu_mj)^2 / var_mj.
// Returns log sum_m exp(logit_m) (= the mixture log-density).  For M == 1, r is not touched.
__device__ __forceinline__ float gmm_pass1(const GmmView& g, int d, const Col& x, const Col& r) {
  if (g.M == 1) {
    float q = 0.f;
    for (int j = 0; j < d; ++j) {
      const float t = x(j) - __ldg(g.mu + j);
      q = fmaf(t * t, __ldg(g.ivar + j), q);
    }
    return __ldg(g.logc) - 0.5f * q;
  }
  float mx = -INFINITY;
  int m = 0;
  for (; m + 4 <= g.M; m += 4) {  // 4 modes at a time so that one x_j load feeds 4 quadratic forms
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    const float* mu = g.mu + (int64_t)m * d;
    const float* iv = g.ivar + (int64_t)m * d;
#pragma unroll 2
    for (int j = 0; j < d; ++j) {
      const float xj = x(j);
      float t;
      t = xj - __ldg(mu + j);         q0 = fmaf(t * t, __ldg(iv + j), q0);
      t = xj - __ldg(mu + d + j);     q1 = fmaf(t * t, __ldg(iv + d + j), q1);
      t = xj - __ldg(mu + 2 * d + j); q2 = fmaf(t * t, __ldg(iv + 2 * d + j), q2);
      t = xj - __ldg(mu + 3 * d + j); q3 = fmaf(t * t, __ldg(iv + 3 * d + j), q3);
    }
    const float l0 = __ldg(g.logc + m) - 0.5f * q0, l1 = __ldg(g.logc + m + 1) - 0.5f * q1;
    const float l2 = __ldg(g.logc + m + 2) - 0.5f * q2, l3 = __ldg(g.logc + m + 3) - 0.5f * q3;
    r(m) = l0; r(m + 1) = l1; r(m + 2) = l2; r(m + 3) = l3;
    mx = fmaxf(fmaxf(mx, fmaxf(l0, l1)), fmaxf(l2, l3));
  }
  for (; m < g.M; ++m) {
    float q = 0.f;
    const float* mu = g.mu + (int64_t)m * d;
    const float* iv = g.ivar + (int64_t)m * d;
    for (int j = 0; j < d; ++j) {
      const float t = x(j) - __ldg(mu + j);
      q = fmaf(t * t, __ldg(iv + j), q);
    }
    const float l = __ldg(g.logc + m) - 0.5f * q;
    r(m) = l;
    mx = fmaxf(mx, l);
  }
  float s = 0.f;
  for (m = 0; m < g.M; ++m) {
    const float e = expf(r(m) - mx);
    r(m) = e;
    s += e;
  }
  const float inv = 1.0f / s;
  for (m = 0; m < g.M; ++m) r(m) *= inv;
  return mx + logf(s);
}

// Pass 2 for dims [j0, j0+JC): score_j = -sum_m r_m (x_j - mu_mj) / var_mj  (zero for j >= d)
__device__ __forceinline__ void gmm_score_chunk(const GmmView& g, int d, const float (&xr)[JC], const Col& r, int j0,
                                                float (&out)[JC]) {
#pragma unroll
  for (int c = 0; c < JC; ++c) out[c] = 0.f;
  if (g.M == 1) {
#pragma unroll
    for (int c = 0; c < JC; ++c)
      if (j0 + c < d) out[c] = -((xr[c] - __ldg(g.mu + j0 + c)) * __ldg(g.ivar + j0 + c));
    return;
  }
  for (int m = 0; m < g.M; ++m) {
    const float rm = r(m);
    const float* mu = g.mu + (int64_t)m * d + j0;
    const float* iv = g.ivar + (int64_t)m * d + j0;
#pragma unroll
    for (int c = 0; c < JC; ++c)
      if (j0 + c < d) out[c] = fmaf(-rm, (xr[c] - __ldg(mu + c)) * __ldg(iv + c), out[c]);
  }
}

// ---------------------------------------------------------------------------------------------------
// PhiFour (1-D lattice, Dirichlet-0)
// ---------------------------------------------------------------------------------------------------
// score_j = -beta * [ (b - x_j (1 - x_j^2)) / coef + coef (2 x_j - x_{j+1} - x_{j-1}) ]
__device__ __forceinline__ float phi4_score_1(const lrds_phi4& p, float coef, float xm, float x0, float xp) {
  const float ret = (p.b - x0 * (1.0f - x0 * x0)) / coef + coef * (2.0f * x0 - xp - xm);
  return -p.beta * ret;
}

// log-density -beta * U(x), U = coef * sum_{i=0..d} (x_{i+1}-x_i)^2/2 + sum((1-x^2)^2/4 + b x)/coef
__device__ __forceinline__ float phi4_logp(const lrds_phi4& p, int d, const Col& x) {
  const float coef = p.a * (float)d;
  float grad = 0.f, v = 0.f, prev = 0.f;
  for (int j = 0; j < d; ++j) {
    const float xj = x(j);
    const float df = xj - prev;
    grad += df * df * 0.5f;
    const float w = 1.0f - xj * xj;
    v += w * w * 0.25f + p.b * xj;
    prev = xj;
  }
  grad += prev * prev * 0.5f;
  return -p.beta * (grad * coef + v / coef);
}

// ---------------------------------------------------------------------------------------------------
// Bayesian logistic regression
// ---------------------------------------------------------------------------------------------------
// Pass 1: g(n) = m_n (y_n - sigma(z_n)), z_n = X_n . w + intercept, m_n = [eps <= sigma(z_n) <= 1 - eps].
// With want_logp the log-posterior (likelihood + Normal priors) is returned.
__device__ __forceinline__ float logreg_pass1(const lrds_logreg& L, int d, const Col& x, const Col& g, bool want_logp) {
  const float icpt = x(d - 1);
  const float hi = 1.0f - L.eps;
  float ll = 0.f;
  for (int n = 0; n < L.N; n += 4) {
    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
    const float* xt = L.Xt + n;
#pragma unroll 2
    for (int j = 0; j < L.p; ++j) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xt + (int64_t)j * L.n_pad));
      const float xj = x(j);
      z0 = fmaf(v.x, xj, z0); z1 = fmaf(v.y, xj, z1); z2 = fmaf(v.z, xj, z2); z3 = fmaf(v.w, xj, z3);
    }
    const float zz[4] = {z0 + icpt, z1 + icpt, z2 + icpt, z3 + icpt};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (n + i < L.N) {
        const float yn = __ldg(L.y + n + i);
        const float sig = 1.0f / (1.0f + expf(-zz[i]));
        const bool inside = (sig >= L.threshold) && (sig >= L.eps) && (sig <= hi);
        g(n + i) = inside ? (yn - sig) : 0.f;
        if (want_logp) {
          const float ps = fminf(fmaxf(fmaxf(sig, L.threshold), L.eps), hi);
          ll += yn * logf(ps) + (1.0f - yn) * log1pf(-ps);
        }
      }
    }
  }
  if (!want_logp) return 0.f;
  // Normal priors, logistic_regression.py:27-39, 51-53
  const float hl2pi = 0.91893853320467274178f;
  float pr = 0.f;
  const float iw = 1.0f / (2.0f * L.weight_scale * L.weight_scale);
  for (int j = 0; j < L.p; ++j) {
    const float w = x(j);
    pr += -(w * w) * iw;
  }
  pr += (float)L.p * (-logf(L.weight_scale) - hl2pi);
  const float di = icpt - L.intercept_mean;
  pr += -(di * di) / (2.0f * L.intercept_scale * L.intercept_scale) - logf(L.intercept_scale) - hl2pi;
  return ll + pr;
}

// Pass 2 for dims [j0, j0+JC): score_j = sum_n g_n X_nj - w_j / s_w^2 ; intercept: sum_n g_n - (b - m)/s_b^2
__device__ __forceinline__ void logreg_score_chunk(const lrds_logreg& L, int d, int d_pad, const float (&xr)[JC],
                                                   const Col& g, int j0, float (&out)[JC]) {
  float acc[JC];
  float gs = 0.f;
#pragma unroll
  for (int c = 0; c < JC; ++c) acc[c] = 0.f;
  for (int n = 0; n < L.N; ++n) {
    const float gn = g(n);
    const float4* row = reinterpret_cast<const float4*>(L.X + (int64_t)n * d_pad + j0);
    const float4 a = __ldg(row), b = __ldg(row + 1);
    acc[0] = fmaf(gn, a.x, acc[0]); acc[1] = fmaf(gn, a.y, acc[1]); acc[2] = fmaf(gn, a.z, acc[2]);
    acc[3] = fmaf(gn, a.w, acc[3]); acc[4] = fmaf(gn, b.x, acc[4]); acc[5] = fmaf(gn, b.y, acc[5]);
    acc[6] = fmaf(gn, b.z, acc[6]); acc[7] = fmaf(gn, b.w, acc[7]);
    gs += gn;
  }
  const float iw = 1.0f / (L.weight_scale * L.weight_scale);
#pragma unroll
  for (int c = 0; c < JC; ++c) {
    const int j = j0 + c;
    if (j < L.p) out[c] = acc[c] - xr[c] * iw;
    else if (j == L.p) out[c] = gs - (xr[c] - L.intercept_mean) / (L.intercept_scale * L.intercept_scale);
    else out[c] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------
// FourierMLP, thread-per-particle (fp32 SIMT).  Weights are read through the read-only path: every lane
// of a warp reads the same address, so each 16-byte load is one broadcast transaction.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fma_row64(float (&acc)[C], const float* __restrict__ wrow, float a) {
  const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
  for (int q = 0; q < C / 4; ++q) {
    const float4 v = __ldg(w4 + q);
    acc[4 * q + 0] = fmaf(v.x, a, acc[4 * q + 0]);
    acc[4 * q + 1] = fmaf(v.y, a, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(v.z, a, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(v.w, a, acc[4 * q + 3]);
  }
}

// hidden activations GELU(h_L) of the particle -> act(0..63)   (models/mlp.py:136-142)
__device__ __forceinline__ void mlp_hidden(const lrds_mlp& w, const float* __restrict__ bias1, const Col& x,
                                           const Col& act) {
  float acc[C];
#pragma unroll
  for (int n = 0; n < C; ++n) acc[n] = __ldg(bias1 + n);
  for (int k = 0; k < w.d; ++k) fma_row64(acc, w.w_in_t + (int64_t)k * C, x(k));
#pragma unroll
  for (int n = 0; n < C; ++n) act(n) = gelu_exact(acc[n]);
  for (int l = 0; l < w.num_hidden; ++l) {
    const float* wl = w.w_hid_t + (int64_t)l * C * C;
#pragma unroll
    for (int n = 0; n < C; ++n) acc[n] = __ldg(w.b_hid + l * C + n);
#pragma unroll 2
    for (int k = 0; k < C; ++k) fma_row64(acc, wl + k * C, act(k));
#pragma unroll
    for (int n = 0; n < C; ++n) act(n) = gelu_exact(acc[n]);
  }
}

// out_layer for dims [j0, j0+JC)   (models/mlp.py:143)
__device__ __forceinline__ void mlp_out_chunk(const lrds_mlp& w, const Col& act, int j0, float (&out)[JC]) {
  {
    const float4* b4 = reinterpret_cast<const float4*>(w.b_out + j0);
    const float4 a = __ldg(b4), b = __ldg(b4 + 1);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
  }
#pragma unroll 4
  for (int k = 0; k < C; ++k) {
    const float ak = act(k);
    const float4* row = reinterpret_cast<const float4*>(w.w_out_t + (int64_t)k * w.d_pad + j0);
    const float4 a = __ldg(row), b = __ldg(row + 1);
    out[0] = fmaf(a.x, ak, out[0]); out[1] = fmaf(a.y, ak, out[1]); out[2] = fmaf(a.z, ak, out[2]);
    out[3] = fmaf(a.w, ak, out[3]); out[4] = fmaf(b.x, ak, out[4]); out[5] = fmaf(b.y, ak, out[5]);
    out[6] = fmaf(b.z, ak, out[6]); out[7] = fmaf(b.w, ak, out[7]);
  }
}

}  // namespace lrds
