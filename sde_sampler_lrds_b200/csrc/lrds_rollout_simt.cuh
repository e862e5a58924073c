// The rollout body: one thread integrates one particle through all K steps.  Particle state (x, mixture
// responsibilities, CMCD carry-over) lives in shared-memory columns for the whole rollout; HBM is touched only
// for x0, the optional recorded noise / trajectory, and the results.  Per-step tables and distribution
// parameters are warp-uniform broadcast loads served by L1/L2.
//
// The drift network is a policy (template parameter MLP) with two hooks:
//   mlp.hidden(bias1, x)      run the network up to (and, on tensor cores, including) the output GEMM
//   mlp.out_chunk(j0, out)    out_layer(x)[j0 .. j0+8) including its bias
// SimtMlp (below) is the fp32 FFMA parity anchor (LRDS_PRECISION_FP32_SIMT); TcMlp (lrds_rollout_tc.cuh) runs the
// four GEMMs on tcgen05 with TMEM-resident operands.  Everything else is shared between the two.
#pragma once
#include "lrds_device.cuh"

namespace lrds {

struct RolloutArgs {
  lrds_spec s;
  const float* x0;
  const float* noise;
  uint64_t seed;
  uint64_t particle_offset;
  float* x_out;
  float* rnd_out;
  float* traj_out;
};

// shared-memory floats per particle for a spec (host + device agree through this one function)
struct ColLayout {
  int x, act, rt, rr, g, us, tsd, db, total;
};

__host__ __device__ inline ColLayout col_layout(const lrds_spec& s) {
  ColLayout L{};
  int off = 0;
  const int dp = s.mlp.d_pad;
  L.x = off; off += dp;
  L.act = off;
  if (s.precision == LRDS_PRECISION_FP32_SIMT) off += C;  // hidden activations: tensor-core backends keep them in TMEM
  L.rt = off;
  if (s.target.kind == LRDS_DISTR_GMM && s.target.gmm.M > 1) off += s.target.gmm.M;
  L.rr = off;
  {
    int m = 0;
    if (s.has_ref_ctrl && s.ref_t.M > 1) m = s.ref_t.M;
    if (s.ref_0.M > 1 && s.ref_0.M > m) m = s.ref_0.M;
    off += m;
  }
  L.g = off;
  if (s.target.kind == LRDS_DISTR_LOGREG) off += s.target.logreg.n_pad;
  L.us = off; L.tsd = off; L.db = off;
  if (s.kind == LRDS_ROLLOUT_CMCD || s.kind == LRDS_ROLLOUT_EUBO_CMCD) {
    L.us = off; off += dp;
    L.tsd = off; off += dp;
    L.db = off; off += dp;
  } else if (s.kind == LRDS_ROLLOUT_EUBO_LINEAR) {
    L.db = off; off += dp;
  }
  L.total = off;
  return L;
}

struct Particle {
  Col x, rt, rr, g, us, tsd, db;
};

// fp32 FFMA drift network: hidden activations in 64 shared-memory columns per particle
struct SimtMlp {
  const lrds_mlp& w;
  Col act;
  __device__ __forceinline__ void hidden(const float* __restrict__ bias1, const Col& x) { mlp_hidden(w, bias1, x, act); }
  __device__ __forceinline__ void out_chunk(int j0, float (&out)[JC]) { mlp_out_chunk(w, act, j0, out); }
};

// ---- target helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ float target_pass1(const lrds_spec& s, const Particle& P, bool want_logp) {
  const lrds_distr& t = s.target;
  if (t.kind == LRDS_DISTR_GMM) return gmm_pass1(gmm_at(t.gmm, 0), s.d, P.x, P.rt);
  if (t.kind == LRDS_DISTR_LOGREG) return logreg_pass1(t.logreg, s.d, P.x, P.g, want_logp);
  if (t.kind == LRDS_DISTR_PHI4) return want_logp ? phi4_logp(t.phi4, s.d, P.x) : 0.f;
  return 0.f;
}

// raw target score for dims [j0, j0+JC); xm / xp are x_{j0-1} / x_{j0+JC} of the SAME state as xr
__device__ __forceinline__ void target_score_chunk(const lrds_spec& s, const Particle& P, const float (&xr)[JC],
                                                   float xm, float xp, int j0, float (&out)[JC]) {
  const lrds_distr& t = s.target;
  if (t.kind == LRDS_DISTR_GMM) {
    gmm_score_chunk(gmm_at(t.gmm, 0), s.d, xr, P.rt, j0, out);
  } else if (t.kind == LRDS_DISTR_LOGREG) {
    logreg_score_chunk(t.logreg, s.d, s.mlp.d_pad, xr, P.g, j0, out);
  } else if (t.kind == LRDS_DISTR_PHI4) {
    const float coef = t.phi4.a * (float)s.d;
#pragma unroll
    for (int c = 0; c < JC; ++c) {
      const float left = (c == 0) ? xm : xr[c - 1];
      const float right = (c == JC - 1) ? xp : xr[c + 1];
      const int j = j0 + c;
      // Dirichlet-0: the neighbour beyond the last site is 0; xr is zero padded beyond d
      out[c] = (j < s.d) ? phi4_score_1(t.phi4, coef, left, xr[c], (j == s.d - 1) ? 0.f : right) : 0.f;
    }
  } else {
#pragma unroll
    for (int c = 0; c < JC; ++c) out[c] = 0.f;
  }
}

__device__ __forceinline__ void load_chunk(const Col& v, int j0, float (&out)[JC]) {
#pragma unroll
  for (int c = 0; c < JC; ++c) out[c] = v(j0 + c);
}

// control u = generative_ctrl(tau, x) for dims [j0, j0+JC), given act = hidden activations at (tau, x),
// the raw target score chunk `ts` (ScoreCtrl) and gamma = clip(score_model(tau)).
template <class MLP>
__device__ __forceinline__ void ctrl_chunk(const lrds_spec& s, MLP& mlp, int j0, const float (&ts)[JC], float gamma,
                                           float (&u)[JC]) {
  mlp.out_chunk(j0, u);
#pragma unroll
  for (int c = 0; c < JC; ++c) {
    float v = clipf(u[c], s.clip_model);
    if (s.ctrl_kind == LRDS_CTRL_SCORE) v = v + (s.scale_score * clipf(ts[c], s.clip_score)) * gamma;
    u[c] = (j0 + c < s.d) ? v : 0.f;
  }
}

__device__ __forceinline__ void noise_chunk(const RolloutArgs& a, int step, int b, int j0, float (&z)[JC]) {
  if (a.noise != nullptr) {
    const float* p = a.noise + ((int64_t)step * a.s.B + b) * a.s.d + j0;
#pragma unroll
    for (int c = 0; c < JC; ++c) z[c] = (j0 + c < a.s.d) ? __ldg(p + c) : 0.f;
  } else {
    float z4[4];
    const uint32_t pidx = (uint32_t)(a.particle_offset + (uint64_t)b);
    normals4(a.seed, pidx, (uint32_t)step, (uint32_t)(j0 >> 2), 0u, z4);
    z[0] = z4[0]; z[1] = z4[1]; z[2] = z4[2]; z[3] = z4[3];
    normals4(a.seed, pidx, (uint32_t)step, (uint32_t)(j0 >> 2) + 1u, 0u, z4);
    z[4] = z4[0]; z[5] = z4[1]; z[6] = z4[2]; z[7] = z4[3];
#pragma unroll
    for (int c = 0; c < JC; ++c)
      if (j0 + c >= a.s.d) z[c] = 0.f;
  }
}

__device__ __forceinline__ void store_traj(const RolloutArgs& a, int row, int b, int j0, const float (&v)[JC]) {
  float* p = a.traj_out + ((int64_t)row * a.s.B + b) * a.s.d + j0;
#pragma unroll
  for (int c = 0; c < JC; ++c)
    if (j0 + c < a.s.d) p[c] = v[c];
}

// ControlledLangevinSDE.drift (eq/sdes.py:101-110) for one coordinate
__device__ __forceinline__ float langevin_drift(const lrds_spec& s, float ts, float ps, float frac) {
  float dr = ts * frac + ps * (1.0f - frac);
  dr *= 0.5f * s.cmcd_diff * s.cmcd_diff;
  return clipf(dr, s.cmcd_clip);
}

// `smem` = this CTA's column area (col_layout(s).total * blockDim.x floats); every thread of the CTA runs the body
// with uniform control flow (idle lanes shadow the last particle), which the tensor-core policy relies on.
template <int KIND, class MLP>
__device__ __forceinline__ void rollout_body(const RolloutArgs& a, float* smem, MLP& mlp) {
  const lrds_spec& s = a.s;
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  const int b_raw = blockIdx.x * NT + tid;
  const bool live = b_raw < s.B;
  const int b = live ? b_raw : s.B - 1;  // idle lanes shadow the last particle, results are not stored
  const int d = s.d, dp = s.mlp.d_pad, K = s.K;
  const ColLayout L = col_layout(s);
  Particle P;
  P.x = Col{smem + L.x * NT + tid, NT};
  P.rt = Col{smem + L.rt * NT + tid, NT};
  P.rr = Col{smem + L.rr * NT + tid, NT};
  P.g = Col{smem + L.g * NT + tid, NT};
  P.us = Col{smem + L.us * NT + tid, NT};
  P.tsd = Col{smem + L.tsd * NT + tid, NT};
  P.db = Col{smem + L.db * NT + tid, NT};

  for (int j = 0; j < dp; ++j) P.x(j) = (j < d) ? __ldg(a.x0 + (int64_t)b * d + j) : 0.f;
  if (a.traj_out != nullptr && live)
    for (int j = 0; j < d; ++j) a.traj_out[(int64_t)b * d + j] = P.x(j);

  const bool score_ctrl = s.ctrl_kind == LRDS_CTRL_SCORE;
  float rnd = 0.f;

  if constexpr (KIND == LRDS_ROLLOUT_LINEAR) {
    for (int k = 0; k < K; ++k) {
      const float* row = s.steps + (int64_t)k * LRDS_STEP_STRIDE;
      const float A = __ldg(row + LRDS_STEP_A), Bc = __ldg(row + LRDS_STEP_B), Cc = __ldg(row + LRDS_STEP_C);
      const float dt = __ldg(row + LRDS_STEP_DT), sqdt = __ldg(row + LRDS_STEP_SQRT_DT);
      const float wcost = __ldg(row + LRDS_STEP_W_COST), wito = __ldg(row + LRDS_STEP_W_ITO);
      const float gamma = __ldg(row + LRDS_STEP_GAMMA), sigu = __ldg(row + LRDS_STEP_SIGU);
      if (score_ctrl) target_pass1(s, P, false);
      GmmView rv{};
      if (s.has_ref_ctrl) {
        rv = gmm_at(s.ref_t, k);
        if (rv.M > 1) gmm_pass1(rv, d, P.x, P.rr);
      }
      mlp.hidden(row + LRDS_STEP_BIAS1, P.x);
      float su2 = 0.f, sito = 0.f, xm = 0.f;
      for (int j0 = 0; j0 < dp; j0 += JC) {
        float xr[JC], ts[JC], u[JC], rs[JC], z[JC], xn[JC];
        load_chunk(P.x, j0, xr);
        const float xp = (j0 + JC < dp) ? P.x(j0 + JC) : 0.f;
        if (score_ctrl) target_score_chunk(s, P, xr, xm, xp, j0, ts);
        ctrl_chunk(s, mlp, j0, ts, gamma, u);
        if (s.has_ref_ctrl) gmm_score_chunk(rv, d, xr, P.rr, j0, rs);
        noise_chunk(a, k, b, j0, z);
#pragma unroll
        for (int c = 0; c < JC; ++c) {
          const float r = s.has_ref_ctrl ? rs[c] : 0.f;
          su2 = fmaf(u[c], u[c], su2);
          if (s.update_form == LRDS_UPDATE_AXPY) {
            xn[c] = (A * xr[c] + Bc * (r + u[c])) + Cc * z[c];
          } else {  // EM: A = f(tau), Bc = sigma, Cc = sigma^2
            float drift = -(A * xr[c]);
            if (s.has_ref_ctrl) drift += Cc * r;
            xn[c] = xr[c] + (drift + Bc * u[c]) * dt + Bc * (z[c] * sqdt);
          }
          if (s.ito_form == LRDS_ITO_SCALED) sito = fmaf(u[c], z[c], sito);
          else if (s.ito_form == LRDS_ITO_EM) sito = fmaf(u[c], z[c] * sqdt, sito);
          else if (s.ito_form == LRDS_ITO_DDS) sito += ((sigu * u[c]) * z[c]) * wito;
          if (j0 + c >= d) xn[c] = 0.f;
        }
        xm = xr[JC - 1];
#pragma unroll
        for (int c = 0; c < JC; ++c) P.x(j0 + c) = xn[c];
        if (a.traj_out != nullptr && live) store_traj(a, k + 1, b, j0, xn);
      }
      rnd += wcost * su2;
      if (s.ito_form == LRDS_ITO_SCALED) rnd += wito * sito;
      else if (s.ito_form != LRDS_ITO_NONE) rnd += sito;
    }
    // terminal cost: rnd += reference_log_prob(x) - terminal_unnorm_log_prob(x)   (oc.py:290, 505, 645, 1389)
    const float lref = gmm_pass1(gmm_at(s.ref_0, 0), d, P.x, P.rr);
    const float ltgt = clipf(target_pass1(s, P, true), s.clip_target);
    rnd += lref - ltgt;
  }

  if constexpr (KIND == LRDS_ROLLOUT_EUBO_LINEAR) {
    {  // rnd = reference_log_prob(x) - terminal_unnorm_log_prob(x)   (oc.py:321, 536)
      const float lref = gmm_pass1(gmm_at(s.ref_0, 0), d, P.x, P.rr);
      const float ltgt = clipf(target_pass1(s, P, true), s.clip_target);
      rnd = lref - ltgt;
    }
    for (int k = 0; k < K; ++k) {  // rows are stored in loop order (reversed time)
      const float* row = s.steps + (int64_t)k * LRDS_STEP_STRIDE;
      const float mean = __ldg(row + LRDS_STEP_EU_A), stdf = __ldg(row + LRDS_STEP_EU_B);
      const float wx = __ldg(row + LRDS_STEP_EU_C), sig = __ldg(row + LRDS_STEP_B);
      const float wcost = __ldg(row + LRDS_STEP_W_COST), wito = __ldg(row + LRDS_STEP_W_ITO);
      const float gamma = __ldg(row + LRDS_STEP_GAMMA);
      for (int j0 = 0; j0 < dp; j0 += JC) {  // x <- mean x + std z   (oc.py:335-337, 550-552)
        float z[JC];
        noise_chunk(a, k, b, j0, z);
#pragma unroll
        for (int c = 0; c < JC; ++c) {
          P.db(j0 + c) = z[c];
          P.x(j0 + c) = (j0 + c < d) ? fmaf(stdf, z[c], P.x(j0 + c) * mean) : 0.f;
        }
      }
      if (score_ctrl) target_pass1(s, P, false);
      const GmmView rv = gmm_at(s.ref_t, k);
      if (rv.M > 1) gmm_pass1(rv, d, P.x, P.rr);
      mlp.hidden(row + LRDS_STEP_BIAS1, P.x);
      float cost = 0.f, gx = 0.f, gz = 0.f, xm = 0.f;
      for (int j0 = 0; j0 < dp; j0 += JC) {
        float xr[JC], ts[JC], u[JC], rs[JC];
        load_chunk(P.x, j0, xr);
        const float xp = (j0 + JC < dp) ? P.x(j0 + JC) : 0.f;
        if (score_ctrl) target_score_chunk(s, P, xr, xm, xp, j0, ts);
        ctrl_chunk(s, mlp, j0, ts, gamma, u);
        gmm_score_chunk(rv, d, xr, P.rr, j0, rs);
#pragma unroll
        for (int c = 0; c < JC; ++c) {
          const float g = (s.update_form == LRDS_UPDATE_EM) ? u[c] / sig : u[c];
          cost = fmaf(g, rs[c] + 0.5f * g, cost);
          gx = fmaf(g, xr[c], gx);
          gz = fmaf(g, P.db(j0 + c), gz);
        }
        xm = xr[JC - 1];
      }
      rnd -= cost * wcost;
      if (s.update_form == LRDS_UPDATE_EM) rnd += gx * wx;
      rnd -= gz * wito;
    }
  }

  if constexpr (KIND == LRDS_ROLLOUT_CMCD || KIND == LRDS_ROLLOUT_EUBO_CMCD) {
    constexpr bool EUBO = (KIND == LRDS_ROLLOUT_EUBO_CMCD);
    const GmmView prior = gmm_at(s.ref_0, 0);
    const float sg = s.cmcd_diff;
    // control and raw target score at the starting point (row 0 forward, row K for the noising direction)
    auto eval_point = [&](int rowi, bool first, float dtk, float frac_for_drift, float& c2, float& cdb) {
      const float* row = s.steps + (int64_t)rowi * LRDS_STEP_STRIDE;
      const float gamma = __ldg(row + LRDS_STEP_GAMMA);
      target_pass1(s, P, false);
      mlp.hidden(row + LRDS_STEP_BIAS1, P.x);
      float xm = 0.f;
      for (int j0 = 0; j0 < dp; j0 += JC) {
        float xr[JC], ts[JC], u[JC];
        load_chunk(P.x, j0, xr);
        const float xp = (j0 + JC < dp) ? P.x(j0 + JC) : 0.f;
        target_score_chunk(s, P, xr, xm, xp, j0, ts);
        ctrl_chunk(s, mlp, j0, ts, gamma, u);
#pragma unroll
        for (int c = 0; c < JC; ++c) {
          const int j = j0 + c;
          if (!first && j < d) {
            const float ps = -((xr[c] - __ldg(prior.mu + j)) * __ldg(prior.ivar + j));
            const float dnew = langevin_drift(s, ts[c], ps, frac_for_drift);
            // forward: cost = (drift_s + drift_t)/sig + u_s - u_t ; noising: (drift_s + drift_t)/sig + u_s - u_t
            // with (u_s, drift_s) the NEW point there (oc.py:737, 816)
            const float cst = EUBO ? ((dnew + P.tsd(j)) / sg + u[c] - P.us(j))
                                   : ((P.tsd(j) + dnew) / sg + P.us(j) - u[c]);
            c2 = fmaf(cst, cst, c2);
            cdb = fmaf(cst, P.db(j), cdb);
          }
          P.us(j) = u[c];
          P.tsd(j) = (j < d) ? ts[c] : 0.f;
        }
        xm = xr[JC - 1];
      }
      (void)dtk;
    };
    float c2 = 0.f, cdb = 0.f;
    if constexpr (!EUBO) {
      rnd = gmm_pass1(prior, d, P.x, P.rr);  // initial_log_prob(x), oc.py:698
      eval_point(0, true, 0.f, 0.f, c2, cdb);
      for (int k = 0; k < K; ++k) {
        const float* row = s.steps + (int64_t)k * LRDS_STEP_STRIDE;
        const float dt = __ldg(row + LRDS_STEP_DT), sqdt = __ldg(row + LRDS_STEP_SQRT_DT);
        const float fs = __ldg(row + LRDS_STEP_FRAC), ft = __ldg(row + LRDS_STEP_STRIDE + LRDS_STEP_FRAC);
        for (int j0 = 0; j0 < dp; j0 += JC) {  // y = x + (drift_s + u_s sig) dt + sig db   (oc.py:722-724)
          float z[JC], xn[JC];
          noise_chunk(a, k, b, j0, z);
#pragma unroll
          for (int c = 0; c < JC; ++c) {
            const int j = j0 + c;
            float y = 0.f;
            if (j < d) {
              const float xj = P.x(j);
              const float ps = -((xj - __ldg(prior.mu + j)) * __ldg(prior.ivar + j));
              const float ds = langevin_drift(s, P.tsd(j), ps, fs);
              const float db = sqdt * z[c];
              y = xj + (ds + P.us(j) * sg) * dt + sg * db;
              P.tsd(j) = ds;
              P.db(j) = db;
            }
            xn[c] = y;
            P.x(j) = y;
          }
          if (a.traj_out != nullptr && live) store_traj(a, k + 1, b, j0, xn);
        }
        c2 = 0.f; cdb = 0.f;
        eval_point(k + 1, false, dt, ft, c2, cdb);
        rnd += 0.5f * c2 * dt;
        rnd += cdb;
      }
      rnd -= clipf(target_pass1(s, P, true), s.clip_target);  // oc.py:750
    } else {
      rnd = -clipf(target_pass1(s, P, true), s.clip_target);  // oc.py:782
      eval_point(K, true, 0.f, 0.f, c2, cdb);
      for (int i = 0; i < K; ++i) {
        const int kt = K - i, ks = K - 1 - i;  // t = ts[kt], s = ts[ks]
        const float* rows = s.steps + (int64_t)ks * LRDS_STEP_STRIDE;
        const float dt = __ldg(rows + LRDS_STEP_DT), sqdt = __ldg(rows + LRDS_STEP_SQRT_DT);
        const float ft = __ldg(s.steps + (int64_t)kt * LRDS_STEP_STRIDE + LRDS_STEP_FRAC);
        for (int j0 = 0; j0 < dp; j0 += JC) {  // y = x + (drift_t - u_t sig) dt + sig db   (oc.py:802-804)
          float z[JC];
          noise_chunk(a, i, b, j0, z);
#pragma unroll
          for (int c = 0; c < JC; ++c) {
            const int j = j0 + c;
            float y = 0.f;
            if (j < d) {
              const float xj = P.x(j);
              const float ps = -((xj - __ldg(prior.mu + j)) * __ldg(prior.ivar + j));
              const float dtt = langevin_drift(s, P.tsd(j), ps, ft);
              const float db = sqdt * z[c];
              y = xj + (dtt - P.us(j) * sg) * dt + sg * db;
              P.tsd(j) = dtt;
              P.db(j) = db;
            }
            P.x(j) = y;
          }
        }
        c2 = 0.f; cdb = 0.f;
        eval_point(ks, false, dt, ft, c2, cdb);  // drift_s uses time t (reference quirk, oc.py:807)
        rnd -= 0.5f * c2 * dt;
        rnd -= cdb;
      }
      rnd += gmm_pass1(prior, d, P.x, P.rr);  // oc.py:825
    }
  }

  if (live) {
    a.rnd_out[b] = rnd;
    if (a.x_out != nullptr)
      for (int j = 0; j < d; ++j) a.x_out[(int64_t)b * d + j] = P.x(j);
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) rollout_simt_kernel(const RolloutArgs a) {
  extern __shared__ float smem[];
  const ColLayout L = col_layout(a.s);
  SimtMlp mlp{a.s.mlp, Col{smem + L.act * blockDim.x + threadIdx.x, (int)blockDim.x}};
  rollout_body<KIND>(a, smem, mlp);
}

}  // namespace lrds
