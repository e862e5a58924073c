// The mixture tensor-core kernel (lrds_rollout_mix.cuh): the benchmark configuration and its EUBO / Gaussian-reference variants (own translation unit: the
// configurations compile in parallel).
#include <cstdio>

#include "lrds_internal.h"
#include "lrds_rollout_mix.cuh"

namespace lrds {

template <class CFG>
static int launch_mix_cfg(const RolloutArgs& a, const TcPlan& p, cudaStream_t st, char* err, size_t n) {
  auto kernel = rollout_mix_kernel<LRDS_PRECISION_F16X3, CFG>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e == cudaSuccess) {
    // a partial last tile gets the auxiliary issuer warp (lrds_rollout_mix.cuh) when the CTA has room for it
    const int aux = (p.warps % 4 != 0 && p.warps < MIX_MAX_WARPS) ? 1 : 0;
    kernel<<<p.grid, (p.warps + aux) * 32, p.smem, st>>>(a, static_cast<const uint8_t*>(a.s.mlp.tc_image), p.tmem_cols, p.warps);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, n, "mixture tensor-core rollout launch (grid %d x %d threads, %zu B smem, %u TMEM cols): %s", p.grid,
             p.warps * 32, p.smem, p.tmem_cols, cudaGetErrorString(e));
    return LRDS_ERR_CUDA;
  }
  return LRDS_OK;
}

int launch_mix_b(int cfg, const RolloutArgs& a, const TcPlan& p, cudaStream_t st, char* err, size_t n);  // lrds_tc_mix_b.cu
int launch_mix_c(int cfg, const RolloutArgs& a, const TcPlan& p, cudaStream_t st, char* err, size_t n);  // lrds_tc_mix_c.cu

int launch_mix_f16x3(const RolloutArgs& a, const TcPlan& p, cudaStream_t st, char* err, size_t n) {
  const int cfg = mix_tc_config(a.s);
  switch (cfg) {  // MixCfg<EUBO, TGT, REF, EM>
    case 0: return launch_mix_cfg<MixBench>(a, p, st, err, n);
    case 1: return launch_mix_cfg<MixCfg<true, 1, 2, false>>(a, p, st, err, n);
    case 2: return launch_mix_cfg<MixCfg<false, 1, 1, false>>(a, p, st, err, n);
    case 3: return launch_mix_cfg<MixCfg<false, 1, 1, true>>(a, p, st, err, n);
    case 4: case 5: case 6: return launch_mix_b(cfg, a, p, st, err, n);
    case 7: case 8: case 9: case 10: return launch_mix_c(cfg, a, p, st, err, n);
  }
  snprintf(err, n, "mixture tensor-core rollout: configuration not built");
  return LRDS_ERR_UNSUPPORTED;
}
}  // namespace lrds

#ifdef LRDS_MIX_TIMING
extern "C" int lrds_debug_mix_timing(unsigned long long* host_out) {  // tools only: 2 CTAs x 16 warps x 16 phase counters
  return (int)cudaMemcpyFromSymbol(host_out, lrds::g_mix_timing, sizeof(lrds::g_mix_timing));
}
#endif
