// Tensor-core rollout kernels of one precision (own translation unit: the precisions compile in parallel), plus the
// kernel that also runs the mixture-score contractions on the tensor core (lrds_rollout_mix.cuh).
#include "lrds_rollout_mix.cuh"
#include "lrds_tc_launch.cuh"

namespace lrds {
template int launch_prec<LRDS_PRECISION_F16X3>(const RolloutArgs&, const TcPlan&, cudaStream_t, char*, size_t);

int launch_mix_f16x3(const RolloutArgs& a, const TcPlan& p, cudaStream_t st, char* err, size_t n) {
  auto kernel = rollout_mix_kernel<LRDS_PRECISION_F16X3>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e == cudaSuccess) {
    kernel<<<p.grid, p.warps * 32, p.smem, st>>>(a, static_cast<const uint8_t*>(a.s.mlp.tc_image), p.tmem_cols);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, n, "mixture tensor-core rollout launch (grid %d x %d threads, %zu B smem, %u TMEM cols): %s", p.grid,
             p.warps * 32, p.smem, p.tmem_cols, cudaGetErrorString(e));
    return LRDS_ERR_CUDA;
  }
  return LRDS_OK;
}
}  // namespace lrds
