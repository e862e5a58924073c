// tcgen05 tensor-core rollout (LRDS_PRECISION_TF32X3 / BF16).  Placeholder until the kernel lands:
// every spec is reported as unsupported so that callers fail loudly instead of silently using SIMT.
#pragma once
#include "lrds_rollout_simt.cuh"

namespace lrds {
inline const char* tc_unsupported_reason() { return "tensor-core path not built yet"; }
inline int launch_rollout_tc(const RolloutArgs&, cudaStream_t) { return LRDS_ERR_UNSUPPORTED; }
}  // namespace lrds
