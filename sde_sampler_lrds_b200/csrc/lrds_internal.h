// Interfaces between the translation units of liblrds_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/lrds_b200.h"

namespace lrds {
struct RolloutArgs;
// lrds_tc.cu; `err` receives the message on failure
int launch_rollout_tc(const RolloutArgs& a, cudaStream_t st, char* err, size_t n);
size_t tc_image_bytes(int d, int num_hidden, int precision);
int pack_tc_image(const lrds_mlp& w, int precision, void* image, cudaStream_t st, char* err, size_t n);
}  // namespace lrds
