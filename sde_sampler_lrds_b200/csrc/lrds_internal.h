// Interfaces between the translation units of liblrds_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/lrds_b200.h"

namespace lrds {
struct RolloutArgs;
// lrds_tc.cu; `err` receives the message on failure
int launch_rollout_tc(const RolloutArgs& a, cudaStream_t st, char* err, size_t n);
size_t tc_image_bytes(int d, int num_hidden, int precision);
int pack_tc_image(const lrds_mlp& w, int precision, void* image, cudaStream_t st, char* err, size_t n);
// lrds_mlp_grad.cu
bool mlp_grad_applicable(int d, int num_hidden);
int mlp_grad_params(int d, int num_hidden);
int64_t mlp_grad_scratch_floats(int d, int num_hidden, int S, int B);
int launch_mlp_grad(const lrds_mlp& mlp, const float* bias1, const float* x, const float* cot, const float* step_w,
                    const float* row_w, float clip, float cot_scale, const float* cot_scale_dev, int S, int B,
                    float* grads, float* dbias1, float* scratch, cudaStream_t st, char* err, size_t n);
}  // namespace lrds
