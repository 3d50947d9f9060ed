#pragma once
#include "common.cuh"
#include "tc_conv_geom.h"

// p = {conv1.weight, conv1.bias, conv2.weight, conv2.bias, conv3.weight, conv3.bias} (reference layouts)
int tc_conv_supported(int C, int H, int W);
long long tc_conv_workspace_floats(int N, int C, int H, int W);
// repack = 0 reuses the tensor-core-format weights already in ws (tc_conv_pack_weights, or an earlier forward with repack = 1)
int tc_conv_pack_weights(cudaStream_t st, const float* const* p, int N, int C, int H, int W, float* ws);
int tc_conv_forward(cudaStream_t st, const float* const* p, const float* obs, const long long* sample_index, int N, int C, int H,
                    int W, float* ws, float* feat, int repack);
// g = gradient slots in the same order as p; overwritten (not accumulated).  ws must still hold the forward's state.
int tc_conv_backward(cudaStream_t st, float* const* g, int N, int C, int H, int W, float* ws, const float* dfeat);
