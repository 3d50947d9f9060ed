#pragma once
#include "common.cuh"

long long conv_encoder_workspace_floats(int N, int C, int H, int W);
int conv_encoder_forward(cudaStream_t st, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                         const float* b3, const float* obs, int N, int C, int H, int W, float* ws, float* feat);
