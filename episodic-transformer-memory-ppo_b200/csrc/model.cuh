#pragma once
#include <vector>

#include "../../include/trxl_ppo.h"
#include "common.cuh"

struct ModelIO {
    int N = 0;
    const float* feat = nullptr;
    const float* table = nullptr; long long slots = 0;
    const long long* ep_index = nullptr;
    const long long* win_index = nullptr;
    const unsigned char* mask = nullptr;
    const long long* pe_index = nullptr;
    const long long* sample_index = nullptr;
    const float* pe_table = nullptr;
    // optional episode grouping (trxl_attn_groups): enables the tensor-core attention path for post-/no-LayerNorm blocks with a
    // parameter-free positional table
    const float* table_pe = nullptr;            // (E, slots, B, D) table + positional rows
    int n_episodes = 0;
    const int4* tiles = nullptr; int n_tiles = 0;
    const int4* ranges = nullptr;
};

int model_layout(const trxl_model_config* cfg, std::vector<trxl_param_entry>& out, long long* total, int* groups);
long long model_workspace_floats(const trxl_model_config* cfg, int N);
int model_fused_supported(const trxl_model_config* cfg);
int model_forward(const trxl_model_config* c, const float* P, const ModelIO& io, float* ws, float* logits, float* value,
                  float* out_mem, cudaStream_t st);
int model_backward(const trxl_model_config* c, const float* P, float* G, const ModelIO& io, float* ws, const float* out_mem,
                   const float* dlogits, const float* dvalue, float* dfeat, cudaStream_t st);
