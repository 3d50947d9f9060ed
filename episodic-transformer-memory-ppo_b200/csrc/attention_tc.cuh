#pragma once
#include "common.cuh"

// Episode-grouped tensor-core attention (attention_tc.cu)
struct AttnTcArgs {
    int N = 0, L = 0, D = 0, H = 0;
    int B = 1, blk = 0;                         // blocks stored per table slot / which one to read
    const float* table_pe = nullptr;            // (E, slots, B, D): episodic memory + positional rows
    long long slots = 0;                        // slots per episode
    int n_episodes = 0;                         // E
    const int4* tiles = nullptr; int n_tiles = 0;   // {first (sample, head) row, rows, episode, 0}: rows of one tile share the episode
    const int4* ranges = nullptr;               // (N,) {first visible slot, visible slots, uniform, episode}
    float scale = 1.f;                          // sqrt(embed_dim)
};

long long attn_tc_row_floats(long long slots);
bool attn_tc_supported(int D, int H, long long slots, int B);
int attn_tc_ranges(const unsigned char* mask, const long long* win_index, const long long* ep_index, const long long* sample_index,
                   int N, int L, int4* ranges, cudaStream_t st);
int attn_tc_table_add_pe(const float* table, const float* pe, float* out, long long E, int M, int B, int D, int layer_norm,
                         cudaStream_t st);
int attn_tc_forward(const AttnTcArgs& a, const float* qk, float* P, float* ctx, cudaStream_t st);
int attn_tc_backward(const AttnTcArgs& a, const float* P, const float* dctx, float* scratch, float* dqk, cudaStream_t st);
