#pragma once
#include "common.cuh"

struct AttnArgs {
    int N = 0, L = 0, D = 0, H = 0;
    int B = 1, blk = 0;                         // blocks stored per table slot / which one to read
    const float* table = nullptr;               // (E, slots, B, D) fp32 episodic memory
    long long slots = 0;                        // slots per episode in `table`
    const long long* ep_index = nullptr;        // (rows,)   episode of a sample, null -> row
    const long long* win_index = nullptr;       // (rows, L) slot of each window position, null -> l
    const unsigned char* mask = nullptr;        // (rows, L) 1 = attend, null -> all ones
    const long long* pe_index = nullptr;        // (rows, L) positional-table row per window position
    const float* pe = nullptr;                  // (M, D) positional table or null
    const long long* sample_index = nullptr;    // (N,) row of sample n in the arrays above, null -> n
    const float* qk = nullptr;                  // (N, H, D) folded query-key vectors
    const float* qkb = nullptr;                 // (N, H) energy bias (pre-LN fold) or null
    int ln = 0;                                 // 1: LayerNorm (no affine, eps 1e-5) every row on the fly
    float scale = 1.f;                          // sqrt(embed_dim)  (reference transformer.py:69)
    float* probs = nullptr;                     // (N, H, L) attention weights (output of fwd, input of bwd)
    float* ctx = nullptr;                       // (N, H, D) sum_l p x_l      (output of fwd, input of bwd)
};

struct AttnBwdArgs {
    const float* dctx = nullptr;                // (N, H, D)
    float* dqk = nullptr;                       // (N, H, D) out
    float* dqkb = nullptr;                      // (N, H) out (pre-LN) or null
    float* dpe = nullptr;                       // (M, D) accumulated with atomics (learned PE) or null
};

int trxl_window_attn_fwd(const AttnArgs& a, cudaStream_t st);
int trxl_window_attn_bwd(const AttnArgs& a, const AttnBwdArgs& g, cudaStream_t st);
