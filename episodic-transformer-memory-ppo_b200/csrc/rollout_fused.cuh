#pragma once
#include "common.cuh"

struct RfGate { long long Wr, Ur, Ug, bg; };       // [Wr;Wz;Wg] at Wr, [Ur;Uz] at Ur (float offsets into the arena)
struct RfBlock {
    long long Wv, Wk, Wq, Wo, bo;
    RfGate g1, g2;
    long long n1w, n1b, n2w, n2b, nkw, nkb, Wff, bff;
};
struct RfArgs {
    int N = 0, D = 0, H = 0, B = 0, L = 0, hid = 0, feat = 0, sumA = 0;
    int ln = 0, pe_mode = 0, gtrxl = 0;
    const float* P = nullptr;
    long long Wh = 0, bh = 0, We = 0, be = 0, pos = 0, blk_stride = 0;
    RfBlock b0;
    long long Wp = 0, bp = 0, Wlv = 0, blv = 0, Wbr = 0, bbr = 0, wval = 0, bval = 0;
    const float* feat_in = nullptr;
    const float* table = nullptr; long long slots = 0;
    const long long* ep_index = nullptr;
    const long long* win_index = nullptr;
    const unsigned char* mask = nullptr;
    const long long* pe_index = nullptr;
    const long long* sample_index = nullptr;
    const float* pe_table = nullptr;
    float* logits = nullptr;
    float* value = nullptr;
    float* out_mem = nullptr;
    long long* trace = nullptr;    // debug (TRXL_RF_TRACE=1): SM clock of cluster 0 at every phase boundary
    int cache_window = 0;      // set by the launcher: the head's window (L x D floats) is kept in shared memory between the two passes
};

size_t rollout_fused_smem_bytes(const RfArgs& a);
bool rollout_fused_supported(const RfArgs& a);
int rollout_fused_forward(const RfArgs& a, cudaStream_t st);
