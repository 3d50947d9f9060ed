#pragma once
#include "common.cuh"

#define TRXL_MAX_BRANCHES 8

struct BranchSpec {               // multi-discrete action head layout inside the concatenated logits
    int n = 0;
    int off[TRXL_MAX_BRANCHES] = {0};
    int size[TRXL_MAX_BRANCHES] = {0};
};

struct PpoLossArgs {
    int N = 0, sumA = 0;
    BranchSpec bs;
    const float* logits = nullptr;        // (N, sumA) raw policy logits
    const float* value = nullptr;         // (N,)
    const long long* actions = nullptr;   // (rows, n_branches)
    const float* old_logp = nullptr;      // (rows, n_branches)
    const float* old_values = nullptr;    // (rows,)
    const float* adv = nullptr;           // (rows,)
    const long long* sidx = nullptr;      // (N,) row of each sample, null -> n
    const double* advstats = nullptr;     // {sum a, sum a^2, count} over the (global) minibatch
    float clip = 0.f, clip_lo = 0.f, clip_hi = 0.f, beta = 0.f, vf_coef = 0.f;
    float* dlogits = nullptr;             // (N, sumA) d loss / d logits, or null
    float* dvalue = nullptr;              // (N,)
    float* partial = nullptr;             // scratch, ppo_loss_partial_floats(N)
};

struct AdamWArgs {
    float decay = 1.f;          // 1 - lr * weight_decay
    float one_minus_b1 = 0.1f, b2 = 0.999f, one_minus_b2 = 0.001f;
    float bc2_sqrt = 1.f, eps = 1e-8f, step_size = 0.f;   // step_size = lr / (1 - b1^t)
};

int ppo_gae(cudaStream_t st, const float* rewards, const unsigned char* dones, const float* values, const float* last_value,
            float* adv, int W, int T, double gamma, double lamda);
int ppo_gather_rows(cudaStream_t st, const float* src, const long long* idx, float* dst, long long rows, long long row_floats);
int ppo_gather_window(cudaStream_t st, const float* in, const long long* idx, float* out, long long N, int L, long long slots,
                      long long inner);
int ppo_memory_scatter(cudaStream_t st, float* table, const long long* ep, const long long* step, const float* new_mem, int W,
                       long long slots, long long inner);
int ppo_rollout_store(cudaStream_t st, float* table, float* table_pe, const float* pe_table, const long long* ep, const long long* step,
                      const float* new_mem, int W, long long slots, int blocks, int D, const float* value, float* value_dst,
                      long long value_stride);
int ppo_rollout_prepare(cudaStream_t st, const long long* step, const long long* ep, const unsigned char* mask_table,
                        const long long* index_table, unsigned char* mask_out, long long mask_stride, long long* idx_out,
                        long long idx_stride, long long* ep_out, long long ep_stride, int W, int L);
int ppo_rollout_fetch(cudaStream_t st, const float* obs_src, long long obs_floats, const long long* step_src, const long long* ep_src,
                      float* obs_dev, float* obs_store, long long store_stride_floats, long long* step_dev, long long* ep_dev, int n);
int ppo_sample_actions(cudaStream_t st, const float* logits, int sumA, const float* u, const long long* forced,
                       const BranchSpec& bs, long long* act_out,
                       long long act_stride, float* logp_out, long long logp_stride, long long* act_compact, int W,
                       long long* done_counter = nullptr, long long* done_flag = nullptr);
int ppo_adv_stats(cudaStream_t st, const float* adv, const long long* sidx, int N, double* out);
int ppo_loss(cudaStream_t st, PpoLossArgs a, float* stats);
long long ppo_loss_partial_floats(int N);
int ppo_clip_adamw(cudaStream_t st, float* params, float* grads, float* m, float* v, long long total, const long long* chunks,
                   int nchunks, int ngroups, float max_norm, float* partial, float* norms, const AdamWArgs& h);
