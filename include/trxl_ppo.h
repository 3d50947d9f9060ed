/* libtrxlppo -- C ABI of the B200-native PPO + TransformerXL hot path.
 *
 * The reference (MarcoMeter/episodic-transformer-memory-ppo) is pure Python/PyTorch and has no FFI;
 * each entry point below names the reference code it replaces (file:line under /root/reference).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = error (TRXL_ERR_*); trxl_last_error() gives a
 *     thread-local message.  Nothing throws across this boundary.
 *   - all tensor arguments are raw DEVICE pointers into memory the caller owns (fp32 `float`,
 *     int64 `long long`, bool `unsigned char`), row-major, 16-byte aligned unless stated otherwise.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on that stream and
 *     never synchronises, allocates or frees device memory.
 *   - no CPU fallback exists: these kernels are built for sm_100a only.
 */
#ifndef TRXL_PPO_H
#define TRXL_PPO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRXL_ABI_VERSION 1
#define TRXL_MAX_BRANCHES 8
#define TRXL_COMM_ID_BYTES 128

enum { TRXL_LN_NONE = 0, TRXL_LN_PRE = 1, TRXL_LN_POST = 2 };        /* config["transformer"]["layer_norm"] */
enum { TRXL_PE_NONE = 0, TRXL_PE_RELATIVE = 1, TRXL_PE_LEARNED = 2 }; /* ...["positional_encoding"]          */

/* Model hyper-parameters (reference configs/ *.yaml + environment spaces, model.py:11-69). */
typedef struct trxl_model_config {
    int32_t embed_dim;          /* transformer.embed_dim (multiple of 4)                */
    int32_t num_heads;          /* transformer.num_heads                                */
    int32_t num_blocks;         /* transformer.num_blocks                               */
    int32_t memory_length;      /* transformer.memory_length                            */
    int32_t hidden_size;        /* hidden_layer_size (multiple of 4)                    */
    int32_t feat_dim;           /* inputs of lin_hidden: conv features or vector-obs dim */
    int32_t layer_norm;         /* TRXL_LN_*                                            */
    int32_t pos_enc;            /* TRXL_PE_*                                            */
    int32_t gtrxl;              /* 0/1                                                  */
    int32_t max_episode_steps;  /* rows of the positional table                         */
    int32_t num_branches;       /* len(action_space_shape)                              */
    int32_t branch_sizes[TRXL_MAX_BRANCHES];
    int32_t conv_in_channels;   /* >0: reserve conv1..3 parameters (visual obs) in the arena */
} trxl_model_config;

/* One parameter of the flat arena, named as in the reference state_dict. */
typedef struct trxl_param_entry {
    char name[96];
    int64_t offset;             /* in floats from the arena base (multiple of 4)        */
    int32_t ndim;
    int64_t shape[4];
    int32_t group;              /* gradient-norm group, see trxl_layout_groups          */
} trxl_param_entry;

const char* trxl_last_error(void);
int trxl_abi_version(void);
/* kernels launched by this library since load (bench.py reports the delta as gpu_launches) */
int64_t trxl_launch_count(void);
/* of which launches of the tcgen05 3xTF32 GEMM (tc_gemm.cu; enabled with TRXL_TCGEN05=1) */
int64_t trxl_tc_gemm_launches(void);
/* Measurement aid: when enabled, every window-attention launch is bracketed by CUDA events on its own
 * stream.  trxl_profile_read sums the durations of launches of `kind` (0 forward, 1 backward) that
 * processed at least min_samples samples; synchronises on those events only. */
int trxl_profile_enable(int on);
int trxl_profile_read(int kind, int min_samples, double* total_ms, int64_t* launches, int64_t* samples);
/* sum of the counts attached to the timed launches of `kind` (kinds 2 / 3 = episode-grouped attention forward / backward:
 * number of 128-row tiles) */
int64_t trxl_profile_aux(int kind, int min_samples);

/* Capture / replay of a sequence of this library's launches as a CUDA graph (the ~45 launches of one rollout step).
 * `stream` must be a non-default stream for begin/end; only calls of this library may be issued in between. */
int trxl_graph_begin(void* stream);
int trxl_graph_end(void* stream, void** graph_exec_out);
int trxl_graph_launch(void* graph_exec, void* stream);
int trxl_graph_destroy(void* graph_exec);
/* dst[r, 0:row_bytes] = src[r, 0:row_bytes] for `rows` strided rows (device to device) */
int trxl_copy_rows(const void* src, void* dst, int64_t rows, int64_t row_bytes, int64_t src_stride_bytes, int64_t dst_stride_bytes,
                   void* stream);
/* plain asynchronous copy of `bytes` bytes; either side may be pinned / cudaHostRegister'ed host memory (the rollout's
 * observation slab, episode cursors, sampled actions: trainer.py:163,189 of the reference do these as torch copies).
 * Capturable into a CUDA graph. */
int trxl_copy_async(const void* src, void* dst, int64_t bytes, void* stream);

/* ---- parameter arena layout ------------------------------------------------------------------ */
/* Number of entries / total floats of the arena for a config (<0 on invalid config). */
int trxl_layout_num_entries(const trxl_model_config* cfg);
int64_t trxl_layout_total_floats(const trxl_model_config* cfg);
int trxl_layout_entry(const trxl_model_config* cfg, int index, trxl_param_entry* out);
/* Number of gradient-norm groups G (reference model.py:128-151): 0 encoder, 1 linear_layer,
 * 2..2+B-1 transformer_block_i, then policy_head_k, lin_policy, lin_value, value (head), other. */
int trxl_layout_groups(const trxl_model_config* cfg);

/* ---- model forward / backward (everything after the CNN) -------------------------------------- */
/* floats of workspace needed by trxl_model_forward/backward for N samples */
int64_t trxl_workspace_floats(const trxl_model_config* cfg, int N);
/* 1 if trxl_model_forward accepts workspace == NULL for this config: inference-only forward that runs the
 * whole trunk in ONE launch (one CTA per sample, activations in shared memory) -- the rollout path. */
int trxl_fused_forward_supported(const trxl_model_config* cfg);

/* Replaces ActorCriticModel.forward from lin_hidden on (model.py:97-110) and Transformer.forward
 * (transformer.py:222-253) including the window gather (utils.py:52-75, buffer.py:90,
 * trainer.py:168,233,271) -- the window is read in place from the episode table.
 *   feat        (N, feat_dim)        encoder features (flattened conv output or the vector obs)
 *   table       (E, slots, B, D)     episodic memory; sample n reads episode ep_index[row]
 *   ep_index    (rows,)  or NULL -> row        win_index (rows, L) or NULL -> 0..L-1
 *   mask        (rows, L) bool       pe_index (rows, L)      sample_index (N,) or NULL -> n
 *     (row = sample_index[n]; lets a minibatch address the flat rollout buffer without copies)
 *   pe_table    (max_episode_steps, D) for TRXL_PE_RELATIVE (host-built sinusoid), else NULL
 * Outputs: logits (N, sum A) raw, value (N,), out_mem (N, B, D) = inputs of every block.
 * workspace == NULL selects the fused inference path (see trxl_fused_forward_supported); it saves no
 * activations, so trxl_model_backward cannot follow it.  On that path pe_index == NULL means that `table` already carries
 * the positional rows (see trxl_rollout_store): no positional add, and the window rows are prefetched block by block. */
int trxl_model_forward(const trxl_model_config* cfg, const float* params, const float* feat, const float* table,
                       int64_t slots, const int64_t* ep_index, const int64_t* win_index, const uint8_t* mask,
                       const int64_t* pe_index, const int64_t* sample_index, const float* pe_table, int N,
                       float* workspace, float* logits, float* value, float* out_mem, void* stream);

/* Backward of the above given d loss/d logits (N, sum A) and d loss/d value (N,): writes every
 * parameter gradient at its arena offset in `grads` (overwrite, except pos_embedding which is
 * accumulated with atomics and must be zeroed by the caller) and, if dfeat != NULL, d loss/d feat.
 * Must follow a trxl_model_forward with the same arguments, workspace and out_mem (autograd of
 * trainer.py:310 for this part of the graph). */
int trxl_model_backward(const trxl_model_config* cfg, const float* params, float* grads, const float* feat,
                        const float* table, int64_t slots, const int64_t* ep_index, const int64_t* win_index,
                        const uint8_t* mask, const int64_t* pe_index, const int64_t* sample_index,
                        const float* pe_table, int N, float* workspace, const float* out_mem, const float* dlogits,
                        const float* dvalue, float* dfeat, void* stream);

/* Inference path of the CNN encoder (model.py:87-94: conv 8/4 -> ReLU -> conv 4/2 -> ReLU -> conv 3/1 -> ReLU -> flatten)
 * as im2col + GEMM with fused bias/ReLU; used for the rollout forwards.  obs (N, C, H, W) -> feat (N, 64*oh*ow) in the
 * reference's NCHW flatten order.  workspace >= trxl_conv_encoder_workspace_floats(cfg, N, H, W) floats. */
int64_t trxl_conv_encoder_workspace_floats(const trxl_model_config* cfg, int N, int H, int W);
int trxl_conv_encoder_forward(const trxl_model_config* cfg, const float* params, const float* obs, int N, int H, int W,
                              float* workspace, float* feat, void* stream);

/* Training path of the CNN encoder on the tcgen05 tensor cores (3xTF32 implicit GEMMs, csrc/tc_conv.cu): replaces the
 * reference's conv forward (model.py:87-94) and the autograd backward of the three convolutions (trainer.py:310) for
 * observations with <= 4 channels.  forward: obs rows (NCHW; row i of the batch is obs[sample_index[i]], or obs[i] when
 * sample_index is NULL) -> feat (N, 64*oh*ow) in the reference's flatten order; the workspace keeps the activations.
 * backward: dfeat (N, 64*oh*ow) -> the six conv gradient slices of `grads` (overwritten), using the same workspace.
 * The weights are converted to tensor-core format (pre-split TF32 planes) into the workspace by pack_weights, or by a
 * forward with repack_weights = 1; rollout forwards (weights frozen) pack once and pass 0.
 * trxl_conv_train_supported returns 1 when the shape is covered. */
int trxl_conv_train_supported(const trxl_model_config* cfg, int H, int W);
int64_t trxl_conv_train_workspace_floats(const trxl_model_config* cfg, int N, int H, int W);
int trxl_conv_train_pack_weights(const trxl_model_config* cfg, const float* params, int N, int H, int W, float* workspace,
                                 void* stream);
int trxl_conv_train_forward(const trxl_model_config* cfg, const float* params, const float* obs, const int64_t* sample_index,
                            int N, int H, int W, float* workspace, float* feat, int repack_weights, void* stream);
int trxl_conv_train_backward(const trxl_model_config* cfg, float* grads, int N, int H, int W, float* workspace,
                             const float* dfeat, void* stream);

/* ---- the hot kernel on its own ---------------------------------------------------------------- */
/* Fused window gather + PE add + [LayerNorm] + q.K + mask + softmax(/sqrt(D)) + P.V with the K/V
 * projections folded onto the query side (MultiHeadAttention.forward transformer.py:31-86 for query
 * length 1).  qk (N,H,D) = per-head Q_h Wk_h; outputs probs (N,H,L) and ctx (N,H,D) = sum_l p x_l. */
int trxl_window_attention_forward(const float* table, int64_t slots, int num_blocks, int block, const int64_t* ep_index,
                                  const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index,
                                  const int64_t* sample_index, const float* pe_table, const float* qk, const float* qkb,
                                  int layer_norm_rows, int N, int L, int D, int H, float* probs, float* ctx, void* stream);
int trxl_window_attention_backward(const float* table, int64_t slots, int num_blocks, int block, const int64_t* ep_index,
                                   const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index,
                                   const int64_t* sample_index, const float* pe_table, const float* qk, const float* probs,
                                   const float* ctx, const float* dctx, int layer_norm_rows, int N, int L, int D, int H,
                                   float* dqk, float* dqkb, float* dpe, void* stream);

/* ---- building blocks (exported for tests and for callers that keep their own graph) ----------- */
/* y (M,N) = relu?(x (M,K) W(N,K)^T + bias) : nn.Linear forward */
int trxl_linear_forward(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, int relu, void* stream);
/* dx = dy W ; dW = dy^T x ; db = colsum(dy)   (any of dx/dW/db may be NULL); scratch >= 64*N+64 floats */
int trxl_linear_backward(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db, int M, int N, int K,
                         float* scratch, void* stream);
/* nn.LayerNorm(D) forward, eps 1e-5; saves mean/rstd (rows,).  backward scratch >= 128*D+64 floats */
int trxl_layernorm_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int rows,
                           int D, void* stream);
int trxl_layernorm_backward(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dx,
                            float* dgamma, float* dbeta, float* scratch, int rows, int D, void* stream);
/* batched_index_select(input, 1, index) (utils.py:52-75): out (N,L,inner) = in (N,slots,inner)[n, idx[n,l]] */
int trxl_gather_window(const float* in, const int64_t* index, float* out, int64_t N, int L, int64_t slots, int64_t inner, void* stream);
/* dst (rows, row_floats) = src[index[r]]  (minibatch gather of observations, buffer.py:92) */
int trxl_gather_rows(const float* src, const int64_t* index, float* dst, int64_t rows, int64_t row_floats, void* stream);

/* ---- PPO data path ---------------------------------------------------------------------------- */
/* Buffer.calc_advantages (buffer.py:95-113): rewards/values/adv (W,T) fp32, dones (W,T) bool.
 * Bit-identical to the reference's fp32 CPU result. */
int trxl_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_value, float* advantages,
             int W, int T, double gamma, double lamda, void* stream);
/* trainer.py:165-166: write mask/window-index rows of every worker's current episode step into
 * the rollout buffer (strided destinations), plus the episode id. */
int trxl_rollout_prepare(const int64_t* step, const int64_t* ep, const uint8_t* mask_table, const int64_t* index_table,
                         uint8_t* mask_out, int64_t mask_stride, int64_t* idx_out, int64_t idx_stride, int64_t* ep_out,
                         int64_t ep_stride, int W, int L, void* stream);
/* trainer.py:163 + the host->device hand-over of a rollout step in one kernel: observations (n, obs_floats) and the episode
 * cursors (n,) are read from `*_src` -- device memory, or pinned / cudaHostRegister'ed HOST memory read in place over PCIe
 * (pass the pointer from trxl_host_device_pointer) -- into the device staging buffers, and the observations also into their
 * rows of the rollout buffer (row w at obs_store + w * store_stride_floats). */
int trxl_rollout_fetch(const float* obs_src, int64_t obs_floats, const int64_t* step_src, const int64_t* ep_src, float* obs_dev,
                       float* obs_store, int64_t store_stride_floats, int64_t* step_dev, int64_t* ep_dev, int n, void* stream);
/* device-side address of a pinned / registered host buffer (cudaHostGetDevicePointer), for kernels that read or write host
 * memory in place: trxl_rollout_fetch sources, and `actions_compact` of trxl_sample_actions (actions land in host memory
 * without a copy node) */
int trxl_host_device_pointer(const void* host_ptr, void** device_ptr_out);
/* trainer.py:174: table[ep[w], step[w]] = new_mem[w] (inner = B*D floats) */
int trxl_memory_scatter(float* table, const int64_t* ep, const int64_t* step, const float* new_mem, int W, int64_t slots,
                        int64_t inner, void* stream);
/* The stores that end a rollout step in ONE kernel: the scatter above (trainer.py:174); if table_pe != NULL also
 * table_pe[ep[w], step[w], b, :] = new_mem[w, b, :] + pe_table[step[w], :] -- a second table that already carries the positional
 * rows (transformer.py:213-214 adds them to every window row of every forward), kept in step with the first, which
 * trxl_model_forward's inference path reads with pe_index == NULL (its window rows then come in by bulk async copies); and if
 * value_dst != NULL value_dst[w * value_stride] = value[w] (trainer.py:186: buffer.values[:, t] = value). */
int trxl_rollout_store(float* table, float* table_pe, const float* pe_table, const int64_t* ep, const int64_t* step,
                       const float* new_mem, int W, int64_t slots, int blocks, int dim, const float* value, float* value_dst,
                       int64_t value_stride, void* stream);
/* trainer.py:177-186: sample each branch from softmax(logits) with caller-supplied uniforms u (W, nb)
 * (inverse CDF); if forced_actions (W, nb) != NULL those actions are taken instead (trajectory replay)
 * and only their log-probabilities are computed. */
int trxl_sample_actions(const float* logits, const float* u, const int64_t* forced_actions, const int32_t* branch_sizes,
                        int num_branches, int64_t* actions, int64_t act_stride, float* log_probs, int64_t logp_stride,
                        int64_t* actions_compact, int W, void* stream);
/* Same, with a completion signal the host can poll in plain memory, without a CUDA call: the kernel takes the launch sequence
 * number seq = *done_counter + 1 (device memory; stored back), and every word of actions_compact (e.g. pinned host memory, see
 * trxl_host_device_pointer) is written as seq << 32 | action -- when all W * num_branches words carry seq, the step's actions are
 * there (8-byte stores are single-copy atomic; no system-wide fence on the step's critical path).  If done_flag != NULL (pinned
 * host memory) seq is additionally published there behind system-wide fences.  W * num_branches <= 1024. */
int trxl_sample_actions_notify(const float* logits, const float* u, const int64_t* forced_actions, const int32_t* branch_sizes,
                               int num_branches, int64_t* actions, int64_t act_stride, float* log_probs, int64_t logp_stride,
                               int64_t* actions_compact, int W, int64_t* done_counter, int64_t* done_flag, void* stream);
/* {sum a, sum a^2, count} of the minibatch advantages as doubles (all-reduce these across ranks) */
int trxl_adv_stats(const float* advantages, const int64_t* sample_index, int N, double* out3, void* stream);
/* trainer.py:277-304,315-316 forward + backward: stats6 = [policy_loss, vf_loss, loss, entropy,
 * approx_kl, clip_fraction]; dlogits (N, sum A), dvalue (N). scratch >= N/128*5+16 floats */
int trxl_ppo_loss(const float* logits, const float* value, const int64_t* actions, const float* old_log_probs,
                  const float* old_values, const float* advantages, const int64_t* sample_index, const double* adv_stats3,
                  const int32_t* branch_sizes, int num_branches, int N, double clip_range, double beta, double vf_coef,
                  float* dlogits, float* dvalue, float* stats6, float* scratch, void* stream);
/* clip_grad_norm_ + AdamW.step (trainer.py:311-312) + get_grad_norm sums (model.py:128-151) over the
 * flat arena.  chunks (nchunks,3) int64 {start,len,group}; norms (G+2): per-group norms of the
 * UNCLIPPED grads, total norm, clip coefficient.  Grads are overwritten with the clipped grads. */
int trxl_clip_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t total_floats,
                         const int64_t* chunks, int nchunks, int ngroups, double max_grad_norm, double lr, double beta1,
                         double beta2, double eps, double weight_decay, int64_t step, float* partial, float* norms, void* stream);

/* ---- episode-grouped attention on the tensor cores ------------------------------------------------------------------
 * transformer.py:59,73: for a minibatch sorted by episode, the energies / context contractions of all samples of one
 * episode are dense GEMMs against that episode's memory rows (fetched once per 128-row tile by TMA from the (E, M, B, D)
 * table) and run as 3xTF32 tcgen05 GEMMs; the per-row window softmax sits between them.  Applies to relative or no positional
 * encoding (a learned table needs gradients into the rows); other configurations ignore the grouping and use the per-sample
 * kernel. */
typedef struct trxl_attn_groups {
    const float* table_pe;     /* (E, M, B, D) from trxl_table_add_pe (layer_norm = 1 for pre-LN models); the table itself if   */
                               /*   there is neither a positional table nor a pre-LayerNorm                                    */
    int32_t n_episodes;        /* E                                                                                           */
    const int32_t* tiles;      /* (n_tiles, 4) {first (sample, head) row, rows, episode, 0}; rows of a tile share the episode, */
    int32_t n_tiles;           /*   a tile has at most 128 rows and starts at a multiple of num_heads                         */
    const int32_t* ranges;     /* (N, 4) from trxl_attention_ranges                                                           */
} trxl_attn_groups;
int trxl_grouped_attention_supported(const trxl_model_config* cfg);
/* out[e, m, b, :] = table[e, m, b, :] + pe_table[m, :] (pe_table may be NULL), and with layer_norm != 0 additionally normalised
 * per row (LayerNorm without affine, eps 1e-5: the norm_kv of a pre-LayerNorm block, transformer.py:131, whose gamma / beta the
 * model folds into Wk / Wv).  Once per update: the table is frozen during the optimisation epochs. */
int trxl_table_add_pe(const float* table, const float* pe_table, float* out, int64_t E, int M, int B, int D, int layer_norm,
                      void* stream);
/* ranges[n] = {first visible slot, visible slots, fully-masked flag, episode} of sample n (visible slots are contiguous:
 * trainer.py:78-90 builds masks as lower-triangular rows and windows as consecutive slots) */
int trxl_attention_ranges(const uint8_t* mask, const int64_t* win_index, const int64_t* ep_index, const int64_t* sample_index,
                          int N, int L, int32_t* ranges4, void* stream);
/* trxl_model_forward / trxl_model_backward with the grouping (samples must be ordered so that every tile's rows are contiguous) */
int trxl_model_forward_grouped(const trxl_model_config* cfg, const float* params, const float* feat, const float* table, int64_t slots,
                               const int64_t* ep_index, const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index,
                               const int64_t* sample_index, const float* pe_table, int N, float* workspace, float* logits,
                               float* value, float* out_mem, const trxl_attn_groups* groups, void* stream);
int trxl_model_backward_grouped(const trxl_model_config* cfg, const float* params, float* grads, const float* feat, const float* table,
                                int64_t slots, const int64_t* ep_index, const int64_t* win_index, const uint8_t* mask,
                                const int64_t* pe_index, const int64_t* sample_index, const float* pe_table, int N, float* workspace,
                                const float* out_mem, const float* dlogits, const float* dvalue, float* dfeat,
                                const trxl_attn_groups* groups, void* stream);

/* ---- multi-GPU exchange (SURVEY.md §8b "trxl_allreduce_grads", §8e) -------------------------------
 * The reference is single-process; data-parallel sharding over workers (trainer.py:145-323 per rank) needs ONE
 * exchange per optimiser step: the in-place sum of the flat fp32 gradient arena (+ the 6 loss statistics appended to
 * its tail).  NCCL is bound with dlopen at first use (libnccl.so.2; TRXL_NCCL_LIB overrides), so single-GPU users never
 * load it.  Rendezvous: rank 0 calls trxl_comm_unique_id and ships the TRXL_COMM_ID_BYTES host bytes to the other ranks
 * by any means (torch.distributed, a file, MPI); every rank then calls trxl_comm_create with its CUDA device current. */
int trxl_comm_unique_id(void* id_out);
int trxl_comm_create(const void* id_bytes, int rank, int world_size, void** comm_out);
int trxl_comm_destroy(void* comm);
/* number of collectives enqueued through this communicator so far (bench/tests count them) */
int64_t trxl_comm_calls(void* comm);
/* NCCL version code of the bound library (e.g. 22809), or -1 if NCCL cannot be loaded */
int trxl_comm_nccl_version(void);
/* in-place sum over ranks on `stream` (asynchronous; ordered with the kernels already enqueued on it) */
int trxl_allreduce_grads(void* comm, float* buf, int64_t count, void* stream);
int trxl_allreduce_f64(void* comm, double* buf, int64_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRXL_PPO_H */
