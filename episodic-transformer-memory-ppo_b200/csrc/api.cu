// extern "C" surface of libtrxlppo (declared in include/trxl_ppo.h).  Thin argument marshalling only.
#include <mutex>
#include <unordered_map>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/trxl_ppo.h"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "conv.cuh"
#include "tc_conv.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "model.cuh"
#include "ppo.cuh"

static thread_local char g_err[512] = "";

void trxl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

long long g_trxl_launches = 0;
static std::unordered_map<void*, long long> g_graph_nodes;      // kernel / copy nodes of every instantiated graph
static std::mutex g_graph_nodes_mutex;

// ---- optional launch timing of the two attention kernels (bench.py roofline) ----
namespace {
constexpr int PROF_KINDS = 4, PROF_CAP = 8192;
struct ProfSlot { cudaEvent_t a, b; int n; long long aux; };
bool g_prof_on = false;
ProfSlot* g_prof[PROF_KINDS] = {nullptr, nullptr, nullptr, nullptr};
int g_prof_count[PROF_KINDS] = {0, 0, 0, 0};
}  // namespace

static bool stream_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}

void trxl_prof_begin(int kind, int n, cudaStream_t st) {
    if (!g_prof_on || g_prof_count[kind] >= PROF_CAP || stream_capturing(st)) return;
    if (!g_prof[kind]) {
        g_prof[kind] = new ProfSlot[PROF_CAP];
        for (int i = 0; i < PROF_CAP; ++i) { cudaEventCreate(&g_prof[kind][i].a); cudaEventCreate(&g_prof[kind][i].b); }
    }
    ProfSlot& s = g_prof[kind][g_prof_count[kind]];
    s.n = n;
    s.aux = 0;
    cudaEventRecord(s.a, st);
}
void trxl_prof_aux(int kind, long long aux) {
    if (!g_prof_on || g_prof_count[kind] >= PROF_CAP || !g_prof[kind]) return;
    g_prof[kind][g_prof_count[kind]].aux = aux;
}
void trxl_prof_end(int kind, cudaStream_t st) {
    if (!g_prof_on || g_prof_count[kind] >= PROF_CAP || !g_prof[kind] || stream_capturing(st)) return;
    cudaEventRecord(g_prof[kind][g_prof_count[kind]].b, st);
    ++g_prof_count[kind];
}

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
typedef const long long* cll;

static int branch_spec(const int32_t* sizes, int n, BranchSpec& bs, int* sumA) {
    TRXL_CHECK_ARG(sizes && n >= 1 && n <= TRXL_MAX_BRANCHES, "bad branch spec (n=%d)", n);
    bs.n = n;
    int off = 0;
    for (int k = 0; k < n; ++k) {
        TRXL_CHECK_ARG(sizes[k] > 0, "branch %d has no actions", k);
        bs.off[k] = off; bs.size[k] = sizes[k]; off += sizes[k];
    }
    *sumA = off;
    return TRXL_OK;
}

extern "C" {

const char* trxl_last_error(void) { return g_err; }
int trxl_abi_version(void) { return TRXL_ABI_VERSION; }
int64_t trxl_launch_count(void) { return g_trxl_launches; }
extern long long g_trxl_tc_launches;
int64_t trxl_tc_gemm_launches(void) { return g_trxl_tc_launches; }

// ---- CUDA graph capture of a sequence of this library's launches (rollout step replay) ----
int trxl_graph_begin(void* stream) {
    cudaError_t e = cudaStreamBeginCapture(S(stream), cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) { trxl_set_error("graph_begin: %s", cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
    return TRXL_OK;
}
int trxl_graph_end(void* stream, void** graph_exec_out) {
    TRXL_CHECK_ARG(graph_exec_out, "graph_end: null output");
    *graph_exec_out = nullptr;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(S(stream), &graph);
    if (e != cudaSuccess || !graph) {
        cudaGetLastError();
        trxl_set_error("graph_end: capture failed: %s", cudaGetErrorString(e));
        return TRXL_ERR_CUDA;
    }
    size_t nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &nodes);
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { trxl_set_error("graph_end: instantiate failed: %s", cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
    {
        // trxl_launch_count() counted the captured launches once, at capture time; every replay adds the graph's node count
        std::lock_guard<std::mutex> lock(g_graph_nodes_mutex);
        g_graph_nodes[exec] = (long long)nodes;
    }
    *graph_exec_out = exec;
    return TRXL_OK;
}
int trxl_graph_launch(void* graph_exec, void* stream) {
    TRXL_CHECK_ARG(graph_exec, "graph_launch: null graph");
    cudaError_t e = cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), S(stream));
    if (e != cudaSuccess) { trxl_set_error("graph_launch: %s", cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
    {
        std::lock_guard<std::mutex> lock(g_graph_nodes_mutex);
        auto it = g_graph_nodes.find(graph_exec);
        if (it != g_graph_nodes.end()) g_trxl_launches += it->second;
    }
    return TRXL_OK;
}
int trxl_graph_destroy(void* graph_exec) {
    if (graph_exec) {
        {
            std::lock_guard<std::mutex> lock(g_graph_nodes_mutex);
            g_graph_nodes.erase(graph_exec);
        }
        cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec));
    }
    return TRXL_OK;
}
// dst[r, :row_bytes] = src[r, :row_bytes] for strided rows (buffer[:, t] = x without a torch op inside a capture)
int trxl_copy_rows(const void* src, void* dst, int64_t rows, int64_t row_bytes, int64_t src_stride_bytes, int64_t dst_stride_bytes,
                   void* stream) {
    TRXL_CHECK_ARG(src && dst && rows >= 0 && row_bytes >= 0, "copy_rows: bad arguments");
    if (rows == 0 || row_bytes == 0) return TRXL_OK;
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dst_stride_bytes, src, (size_t)src_stride_bytes, (size_t)row_bytes, (size_t)rows,
                                      cudaMemcpyDeviceToDevice, S(stream));
    ++g_trxl_launches;
    if (e != cudaSuccess) { trxl_set_error("copy_rows: %s", cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
    return TRXL_OK;
}

int trxl_copy_async(const void* src, void* dst, int64_t bytes, void* stream) {
    TRXL_CHECK_ARG(src && dst && bytes >= 0, "copy_async: bad arguments");
    if (bytes == 0) return TRXL_OK;
    // cudaMemcpyDefault: the direction comes from unified addressing, so pinned (or cudaHostRegister'ed) host buffers and
    // device buffers can be mixed; capturable into a CUDA graph as a memcpy node
    cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, S(stream));
    ++g_trxl_launches;
    if (e != cudaSuccess) { trxl_set_error("copy_async: %s", cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
    return TRXL_OK;
}

int trxl_profile_enable(int on) {
    g_prof_on = on != 0;
    if (on) { for (int k = 0; k < PROF_KINDS; ++k) g_prof_count[k] = 0; }
    return TRXL_OK;
}
int trxl_profile_read(int kind, int min_samples, double* total_ms, int64_t* launches, int64_t* samples) {
    TRXL_CHECK_ARG(kind >= 0 && kind < PROF_KINDS && total_ms && launches && samples, "profile_read: bad arguments");
    *total_ms = 0.0; *launches = 0; *samples = 0;
    for (int i = 0; i < g_prof_count[kind]; ++i) {
        ProfSlot& s = g_prof[kind][i];
        if (s.n < min_samples) continue;
        if (cudaEventSynchronize(s.b) != cudaSuccess) { trxl_set_error("profile_read: event sync failed"); return TRXL_ERR_CUDA; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) { trxl_set_error("profile_read: elapsed failed"); return TRXL_ERR_CUDA; }
        *total_ms += ms; *launches += 1; *samples += s.n;
    }
    return TRXL_OK;
}

int64_t trxl_profile_aux(int kind, int min_samples) {
    if (kind < 0 || kind >= PROF_KINDS) return -1;
    long long total = 0;
    for (int i = 0; i < g_prof_count[kind]; ++i)
        if (g_prof[kind][i].n >= min_samples) total += g_prof[kind][i].aux;
    return total;
}

int trxl_layout_num_entries(const trxl_model_config* cfg) {
    std::vector<trxl_param_entry> e;
    int rc = model_layout(cfg, e, nullptr, nullptr);
    return rc == TRXL_OK ? (int)e.size() : rc;
}
int64_t trxl_layout_total_floats(const trxl_model_config* cfg) {
    std::vector<trxl_param_entry> e;
    long long total = 0;
    int rc = model_layout(cfg, e, &total, nullptr);
    return rc == TRXL_OK ? total : rc;
}
int trxl_layout_entry(const trxl_model_config* cfg, int index, trxl_param_entry* out) {
    std::vector<trxl_param_entry> e;
    TRXL_PROPAGATE(model_layout(cfg, e, nullptr, nullptr));
    TRXL_CHECK_ARG(out && index >= 0 && index < (int)e.size(), "layout entry %d out of range", index);
    *out = e[index];
    return TRXL_OK;
}
int trxl_layout_groups(const trxl_model_config* cfg) {
    std::vector<trxl_param_entry> e;
    int g = 0;
    int rc = model_layout(cfg, e, nullptr, &g);
    return rc == TRXL_OK ? g : rc;
}
int64_t trxl_workspace_floats(const trxl_model_config* cfg, int N) { return model_workspace_floats(cfg, N); }
int trxl_fused_forward_supported(const trxl_model_config* cfg) { return model_fused_supported(cfg); }

int trxl_model_forward(const trxl_model_config* cfg, const float* params, const float* feat, const float* table, int64_t slots,
                       const int64_t* ep_index, const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index,
                       const int64_t* sample_index, const float* pe_table, int N, float* workspace, float* logits, float* value,
                       float* out_mem, void* stream) {
    ModelIO io;
    io.N = N; io.feat = feat; io.table = table; io.slots = slots; io.ep_index = (cll)ep_index; io.win_index = (cll)win_index;
    io.mask = mask; io.pe_index = (cll)pe_index; io.sample_index = (cll)sample_index; io.pe_table = pe_table;
    return model_forward(cfg, params, io, workspace, logits, value, out_mem, S(stream));
}

static void set_groups(ModelIO& io, const trxl_attn_groups* g) {
    if (!g) return;
    io.table_pe = g->table_pe; io.n_episodes = g->n_episodes; io.tiles = reinterpret_cast<const int4*>(g->tiles);
    io.n_tiles = g->n_tiles; io.ranges = reinterpret_cast<const int4*>(g->ranges);
}

int trxl_model_forward_grouped(const trxl_model_config* cfg, const float* params, const float* feat, const float* table, int64_t slots,
                               const int64_t* ep_index, const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index,
                               const int64_t* sample_index, const float* pe_table, int N, float* workspace, float* logits,
                               float* value, float* out_mem, const trxl_attn_groups* groups, void* stream) {
    ModelIO io;
    io.N = N; io.feat = feat; io.table = table; io.slots = slots; io.ep_index = (cll)ep_index; io.win_index = (cll)win_index;
    io.mask = mask; io.pe_index = (cll)pe_index; io.sample_index = (cll)sample_index; io.pe_table = pe_table;
    set_groups(io, groups);
    return model_forward(cfg, params, io, workspace, logits, value, out_mem, S(stream));
}

int trxl_model_backward_grouped(const trxl_model_config* cfg, const float* params, float* grads, const float* feat, const float* table,
                                int64_t slots, const int64_t* ep_index, const int64_t* win_index, const uint8_t* mask,
                                const int64_t* pe_index, const int64_t* sample_index, const float* pe_table, int N, float* workspace,
                                const float* out_mem, const float* dlogits, const float* dvalue, float* dfeat,
                                const trxl_attn_groups* groups, void* stream) {
    ModelIO io;
    io.N = N; io.feat = feat; io.table = table; io.slots = slots; io.ep_index = (cll)ep_index; io.win_index = (cll)win_index;
    io.mask = mask; io.pe_index = (cll)pe_index; io.sample_index = (cll)sample_index; io.pe_table = pe_table;
    set_groups(io, groups);
    return model_backward(cfg, params, grads, io, workspace, out_mem, dlogits, dvalue, dfeat, S(stream));
}

int trxl_attention_ranges(const uint8_t* mask, const int64_t* win_index, const int64_t* ep_index, const int64_t* sample_index, int N,
                          int L, int32_t* ranges4, void* stream) {
    return attn_tc_ranges(mask, (cll)win_index, (cll)ep_index, (cll)sample_index, N, L, reinterpret_cast<int4*>(ranges4), S(stream));
}

int trxl_table_add_pe(const float* table, const float* pe_table, float* out, int64_t E, int M, int B, int D, int layer_norm,
                      void* stream) {
    return attn_tc_table_add_pe(table, pe_table, out, E, M, B, D, layer_norm, S(stream));
}

int trxl_grouped_attention_supported(const trxl_model_config* cfg) {
    return cfg && cfg->pos_enc != TRXL_PE_LEARNED &&
           attn_tc_supported(cfg->embed_dim, cfg->num_heads, cfg->max_episode_steps, cfg->num_blocks) ? 1 : 0;
}

int trxl_model_backward(const trxl_model_config* cfg, const float* params, float* grads, const float* feat, const float* table,
                        int64_t slots, const int64_t* ep_index, const int64_t* win_index, const uint8_t* mask,
                        const int64_t* pe_index, const int64_t* sample_index, const float* pe_table, int N, float* workspace,
                        const float* out_mem, const float* dlogits, const float* dvalue, float* dfeat, void* stream) {
    ModelIO io;
    io.N = N; io.feat = feat; io.table = table; io.slots = slots; io.ep_index = (cll)ep_index; io.win_index = (cll)win_index;
    io.mask = mask; io.pe_index = (cll)pe_index; io.sample_index = (cll)sample_index; io.pe_table = pe_table;
    return model_backward(cfg, params, grads, io, workspace, out_mem, dlogits, dvalue, dfeat, S(stream));
}

int64_t trxl_conv_encoder_workspace_floats(const trxl_model_config* cfg, int N, int H, int W) {
    if (!cfg || cfg->conv_in_channels <= 0 || N < 0) return -1;
    return conv_encoder_workspace_floats(N, cfg->conv_in_channels, H, W);
}

int trxl_conv_encoder_forward(const trxl_model_config* cfg, const float* params, const float* obs, int N, int H, int W,
                              float* workspace, float* feat, void* stream) {
    TRXL_CHECK_ARG(cfg && cfg->conv_in_channels > 0, "conv_encoder: config has no convolutional encoder");
    TRXL_CHECK_ARG(params && obs && workspace && feat, "conv_encoder: null pointer");
    std::vector<trxl_param_entry> e;
    TRXL_PROPAGATE(model_layout(cfg, e, nullptr, nullptr));
    long long off[6] = {-1, -1, -1, -1, -1, -1};
    const char* names[6] = {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias"};
    for (const auto& en : e)
        for (int i = 0; i < 6; ++i)
            if (strcmp(en.name, names[i]) == 0) off[i] = en.offset;
    for (int i = 0; i < 6; ++i) TRXL_CHECK_ARG(off[i] >= 0, "conv_encoder: parameter %s missing from the layout", names[i]);
    return conv_encoder_forward(S(stream), params + off[0], params + off[1], params + off[2], params + off[3], params + off[4],
                                params + off[5], obs, N, cfg->conv_in_channels, H, W, workspace, feat);
}

static int conv_param_offsets(const trxl_model_config* cfg, long long (&off)[6]) {
    std::vector<trxl_param_entry> e;
    TRXL_PROPAGATE(model_layout(cfg, e, nullptr, nullptr));
    const char* names[6] = {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias"};
    for (int i = 0; i < 6; ++i) off[i] = -1;
    for (const auto& en : e)
        for (int i = 0; i < 6; ++i)
            if (strcmp(en.name, names[i]) == 0) off[i] = en.offset;
    for (int i = 0; i < 6; ++i) TRXL_CHECK_ARG(off[i] >= 0, "conv_train: parameter %s missing from the layout", names[i]);
    return TRXL_OK;
}

int trxl_conv_train_supported(const trxl_model_config* cfg, int H, int W) {
    if (!cfg || cfg->conv_in_channels <= 0) return 0;
    return tc_conv_supported(cfg->conv_in_channels, H, W);
}

int64_t trxl_conv_train_workspace_floats(const trxl_model_config* cfg, int N, int H, int W) {
    if (!cfg || cfg->conv_in_channels <= 0 || N < 0) return -1;
    return tc_conv_workspace_floats(N, cfg->conv_in_channels, H, W);
}

int trxl_conv_train_pack_weights(const trxl_model_config* cfg, const float* params, int N, int H, int W, float* workspace,
                                 void* stream) {
    TRXL_CHECK_ARG(cfg && cfg->conv_in_channels > 0, "conv_train: config has no convolutional encoder");
    TRXL_CHECK_ARG(params && workspace, "conv_train: null pointer");
    long long off[6];
    TRXL_PROPAGATE(conv_param_offsets(cfg, off));
    const float* p[6];
    for (int i = 0; i < 6; ++i) p[i] = params + off[i];
    return tc_conv_pack_weights(S(stream), p, N, cfg->conv_in_channels, H, W, workspace);
}

int trxl_conv_train_forward(const trxl_model_config* cfg, const float* params, const float* obs, const int64_t* sample_index, int N,
                            int H, int W, float* workspace, float* feat, int repack_weights, void* stream) {
    TRXL_CHECK_ARG(cfg && cfg->conv_in_channels > 0, "conv_train: config has no convolutional encoder");
    TRXL_CHECK_ARG(params && obs && workspace && feat, "conv_train: null pointer");
    long long off[6];
    TRXL_PROPAGATE(conv_param_offsets(cfg, off));
    const float* p[6];
    for (int i = 0; i < 6; ++i) p[i] = params + off[i];
    return tc_conv_forward(S(stream), p, obs, (cll)sample_index, N, cfg->conv_in_channels, H, W, workspace, feat, repack_weights);
}

int trxl_conv_train_backward(const trxl_model_config* cfg, float* grads, int N, int H, int W, float* workspace, const float* dfeat,
                             void* stream) {
    TRXL_CHECK_ARG(cfg && cfg->conv_in_channels > 0, "conv_train: config has no convolutional encoder");
    TRXL_CHECK_ARG(grads && workspace && dfeat, "conv_train: null pointer");
    long long off[6];
    TRXL_PROPAGATE(conv_param_offsets(cfg, off));
    float* g[6];
    for (int i = 0; i < 6; ++i) g[i] = grads + off[i];
    return tc_conv_backward(S(stream), g, N, cfg->conv_in_channels, H, W, workspace, dfeat);
}

static AttnArgs make_attn(const float* table, int64_t slots, int num_blocks, int block, const int64_t* ep_index,
                          const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index, const int64_t* sample_index,
                          const float* pe_table, const float* qk, const float* qkb, int ln, int N, int L, int D, int H,
                          float* probs, float* ctx) {
    AttnArgs a;
    a.N = N; a.L = L; a.D = D; a.H = H; a.B = num_blocks; a.blk = block; a.table = table; a.slots = slots;
    a.ep_index = (cll)ep_index; a.win_index = (cll)win_index; a.mask = mask; a.pe_index = pe_table ? (cll)pe_index : nullptr;
    a.pe = pe_table; a.sample_index = (cll)sample_index; a.qk = qk; a.qkb = qkb; a.ln = ln;
    a.scale = (float)sqrt((double)D); a.probs = probs; a.ctx = ctx;
    return a;
}

int trxl_window_attention_forward(const float* table, int64_t slots, int num_blocks, int block, const int64_t* ep_index,
                                  const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index,
                                  const int64_t* sample_index, const float* pe_table, const float* qk, const float* qkb,
                                  int layer_norm_rows, int N, int L, int D, int H, float* probs, float* ctx, void* stream) {
    TRXL_CHECK_ARG(block >= 0 && block < num_blocks, "window_attention: block %d out of range", block);
    return trxl_window_attn_fwd(make_attn(table, slots, num_blocks, block, ep_index, win_index, mask, pe_index, sample_index,
                                          pe_table, qk, qkb, layer_norm_rows, N, L, D, H, probs, ctx), S(stream));
}

int trxl_window_attention_backward(const float* table, int64_t slots, int num_blocks, int block, const int64_t* ep_index,
                                   const int64_t* win_index, const uint8_t* mask, const int64_t* pe_index,
                                   const int64_t* sample_index, const float* pe_table, const float* qk, const float* probs,
                                   const float* ctx, const float* dctx, int layer_norm_rows, int N, int L, int D, int H,
                                   float* dqk, float* dqkb, float* dpe, void* stream) {
    TRXL_CHECK_ARG(block >= 0 && block < num_blocks, "window_attention: block %d out of range", block);
    AttnArgs a = make_attn(table, slots, num_blocks, block, ep_index, win_index, mask, pe_index, sample_index, pe_table, qk,
                           nullptr, layer_norm_rows, N, L, D, H, const_cast<float*>(probs), const_cast<float*>(ctx));
    AttnBwdArgs g;
    g.dctx = dctx; g.dqk = dqk; g.dqkb = dqkb; g.dpe = dpe;
    return trxl_window_attn_bwd(a, g, S(stream));
}

int trxl_linear_forward(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, int relu, void* stream) {
    TRXL_CHECK_ARG(x && W && y, "linear_forward: null pointer");
    return gemm_nt(S(stream), M, N, K, x, K, W, K, y, N, bias, relu);
}

int trxl_linear_backward(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db, int M, int N, int K,
                         float* scratch, void* stream) {
    TRXL_CHECK_ARG(dy, "linear_backward: null dy");
    if (dx) { TRXL_CHECK_ARG(W, "linear_backward: dx needs W"); TRXL_PROPAGATE(gemm_nn(S(stream), M, K, N, dy, N, W, K, dx, K)); }
    if (dW) { TRXL_CHECK_ARG(x, "linear_backward: dW needs x"); TRXL_PROPAGATE(gemm_tn(S(stream), N, K, M, dy, N, x, K, dW, K)); }
    if (db) { TRXL_CHECK_ARG(scratch, "linear_backward: db needs scratch"); TRXL_PROPAGATE(ew_colsum(S(stream), dy, N, db, M, N, 1.f, 0, scratch)); }
    return TRXL_OK;
}

int trxl_layernorm_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int rows,
                           int D, void* stream) {
    TRXL_CHECK_ARG(x && gamma && beta && y, "layernorm_forward: null pointer");
    return ew_layernorm_fwd(S(stream), x, D, nullptr, 0, gamma, beta, y, D, nullptr, 0, mean, rstd, rows, D);
}

int trxl_layernorm_backward(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dx,
                            float* dgamma, float* dbeta, float* scratch, int rows, int D, void* stream) {
    TRXL_CHECK_ARG(dy && x && mean && rstd && gamma && dx, "layernorm_backward: null pointer");
    TRXL_CHECK_ARG(!dgamma || (dbeta && scratch), "layernorm_backward: dgamma needs dbeta and scratch");
    return ew_layernorm_bwd(S(stream), dy, D, x, D, mean, rstd, gamma, dx, D, 0, dgamma, dbeta, 0, scratch, rows, D);
}

int trxl_gather_window(const float* in, const int64_t* index, float* out, int64_t N, int L, int64_t slots, int64_t inner, void* stream) {
    TRXL_CHECK_ARG(in && index && out, "gather_window: null pointer");
    return ppo_gather_window(S(stream), in, (cll)index, out, N, L, slots, inner);
}
int trxl_gather_rows(const float* src, const int64_t* index, float* dst, int64_t rows, int64_t row_floats, void* stream) {
    TRXL_CHECK_ARG(src && index && dst, "gather_rows: null pointer");
    return ppo_gather_rows(S(stream), src, (cll)index, dst, rows, row_floats);
}

int trxl_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_value, float* advantages, int W,
             int T, double gamma, double lamda, void* stream) {
    TRXL_CHECK_ARG(rewards && dones && values && last_value && advantages, "gae: null pointer");
    return ppo_gae(S(stream), rewards, dones, values, last_value, advantages, W, T, gamma, lamda);
}

int trxl_rollout_prepare(const int64_t* step, const int64_t* ep, const uint8_t* mask_table, const int64_t* index_table,
                         uint8_t* mask_out, int64_t mask_stride, int64_t* idx_out, int64_t idx_stride, int64_t* ep_out,
                         int64_t ep_stride, int W, int L, void* stream) {
    TRXL_CHECK_ARG(step && mask_table && index_table && mask_out && idx_out, "rollout_prepare: null pointer");
    return ppo_rollout_prepare(S(stream), (cll)step, (cll)ep, mask_table, (cll)index_table, mask_out, mask_stride,
                               (long long*)idx_out, idx_stride, (long long*)ep_out, ep_stride, W, L);
}

int trxl_rollout_fetch(const float* obs_src, int64_t obs_floats, const int64_t* step_src, const int64_t* ep_src, float* obs_dev,
                       float* obs_store, int64_t store_stride_floats, int64_t* step_dev, int64_t* ep_dev, int n, void* stream) {
    TRXL_CHECK_ARG(obs_src && step_src && ep_src && obs_dev && obs_store && step_dev && ep_dev, "rollout_fetch: null pointer");
    return ppo_rollout_fetch(S(stream), obs_src, obs_floats, (cll)step_src, (cll)ep_src, obs_dev, obs_store, store_stride_floats,
                             (long long*)step_dev, (long long*)ep_dev, n);
}

int trxl_host_device_pointer(const void* host_ptr, void** device_ptr_out) {
    TRXL_CHECK_ARG(host_ptr && device_ptr_out, "host_device_pointer: null pointer");
    void* d = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&d, const_cast<void*>(host_ptr), 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        trxl_set_error("host_device_pointer: %s (is the buffer pinned or cudaHostRegister'ed?)", cudaGetErrorString(e));
        return TRXL_ERR_CUDA;
    }
    *device_ptr_out = d;
    return TRXL_OK;
}

int trxl_memory_scatter(float* table, const int64_t* ep, const int64_t* step, const float* new_mem, int W, int64_t slots,
                        int64_t inner, void* stream) {
    TRXL_CHECK_ARG(table && ep && step && new_mem, "memory_scatter: null pointer");
    return ppo_memory_scatter(S(stream), table, (cll)ep, (cll)step, new_mem, W, slots, inner);
}

int trxl_rollout_store(float* table, float* table_pe, const float* pe_table, const int64_t* ep, const int64_t* step,
                       const float* new_mem, int W, int64_t slots, int blocks, int dim, const float* value, float* value_dst,
                       int64_t value_stride, void* stream) {
    TRXL_CHECK_ARG(table && ep && step && new_mem, "rollout_store: null pointer");
    TRXL_CHECK_ARG(!table_pe || pe_table, "rollout_store: table_pe given without pe_table");
    TRXL_CHECK_ARG(!value_dst || value, "rollout_store: value_dst given without value");
    TRXL_CHECK_ARG(blocks > 0 && dim > 0 && slots > 0, "rollout_store: bad shape");
    return ppo_rollout_store(S(stream), table, table_pe, pe_table, (cll)ep, (cll)step, new_mem, W, slots, blocks, dim, value, value_dst,
                             value_stride);
}

int trxl_sample_actions(const float* logits, const float* u, const int64_t* forced_actions, const int32_t* branch_sizes,
                        int num_branches, int64_t* actions, int64_t act_stride, float* log_probs, int64_t logp_stride,
                        int64_t* actions_compact, int W, void* stream) {
    TRXL_CHECK_ARG(logits && (u || forced_actions) && actions && log_probs, "sample_actions: null pointer");
    BranchSpec bs; int sumA = 0;
    TRXL_PROPAGATE(branch_spec(branch_sizes, num_branches, bs, &sumA));
    return ppo_sample_actions(S(stream), logits, sumA, u, (cll)forced_actions, bs, (long long*)actions, act_stride, log_probs, logp_stride,
                              (long long*)actions_compact, W);
}

int trxl_sample_actions_notify(const float* logits, const float* u, const int64_t* forced_actions, const int32_t* branch_sizes,
                               int num_branches, int64_t* actions, int64_t act_stride, float* log_probs, int64_t logp_stride,
                               int64_t* actions_compact, int W, int64_t* done_counter, int64_t* done_flag, void* stream) {
    TRXL_CHECK_ARG(logits && (u || forced_actions) && actions && log_probs && done_counter && actions_compact, "sample_actions_notify: null pointer");
    BranchSpec bs;
    int sumA = 0;
    TRXL_PROPAGATE(branch_spec(branch_sizes, num_branches, bs, &sumA));
    return ppo_sample_actions(S(stream), logits, sumA, u, (cll)forced_actions, bs, (long long*)actions, act_stride, log_probs, logp_stride,
                              (long long*)actions_compact, W, (long long*)done_counter, (long long*)done_flag);
}

int trxl_adv_stats(const float* advantages, const int64_t* sample_index, int N, double* out3, void* stream) {
    TRXL_CHECK_ARG(advantages && out3 && N > 0, "adv_stats: bad arguments");
    return ppo_adv_stats(S(stream), advantages, (cll)sample_index, N, out3);
}

int trxl_ppo_loss(const float* logits, const float* value, const int64_t* actions, const float* old_log_probs,
                  const float* old_values, const float* advantages, const int64_t* sample_index, const double* adv_stats3,
                  const int32_t* branch_sizes, int num_branches, int N, double clip_range, double beta, double vf_coef,
                  float* dlogits, float* dvalue, float* stats6, float* scratch, void* stream) {
    TRXL_CHECK_ARG(logits && value && actions && old_log_probs && old_values && advantages && adv_stats3 && stats6 && scratch,
                   "ppo_loss: null pointer");
    PpoLossArgs a;
    TRXL_PROPAGATE(branch_spec(branch_sizes, num_branches, a.bs, &a.sumA));
    a.N = N; a.logits = logits; a.value = value; a.actions = (cll)actions; a.old_logp = old_log_probs; a.old_values = old_values;
    a.adv = advantages; a.sidx = (cll)sample_index; a.advstats = adv_stats3;
    a.clip = (float)clip_range; a.clip_lo = (float)(1.0 - clip_range); a.clip_hi = (float)(1.0 + clip_range);
    a.beta = (float)beta; a.vf_coef = (float)vf_coef; a.dlogits = dlogits; a.dvalue = dvalue; a.partial = scratch;
    return ppo_loss(S(stream), a, stats6);
}

int trxl_clip_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t total_floats,
                         const int64_t* chunks, int nchunks, int ngroups, double max_grad_norm, double lr, double beta1,
                         double beta2, double eps, double weight_decay, int64_t step, float* partial, float* norms, void* stream) {
    TRXL_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && total_floats > 0 && step >= 1, "clip_adamw: bad arguments");
    TRXL_CHECK_ARG(nchunks == 0 || (chunks && partial && norms), "clip_adamw: clipping needs chunks/partial/norms");
    AdamWArgs h;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    h.decay = (float)(1.0 - lr * weight_decay);
    h.one_minus_b1 = (float)(1.0 - beta1);
    h.b2 = (float)beta2;
    h.one_minus_b2 = (float)(1.0 - beta2);
    h.bc2_sqrt = (float)sqrt(bc2);
    h.eps = (float)eps;
    h.step_size = (float)(lr / bc1);
    return ppo_clip_adamw(S(stream), params, grads, exp_avg, exp_avg_sq, total_floats, (cll)chunks, nchunks, ngroups,
                          (float)max_grad_norm, partial, norms, h);
}

}  // extern "C"
