// Parameter-arena layout, activation workspace and the forward/backward launch sequences of the
// TrXL actor-critic trunk (everything after the CNN encoder).
//
// Reference graph being replaced: model.py:97-110 (lin_hidden, heads), transformer.py:222-253
// (embedding, PE, block loop), transformer.py:117-172 (block), transformer.py:287-298 (GRU gate),
// and its autograd backward (trainer.py:310).  Memory windows are read in place from the episode
// table by the attention kernel (attention.cu); see that file for the query-side fold.
#include <string.h>

#include <string>
#include <vector>

#include "../../include/trxl_ppo.h"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "model.cuh"
#include "rollout_fused.cuh"

namespace {

struct GateP { long long Wr, Ur, Ug, bg; };      // [Wr;Wz;Wg] contiguous at Wr, [Ur;Uz] contiguous at Ur
struct BlockP {
    long long Wv, Wk, Wq, Wo, bo;
    GateP g1, g2;
    long long n1w, n1b, n2w, n2b, nkw, nkb;
    long long Wff, bff;
};
struct Layout {
    long long conv[6];
    long long Wh, bh, We, be, pos;
    std::vector<BlockP> blk;
    long long Wp, bp, Wlv, blv, Wbr, bbr, wval, bval;
    long long total;
    int sumA, groups;
    std::vector<trxl_param_entry> entries;
};

int validate(const trxl_model_config* c) {
    TRXL_CHECK_ARG(c != nullptr, "config is NULL");
    TRXL_CHECK_ARG(c->embed_dim > 0 && c->embed_dim % 4 == 0 && c->embed_dim <= 1024, "embed_dim must be a multiple of 4 in (0,1024], got %d", c->embed_dim);
    TRXL_CHECK_ARG(c->num_heads > 0 && c->embed_dim % c->num_heads == 0, "embed_dim %d not divisible by num_heads %d", c->embed_dim, c->num_heads);
    TRXL_CHECK_ARG(c->num_blocks > 0 && c->num_blocks <= 64, "num_blocks out of range: %d", c->num_blocks);
    TRXL_CHECK_ARG(c->memory_length > 0, "memory_length must be positive");
    TRXL_CHECK_ARG(c->hidden_size > 0 && c->hidden_size % 4 == 0, "hidden_layer_size must be a positive multiple of 4, got %d", c->hidden_size);
    TRXL_CHECK_ARG(c->feat_dim > 0, "feat_dim must be positive");
    TRXL_CHECK_ARG(c->layer_norm >= 0 && c->layer_norm <= 2, "bad layer_norm mode %d", c->layer_norm);
    TRXL_CHECK_ARG(c->pos_enc >= 0 && c->pos_enc <= 2, "bad positional_encoding mode %d", c->pos_enc);
    TRXL_CHECK_ARG(c->num_branches >= 1 && c->num_branches <= TRXL_MAX_BRANCHES, "num_branches out of range: %d", c->num_branches);
    for (int k = 0; k < c->num_branches; ++k) TRXL_CHECK_ARG(c->branch_sizes[k] > 0, "branch %d has no actions", k);
    TRXL_CHECK_ARG(c->pos_enc == TRXL_PE_NONE || c->max_episode_steps > 0, "max_episode_steps must be positive");
    return TRXL_OK;
}

int build_layout(const trxl_model_config* c, Layout& L) {
    TRXL_PROPAGATE(validate(c));
    const int D = c->embed_dim, B = c->num_blocks, nb = c->num_branches, hid = c->hidden_size;
    long long cur = 0;
    const int g_enc = 0, g_lin = 1, g_blk0 = 2, g_head0 = 2 + B, g_lp = 2 + B + nb, g_lv = 3 + B + nb, g_vh = 4 + B + nb,
              g_other = 5 + B + nb;
    L.groups = 6 + B + nb;
    auto add = [&](const std::string& name, int group, std::initializer_list<long long> shape) -> long long {
        trxl_param_entry e;
        memset(&e, 0, sizeof(e));
        strncpy(e.name, name.c_str(), sizeof(e.name) - 1);
        long long n = 1;
        e.ndim = 0;
        for (long long s : shape) { e.shape[e.ndim++] = s; n *= s; }
        e.offset = cur;
        e.group = group;
        L.entries.push_back(e);
        const long long off = cur;
        cur += (n + 3) / 4 * 4;
        return off;
    };
    for (int i = 0; i < 6; ++i) L.conv[i] = -1;
    if (c->conv_in_channels > 0) {
        L.conv[0] = add("conv1.weight", g_enc, {32, c->conv_in_channels, 8, 8});
        L.conv[1] = add("conv1.bias", g_enc, {32});
        L.conv[2] = add("conv2.weight", g_enc, {64, 32, 4, 4});
        L.conv[3] = add("conv2.bias", g_enc, {64});
        L.conv[4] = add("conv3.weight", g_enc, {64, 64, 3, 3});
        L.conv[5] = add("conv3.bias", g_enc, {64});
    }
    L.Wh = add("lin_hidden.weight", g_lin, {D, c->feat_dim});
    L.bh = add("lin_hidden.bias", g_lin, {D});
    L.We = add("transformer.linear_embedding.weight", g_other, {D, D});
    L.be = add("transformer.linear_embedding.bias", g_other, {D});
    L.pos = -1;
    if (c->pos_enc == TRXL_PE_LEARNED) L.pos = add("transformer.pos_embedding", g_other, {c->max_episode_steps, D});
    L.blk.resize(B);
    for (int i = 0; i < B; ++i) {
        BlockP& b = L.blk[i];
        const std::string p = "transformer.transformer_blocks." + std::to_string(i) + ".";
        const int g = g_blk0 + i;
        b.Wv = add(p + "attention.values.weight", g, {D, D});
        b.Wk = add(p + "attention.keys.weight", g, {D, D});
        b.Wq = add(p + "attention.queries.weight", g, {D, D});
        b.Wo = add(p + "attention.fc_out.weight", g, {D, D});
        b.bo = add(p + "attention.fc_out.bias", g, {D});
        if (c->gtrxl) {
            for (int k = 0; k < 2; ++k) {
                GateP& gp = k ? b.g2 : b.g1;
                const std::string q = p + (k ? "gate2." : "gate1.");
                gp.Wr = add(q + "Wr.weight", g, {D, D});
                add(q + "Wz.weight", g, {D, D});
                add(q + "Wg.weight", g, {D, D});
                gp.Ur = add(q + "Ur.weight", g, {D, D});
                add(q + "Uz.weight", g, {D, D});
                gp.Ug = add(q + "Ug.weight", g, {D, D});
                gp.bg = add(q + "bg", g, {D});
            }
        }
        b.n1w = add(p + "norm1.weight", g, {D});
        b.n1b = add(p + "norm1.bias", g, {D});
        b.n2w = add(p + "norm2.weight", g, {D});
        b.n2b = add(p + "norm2.bias", g, {D});
        b.nkw = b.nkb = -1;
        if (c->layer_norm == TRXL_LN_PRE) {
            b.nkw = add(p + "norm_kv.weight", g, {D});
            b.nkb = add(p + "norm_kv.bias", g, {D});
        }
        b.Wff = add(p + "fc.0.weight", g, {D, D});
        b.bff = add(p + "fc.0.bias", g, {D});
    }
    L.Wp = add("lin_policy.weight", g_lp, {hid, D});
    L.bp = add("lin_policy.bias", g_lp, {hid});
    L.Wlv = add("lin_value.weight", g_lv, {hid, D});
    L.blv = add("lin_value.bias", g_lv, {hid});
    L.sumA = 0;
    for (int k = 0; k < nb; ++k) {           // branch weights contiguous -> one (sumA, hid) matrix
        const long long off = add("policy_branches." + std::to_string(k) + ".weight", g_head0 + k, {c->branch_sizes[k], hid});
        if (k == 0) L.Wbr = off;
        L.sumA += c->branch_sizes[k];
    }
    {
        // biases packed without padding so logits = hp Wbr^T + bbr is one GEMM
        long long packed = cur;
        for (int k = 0; k < nb; ++k) {
            trxl_param_entry e;
            memset(&e, 0, sizeof(e));
            const std::string name = "policy_branches." + std::to_string(k) + ".bias";
            strncpy(e.name, name.c_str(), sizeof(e.name) - 1);
            e.ndim = 1; e.shape[0] = c->branch_sizes[k]; e.offset = packed; e.group = g_head0 + k;
            L.entries.push_back(e);
            if (k == 0) L.bbr = packed;
            packed += c->branch_sizes[k];
        }
        cur = (packed + 3) / 4 * 4;
    }
    L.wval = add("value.weight", g_vh, {1, hid});
    L.bval = add("value.bias", g_vh, {1});
    L.total = cur;
    return TRXL_OK;
}

// ---------------------------------------------------------------------------------------------
struct GateA { float *G1, *G2, *r, *z, *rx, *G3, *hc; };
struct BlockA {
    float *q_in, *m1, *r1, *Q, *qk, *qkb, *probs, *ctx, *att_o, *att, *h1pre, *h1, *h_, *m2, *r2, *f, *out_pre;
    float *Wkg, *kb, *Wvg, *bv;
    GateA g1, g2;
};
struct Acts {
    float *h0, *h_final, *hp, *hv;
    std::vector<BlockA> blk;
    // backward scratch
    float *pool[2][9], *dH0;
    float *dctx, *dqk, *dqkb, *dscore, *dA1, *dz, *drx, *dWg, *dvec, *dhp, *dhv, *ew;
    float* gemm_ws; long long gemm_ws_n;      // split-K partials of the weight-gradient GEMMs
    long long total;
};

struct GemmWsGuard {       // the split-K workspace is only valid while this forward/backward is being enqueued
    GemmWsGuard(float* p, long long n) { trxl_gemm_set_workspace(p, n); }
    ~GemmWsGuard() { trxl_gemm_set_workspace(nullptr, 0); }
};

struct Bump {
    float* base; long long cur = 0;
    float* take(long long n) { float* p = base ? base + cur : nullptr; cur += (n + 3) / 4 * 4; return p; }
};

void carve(const trxl_model_config* c, int N, float* ws, Acts& A) {
    const long long D = c->embed_dim, H = c->num_heads, L = c->memory_length, hid = c->hidden_size;
    const bool pre = c->layer_norm == TRXL_LN_PRE, post = c->layer_norm == TRXL_LN_POST;
    int sumA = 0;
    for (int k = 0; k < c->num_branches; ++k) sumA += c->branch_sizes[k];
    Bump b{ws};
    const long long ND = (long long)N * D;
    A.h0 = b.take(ND); A.h_final = b.take(ND); A.hp = b.take(N * hid); A.hv = b.take(N * hid);
    A.blk.resize(c->num_blocks);
    for (auto& k : A.blk) {
        memset(&k, 0, sizeof(k));
        if (pre) { k.q_in = b.take(ND); k.m1 = b.take(N); k.r1 = b.take(N); k.h_ = b.take(ND); k.m2 = b.take(N); k.r2 = b.take(N);
                   k.qkb = b.take((long long)N * H); k.Wkg = b.take(D * D); k.kb = b.take(D); k.Wvg = b.take(D * D); k.bv = b.take(D); }
        if (post) { k.h1 = b.take(ND); k.m1 = b.take(N); k.r1 = b.take(N); k.out_pre = b.take(ND); k.m2 = b.take(N); k.r2 = b.take(N); }
        // probs: (N, H, L) for the per-sample kernel; the episode-grouped tensor-core path keeps P over all slots of the episode
        long long probs_row = L;
        if (c->pos_enc != TRXL_PE_LEARNED && attn_tc_row_floats(c->max_episode_steps) > probs_row)
            probs_row = attn_tc_row_floats(c->max_episode_steps);
        k.Q = b.take(ND); k.qk = b.take(ND * H); k.probs = b.take((long long)N * H * probs_row); k.ctx = b.take(ND * H);
        k.att_o = b.take(ND); k.h1pre = b.take(ND); k.f = b.take(ND);
        if (c->gtrxl) {
            k.att = b.take(ND);
            for (GateA* g : {&k.g1, &k.g2}) {
                g->G1 = b.take(3 * ND); g->G2 = b.take(2 * ND); g->r = b.take(ND); g->z = b.take(ND); g->rx = b.take(ND);
                g->G3 = b.take(ND); g->hc = b.take(ND);
            }
        }
    }
    for (int p = 0; p < 2; ++p) for (int i = 0; i < 9; ++i) A.pool[p][i] = b.take(ND);
    A.dH0 = b.take(ND);
    A.dctx = b.take(ND * H); A.dqk = b.take(ND * H); A.dqkb = b.take((long long)N * H);
    A.dscore = b.take((long long)N * H * attn_tc_row_floats(c->max_episode_steps > 0 ? c->max_episode_steps : 1));
    A.dA1 = b.take(3 * ND); A.dz = b.take(ND); A.drx = b.take(ND);
    A.dWg = b.take(D * D); A.dvec = b.take(D);
    A.dhp = b.take(N * hid); A.dhv = b.take(N * hid);
    long long widest = 3 * D; if (hid > widest) widest = hid; if (sumA > widest) widest = sumA; if (c->feat_dim > widest) widest = c->feat_dim;
    A.ew = b.take(ew_scratch_floats(N, (int)widest));
    A.gemm_ws_n = 32LL * 74 * 4096;     // 32 splits x (at most 74 output tiles of 64x64): the largest split-K GEMM trxl_gemm picks
    A.gemm_ws = b.take(A.gemm_ws_n);
    A.total = b.cur;
}

AttnArgs attn_args(const trxl_model_config* c, const ModelIO& io, int blk, const BlockA& a, const float* pe) {
    AttnArgs t;
    t.N = io.N; t.L = c->memory_length; t.D = c->embed_dim; t.H = c->num_heads; t.B = c->num_blocks; t.blk = blk;
    t.table = io.table; t.slots = io.slots; t.ep_index = io.ep_index; t.win_index = io.win_index; t.mask = io.mask;
    t.pe_index = pe ? io.pe_index : nullptr; t.pe = pe; t.sample_index = io.sample_index;
    t.qk = a.qk; t.qkb = (c->layer_norm == TRXL_LN_PRE) ? a.qkb : nullptr; t.ln = c->layer_norm == TRXL_LN_PRE;
    t.scale = (float)sqrt((double)c->embed_dim);
    t.probs = a.probs; t.ctx = a.ctx;
    return t;
}

// the episode-grouped tensor-core attention applies when the caller supplied the grouping and the block needs neither the
// pre-LayerNorm fold nor gradients into a learned positional table
bool use_grouped_attention(const trxl_model_config* c, const ModelIO& io) {
    return io.tiles && io.ranges && io.table_pe && io.n_tiles > 0 && c->pos_enc != TRXL_PE_LEARNED &&
           io.slots == c->max_episode_steps && attn_tc_supported(c->embed_dim, c->num_heads, io.slots, c->num_blocks);
}

AttnTcArgs attn_tc_args(const trxl_model_config* c, const ModelIO& io, int blk) {
    AttnTcArgs t;
    t.N = io.N; t.L = c->memory_length; t.D = c->embed_dim; t.H = c->num_heads; t.B = c->num_blocks; t.blk = blk;
    t.table_pe = io.table_pe; t.slots = io.slots; t.n_episodes = io.n_episodes; t.tiles = io.tiles; t.n_tiles = io.n_tiles;
    t.ranges = io.ranges; t.scale = (float)sqrt((double)c->embed_dim);
    return t;
}

int gate_forward(cudaStream_t st, const float* P, const GateP& gp, const GateA& ga, const float* x, long long ldx,
                 const float* y, long long ldy, float* out, long long ldo, int N, int D) {
    TRXL_PROPAGATE(gemm_nt(st, N, 3 * D, D, y, ldy, P + gp.Wr, D, ga.G1, 3 * D));
    TRXL_PROPAGATE(gemm_nt(st, N, 2 * D, D, x, ldx, P + gp.Ur, D, ga.G2, 2 * D));
    TRXL_PROPAGATE(ew_gate_fwd_a(st, ga.G1, ga.G2, P + gp.bg, x, ldx, ga.r, ga.z, ga.rx, N, D));
    TRXL_PROPAGATE(gemm_nt(st, N, D, D, ga.rx, D, P + gp.Ug, D, ga.G3, D));
    TRXL_PROPAGATE(ew_gate_fwd_b(st, ga.G1, ga.G3, x, ldx, ga.z, ga.hc, out, ldo, N, D));
    return TRXL_OK;
}

// dx (N,D contiguous, overwritten), dy (N,D contiguous, overwritten)
int gate_backward(cudaStream_t st, const float* P, float* G, const GateP& gp, const GateA& ga, const Acts& A, const float* dout,
                  const float* x, long long ldx, const float* y, long long ldy, float* dx, float* dy, int N, int D) {
    const long long D3 = 3LL * D;
    TRXL_PROPAGATE(ew_gate_bwd_a(st, dout, D, x, ldx, ga.z, ga.hc, A.dA1, A.dz, dx, D, 0, N, D));
    TRXL_PROPAGATE(gemm_tn(st, D, D, N, A.dA1 + 2 * D, D3, ga.rx, D, G + gp.Ug, D, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(gemm_nn(st, N, D, D, A.dA1 + 2 * D, D3, P + gp.Ug, D, A.drx, D));
    TRXL_PROPAGATE(ew_gate_bwd_b(st, A.drx, x, ldx, ga.r, ga.z, A.dz, A.dA1, dx, D, N, D));
    TRXL_PROPAGATE(gemm_tn(st, 3 * D, D, N, A.dA1, D3, y, ldy, G + gp.Wr, D, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(gemm_tn(st, 2 * D, D, N, A.dA1, D3, x, ldx, G + gp.Ur, D, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(ew_colsum(st, A.dA1 + D, D3, G + gp.bg, N, D, -1.f, 0, A.ew));
    TRXL_PROPAGATE(gemm_nn(st, N, D, 3 * D, A.dA1, D3, P + gp.Wr, D, dy, D));
    TRXL_PROPAGATE(gemm_nn(st, N, D, 2 * D, A.dA1, D3, P + gp.Ur, D, dx, D, 1));
    return TRXL_OK;
}

int per_head_gemm(cudaStream_t st, int M, int Nn, int K, const float* A, long long lda, int a_kc, long long sA, const float* B,
                  long long ldb, int b_kc, long long sB, float* C, long long ldc, long long sC, int H, const float* bias = nullptr,
                  long long sBias = 0, float* ws = nullptr, long long ws_n = 0) {
    GemmArgs g;
    g.M = M; g.N = Nn; g.K = K; g.A = A; g.lda = lda; g.a_kc = a_kc; g.sA = sA; g.B = B; g.ldb = ldb; g.b_kc = b_kc; g.sB = sB;
    g.C = C; g.ldc = ldc; g.sC = sC; g.batch = H; g.bias = bias; g.sBias = sBias; g.ws = ws; g.ws_floats = ws_n;
    return trxl_gemm(g, st);
}

}  // namespace

int model_layout(const trxl_model_config* cfg, std::vector<trxl_param_entry>& out, long long* total, int* groups) {
    Layout L;
    TRXL_PROPAGATE(build_layout(cfg, L));
    out = L.entries;
    if (total) *total = L.total;
    if (groups) *groups = L.groups;
    return TRXL_OK;
}

long long model_workspace_floats(const trxl_model_config* cfg, int N) {
    if (validate(cfg) != TRXL_OK || N < 0) return -1;
    Acts A;
    carve(cfg, N, nullptr, A);
    return A.total;
}

static void fill_fused_args(const trxl_model_config* c, const Layout& L, RfArgs& a) {
    a.D = c->embed_dim; a.H = c->num_heads; a.B = c->num_blocks; a.L = c->memory_length; a.hid = c->hidden_size;
    a.feat = c->feat_dim; a.sumA = L.sumA; a.ln = c->layer_norm; a.pe_mode = c->pos_enc; a.gtrxl = c->gtrxl;
    a.Wh = L.Wh; a.bh = L.bh; a.We = L.We; a.be = L.be; a.pos = L.pos < 0 ? 0 : L.pos;
    a.blk_stride = c->num_blocks > 1 ? L.blk[1].Wv - L.blk[0].Wv : 0;
    const BlockP& b = L.blk[0];
    a.b0.Wv = b.Wv; a.b0.Wk = b.Wk; a.b0.Wq = b.Wq; a.b0.Wo = b.Wo; a.b0.bo = b.bo;
    a.b0.g1 = RfGate{b.g1.Wr, b.g1.Ur, b.g1.Ug, b.g1.bg};
    a.b0.g2 = RfGate{b.g2.Wr, b.g2.Ur, b.g2.Ug, b.g2.bg};
    a.b0.n1w = b.n1w; a.b0.n1b = b.n1b; a.b0.n2w = b.n2w; a.b0.n2b = b.n2b; a.b0.nkw = b.nkw; a.b0.nkb = b.nkb;
    a.b0.Wff = b.Wff; a.b0.bff = b.bff;
    a.Wp = L.Wp; a.bp = L.bp; a.Wlv = L.Wlv; a.blv = L.blv; a.Wbr = L.Wbr; a.bbr = L.bbr; a.wval = L.wval; a.bval = L.bval;
}

int model_fused_supported(const trxl_model_config* c) {
    Layout L;
    if (build_layout(c, L) != TRXL_OK) return 0;
    RfArgs a;
    fill_fused_args(c, L, a);
    return rollout_fused_supported(a) ? 1 : 0;
}

// Inference-only forward (no activations saved): the whole trunk in one launch, one CTA per sample.
static int model_forward_fused(const trxl_model_config* c, const Layout& L, const float* P, const ModelIO& io, float* logits,
                               float* value, float* out_mem, cudaStream_t st) {
    RfArgs a;
    fill_fused_args(c, L, a);
    a.N = io.N; a.P = P; a.feat_in = io.feat; a.table = io.table; a.slots = io.slots; a.ep_index = io.ep_index;
    a.win_index = io.win_index; a.mask = io.mask; a.pe_index = c->pos_enc == TRXL_PE_NONE ? nullptr : io.pe_index;
    a.sample_index = io.sample_index; a.pe_table = io.pe_table; a.logits = logits; a.value = value; a.out_mem = out_mem;
    return rollout_fused_forward(a, st);
}

int model_forward(const trxl_model_config* c, const float* P, const ModelIO& io, float* ws, float* logits, float* value,
                  float* out_mem, cudaStream_t st) {
    Layout L;
    TRXL_PROPAGATE(build_layout(c, L));
    TRXL_CHECK_ARG(P && io.feat && io.table && logits && value && out_mem, "model_forward: null pointer");
    if (ws == nullptr) {        // no workspace: inference-only fused path
        TRXL_CHECK_ARG(io.N > 0, "model_forward: N must be positive");
        // pe_index == NULL on this path: the table's rows already carry their positional rows (trxl_rollout_store)
        TRXL_CHECK_ARG(c->pos_enc != TRXL_PE_RELATIVE || io.pe_table || !io.pe_index, "model_forward: relative PE needs pe_table");
        return model_forward_fused(c, L, P, io, logits, value, out_mem, st);
    }
    TRXL_CHECK_ARG(io.N > 0, "model_forward: N must be positive");
    TRXL_CHECK_ARG(c->pos_enc != TRXL_PE_RELATIVE || io.pe_table, "model_forward: relative PE needs pe_table");
    TRXL_CHECK_ARG(c->pos_enc == TRXL_PE_NONE || io.pe_index, "model_forward: positional encoding needs pe_index");
    Acts A;
    carve(c, io.N, ws, A);
    GemmWsGuard ws_guard(A.gemm_ws, A.gemm_ws_n);
    const int N = io.N, D = c->embed_dim, H = c->num_heads, B = c->num_blocks, dh = D / H, hid = c->hidden_size;
    const long long BD = (long long)B * D;
    const bool pre = c->layer_norm == TRXL_LN_PRE, post = c->layer_norm == TRXL_LN_POST;
    const float* pe = c->pos_enc == TRXL_PE_RELATIVE ? io.pe_table : (c->pos_enc == TRXL_PE_LEARNED ? P + L.pos : nullptr);

    const bool grouped = use_grouped_attention(c, io);
    TRXL_PROPAGATE(gemm_nt(st, N, D, c->feat_dim, io.feat, c->feat_dim, P + L.Wh, c->feat_dim, A.h0, D, P + L.bh, 1));
    TRXL_PROPAGATE(gemm_nt(st, N, D, D, A.h0, D, P + L.We, D, out_mem, BD, P + L.be, 1));
    for (int i = 0; i < B; ++i) {
        const BlockP& p = L.blk[i];
        BlockA& a = A.blk[i];
        const float* h_in = out_mem + (long long)i * D;
        float* h_out = (i + 1 < B) ? out_mem + (long long)(i + 1) * D : A.h_final;
        const long long ld_out = (i + 1 < B) ? BD : D;
        const float* q_in = h_in; long long ld_q = BD;
        if (pre) {
            TRXL_PROPAGATE(ew_layernorm_fwd(st, h_in, BD, nullptr, 0, P + p.n1w, P + p.n1b, a.q_in, D, nullptr, 0, a.m1, a.r1, N, D));
            q_in = a.q_in; ld_q = D;
        }
        TRXL_PROPAGATE(gemm_nt(st, N, D, D, q_in, ld_q, P + p.Wq, D, a.Q, D));
        const float *Wkg = P + p.Wk, *Wvg = P + p.Wv, *bv = nullptr;
        if (pre) {
            TRXL_PROPAGATE(ew_scale_cols(st, P + p.Wk, P + p.nkw, a.Wkg, D, D));
            TRXL_PROPAGATE(ew_matvec(st, P + p.Wk, P + p.nkb, a.kb, D, D));
            TRXL_PROPAGATE(ew_scale_cols(st, P + p.Wv, P + p.nkw, a.Wvg, D, D));
            TRXL_PROPAGATE(ew_matvec(st, P + p.Wv, P + p.nkb, a.bv, D, D));
            TRXL_PROPAGATE(ew_head_dot(st, a.Q, a.kb, a.qkb, N, H, dh));
            Wkg = a.Wkg; Wvg = a.Wvg; bv = a.bv;
        }
        // qk[n,h,:] = Q[n, h*dh:(h+1)*dh] @ Wkg[h*dh:(h+1)*dh, :]
        TRXL_PROPAGATE(per_head_gemm(st, N, D, dh, a.Q, D, 1, dh, Wkg, D, 0, (long long)dh * D, a.qk, (long long)H * D, D, H));
        if (grouped) TRXL_PROPAGATE(attn_tc_forward(attn_tc_args(c, io, i), a.qk, a.probs, a.ctx, st));
        else TRXL_PROPAGATE(trxl_window_attn_fwd(attn_args(c, io, i, a, pe), st));
        // att_o[n, h*dh + j] = ctx[n,h,:] . Wvg[h*dh + j, :] (+ bv)
        TRXL_PROPAGATE(per_head_gemm(st, N, dh, D, a.ctx, (long long)H * D, 1, D, Wvg, D, 1, (long long)dh * D, a.att_o, D, dh, H, bv, dh));
        if (c->gtrxl) {
            TRXL_PROPAGATE(gemm_nt(st, N, D, D, a.att_o, D, P + p.Wo, D, a.att, D, P + p.bo));
            TRXL_PROPAGATE(gate_forward(st, P, p.g1, a.g1, h_in, BD, a.att, D, a.h1pre, D, N, D));
        } else {
            TRXL_PROPAGATE(gemm_nt(st, N, D, D, a.att_o, D, P + p.Wo, D, a.h1pre, D, P + p.bo, 0, h_in, BD));
        }
        const float* h1 = a.h1pre;
        if (post) {
            TRXL_PROPAGATE(ew_layernorm_fwd(st, a.h1pre, D, nullptr, 0, P + p.n1w, P + p.n1b, a.h1, D, nullptr, 0, a.m1, a.r1, N, D));
            h1 = a.h1;
        }
        const float* h_ = h1;
        if (pre) {
            TRXL_PROPAGATE(ew_layernorm_fwd(st, h1, D, nullptr, 0, P + p.n2w, P + p.n2b, a.h_, D, nullptr, 0, a.m2, a.r2, N, D));
            h_ = a.h_;
        }
        TRXL_PROPAGATE(gemm_nt(st, N, D, D, h_, D, P + p.Wff, D, a.f, D, P + p.bff, 1));
        if (c->gtrxl) {
            float* dst = post ? a.out_pre : h_out;
            TRXL_PROPAGATE(gate_forward(st, P, p.g2, a.g2, h1, D, a.f, D, dst, post ? D : ld_out, N, D));
            if (post) TRXL_PROPAGATE(ew_layernorm_fwd(st, a.out_pre, D, nullptr, 0, P + p.n2w, P + p.n2b, h_out, ld_out, nullptr, 0, a.m2, a.r2, N, D));
        } else if (post) {
            TRXL_PROPAGATE(ew_layernorm_fwd(st, a.f, D, h1, D, P + p.n2w, P + p.n2b, h_out, ld_out, a.out_pre, D, a.m2, a.r2, N, D));
        } else {
            TRXL_PROPAGATE(ew_add(st, a.f, D, h1, D, h_out, ld_out, N, D));
        }
    }
    TRXL_PROPAGATE(gemm_nt(st, N, hid, D, A.h_final, D, P + L.Wp, D, A.hp, hid, P + L.bp, 1));
    TRXL_PROPAGATE(gemm_nt(st, N, hid, D, A.h_final, D, P + L.Wlv, D, A.hv, hid, P + L.blv, 1));
    TRXL_PROPAGATE(gemm_nt(st, N, L.sumA, hid, A.hp, hid, P + L.Wbr, hid, logits, L.sumA, P + L.bbr));
    TRXL_PROPAGATE(gemm_nt(st, N, 1, hid, A.hv, hid, P + L.wval, hid, value, 1, P + L.bval));
    return TRXL_OK;
}

int model_backward(const trxl_model_config* c, const float* P, float* G, const ModelIO& io, float* ws, const float* out_mem,
                   const float* dlogits, const float* dvalue, float* dfeat, cudaStream_t st) {
    Layout L;
    TRXL_PROPAGATE(build_layout(c, L));
    TRXL_CHECK_ARG(P && G && io.feat && io.table && ws && out_mem && dlogits && dvalue, "model_backward: null pointer");
    Acts A;
    carve(c, io.N, ws, A);
    GemmWsGuard ws_guard(A.gemm_ws, A.gemm_ws_n);
    const int N = io.N, D = c->embed_dim, H = c->num_heads, B = c->num_blocks, dh = D / H, hid = c->hidden_size;
    const long long BD = (long long)B * D, HD = (long long)H * D;
    const bool pre = c->layer_norm == TRXL_LN_PRE, post = c->layer_norm == TRXL_LN_POST;
    const float* pe = c->pos_enc == TRXL_PE_RELATIVE ? io.pe_table : (c->pos_enc == TRXL_PE_LEARNED ? P + L.pos : nullptr);

    const bool grouped = use_grouped_attention(c, io);
    // ---- heads ----
    float* dH = A.dH0;        // gradient w.r.t. the current block output; never inside the pool being used
    TRXL_PROPAGATE(gemm_tn(st, L.sumA, hid, N, dlogits, L.sumA, A.hp, hid, G + L.Wbr, hid, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(ew_colsum(st, dlogits, L.sumA, G + L.bbr, N, L.sumA, 1.f, 0, A.ew));
    TRXL_PROPAGATE(gemm_nn(st, N, hid, L.sumA, dlogits, L.sumA, P + L.Wbr, hid, A.dhp, hid));
    TRXL_PROPAGATE(ew_relu_bwd(st, A.dhp, hid, A.hp, hid, A.dhp, hid, N, hid, 0));
    TRXL_PROPAGATE(gemm_tn(st, hid, D, N, A.dhp, hid, A.h_final, D, G + L.Wp, D, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(ew_colsum(st, A.dhp, hid, G + L.bp, N, hid, 1.f, 0, A.ew));
    TRXL_PROPAGATE(gemm_nn(st, N, D, hid, A.dhp, hid, P + L.Wp, D, dH, D));
    TRXL_PROPAGATE(gemm_tn(st, 1, hid, N, dvalue, 1, A.hv, hid, G + L.wval, hid, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(ew_colsum(st, dvalue, 1, G + L.bval, N, 1, 1.f, 0, A.ew));
    TRXL_PROPAGATE(gemm_nn(st, N, hid, 1, dvalue, 1, P + L.wval, hid, A.dhv, hid));
    TRXL_PROPAGATE(ew_relu_bwd(st, A.dhv, hid, A.hv, hid, A.dhv, hid, N, hid, 0));
    TRXL_PROPAGATE(gemm_tn(st, hid, D, N, A.dhv, hid, A.h_final, D, G + L.Wlv, D, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(ew_colsum(st, A.dhv, hid, G + L.blv, N, hid, 1.f, 0, A.ew));
    TRXL_PROPAGATE(gemm_nn(st, N, D, hid, A.dhv, hid, P + L.Wlv, D, dH, D, 1));

    // ---- blocks, last to first ----
    for (int i = B - 1; i >= 0; --i) {
        const BlockP& p = L.blk[i];
        BlockA& a = A.blk[i];
        float** S = A.pool[i & 1];              // incoming dH lives in dH0 or in the other pool
        const float* h_in = out_mem + (long long)i * D;
        const float* q_in = pre ? a.q_in : h_in; const long long ld_q = pre ? D : BD;
        const float* h1 = post ? a.h1 : a.h1pre;
        const float* h_ = pre ? a.h_ : h1;
        // d(out_pre)
        float* d_outpre = dH;
        if (post) {
            TRXL_PROPAGATE(ew_layernorm_bwd(st, dH, D, a.out_pre, D, a.m2, a.r2, P + p.n2w, S[0], D, 0, G + p.n2w, G + p.n2b, 0, A.ew, N, D));
            d_outpre = S[0];
        }
        float *dH1, *dF;
        if (c->gtrxl) {
            dH1 = S[1]; dF = S[2];
            TRXL_PROPAGATE(gate_backward(st, P, G, p.g2, a.g2, A, d_outpre, h1, D, a.f, D, dH1, dF, N, D));
        } else {
            dH1 = d_outpre; dF = d_outpre;
        }
        // f = relu(h_ Wff^T + bff)
        TRXL_PROPAGATE(ew_relu_bwd(st, dF, D, a.f, D, S[3], D, N, D, 0));
        TRXL_PROPAGATE(gemm_tn(st, D, D, N, S[3], D, h_, D, G + p.Wff, D, 0, A.gemm_ws, A.gemm_ws_n));
        TRXL_PROPAGATE(ew_colsum(st, S[3], D, G + p.bff, N, D, 1.f, 0, A.ew));
        if (pre) {
            TRXL_PROPAGATE(gemm_nn(st, N, D, D, S[3], D, P + p.Wff, D, S[4], D));
            TRXL_PROPAGATE(ew_layernorm_bwd(st, S[4], D, h1, D, a.m2, a.r2, P + p.n2w, dH1, D, 1, G + p.n2w, G + p.n2b, 0, A.ew, N, D));
        } else {
            TRXL_PROPAGATE(gemm_nn(st, N, D, D, S[3], D, P + p.Wff, D, dH1, D, 1));
        }
        float* dH1pre = dH1;
        if (post) {
            TRXL_PROPAGATE(ew_layernorm_bwd(st, dH1, D, a.h1pre, D, a.m1, a.r1, P + p.n1w, S[5], D, 0, G + p.n1w, G + p.n1b, 0, A.ew, N, D));
            dH1pre = S[5];
        }
        float *dHin, *dAtt;
        if (c->gtrxl) {
            dHin = S[6]; dAtt = S[7];
            TRXL_PROPAGATE(gate_backward(st, P, G, p.g1, a.g1, A, dH1pre, h_in, BD, a.att, D, dHin, dAtt, N, D));
        } else {
            dHin = dH1pre; dAtt = dH1pre;
        }
        // att = att_o Wo^T + bo
        TRXL_PROPAGATE(gemm_tn(st, D, D, N, dAtt, D, a.att_o, D, G + p.Wo, D, 0, A.gemm_ws, A.gemm_ws_n));
        TRXL_PROPAGATE(ew_colsum(st, dAtt, D, G + p.bo, N, D, 1.f, 0, A.ew));
        float* dAtto = S[3];
        TRXL_PROPAGATE(gemm_nn(st, N, D, D, dAtt, D, P + p.Wo, D, dAtto, D));
        const float* Wkg = pre ? a.Wkg : P + p.Wk;
        const float* Wvg = pre ? a.Wvg : P + p.Wv;
        // dWvg[h*dh + j, :] = sum_n dAtto[n, h*dh + j] ctx[n,h,:]
        TRXL_PROPAGATE(per_head_gemm(st, dh, D, N, dAtto, D, 0, dh, a.ctx, HD, 0, D, pre ? A.dWg : G + p.Wv, D, (long long)dh * D, H, nullptr, 0, A.gemm_ws, A.gemm_ws_n));
        // dctx[n,h,:] = sum_j dAtto[n, h*dh + j] Wvg[h*dh + j, :]
        TRXL_PROPAGATE(per_head_gemm(st, N, D, dh, dAtto, D, 1, dh, Wvg, D, 0, (long long)dh * D, A.dctx, HD, D, H));
        if (pre) {
            TRXL_PROPAGATE(ew_colsum(st, dAtto, D, A.dvec, N, D, 1.f, 0, A.ew));
            TRXL_PROPAGATE(ew_unfold(st, A.dWg, A.dvec, P + p.Wv, P + p.nkw, P + p.nkb, G + p.Wv, G + p.nkw, G + p.nkb, D, D, 0));
        }
        if (grouped) {
            // (pre-LN: the energy bias qkb is constant over a row's window, so it cancels in the softmax and gets no gradient)
            if (pre) cudaMemsetAsync(A.dqkb, 0, sizeof(float) * (size_t)N * H, st);
            TRXL_PROPAGATE(attn_tc_backward(attn_tc_args(c, io, i), a.probs, A.dctx, A.dscore, A.dqk, st));
        } else {
            AttnBwdArgs ab;
            ab.dctx = A.dctx; ab.dqk = A.dqk; ab.dqkb = pre ? A.dqkb : nullptr;
            ab.dpe = (c->pos_enc == TRXL_PE_LEARNED) ? G + L.pos : nullptr;
            TRXL_PROPAGATE(trxl_window_attn_bwd(attn_args(c, io, i, a, pe), ab, st));
        }
        // dQ[n, h*dh + j] = dqk[n,h,:] . Wkg[h*dh + j, :]
        float* dQ = S[4];
        TRXL_PROPAGATE(per_head_gemm(st, N, dh, D, A.dqk, HD, 1, D, Wkg, D, 1, (long long)dh * D, dQ, D, dh, H));
        // dWkg[h*dh + j, :] = sum_n Q[n, h*dh + j] dqk[n,h,:]
        TRXL_PROPAGATE(per_head_gemm(st, dh, D, N, a.Q, D, 0, dh, A.dqk, HD, 0, D, pre ? A.dWg : G + p.Wk, D, (long long)dh * D, H, nullptr, 0, A.gemm_ws, A.gemm_ws_n));
        if (pre) {
            TRXL_PROPAGATE(ew_head_dot_bwd(st, a.Q, A.dqkb, a.kb, dQ, A.dvec, N, H, dh, A.ew));
            TRXL_PROPAGATE(ew_unfold(st, A.dWg, A.dvec, P + p.Wk, P + p.nkw, P + p.nkb, G + p.Wk, G + p.nkw, G + p.nkb, D, D, 1));
        }
        // Q = q_in Wq^T
        TRXL_PROPAGATE(gemm_tn(st, D, D, N, dQ, D, q_in, ld_q, G + p.Wq, D, 0, A.gemm_ws, A.gemm_ws_n));
        if (pre) {
            TRXL_PROPAGATE(gemm_nn(st, N, D, D, dQ, D, P + p.Wq, D, S[8], D));
            TRXL_PROPAGATE(ew_layernorm_bwd(st, S[8], D, h_in, BD, a.m1, a.r1, P + p.n1w, dHin, D, 1, G + p.n1w, G + p.n1b, 0, A.ew, N, D));
        } else {
            TRXL_PROPAGATE(gemm_nn(st, N, D, D, dQ, D, P + p.Wq, D, dHin, D, 1));
        }
        dH = dHin;
    }
    // ---- embedding + lin_hidden ----
    float** S = A.pool[1];                      // block 0 used pool[0]; pool[1] is free again
    float* dE = S[0];
    TRXL_PROPAGATE(ew_relu_bwd(st, dH, D, out_mem, BD, dE, D, N, D, 0));
    TRXL_PROPAGATE(gemm_tn(st, D, D, N, dE, D, A.h0, D, G + L.We, D, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(ew_colsum(st, dE, D, G + L.be, N, D, 1.f, 0, A.ew));
    float* dh0 = S[1];
    TRXL_PROPAGATE(gemm_nn(st, N, D, D, dE, D, P + L.We, D, dh0, D));
    TRXL_PROPAGATE(ew_relu_bwd(st, dh0, D, A.h0, D, dh0, D, N, D, 0));
    TRXL_PROPAGATE(gemm_tn(st, D, c->feat_dim, N, dh0, D, io.feat, c->feat_dim, G + L.Wh, c->feat_dim, 0, A.gemm_ws, A.gemm_ws_n));
    TRXL_PROPAGATE(ew_colsum(st, dh0, D, G + L.bh, N, D, 1.f, 0, A.ew));
    if (dfeat) TRXL_PROPAGATE(gemm_nn(st, N, c->feat_dim, D, dh0, D, P + L.Wh, c->feat_dim, dfeat, c->feat_dim));
    return TRXL_OK;
}
