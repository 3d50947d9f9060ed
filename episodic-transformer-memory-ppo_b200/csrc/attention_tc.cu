// Episode-grouped window attention on the tensor cores (query length 1), forward and backward.
//
// Reference contractions: transformer.py:59 (einsum nqhd,nkhd->nhqk) and :73 (nhql,nlhd->nqhd), after the query-side fold
// of attention.cu:  energy[n,h,t] = qk[n,h,:] . (x_t + pe_t),   ctx[n,h,:] = sum_t p[n,h,t] (x_t + pe_t).
//
// The per-sample kernel of attention.cu streams every sample's window separately on the FMA pipe.  Here the minibatch is
// sorted by episode (the host does that once per epoch, trainer.py) and cut into row tiles of 128 (sample, head) rows that
// belong to ONE episode.  For such a tile both contractions are dense GEMMs against that episode's memory rows of the block
//     S   (128 x M)  = QK_tile (128 x D)  .  Xpe_e^T (D x M)          all M slots of the episode, window applied afterwards
//     ctx (128 x D)  = P_tile  (128 x M)  .  Xpe_e   (M x D)
// and run on the TMA + tcgen05 3xTF32 GEMM of tc_gemm.cu in its grouped mode: the episode's rows are fetched ONCE per tile by
// TMA straight from the (E, M, B, D) table (a strided rank-3 tensor map: row stride B*D, batch stride M*B*D -- no gather, no
// copy), instead of once per sample.  The softmax over each row's window (visible slots form one contiguous slot range
// [lo, lo+cnt); a fully masked row attends uniformly over its L window slots, as the reference's finite -1e20 fill does) is a
// small warp-per-row kernel between the two GEMMs.  Backward: dP = dctx . Xpe_e^T, dS = P (dP - P.dP) / sqrt(D) on the window,
// dqk = dS . Xpe_e -- the same two GEMM shapes.  The memory rows carry no gradient (transformer.py:248).
//
// Xpe = table + positional row is materialised once per update (trxl_table_add_pe): the table is frozen during the
// optimisation epochs and the sinusoidal table has no parameters.  For pre-LayerNorm models the rows are also normalised there
// (norm_kv without its affine part, which the model folds into Wk / Wv): the kernels below then see plain rows; the energy
// bias of the fold (qkb) is constant over a row's window, cancels in the softmax and receives no gradient.  Learned positional
// tables (gradients flow into the rows) stay on the per-sample kernel (attention.cu).
#include "attention_tc.cuh"

#include <stdlib.h>

#include "gemm.cuh"

int trxl_tc_gemm(const GemmArgs& g, int bn, cudaStream_t st);      // tc_gemm.cu
bool trxl_tc_gemm_eligible(const GemmArgs& g);

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ranges[n] = {first visible slot, number of visible slots, uniform (row fully masked), episode}
__global__ void attn_ranges_kernel(const unsigned char* __restrict__ mask, const long long* __restrict__ win_index,
                                   const long long* __restrict__ ep_index, const long long* __restrict__ sample_index, int N, int L,
                                   int4* __restrict__ ranges) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const long long row = sample_index ? sample_index[n] : n;
    int cnt = 0, first = L;
    for (int l = lane; l < L; l += 32) {
        if (mask[row * L + l]) { ++cnt; first = min(first, l); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(FULL, cnt, o);
        first = min(first, __shfl_xor_sync(FULL, first, o));
    }
    if (lane == 0) {
        int4 r;
        if (cnt == 0) { r.x = (int)win_index[row * L]; r.y = L; r.z = 1; }
        else { r.x = (int)win_index[row * L + first]; r.y = cnt; r.z = 0; }
        r.w = (int)(ep_index ? ep_index[row] : row);
        ranges[n] = r;
    }
}

// in place: S (rows, ld) energies over all slots -> P: softmax over [lo, lo+cnt) of S / scale, zero elsewhere.  One warp per row;
// the row lives in registers between the passes (PL values per lane, M <= 32 PL): one read and one write of the matrix.
template <int PL>
__global__ void attn_softmax_kernel(float* __restrict__ S, long long ld, int M, const int4* __restrict__ ranges, int H, int rows,
                                    float scale, int L) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int4 rg = ranges[r / H];
    float* row = S + (long long)r * ld;
    const int lo = rg.x, hi = rg.x + rg.y;
    if (rg.z) {                                   // fully masked: uniform over the L window slots
        const float u = 1.f / (float)L;
        for (int t = lane; t < M; t += 32) row[t] = (t >= lo && t < hi) ? u : 0.f;
        return;
    }
    float e[PL];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const int t = lane + 32 * i;
        e[i] = (t >= lo && t < hi) ? __fdiv_rn(row[t], scale) : -INFINITY;
        m = fmaxf(m, e[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        e[i] = __expf(e[i] - m);                  // slots outside the window hold -inf -> 0
        s += e[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    const float inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const int t = lane + 32 * i;
        if (t < M) row[t] = e[i] * inv;
    }
}

// in place: dP (rows, ld) -> dS = P (dP - sum_t P dP) / scale on the window, zero elsewhere (and everywhere for uniform rows)
template <int PL>
__global__ void attn_dscore_kernel(float* __restrict__ dP, const float* __restrict__ P, long long ld, int M,
                                   const int4* __restrict__ ranges, int H, int rows, float inv_scale) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int4 rg = ranges[r / H];
    float* d = dP + (long long)r * ld;
    const float* p = P + (long long)r * ld;
    const int lo = rg.x, hi = rg.x + rg.y;
    if (rg.z) {
        for (int t = lane; t < M; t += 32) d[t] = 0.f;
        return;
    }
    float pv[PL], dv[PL];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const int t = lane + 32 * i;
        const bool in = t >= lo && t < hi;
        pv[i] = in ? p[t] : 0.f;
        dv[i] = in ? d[t] : 0.f;
        dot = fmaf(pv[i], dv[i], dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const int t = lane + 32 * i;
        if (t < M) d[t] = pv[i] * (dv[i] - dot) * inv_scale;      // pv = 0 outside the window
    }
}

// out[e, m, b, :] = table[e, m, b, :] + pe[m, :]
__global__ void table_add_pe_kernel(const float4* __restrict__ table, const float4* __restrict__ pe, float4* __restrict__ out,
                                    long long total4, int M, int B, int D4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D4);
        const int m = (int)((i / ((long long)D4 * B)) % M);
        float4 v = table[i];
        const float4 p = pe[(long long)m * D4 + d];
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        out[i] = v;
    }
}

// out[e, m, b, :] = LayerNorm_noaffine(table[e, m, b, :] + pe[m, :])   (eps 1e-5, biased variance; one warp per row).
// The pre-LayerNorm block normalises every window row with norm_kv (transformer.py:131); gamma / beta are folded into Wk / Wv by
// the caller, and the normalisation itself depends only on the stored row, so it is done once per update here.
__global__ void table_add_pe_ln_kernel(const float* __restrict__ table, const float* __restrict__ pe, float* __restrict__ out,
                                       long long rows, int M, int B, int D) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int m = (int)((r / B) % M);
    const float* x = table + r * D;
    const float* p = pe ? pe + (long long)m * D : nullptr;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s += x[j] + (p ? p[j] : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    const float mu = s / (float)D;
    float v = 0.f;
    for (int j = lane; j < D; j += 32) {
        const float d = x[j] + (p ? p[j] : 0.f) - mu;
        v = fmaf(d, d, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const float rstd = rsqrtf(v / (float)D + 1e-5f);
    for (int j = lane; j < D; j += 32) out[r * D + j] = (x[j] + (p ? p[j] : 0.f) - mu) * rstd;
}

GemmArgs grouped_args(const AttnTcArgs& a, const float* A, long long lda, int K, int Nn, int b_kc, float* C, long long ldc) {
    GemmArgs g;
    g.M = a.N * a.H; g.N = Nn; g.K = K;
    g.A = A; g.lda = lda; g.a_kc = 1;
    g.B = a.table_pe + (long long)a.blk * a.D; g.ldb = (long long)a.B * a.D; g.b_kc = b_kc;
    g.sB = a.slots * a.B * a.D; g.b_batch = a.n_episodes;
    g.C = C; g.ldc = ldc;
    g.tiles = a.tiles; g.n_tiles = a.n_tiles;
    g.ksplit = 1; g.k_per_split = K;
    g.alpha = 1.f;
    return g;
}

// N tile of a grouped GEMM: one 256-wide CTA per row tile when the output is wider than 128 (each CTA converts the tile's A
// operand once instead of twice, and a c3 launch is 136 CTAs = one wave), else the narrowest tile that covers it
int tile_n(int n) { return n > 128 ? 256 : (n > 64 ? 128 : (n > 32 ? 64 : 32)); }

}  // namespace

long long attn_tc_row_floats(long long slots) { return (slots + 3) / 4 * 4; }

bool attn_tc_supported(int D, int H, long long slots, int B) {
    // (slots <= 1024: the row kernels keep a row of energies in registers, 32 values per lane)
    return H > 0 && 128 % H == 0 && D % 4 == 0 && D >= 4 && slots >= 1 && slots <= 1024 && ((long long)B * D) % 4 == 0;
}

int attn_tc_ranges(const unsigned char* mask, const long long* win_index, const long long* ep_index, const long long* sample_index,
                   int N, int L, int4* ranges, cudaStream_t st) {
    TRXL_CHECK_ARG(mask && win_index && ranges && N >= 0 && L > 0, "attention_ranges: bad arguments");
    if (N == 0) return TRXL_OK;
    attn_ranges_kernel<<<trxl_cdiv(N, 8), 256, 0, st>>>(mask, win_index, ep_index, sample_index, N, L, ranges);
    TRXL_CHECK_LAUNCH("attention_ranges");
    return TRXL_OK;
}

int attn_tc_table_add_pe(const float* table, const float* pe, float* out, long long E, int M, int B, int D, int layer_norm,
                         cudaStream_t st) {
    TRXL_CHECK_ARG(table && out && D % 4 == 0 && (pe || layer_norm), "table_add_pe: bad arguments");
    const long long total4 = E * M * B * (D / 4);
    if (total4 == 0) return TRXL_OK;
    if (layer_norm) {
        const long long rows = E * M * B;
        table_add_pe_ln_kernel<<<trxl_cdiv(rows, 8), 256, 0, st>>>(table, pe, out, rows, M, B, D);
        TRXL_CHECK_LAUNCH("table_add_pe_ln");
        return TRXL_OK;
    }
    int blocks = trxl_cdiv(total4, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    table_add_pe_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(table), reinterpret_cast<const float4*>(pe),
                                               reinterpret_cast<float4*>(out), total4, M, B, D / 4);
    TRXL_CHECK_LAUNCH("table_add_pe");
    return TRXL_OK;
}

// qk (N, H, D) -> P (N*H, ld) [saved for the backward], ctx (N, H, D)
int attn_tc_forward(const AttnTcArgs& a, const float* qk, float* P, float* ctx, cudaStream_t st) {
    const long long ld = attn_tc_row_floats(a.slots);
    const int rows = a.N * a.H;
    GemmArgs g1 = grouped_args(a, qk, a.D, a.D, (int)a.slots, 1, P, ld);               // S = QK . Xpe^T     (B K-major: rows = slots)
    TRXL_CHECK_ARG(trxl_tc_gemm_eligible(g1), "attention_tc: operands not TMA-eligible");
    trxl_prof_begin(2, a.N, st);
    trxl_prof_aux(2, a.n_tiles);
    TRXL_PROPAGATE(trxl_tc_gemm(g1, tile_n((int)a.slots), st));
    // (a masked softmax fused into this GEMM's epilogue -- straight from the TMEM accumulators, row halves exchanged between the two
    // warps of a lane quarter -- was measured SLOWER, 83.5 vs 76.3 us per forward on the same box: a CTA's 8 epilogue warps
    // serialise what this kernel spreads over the whole GPU; likewise for the dscore pass, 108 vs 71 us)
    if (a.slots <= 256) attn_softmax_kernel<8><<<trxl_cdiv(rows, 8), 256, 0, st>>>(P, ld, (int)a.slots, a.ranges, a.H, rows, a.scale, a.L);
    else attn_softmax_kernel<32><<<trxl_cdiv(rows, 8), 256, 0, st>>>(P, ld, (int)a.slots, a.ranges, a.H, rows, a.scale, a.L);
    TRXL_CHECK_LAUNCH("attention_softmax");
    GemmArgs g2 = grouped_args(a, P, ld, (int)a.slots, a.D, 0, ctx, a.D);               // ctx = P . Xpe      (B MN-major: k = slot)
    TRXL_PROPAGATE(trxl_tc_gemm(g2, tile_n(a.D), st));
    trxl_prof_end(2, st);
    return TRXL_OK;
}

// dctx (N, H, D), P from the forward -> dqk (N, H, D); scratch (N*H, ld)
int attn_tc_backward(const AttnTcArgs& a, const float* P, const float* dctx, float* scratch, float* dqk, cudaStream_t st) {
    const long long ld = attn_tc_row_floats(a.slots);
    const int rows = a.N * a.H;
    GemmArgs g1 = grouped_args(a, dctx, a.D, a.D, (int)a.slots, 1, scratch, ld);        // dP = dctx . Xpe^T
    trxl_prof_begin(3, a.N, st);
    trxl_prof_aux(3, a.n_tiles);
    TRXL_PROPAGATE(trxl_tc_gemm(g1, tile_n((int)a.slots), st));
    if (a.slots <= 256) attn_dscore_kernel<8><<<trxl_cdiv(rows, 8), 256, 0, st>>>(scratch, P, ld, (int)a.slots, a.ranges, a.H, rows, 1.f / a.scale);
    else attn_dscore_kernel<32><<<trxl_cdiv(rows, 8), 256, 0, st>>>(scratch, P, ld, (int)a.slots, a.ranges, a.H, rows, 1.f / a.scale);
    TRXL_CHECK_LAUNCH("attention_dscore");
    GemmArgs g2 = grouped_args(a, scratch, ld, (int)a.slots, a.D, 0, dqk, a.D);         // dqk = dS . Xpe
    TRXL_PROPAGATE(trxl_tc_gemm(g2, tile_n(a.D), st));
    trxl_prof_end(3, st);
    return TRXL_OK;
}
