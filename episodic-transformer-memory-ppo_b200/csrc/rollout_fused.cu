// Fused trunk forward for rollout-sized batches: ONE launch, one 4-CTA thread-block cluster per sample.
//
// Every layer after the CNN encoder is row-independent (a sample never reads another sample's activations), and at
// rollout batch sizes (one row per env worker) the layered path (model.cu) is pure latency: ~45 dependent launches of
// 3-15 us, each using a handful of SMs.  Here a cluster of four CTAs carries one sample through
// lin_hidden -> embedding -> B transformer blocks -> policy / value heads in a single launch:
//   * the activation vectors (D floats) live replicated in the shared memory of all four CTAs;
//   * every matrix-vector product is cut by output rows: a CTA streams only its quarter of the weight matrix from L2
//     (eight rows per warp, all of a row group's 16-byte loads issued before the first FMA, so a D x D layer costs about
//     one L2 round trip) and writes its quarter of the result into all four CTAs through distributed shared memory,
//     followed by one cluster barrier;
//   * attention heads are dealt to the CTAs (head h -> CTA h % 4): K-fold, the two passes over the episodic-memory
//     window and the V-unfold of a head stay inside one CTA;
//   * the window rows of a block depend only on the episode table, not on the activations, so when they need no
//     positional add (no positional encoding, or a table that already carries it) every CTA pulls the NEXT block's L rows
//     into shared memory with bulk async copies (cp.async.bulk, one per row, completion on an mbarrier) while the current
//     block's projections run: the energies / context passes then read shared memory only (measured on B200 at c3: the
//     energies pass was 4 dependent L2 round trips, 13.9 k of a block's 43 k clocks);
//   * LayerNorms / gates' elementwise parts are recomputed by every CTA on its replicated vectors.
// Same math and the same query-side fold as the layered kernels (attention.cu), fp32 throughout; inference only (nothing is
// saved for a backward pass).  32 samples -> 128 CTAs on 128 SMs.
#include "rollout_fused.cuh"

#include <cooperative_groups.h>

#include "tc_common.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace {

constexpr int RF_THREADS = 256, RF_WARPS = RF_THREADS / 32, RF_CL = 4;
constexpr int RF_PART = 4096;                      // floats of cross-warp / cross-group partial sums (>= RF_WARPS * D)
constexpr float LN_EPS = 1e-5f;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float dot4f(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// one row global -> shared through the bulk-copy engine; `bytes` (multiple of 16) are credited to the mbarrier on arrival
__device__ __forceinline__ void bulk_row_g2s(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

struct Cl {
    int rank, tid, warp, lane;
    float* local;                  // this CTA's dynamic shared memory
    float* peer[RF_CL];            // the same array in every CTA of the cluster (generic addresses into DSMEM)
};
// write v to element `p[j]` (p inside the dynamic shared memory) of every CTA of the cluster
__device__ __forceinline__ void bcast_store(const Cl& c, float* p, int j, float v) {
    const long long off = (p - c.local) + j;
#pragma unroll
    for (int q = 0; q < RF_CL; ++q) c.peer[q][off] = v;
}
// this CTA's share [j0, j1) of n output rows (multiples of 8 so a warp's row group never straddles two CTAs)
__device__ __forceinline__ void slice(const Cl& c, int n, int& j0, int& j1) {
    const int per = (((n + RF_CL - 1) / RF_CL) + 7) & ~7;
    j0 = min(n, c.rank * per);
    j1 = min(n, j0 + per);
}

// y[j] = act( W[j, :] . x + bias[j] ) + resid[j]   for j in [j0, j1).  W row-major (rows x K).  One warp per ROWS rows; every
// lane issues the 2 * ROWS * KU 16-byte loads of its slice of those rows (KU chunks of 256 k) before the first FMA -- the
// products are pure L2 latency, so the depth in flight is what sets their speed (8 x 1: the D x D layers, one round trip;
// 8 x 2: lin_hidden's long rows; 16 x 1: the heads' 2 * hid rows).  The accumulation order per row does not depend on
// ROWS / KU.  The result goes to y[j] of this CTA, or of every CTA in the cluster when bcast (the caller then runs a cluster
// barrier), and/or to y_global[j].
template <int ROWS = 8, int KU = 1>
__device__ void gemv_range(const Cl& c, const float* __restrict__ W, int K, int j0, int j1, const float* x,
                           const float* __restrict__ bias, bool relu, const float* resid, float* y, float* y_global, bool bcast) {
    const bool vec = ((K & 3) == 0) && ((((uintptr_t)W) & 15) == 0) && ((((uintptr_t)x) & 15) == 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jb = j0 + c.warp * ROWS; jb < j1; jb += RF_WARPS * ROWS) {
        const int nrow = min(ROWS, j1 - jb);
        float acc[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
        if (vec) {
            for (int k0 = c.lane * 4; k0 < K; k0 += 256 * KU) {
                float4 w[ROWS][2 * KU];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const float* wr = W + (long long)(jb + (r < nrow ? r : 0)) * K + k0;
#pragma unroll
                    for (int u = 0; u < 2 * KU; ++u) w[r][u] = (k0 + 128 * u < K) ? ldg4(wr + 128 * u) : zero4;
                }
                float4 xv[2 * KU];
#pragma unroll
                for (int u = 0; u < 2 * KU; ++u) xv[u] = (k0 + 128 * u < K) ? *reinterpret_cast<const float4*>(x + k0 + 128 * u) : zero4;
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    if (u > 0 && k0 + 256 * u >= K) break;        // (a lane past the end of the row only holds zeros)
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) acc[r] += dot4f(w[r][2 * u], xv[2 * u]) + dot4f(w[r][2 * u + 1], xv[2 * u + 1]);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                if (r >= nrow) continue;
                const float* wr = W + (long long)(jb + r) * K;
                for (int k = c.lane; k < K; k += 32) acc[r] = fmaf(__ldg(wr + k), x[k], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = wsum(acc[r]);
        if (c.lane < nrow) {
            float v = acc[0];
#pragma unroll
            for (int r = 1; r < ROWS; ++r) if (c.lane == r) v = acc[r];
            const int j = jb + c.lane;
            if (bias) v += bias[j];
            if (relu) v = fmaxf(v, 0.f);
            if (resid) v += resid[j];
            if (y) {
                if (bcast) bcast_store(c, y, j, v);
                else y[j] = v;
            }
            if (y_global) y_global[j] = v;
        }
    }
}

// out[j] = (sum_{d < dh} q[d] * Wk[d * D + j]) * gamma[j]   (Wk points at the head's first row).  thread = (float4 column,
// d-group): a thread's <= 8 loads per step are independent, the d-groups are summed through shared memory.
__device__ void fold_k(const Cl& c, const float* __restrict__ Wk, int D, int dh, const float* q, const float* __restrict__ gamma,
                       float* part, float* out) {
    const int ncol4 = D >> 2;
    int ng = RF_THREADS / ncol4;
    if (ng > dh) ng = dh;
    if (ng * D > RF_PART) ng = RF_PART / D;
    if (ng < 1) ng = 1;
    const int dper = (dh + ng - 1) / ng;
    if (c.tid < ncol4 * ng) {
        const int jq = c.tid % ncol4, dg = c.tid / ncol4;
        const int d0 = dg * dper, d1 = min(dh, d0 + dper);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int d = d0; d < d1; d += 8) {
            float4 w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = (d + u < d1) ? ldg4(Wk + (long long)(d + u) * D + jq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float qv = (d + u < d1) ? q[d + u] : 0.f;
                acc.x = fmaf(qv, w[u].x, acc.x); acc.y = fmaf(qv, w[u].y, acc.y);
                acc.z = fmaf(qv, w[u].z, acc.z); acc.w = fmaf(qv, w[u].w, acc.w);
            }
        }
        *reinterpret_cast<float4*>(part + dg * D + jq * 4) = acc;
    }
    __syncthreads();
    for (int j = c.tid; j < D; j += RF_THREADS) {
        float s = 0.f;
        for (int g = 0; g < ng; ++g) s += part[g * D + j];
        if (gamma) s *= gamma[j];
        out[j] = s;
    }
    __syncthreads();
}

// LayerNorm of a D-vector in shared memory (whole CTA, every CTA on its own replica); y may alias x
__device__ void layer_norm(const Cl& c, const float* x, const float* x2, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float* y, int D, float* red) {
    float s = 0.f;
    for (int j = c.tid; j < D; j += RF_THREADS) s += x[j] + (x2 ? x2[j] : 0.f);
    s = wsum(s);
    if (c.lane == 0) red[c.warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < RF_WARPS; ++w) tot += red[w];
    const float mu = tot / (float)D;
    __syncthreads();
    float v = 0.f;
    for (int j = c.tid; j < D; j += RF_THREADS) {
        const float d = x[j] + (x2 ? x2[j] : 0.f) - mu;
        v = fmaf(d, d, v);
    }
    v = wsum(v);
    if (c.lane == 0) red[c.warp] = v;
    __syncthreads();
    float var = 0.f;
    for (int w = 0; w < RF_WARPS; ++w) var += red[w];
    const float rs = rsqrtf(var / (float)D + LN_EPS);
    __syncthreads();
    for (int j = c.tid; j < D; j += RF_THREADS) y[j] = (x[j] + (x2 ? x2[j] : 0.f) - mu) * rs * gamma[j] + beta[j];
    __syncthreads();
}

// GRU gate (reference transformer.py:295-298): out = (1 - z) x + z tanh(Wg y + Ug (r x)), r = s(Wr y + Ur x),
// z = s(Wz y + Uz x - bg).  Matrix rows are cut across the cluster; two cluster barriers (after r.x and after out).
__device__ void gru_gate(cg::cluster_group& cluster, const Cl& c, const float* P, const RfGate& g, const float* x, const float* yv,
                         float* out, int D, float* t1, float* t2, float* t3) {
    int j0, j1;
    slice(c, D, j0, j1);
    const long long DD = (long long)D * D;
    gemv_range(c, P + g.Wr, D, j0, j1, yv, nullptr, false, nullptr, t1, nullptr, false);
    gemv_range(c, P + g.Wr + DD, D, j0, j1, yv, nullptr, false, nullptr, t2, nullptr, false);
    __syncthreads();
    gemv_range(c, P + g.Ur, D, j0, j1, x, nullptr, false, t1, t1, nullptr, false);
    gemv_range(c, P + g.Ur + DD, D, j0, j1, x, nullptr, false, t2, t2, nullptr, false);
    __syncthreads();
    // every CTA must be done reading the previous contents of t1 (as a full vector) before r.x overwrites it cluster-wide:
    // t1 is only ever read in full by the Ug product below, which sits behind the barrier that follows
    for (int j = j0 + c.tid; j < j1; j += RF_THREADS) {
        const float r = 1.f / (1.f + expf(-t1[j]));
        const float z = 1.f / (1.f + expf(-(t2[j] - P[g.bg + j])));
        t2[j] = z;
        bcast_store(c, t1, j, r * x[j]);          // r (.) x, needed in full by Ug
    }
    cluster.sync();
    gemv_range(c, P + g.Wr + 2 * DD, D, j0, j1, yv, nullptr, false, nullptr, t3, nullptr, false);      // Wg y
    __syncthreads();
    gemv_range(c, P + g.Ug, D, j0, j1, t1, nullptr, false, t3, t3, nullptr, false);                    // + Ug (r.x)
    __syncthreads();
    for (int j = j0 + c.tid; j < j1; j += RF_THREADS) {
        const float h = tanhf(t3[j]), z = t2[j];
        bcast_store(c, out, j, (1.f - z) * x[j] + z * h);
    }
    cluster.sync();
}

// Start the bulk copies of one block's visible window rows into s_cache (whole CTA calls; the rows were last touched by
// generic-proxy accesses that a __syncthreads has already ordered before this call).
__device__ void window_prefetch(const Cl& c, const float* tab, const int* s_win, const int* s_vis, int L, int D, int nvis,
                                float* s_cache, uint64_t* bar) {
    tc::fence_async_proxy();
    if (c.tid == 0) mbar_expect_tx(bar, (uint32_t)nvis * (uint32_t)D * 4u);
    // a warp issues its lanes' bulk copies one after the other (~50 clocks each, measured), so the rows are dealt across the
    // warps first: row l -> warp l % 8, lane l / 8
    for (int l = c.warp + c.lane * RF_WARPS; l < L; l += RF_THREADS)
        if (s_vis[l]) bulk_row_g2s(s_cache + (long long)l * D, tab + s_win[l], (uint32_t)D * 4u, bar);
}

__global__ void __cluster_dims__(RF_CL, 1, 1) __launch_bounds__(RF_THREADS, 1) rollout_fused_kernel(const RfArgs a) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int D = a.D, H = a.H, L = a.L, dh = D / H, hid = a.hid;
    const int Lp = (L + 3) & ~3;
    // shared-memory carve (floats); identical in every CTA so that offsets address the same vector cluster-wide
    float* s_feat = sm;                               // [featp]
    float* s_h = s_feat + ((a.feat + 3) & ~3);        // [D] current block input / output
    float* s_a = s_h + D;                             // [D] scratch vectors
    float* s_b = s_a + D;
    float* s_c = s_b + D;
    float* s_d = s_c + D;
    float* s_e = s_d + D;
    float* s_f = s_e + D;
    float* s_g = s_f + D;
    float* s_qk = s_g + D;                            // [D] folded query of the head being processed
    float* s_ctx = s_qk + D;                          // [D] context of the head being processed
    float* s_p = s_ctx + D;                           // [Lp] energies -> weights
    float* s_mu = s_p + Lp;                           // [Lp]
    float* s_rs = s_mu + Lp;                          // [Lp]
    float* s_hd = s_rs + Lp;                          // [2*hid] head hiddens
    float* s_red = s_hd + ((2 * hid + 3) & ~3);       // [2 * RF_WARPS + 4]
    float* s_part = s_red + 2 * RF_WARPS + 4;         // [RF_PART]
    int* s_win = reinterpret_cast<int*>(s_part + RF_PART);   // [L]
    int* s_pe = s_win + Lp;                           // [L]
    int* s_vis = s_pe + Lp;                           // [L]
    float* s_cache = reinterpret_cast<float*>(s_vis + Lp);   // [L][D] window rows (+PE) of the head in flight, if it fits
    __shared__ int s_any;
    __shared__ int s_nvis;
    __shared__ __align__(8) uint64_t s_wbar;          // completion of the window prefetch in flight

    int mark_i = 0;
#define RF_MARK() do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[mark_i] = clock64(); ++mark_i; } while (0)
    RF_MARK();
    Cl c;
    c.rank = (int)cluster.block_rank();
    c.tid = threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
    c.local = sm;
#pragma unroll
    for (int q = 0; q < RF_CL; ++q) c.peer[q] = cluster.map_shared_rank(sm, q);
    const int n = blockIdx.x / RF_CL, tid = c.tid, warp = c.warp, lane = c.lane;
    const float* P = a.P;
    const long long row = a.sample_index ? a.sample_index[n] : n;
    const long long ep = a.ep_index ? a.ep_index[row] : row;
    const bool pre = a.ln == 1, post = a.ln == 2;

    // ---- stage the sample's inputs (every CTA its own copy) ----
    if (tid == 0) s_any = 0;
    for (int k = tid; k < a.feat; k += RF_THREADS) s_feat[k] = a.feat_in[(long long)n * a.feat + k];
    __syncthreads();
    int any = 0;
    for (int l = tid; l < L; l += RF_THREADS) {
        const int m = a.mask ? a.mask[row * L + l] : 1;
        s_vis[l] = m;
        any |= m;
        s_win[l] = (int)((a.win_index ? a.win_index[row * L + l] : (long long)l) * a.B * D);
        s_pe[l] = (int)((a.pe_index ? a.pe_index[row * L + l] : 0) * D);
    }
    if (any) s_any = 1;
    if (tid == 0) { tc::mbar_init(&s_wbar, 1); tc::mbar_init_fence(); }
    __syncthreads();
    const bool all_masked = (s_any == 0);
    if (all_masked) {
        for (int l = tid; l < L; l += RF_THREADS) s_vis[l] = 1;
        __syncthreads();
    }
    const float* pe = (a.pe_mode == 1 && a.pe_index) ? a.pe_table : ((a.pe_mode == 2 && a.pe_index) ? P + a.pos : nullptr);
    // window rows through the bulk-copy engine, one block ahead (only CTAs that own a head; rows must need no positional add)
    const bool bulk = a.cache_window && pe == nullptr && c.rank < H;
    if (bulk) {
        int cnt = 0;
        for (int l0 = 0; l0 < L; l0 += RF_THREADS) cnt += __syncthreads_count((l0 + tid < L) && s_vis[l0 + tid]);
        if (tid == 0) s_nvis = cnt;
        window_prefetch(c, a.table + (ep * a.slots) * a.B * (long long)D, s_win, s_vis, L, D, cnt, s_cache, &s_wbar);
    }
    cluster.sync();                                   // every CTA's shared memory is live before the first remote store
    RF_MARK();                                        // 1: inputs staged

    int j0, j1;
    // ---- lin_hidden + embedding ----
    slice(c, D, j0, j1);
    gemv_range<8, 2>(c, P + a.Wh, a.feat, j0, j1, s_feat, P + a.bh, true, nullptr, s_a, nullptr, true);
    cluster.sync();
    RF_MARK();                                        // 2: lin_hidden
    gemv_range(c, P + a.We, D, j0, j1, s_a, P + a.be, true, nullptr, s_h, nullptr, true);
    cluster.sync();
    RF_MARK();                                        // 3: embedding

    const float scale = sqrtf((float)D);
    const int NC = (D + 127) / 128;                   // float4 chunks per lane per row (<= 4)
    for (int blk = 0; blk < a.B; ++blk) {
        const long long bo = (long long)blk * a.blk_stride;
        const float* tab = a.table + ((ep * a.slots) * a.B + blk) * (long long)D;
        if (c.rank == 0)
            for (int j = tid; j < D; j += RF_THREADS) a.out_mem[((long long)n * a.B + blk) * D + j] = s_h[j];
        // q_in
        const float* q_in = s_h;
        if (pre) {
            layer_norm(c, s_h, nullptr, P + a.b0.n1w + bo, P + a.b0.n1b + bo, s_a, D, s_red);
            q_in = s_a;
        }
        // Q = q_in Wq^T  -> s_b (replicated)
        slice(c, D, j0, j1);
        gemv_range(c, P + a.b0.Wq + bo, D, j0, j1, q_in, nullptr, false, nullptr, s_b, nullptr, true);
        cluster.sync();
        RF_MARK();                                    // block + 0: Q
        const float* Wk = P + a.b0.Wk + bo;
        // ---- the heads dealt to this CTA ----
        for (int h = c.rank; h < H; h += RF_CL) {
            // folded query: qk[j] = sum_{d in head h} Q[d] Wk[d, j] (* gamma_kv[j]);  pre-LN extras: qkb, sum_j qk[j]
            fold_k(c, Wk + (long long)h * dh * D, D, dh, s_b + h * dh, pre ? P + a.b0.nkw + bo : nullptr, s_part, s_qk);
            if (h == c.rank) RF_MARK();               // block + 1: K-fold
            float qkb = 0.f, sg = 0.f;
            if (pre) {
                // kb[d] = Wk[d, :] . beta_kv for this head's rows -> s_d[h*dh ..]
                gemv_range(c, Wk, D, h * dh, (h + 1) * dh, P + a.b0.nkb + bo, nullptr, false, nullptr, s_d, nullptr, false);
                __syncthreads();
                float qb = 0.f, s1 = 0.f;
                for (int d = tid; d < dh; d += RF_THREADS) qb = fmaf(s_b[h * dh + d], s_d[h * dh + d], qb);
                for (int j = tid; j < D; j += RF_THREADS) s1 += s_qk[j];
                qb = wsum(qb);
                s1 = wsum(s1);
                if (lane == 0) { s_red[warp] = qb; s_red[RF_WARPS + warp] = s1; }
                __syncthreads();
                for (int w = 0; w < RF_WARPS; ++w) { qkb += s_red[w]; sg += s_red[RF_WARPS + w]; }
                __syncthreads();
            }
            // this lane's slice of the folded query
            float4 qv[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int col = cc * 128 + lane * 4;
                qv[cc] = (cc < NC && col < D) ? *reinterpret_cast<const float4*>(s_qk + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (bulk && h == c.rank) tc::mbar_wait(&s_wbar, (uint32_t)(blk & 1));      // this block's window has landed
            // ---- pass 1: energies ----
            // prefetched window: rows are in shared memory; eight rows per warp and trip, their reductions interleaved, lane u
            // finishes row l0 + u (same per-row arithmetic as the streaming form below)
            if (bulk) {
                for (int l0 = warp * 8; l0 < L; l0 += RF_WARPS * 8) {
                    float d[8], s1[8], s2[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int l = l0 + u;
                        d[u] = 0.f; s1[u] = 0.f; s2[u] = 0.f;
                        if (l < L && s_vis[l]) {                        // warp-uniform
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                const int col = cc * 128 + lane * 4;
                                if (cc < NC && col < D) {
                                    const float4 xv = *reinterpret_cast<const float4*>(s_cache + l * D + col);
                                    d[u] += dot4f(qv[cc], xv);
                                    if (pre) { s1[u] += xv.x + xv.y + xv.z + xv.w; s2[u] += dot4f(xv, xv); }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) d[u] += __shfl_xor_sync(FULL, d[u], o);
                    }
                    if (pre) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                            for (int u = 0; u < 8; ++u) { s1[u] += __shfl_xor_sync(FULL, s1[u], o); s2[u] += __shfl_xor_sync(FULL, s2[u], o); }
                        }
                    }
                    float dv = d[0], s1v = s1[0], s2v = s2[0];
#pragma unroll
                    for (int u = 1; u < 8; ++u) if (lane == u) { dv = d[u]; s1v = s1[u]; s2v = s2[u]; }
                    const int l = l0 + lane;
                    if (lane < 8 && l < L && s_vis[l]) {
                        float e = dv;
                        if (pre) {
                            const float mu = s1v / (float)D;
                            const float rstd = rsqrtf(fmaxf(s2v / (float)D - mu * mu, 0.f) + LN_EPS);
                            s_mu[l] = mu; s_rs[l] = rstd;
                            e = fmaf(rstd, dv - mu * sg, qkb);
                        }
                        s_p[l] = all_masked ? 0.f : __fdiv_rn(e, scale);
                    }
                }
            }
            // streaming form: rows (+ positional rows) come from L2, four window rows in flight per warp
            for (int l0 = warp * 4; l0 < L && !bulk; l0 += RF_WARPS * 4) {
                float4 x[4][4];
                bool vis[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int l = l0 + u;
                    vis[u] = (l < L) && s_vis[l];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int col = cc * 128 + lane * 4;
                        x[u][cc] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (vis[u] && cc < NC && col < D) x[u][cc] = ldg4(tab + s_win[l] + col);
                    }
                }
                if (pe) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int col = cc * 128 + lane * 4;
                            if (vis[u] && cc < NC && col < D) {
                                const float4 p4 = ldg4(pe + s_pe[l0 + u] + col);
                                x[u][cc].x += p4.x; x[u][cc].y += p4.y; x[u][cc].z += p4.z; x[u][cc].w += p4.w;
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!vis[u]) continue;                      // warp-uniform
                    const int l = l0 + u;
                    float d = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int col = cc * 128 + lane * 4;
                        if (a.cache_window && cc < NC && col < D) *reinterpret_cast<float4*>(s_cache + l * D + col) = x[u][cc];
                        d += dot4f(qv[cc], x[u][cc]);
                        s1 += x[u][cc].x + x[u][cc].y + x[u][cc].z + x[u][cc].w;
                        s2 += dot4f(x[u][cc], x[u][cc]);
                    }
                    d = wsum(d);
                    float e = d;
                    if (pre) {
                        s1 = wsum(s1);
                        s2 = wsum(s2);
                        const float mu = s1 / (float)D;
                        const float rstd = rsqrtf(fmaxf(s2 / (float)D - mu * mu, 0.f) + LN_EPS);
                        if (lane == 0) { s_mu[l] = mu; s_rs[l] = rstd; }
                        e = fmaf(rstd, d - mu * sg, qkb);
                    }
                    if (lane == 0) s_p[l] = all_masked ? 0.f : __fdiv_rn(e, scale);
                }
            }
            __syncthreads();
            if (h == c.rank) RF_MARK();               // block + 2: energies
            // ---- softmax over the L energies (warp 0) ----
            if (warp == 0) {
                float m = -INFINITY;
                for (int l = lane; l < L; l += 32) if (s_vis[l]) m = fmaxf(m, s_p[l]);
                m = wmax(m);
                float s = 0.f;
                for (int l = lane; l < L; l += 32) {
                    const float p = s_vis[l] ? __expf(s_p[l] - m) : 0.f;
                    s_p[l] = p;
                    s += p;
                }
                s = wsum(s);
                const float inv = 1.f / s;
                for (int l = lane; l < L; l += 32) s_p[l] *= inv;
            }
            __syncthreads();
            if (h == c.rank) RF_MARK();               // block + 3: softmax
            // ---- pass 2: ctx = sum_l p[l] x_l (rows come back from L2), warps merged in warp order ----
            {
                float4 acc[4];
                float csum = 0.f;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) acc[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
                // window in shared memory: two row pairs (l0, l0 + 1, l0 + 16, l0 + 17) per trip, accumulated in the order of the
                // two-row loop below
                for (int lb = warp * 2; lb < L && a.cache_window; lb += RF_WARPS * 4) {
                    float4 x[4][4];
                    bool vis[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int l = lb + (q >> 1) * (RF_WARPS * 2) + (q & 1);
                        vis[q] = (l < L) && s_vis[l];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int col = cc * 128 + lane * 4;
                            x[q][cc] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (vis[q] && cc < NC && col < D) x[q][cc] = *reinterpret_cast<const float4*>(s_cache + l * D + col);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (!vis[q]) continue;
                        const int l = lb + (q >> 1) * (RF_WARPS * 2) + (q & 1);
                        const float wgt = s_p[l] * (pre ? s_rs[l] : 1.f);
                        if (pre) csum = fmaf(wgt, s_mu[l], csum);
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            acc[cc].x = fmaf(wgt, x[q][cc].x, acc[cc].x); acc[cc].y = fmaf(wgt, x[q][cc].y, acc[cc].y);
                            acc[cc].z = fmaf(wgt, x[q][cc].z, acc[cc].z); acc[cc].w = fmaf(wgt, x[q][cc].w, acc[cc].w);
                        }
                    }
                }
                for (int l0 = warp * 2; l0 < L && !a.cache_window; l0 += RF_WARPS * 2) {
                    float4 x[2][4];
                    bool vis[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int l = l0 + u;
                        vis[u] = (l < L) && s_vis[l];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int col = cc * 128 + lane * 4;
                            x[u][cc] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (vis[u] && cc < NC && col < D) {
                                x[u][cc] = ldg4(tab + s_win[l] + col);
                                if (pe) {
                                    const float4 p4 = ldg4(pe + s_pe[l] + col);
                                    x[u][cc].x += p4.x; x[u][cc].y += p4.y; x[u][cc].z += p4.z; x[u][cc].w += p4.w;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (!vis[u]) continue;
                        const int l = l0 + u;
                        const float wgt = s_p[l] * (pre ? s_rs[l] : 1.f);
                        if (pre) csum = fmaf(wgt, s_mu[l], csum);
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            acc[cc].x = fmaf(wgt, x[u][cc].x, acc[cc].x); acc[cc].y = fmaf(wgt, x[u][cc].y, acc[cc].y);
                            acc[cc].z = fmaf(wgt, x[u][cc].z, acc[cc].z); acc[cc].w = fmaf(wgt, x[u][cc].w, acc[cc].w);
                        }
                    }
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int col = cc * 128 + lane * 4;
                    if (cc < NC && col < D) *reinterpret_cast<float4*>(s_part + warp * D + col) = acc[cc];
                }
                if (lane == 0) s_red[warp] = csum;
                __syncthreads();
                float cs = 0.f;
                if (pre)
                    for (int w = 0; w < RF_WARPS; ++w) cs += s_red[w];
                for (int j = tid; j < D; j += RF_THREADS) {
                    float s = 0.f;
                    for (int w = 0; w < RF_WARPS; ++w) s += s_part[w * D + j];
                    // pre-LN: ctx = gamma (.) (ctx_hat - csum) + beta  (sum_l p = 1)
                    s_ctx[j] = pre ? (s - cs) * P[a.b0.nkw + bo + j] + P[a.b0.nkb + bo + j] : s;
                }
                __syncthreads();
            }
            if (h == c.rank) RF_MARK();               // block + 4: context
            // V-unfold of this head: att[d] = Wv[d, :] . ctx  for d in the head's rows -> s_c of every CTA
            gemv_range(c, P + a.b0.Wv + bo, D, h * dh, (h + 1) * dh, s_ctx, nullptr, false, nullptr, s_c, nullptr, true);
            __syncthreads();
        }
        if (bulk && blk + 1 < a.B) window_prefetch(c, tab + D, s_win, s_vis, L, D, s_nvis, s_cache, &s_wbar);
        cluster.sync();
        RF_MARK();                                    // block + 5: V-unfold
        // fc_out (+ residual when not gated) -> s_b
        slice(c, D, j0, j1);
        gemv_range(c, P + a.b0.Wo + bo, D, j0, j1, s_c, P + a.b0.bo + bo, false, a.gtrxl ? nullptr : s_h, s_b, nullptr, true);
        cluster.sync();
        RF_MARK();                                    // block + 6: fc_out
        float* h1 = s_b;                               // h1pre
        if (a.gtrxl) {
            RfGate g1 = a.b0.g1; g1.Wr += bo; g1.Ur += bo; g1.Ug += bo; g1.bg += bo;
            gru_gate(cluster, c, P, g1, s_h, s_b, s_c, D, s_a, s_d, s_e);        // x = h_in, y = att
            h1 = s_c;
        }
        if (post) {
            layer_norm(c, h1, nullptr, P + a.b0.n1w + bo, P + a.b0.n1b + bo, s_a, D, s_red);
            h1 = s_a;
        }
        // h1 lives in s_a (post), s_c (gated, no post) or s_b (plain)
        const float* h_ = h1;
        if (pre) {
            layer_norm(c, h1, nullptr, P + a.b0.n2w + bo, P + a.b0.n2b + bo, s_f, D, s_red);
            h_ = s_f;
        }
        RF_MARK();                                    // block + 7: gate / norms before the feed-forward
        slice(c, D, j0, j1);
        gemv_range(c, P + a.b0.Wff + bo, D, j0, j1, h_, P + a.b0.bff + bo, true, nullptr, s_g, nullptr, true);
        cluster.sync();
        RF_MARK();                                    // block + 8: feed-forward
        if (a.gtrxl) {
            RfGate g2 = a.b0.g2; g2.Wr += bo; g2.Ur += bo; g2.Ug += bo; g2.bg += bo;
            float* pool[5] = {s_a, s_b, s_c, s_d, s_e};
            float* t[3];
            int nt = 0;
            for (int i = 0; i < 5 && nt < 3; ++i)
                if (pool[i] != h1) t[nt++] = pool[i];
            gru_gate(cluster, c, P, g2, h1, s_g, s_f, D, t[0], t[1], t[2]);       // out_pre -> s_f
            if (post) {
                layer_norm(c, s_f, nullptr, P + a.b0.n2w + bo, P + a.b0.n2b + bo, s_h, D, s_red);
            } else {
                for (int j = tid; j < D; j += RF_THREADS) s_h[j] = s_f[j];
                __syncthreads();
            }
        } else if (post) {
            layer_norm(c, s_g, h1, P + a.b0.n2w + bo, P + a.b0.n2b + bo, s_h, D, s_red);
        } else {
            for (int j = tid; j < D; j += RF_THREADS) s_h[j] = s_g[j] + h1[j];
            __syncthreads();
        }
        // a faster CTA must not start the next block's broadcasts (Q -> s_b, ...) while a slower one still reads this block's
        cluster.sync();
        RF_MARK();                                    // block + 9: gate / norm after the feed-forward
    }
    // ---- heads ----
    slice(c, hid, j0, j1);
    gemv_range<16, 1>(c, P + a.Wp, D, j0, j1, s_h, P + a.bp, true, nullptr, s_hd, nullptr, true);
    gemv_range<16, 1>(c, P + a.Wlv, D, j0, j1, s_h, P + a.blv, true, nullptr, s_hd + hid, nullptr, true);
    cluster.sync();
    if (c.rank == 0) {
        // the sumA logit rows and the value row, one per warp, all in flight together (per-row arithmetic as in gemv_range)
        const bool vec = ((hid & 3) == 0) && ((((uintptr_t)(P + a.Wbr)) & 15) == 0) && ((((uintptr_t)(P + a.wval)) & 15) == 0);
        for (int j = warp; j <= a.sumA; j += RF_WARPS) {
            const bool is_value = (j == a.sumA);
            const float* wr = is_value ? P + a.wval : P + a.Wbr + (long long)j * hid;
            const float* xv = is_value ? s_hd + hid : s_hd;
            float acc = 0.f;
            if (vec) {
                for (int k0 = lane * 4; k0 < hid; k0 += 256) {
                    const bool has2 = (k0 + 128) < hid;
                    const float4 w0 = ldg4(wr + k0), w1 = has2 ? ldg4(wr + k0 + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 x0 = *reinterpret_cast<const float4*>(xv + k0);
                    const float4 x1 = has2 ? *reinterpret_cast<const float4*>(xv + k0 + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc += dot4f(w0, x0) + dot4f(w1, x1);
                }
            } else {
                for (int k = lane; k < hid; k += 32) acc = fmaf(__ldg(wr + k), xv[k], acc);
            }
            acc = wsum(acc);
            if (lane == 0) {
                if (is_value) a.value[n] = acc + P[a.bval];
                else a.logits[(long long)n * a.sumA + j] = acc + P[a.bbr + j];
            }
        }
    }
    RF_MARK();                                        // heads
#undef RF_MARK
}

}  // namespace

static size_t base_smem_bytes(const RfArgs& a) {
    const int Lp = (a.L + 3) & ~3;
    const size_t floats = ((a.feat + 3) & ~3) + 10 * (size_t)a.D + 3 * (size_t)Lp + ((2 * (size_t)a.hid + 3) & ~3) + 2 * RF_WARPS + 4 + RF_PART;
    return floats * 4 + (size_t)3 * Lp * 4 + 64;
}
static bool window_fits(const RfArgs& a) { return base_smem_bytes(a) + (size_t)a.L * a.D * 4 <= 200 * 1024; }
size_t rollout_fused_smem_bytes(const RfArgs& a) {
    return base_smem_bytes(a) + (window_fits(a) ? (size_t)a.L * a.D * 4 : 0);
}

bool rollout_fused_supported(const RfArgs& a) {
    return a.D % 4 == 0 && a.D >= 8 && a.D <= 512 && RF_WARPS * a.D <= RF_PART && a.H >= 1 && a.D % a.H == 0 && a.hid % 4 == 0 &&
           base_smem_bytes(a) <= 200 * 1024;
}

int rollout_fused_forward(const RfArgs& a, cudaStream_t st) {
    TRXL_CHECK_ARG(rollout_fused_supported(a), "rollout_fused: unsupported shape (D=%d H=%d L=%d)", a.D, a.H, a.L);
    if (a.N == 0) return TRXL_OK;
    const size_t smem = rollout_fused_smem_bytes(a);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(rollout_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { trxl_set_error("rollout_fused: cannot reserve %zu bytes of shared memory", smem); return TRXL_ERR_CUDA; }
        attr_smem = smem;
    }
    RfArgs b = a;
    b.cache_window = window_fits(a) ? 1 : 0;
    // TRXL_RF_TRACE present at the first call: every later call whose current value is > 0 prints cluster 0's phase latencies
    // (synchronises the stream; debugging only -- tools/rf_trace.py)
    static int trace_armed = -1;
    if (trace_armed < 0) trace_armed = getenv("TRXL_RF_TRACE") ? 1 : 0;
    int trace = 0;
    if (trace_armed) { const char* e = getenv("TRXL_RF_TRACE"); trace = (e && atoi(e) > 0) ? 1 : 0; }
    static long long* trace_dev = nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (trace) cudaStreamIsCapturing(st, &cap);
    const bool tracing = trace && cap == cudaStreamCaptureStatusNone;
    if (tracing) {
        if (!trace_dev) cudaMalloc(&trace_dev, 256 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, 256 * sizeof(long long), st);
        b.trace = trace_dev;
    }
    rollout_fused_kernel<<<a.N * RF_CL, RF_THREADS, smem, st>>>(b);
    TRXL_CHECK_LAUNCH("rollout_fused");
    if (tracing) {
        long long host[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
        static const char* blk_names[10] = {"Q", "K-fold", "energies", "softmax", "context", "V-unfold", "fc_out", "pre-FF", "FF", "post-FF"};
        const int marks = 4 + 10 * a.B + 1;
        fprintf(stderr, "[rf-trace] N=%d D=%d H=%d L=%d B=%d feat=%d cache=%d total=%lld clk\n", a.N, a.D, a.H, a.L, a.B, a.feat, b.cache_window,
                host[marks - 1] - host[0]);
        for (int i = 1; i < marks && i < 256; ++i) {
            const char* name = i == 1 ? "stage" : (i == 2 ? "lin_hidden" : (i == 3 ? "embedding" : (i == marks - 1 ? "heads" : blk_names[(i - 4) % 10])));
            fprintf(stderr, "[rf-trace]  %2d %-10s %6lld clk\n", i, name, host[i] - host[i - 1]);
        }
    }
    return TRXL_OK;
}
