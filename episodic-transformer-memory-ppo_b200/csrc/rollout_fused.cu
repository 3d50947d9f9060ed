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
//     window (read in place from the episode table) and the V-unfold of a head stay inside one CTA;
//   * LayerNorms / gates' elementwise parts are recomputed by every CTA on its replicated vectors.
// Same math and the same query-side fold as the layered kernels (attention.cu), fp32 throughout; inference only (nothing is
// saved for a backward pass).  32 samples -> 128 CTAs on 128 SMs.
#include "rollout_fused.cuh"

#include <cooperative_groups.h>

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

// y[j] = act( W[j, :] . x + bias[j] ) + resid[j]   for j in [j0, j1).  W row-major (rows x K).  One warp per 8 rows; every
// lane issues the 16 loads of its slice of the 8 rows before the first FMA.  The result goes to y[j] of this CTA, or of
// every CTA in the cluster when bcast (the caller then runs a cluster barrier), and/or to y_global[j].
__device__ void gemv_range(const Cl& c, const float* __restrict__ W, int K, int j0, int j1, const float* x,
                           const float* __restrict__ bias, bool relu, const float* resid, float* y, float* y_global, bool bcast) {
    const bool vec = ((K & 3) == 0) && ((((uintptr_t)W) & 15) == 0) && ((((uintptr_t)x) & 15) == 0);
    for (int jb = j0 + c.warp * 8; jb < j1; jb += RF_WARPS * 8) {
        const int nrow = min(8, j1 - jb);
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.f;
        if (vec) {
            for (int k0 = c.lane * 4; k0 < K; k0 += 256) {
                const bool has2 = (k0 + 128) < K;
                float4 w[8][2];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float* wr = W + (long long)(jb + (r < nrow ? r : 0)) * K + k0;
                    w[r][0] = ldg4(wr);
                    w[r][1] = has2 ? ldg4(wr + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const float4 x0 = *reinterpret_cast<const float4*>(x + k0);
                const float4 x1 = has2 ? *reinterpret_cast<const float4*>(x + k0 + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] += dot4f(w[r][0], x0) + dot4f(w[r][1], x1);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r >= nrow) continue;
                const float* wr = W + (long long)(jb + r) * K;
                for (int k = c.lane; k < K; k += 32) acc[r] = fmaf(__ldg(wr + k), x[k], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = wsum(acc[r]);
        if (c.lane < nrow) {
            float v = acc[0];
#pragma unroll
            for (int r = 1; r < 8; ++r) if (c.lane == r) v = acc[r];
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
    __syncthreads();
    const bool all_masked = (s_any == 0);
    if (all_masked) {
        for (int l = tid; l < L; l += RF_THREADS) s_vis[l] = 1;
        __syncthreads();
    }
    const float* pe = a.pe_mode == 1 ? a.pe_table : (a.pe_mode == 2 ? P + a.pos : nullptr);
    cluster.sync();                                   // every CTA's shared memory is live before the first remote store

    int j0, j1;
    // ---- lin_hidden + embedding ----
    slice(c, D, j0, j1);
    gemv_range(c, P + a.Wh, a.feat, j0, j1, s_feat, P + a.bh, true, nullptr, s_a, nullptr, true);
    cluster.sync();
    gemv_range(c, P + a.We, D, j0, j1, s_a, P + a.be, true, nullptr, s_h, nullptr, true);
    cluster.sync();

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
        const float* Wk = P + a.b0.Wk + bo;
        // ---- the heads dealt to this CTA ----
        for (int h = c.rank; h < H; h += RF_CL) {
            // folded query: qk[j] = sum_{d in head h} Q[d] Wk[d, j] (* gamma_kv[j]);  pre-LN extras: qkb, sum_j qk[j]
            fold_k(c, Wk + (long long)h * dh * D, D, dh, s_b + h * dh, pre ? P + a.b0.nkw + bo : nullptr, s_part, s_qk);
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
            // ---- pass 1: energies (four window rows in flight per warp) ----
            for (int l0 = warp * 4; l0 < L; l0 += RF_WARPS * 4) {
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
            // ---- pass 2: ctx = sum_l p[l] x_l (rows come back from L2), warps merged in warp order ----
            {
                float4 acc[4];
                float csum = 0.f;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) acc[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int l0 = warp * 2; l0 < L; l0 += RF_WARPS * 2) {
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
                                if (a.cache_window) {
                                    x[u][cc] = *reinterpret_cast<const float4*>(s_cache + l * D + col);
                                } else {
                                    x[u][cc] = ldg4(tab + s_win[l] + col);
                                    if (pe) {
                                        const float4 p4 = ldg4(pe + s_pe[l] + col);
                                        x[u][cc].x += p4.x; x[u][cc].y += p4.y; x[u][cc].z += p4.z; x[u][cc].w += p4.w;
                                    }
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
            // V-unfold of this head: att[d] = Wv[d, :] . ctx  for d in the head's rows -> s_c of every CTA
            gemv_range(c, P + a.b0.Wv + bo, D, h * dh, (h + 1) * dh, s_ctx, nullptr, false, nullptr, s_c, nullptr, true);
            __syncthreads();
        }
        cluster.sync();
        // fc_out (+ residual when not gated) -> s_b
        slice(c, D, j0, j1);
        gemv_range(c, P + a.b0.Wo + bo, D, j0, j1, s_c, P + a.b0.bo + bo, false, a.gtrxl ? nullptr : s_h, s_b, nullptr, true);
        cluster.sync();
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
        slice(c, D, j0, j1);
        gemv_range(c, P + a.b0.Wff + bo, D, j0, j1, h_, P + a.b0.bff + bo, true, nullptr, s_g, nullptr, true);
        cluster.sync();
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
    }
    // ---- heads ----
    slice(c, hid, j0, j1);
    gemv_range(c, P + a.Wp, D, j0, j1, s_h, P + a.bp, true, nullptr, s_hd, nullptr, true);
    gemv_range(c, P + a.Wlv, D, j0, j1, s_h, P + a.blv, true, nullptr, s_hd + hid, nullptr, true);
    cluster.sync();
    if (c.rank == 0) {
        gemv_range(c, P + a.Wbr, hid, 0, a.sumA, s_hd, P + a.bbr, false, nullptr, nullptr, a.logits + (long long)n * a.sumA, false);
        gemv_range(c, P + a.wval, hid, 0, 1, s_hd + hid, P + a.bval, false, nullptr, nullptr, a.value + n, false);
    }
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
    rollout_fused_kernel<<<a.N * RF_CL, RF_THREADS, smem, st>>>(b);
    TRXL_CHECK_LAUNCH("rollout_fused");
    return TRXL_OK;
}
