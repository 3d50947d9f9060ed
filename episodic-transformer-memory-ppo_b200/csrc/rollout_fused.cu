// Fused trunk forward for rollout-sized batches: ONE launch, one CTA per sample.
//
// Every layer after the CNN encoder is row-independent (a sample never reads another sample's
// activations), so a CTA can carry its sample through lin_hidden -> embedding -> B transformer blocks
// (query projection, K-fold, window attention over the episode table, V-unfold, fc_out, residual or
// GRU gate, LayerNorms, feed-forward) -> policy / value heads with activations living in shared memory
// and nothing but __syncthreads between stages.  The layered path (model.cu) needs ~45 launches of
// 3-8 us each for a 32-sample step; here the step is bounded by streaming the ~10 MB of weights from
// L2 once per CTA.  Same math and same fold as the layered kernels (attention.cu), fp32 throughout.
#include "rollout_fused.cuh"

namespace {

constexpr int RF_THREADS = 512, RF_WARPS = RF_THREADS / 32;
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

// y[j] = act( W[j, :] . x(j) + bias[j] ) + resid[j],  W row-major (Nout x K), one warp per 4 output rows.
// `xsel` lets the input vector depend on the output row (per-head V-unfold): x = xbase + (j / group) * K.
__device__ void gemv_rows(const float* __restrict__ W, int K, int Nout, const float* xbase, int group, const float* __restrict__ bias,
                          bool relu, const float* resid, float* y, float* y_global, int warp, int lane) {
    const bool vec = ((K & 3) == 0) && ((((uintptr_t)W) & 15) == 0);
    for (int j0 = warp * 4; j0 < Nout; j0 += RF_WARPS * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = j0 + r;
            if (j >= Nout) continue;
            const float* x = xbase + (group > 0 ? (j / group) * K : 0);
            const float* wr = W + (long long)j * K;
            if (vec) {
                for (int k = lane * 4; k < K; k += 128)
                    acc[r] += dot4f(*reinterpret_cast<const float4*>(wr + k), *reinterpret_cast<const float4*>(x + k));
            } else {
                for (int k = lane; k < K; k += 32) acc[r] = fmaf(wr[k], x[k], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = wsum(acc[r]);
        if (lane < 4 && j0 + lane < Nout) {
            const int j = j0 + lane;
            float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
            if (bias) v += bias[j];
            if (relu) v = fmaxf(v, 0.f);
            if (resid) v += resid[j];
            if (y) y[j] = v;
            if (y_global) y_global[j] = v;
        }
    }
}

// LayerNorm of a D-vector in shared memory (block-wide), y may alias x
__device__ void layer_norm(const float* x, const float* x2, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float* y, float* y_global, int D, float* red, int tid, int warp, int lane) {
    float s = 0.f;
    for (int j = tid; j < D; j += RF_THREADS) s += x[j] + (x2 ? x2[j] : 0.f);
    s = wsum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < RF_WARPS; ++w) tot += red[w];
    const float mu = tot / (float)D;
    __syncthreads();
    float v = 0.f;
    for (int j = tid; j < D; j += RF_THREADS) {
        const float d = x[j] + (x2 ? x2[j] : 0.f) - mu;
        v = fmaf(d, d, v);
    }
    v = wsum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float var = 0.f;
    for (int w = 0; w < RF_WARPS; ++w) var += red[w];
    const float rs = rsqrtf(var / (float)D + LN_EPS);
    for (int j = tid; j < D; j += RF_THREADS) {
        const float o = (x[j] + (x2 ? x2[j] : 0.f) - mu) * rs * gamma[j] + beta[j];
        if (y) y[j] = o;
        if (y_global) y_global[j] = o;
    }
    __syncthreads();
}

// GRU gate (reference transformer.py:295-298) on D-vectors in shared memory; out may be global
__device__ void gru_gate(const float* P, const RfGate& g, const float* x, const float* yv, float* out, float* out_global, int D,
                         float* t1, float* t2, float* t3, int tid, int warp, int lane) {
    // t1 = Wr y + Ur x ; t2 = Wz y + Uz x - bg ; r = sigmoid(t1), z = sigmoid(t2)
    gemv_rows(P + g.Wr, D, D, yv, 0, nullptr, false, nullptr, t1, nullptr, warp, lane);
    gemv_rows(P + g.Wr + (long long)D * D, D, D, yv, 0, nullptr, false, nullptr, t2, nullptr, warp, lane);
    __syncthreads();
    gemv_rows(P + g.Ur, D, D, x, 0, nullptr, false, t1, t1, nullptr, warp, lane);
    gemv_rows(P + g.Ur + (long long)D * D, D, D, x, 0, nullptr, false, t2, t2, nullptr, warp, lane);
    __syncthreads();
    for (int j = tid; j < D; j += RF_THREADS) {
        const float r = 1.f / (1.f + expf(-t1[j]));
        const float z = 1.f / (1.f + expf(-(t2[j] - P[g.bg + j])));
        t1[j] = r * x[j];        // r (.) x
        t2[j] = z;
    }
    __syncthreads();
    gemv_rows(P + g.Wr + 2LL * D * D, D, D, yv, 0, nullptr, false, nullptr, t3, nullptr, warp, lane);      // Wg y
    __syncthreads();
    gemv_rows(P + g.Ug, D, D, t1, 0, nullptr, false, t3, t3, nullptr, warp, lane);                         // + Ug (r.x)
    __syncthreads();
    for (int j = tid; j < D; j += RF_THREADS) {
        const float h = tanhf(t3[j]), z = t2[j];
        const float o = (1.f - z) * x[j] + z * h;
        if (out) out[j] = o;
        if (out_global) out_global[j] = o;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RF_THREADS, 1) rollout_fused_kernel(const RfArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int D = a.D, H = a.H, L = a.L, dh = D / H, hid = a.hid;
    const int Lp = (L + 3) & ~3;
    // shared-memory carve (floats)
    float* s_feat = sm;                               // [featp]
    float* s_h = s_feat + ((a.feat + 3) & ~3);        // [D] current block input / output
    float* s_a = s_h + D;                             // [D] scratch vectors
    float* s_b = s_a + D;
    float* s_c = s_b + D;
    float* s_d = s_c + D;
    float* s_e = s_d + D;
    float* s_qk = s_e + D;                            // [H][D]
    float* s_ctx = s_qk + H * D;                      // [H][D]
    float* s_p = s_ctx + H * D;                       // [H][Lp] energies -> weights
    float* s_mu = s_p + H * Lp;                       // [Lp]
    float* s_rs = s_mu + Lp;                          // [Lp]
    float* s_hd = s_rs + Lp;                          // [2*hid] head hiddens
    float* s_red = s_hd + 2 * hid;                    // [RF_WARPS + H*RF_WARPS]
    int* s_win = reinterpret_cast<int*>(s_red + RF_WARPS * (1 + H) + 4);   // [L]
    int* s_pe = s_win + L;                            // [L]
    int* s_vis = s_pe + L;                            // [L]
    __shared__ int s_any;

    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* P = a.P;
    const long long row = a.sample_index ? a.sample_index[n] : n;
    const long long ep = a.ep_index ? a.ep_index[row] : row;
    const bool pre = a.ln == 1, post = a.ln == 2;

    // ---- stage the sample's inputs ----
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

    // ---- lin_hidden + embedding ----
    gemv_rows(P + a.Wh, a.feat, D, s_feat, 0, P + a.bh, true, nullptr, s_a, nullptr, warp, lane);
    __syncthreads();
    gemv_rows(P + a.We, D, D, s_a, 0, P + a.be, true, nullptr, s_h, nullptr, warp, lane);
    __syncthreads();

    const float scale = sqrtf((float)D);
    for (int blk = 0; blk < a.B; ++blk) {
        const long long bo = (long long)blk * a.blk_stride;
        const float* tab = a.table + ((ep * a.slots) * a.B + blk) * (long long)D;
        for (int j = tid; j < D; j += RF_THREADS) a.out_mem[((long long)n * a.B + blk) * D + j] = s_h[j];
        // q_in
        const float* q_in = s_h;
        if (pre) {
            layer_norm(s_h, nullptr, P + a.b0.n1w + bo, P + a.b0.n1b + bo, s_a, nullptr, D, s_red, tid, warp, lane);
            q_in = s_a;
        }
        // Q = q_in Wq^T  -> s_b
        gemv_rows(P + a.b0.Wq + bo, D, D, q_in, 0, nullptr, false, nullptr, s_b, nullptr, warp, lane);
        __syncthreads();
        // qk[h, j] = sum_{d in head h} Q[d] Wk[d, j] (* gamma_kv[j]) ; qkb[h] = sum_d Q[d] (Wk[d,:] . beta_kv)
        const float* Wk = P + a.b0.Wk + bo;
        for (int i = tid; i < H * D; i += RF_THREADS) {
            const int h = i / D, j = i % D;
            float acc = 0.f;
            for (int d = 0; d < dh; ++d) acc = fmaf(s_b[h * dh + d], Wk[(long long)(h * dh + d) * D + j], acc);
            if (pre) acc *= P[a.b0.nkw + bo + j];
            s_qk[i] = acc;
        }
        float* s_kb = s_c;          // [D] kb = Wk beta_kv (pre) ; s_d[h] = qkb, s_d[H + h] = sum_j qkg[h, j]
        if (pre) {
            gemv_rows(Wk, D, D, P + a.b0.nkb + bo, 0, nullptr, false, nullptr, s_kb, nullptr, warp, lane);
        }
        __syncthreads();
        if (pre && warp < H) {
            float qb = 0.f, sg = 0.f;
            for (int d = lane; d < dh; d += 32) qb = fmaf(s_b[warp * dh + d], s_kb[warp * dh + d], qb);
            for (int j = lane; j < D; j += 32) sg += s_qk[warp * D + j];
            qb = wsum(qb);
            sg = wsum(sg);
            if (lane == 0) { s_d[warp] = qb; s_d[H + warp] = sg; }
        }
        __syncthreads();
        // ---- window attention pass 1: energies ----
        for (int l = warp; l < L; l += RF_WARPS) {
            if (!s_vis[l]) continue;
            const float* src = tab + s_win[l];
            const float* per = pe ? pe + s_pe[l] : nullptr;
            float4 x[4];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = c * 128 + lane * 4;
                x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col < D) {
                    x[c] = *reinterpret_cast<const float4*>(src + col);
                    if (per) {
                        const float4 p4 = *reinterpret_cast<const float4*>(per + col);
                        x[c].x += p4.x; x[c].y += p4.y; x[c].z += p4.z; x[c].w += p4.w;
                    }
                    s1 += x[c].x + x[c].y + x[c].z + x[c].w;
                    s2 += dot4f(x[c], x[c]);
                }
            }
            float mu = 0.f, rstd = 1.f;
            if (pre) {
                s1 = wsum(s1);
                s2 = wsum(s2);
                mu = s1 / (float)D;
                rstd = rsqrtf(fmaxf(s2 / (float)D - mu * mu, 0.f) + LN_EPS);
                if (lane == 0) { s_mu[l] = mu; s_rs[l] = rstd; }
            }
            for (int h = 0; h < H; ++h) {
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = c * 128 + lane * 4;
                    if (col < D) d += dot4f(*reinterpret_cast<const float4*>(s_qk + h * D + col), x[c]);
                }
                d = wsum(d);
                if (lane == 0) {
                    float e = pre ? fmaf(rstd, d - mu * s_d[H + h], s_d[h]) : d;
                    s_p[h * Lp + l] = all_masked ? 0.f : __fdiv_rn(e, scale);
                }
            }
        }
        __syncthreads();
        // ---- softmax (warp h -> head h; H <= RF_WARPS) ----
        for (int h = warp; h < H; h += RF_WARPS) {
            float* e = s_p + h * Lp;
            float m = -INFINITY;
            for (int l = lane; l < L; l += 32) if (s_vis[l]) m = fmaxf(m, e[l]);
            m = wmax(m);
            float s = 0.f;
            for (int l = lane; l < L; l += 32) {
                const float p = s_vis[l] ? __expf(e[l] - m) : 0.f;
                e[l] = p;
                s += p;
            }
            s = wsum(s);
            const float inv = 1.f / s;
            for (int l = lane; l < L; l += 32) e[l] *= inv;
        }
        for (int i = tid; i < H * D; i += RF_THREADS) s_ctx[i] = 0.f;
        __syncthreads();
        // ---- pass 2: ctx[h, :] = sum_l p[h, l] x_l ----
        {
            float4 acc[4];
            for (int h = 0; h < H; ++h) {
                float csum = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int l = warp; l < L; l += RF_WARPS) {
                    if (!s_vis[l]) continue;
                    const float* src = tab + s_win[l];
                    const float* per = pe ? pe + s_pe[l] : nullptr;
                    const float wgt = s_p[h * Lp + l] * (pre ? s_rs[l] : 1.f);
                    if (pre) csum = fmaf(wgt, s_mu[l], csum);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = c * 128 + lane * 4;
                        if (col < D) {
                            float4 xv = *reinterpret_cast<const float4*>(src + col);
                            if (per) {
                                const float4 p4 = *reinterpret_cast<const float4*>(per + col);
                                xv.x += p4.x; xv.y += p4.y; xv.z += p4.z; xv.w += p4.w;
                            }
                            acc[c].x = fmaf(wgt, xv.x, acc[c].x); acc[c].y = fmaf(wgt, xv.y, acc[c].y);
                            acc[c].z = fmaf(wgt, xv.z, acc[c].z); acc[c].w = fmaf(wgt, xv.w, acc[c].w);
                        }
                    }
                }
                if (lane == 0) s_red[RF_WARPS + h * RF_WARPS + warp] = csum;
                // merge warps in warp order (deterministic)
                for (int ww = 0; ww < RF_WARPS; ++ww) {
                    if (warp == ww) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int col = c * 128 + lane * 4;
                            if (col < D) {
                                float4* p4 = reinterpret_cast<float4*>(s_ctx + h * D + col);
                                float4 t = *p4;
                                t.x += acc[c].x; t.y += acc[c].y; t.z += acc[c].z; t.w += acc[c].w;
                                *p4 = t;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            if (pre) {
                // ctx = gamma (.) (ctx_hat - csum) + beta  (sum_l p = 1)
                for (int i = tid; i < H * D; i += RF_THREADS) {
                    const int h = i / D, j = i % D;
                    float cs = 0.f;
                    for (int ww = 0; ww < RF_WARPS; ++ww) cs += s_red[RF_WARPS + h * RF_WARPS + ww];
                    s_ctx[i] = (s_ctx[i] - cs) * P[a.b0.nkw + bo + j] + P[a.b0.nkb + bo + j];
                }
                __syncthreads();
            }
        }
        // att_o[d] = Wv[d, :] . ctx[h(d), :]  -> s_c
        gemv_rows(P + a.b0.Wv + bo, D, D, s_ctx, dh, nullptr, false, nullptr, s_c, nullptr, warp, lane);
        __syncthreads();
        // fc_out (+ residual when not gated) -> s_b
        gemv_rows(P + a.b0.Wo + bo, D, D, s_c, 0, P + a.b0.bo + bo, false, a.gtrxl ? nullptr : s_h, s_b, nullptr, warp, lane);
        __syncthreads();
        float* h1 = s_b;                               // h1pre
        if (a.gtrxl) {
            RfGate g1 = a.b0.g1; g1.Wr += bo; g1.Ur += bo; g1.Ug += bo; g1.bg += bo;
            gru_gate(P, g1, s_h, s_b, s_c, nullptr, D, s_a, s_d, s_e, tid, warp, lane);       // x = h_in, y = att
            h1 = s_c;
        }
        if (post) {
            layer_norm(h1, nullptr, P + a.b0.n1w + bo, P + a.b0.n1b + bo, s_a, nullptr, D, s_red, tid, warp, lane);
            h1 = s_a;
        }
        // now h1 lives in s_a (post), s_c (gated, no post) or s_b (plain); pick free scratch for the rest
        float* t_in = (h1 == s_a) ? s_b : s_a;          // LN2 output (pre) else unused
        const float* h_ = h1;
        if (pre) {
            layer_norm(h1, nullptr, P + a.b0.n2w + bo, P + a.b0.n2b + bo, t_in, nullptr, D, s_red, tid, warp, lane);
            h_ = t_in;
        }
        float* f = (h1 == s_d || t_in == s_d) ? s_e : s_d;
        gemv_rows(P + a.b0.Wff + bo, D, D, h_, 0, P + a.b0.bff + bo, true, nullptr, f, nullptr, warp, lane);
        __syncthreads();
        if (a.gtrxl) {
            RfGate g2 = a.b0.g2; g2.Wr += bo; g2.Ur += bo; g2.Ug += bo; g2.bg += bo;
            // three scratch vectors distinct from h1, f
            float* pool[6] = {s_a, s_b, s_c, s_d, s_e, s_qk};
            float* t[3]; int nt = 0;
            for (int i = 0; i < 6 && nt < 3; ++i) if (pool[i] != h1 && pool[i] != f) t[nt++] = pool[i];
            if (post) {
                gru_gate(P, g2, h1, f, s_ctx, nullptr, D, t[0], t[1], t[2], tid, warp, lane);          // out_pre -> s_ctx[0:D]
                layer_norm(s_ctx, nullptr, P + a.b0.n2w + bo, P + a.b0.n2b + bo, s_h, nullptr, D, s_red, tid, warp, lane);
            } else {
                gru_gate(P, g2, h1, f, s_ctx, nullptr, D, t[0], t[1], t[2], tid, warp, lane);
                for (int j = tid; j < D; j += RF_THREADS) s_h[j] = s_ctx[j];
                __syncthreads();
            }
        } else if (post) {
            layer_norm(f, h1, P + a.b0.n2w + bo, P + a.b0.n2b + bo, s_h, nullptr, D, s_red, tid, warp, lane);
        } else {
            for (int j = tid; j < D; j += RF_THREADS) s_h[j] = f[j] + h1[j];
            __syncthreads();
        }
    }
    // ---- heads ----
    gemv_rows(P + a.Wp, D, hid, s_h, 0, P + a.bp, true, nullptr, s_hd, nullptr, warp, lane);
    gemv_rows(P + a.Wlv, D, hid, s_h, 0, P + a.blv, true, nullptr, s_hd + hid, nullptr, warp, lane);
    __syncthreads();
    gemv_rows(P + a.Wbr, hid, a.sumA, s_hd, 0, P + a.bbr, false, nullptr, nullptr, a.logits + (long long)n * a.sumA, warp, lane);
    gemv_rows(P + a.wval, hid, 1, s_hd + hid, 0, P + a.bval, false, nullptr, nullptr, a.value + n, warp, lane);
}

}  // namespace

size_t rollout_fused_smem_bytes(const RfArgs& a) {
    const int Lp = (a.L + 3) & ~3;
    const size_t floats = ((a.feat + 3) & ~3) + 6 * (size_t)a.D + 2 * (size_t)a.H * a.D + (size_t)a.H * Lp + 2 * Lp + 2 * (size_t)a.hid +
                          RF_WARPS * (1 + a.H) + 4;
    return floats * 4 + (size_t)3 * a.L * 4 + 64;
}

bool rollout_fused_supported(const RfArgs& a) {
    return a.D % 4 == 0 && a.D <= 512 && a.H <= RF_WARPS && a.D >= 6 * 0 + 2 * a.H && rollout_fused_smem_bytes(a) <= 200 * 1024;
}

int rollout_fused_forward(const RfArgs& a, cudaStream_t st) {
    TRXL_CHECK_ARG(rollout_fused_supported(a), "rollout_fused: unsupported shape (D=%d H=%d L=%d)", a.D, a.H, a.L);
    const size_t smem = rollout_fused_smem_bytes(a);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(rollout_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { trxl_set_error("rollout_fused: cannot reserve %zu bytes of shared memory", smem); return TRXL_ERR_CUDA; }
        attr_smem = smem;
    }
    rollout_fused_kernel<<<a.N, RF_THREADS, smem, st>>>(a);
    TRXL_CHECK_LAUNCH("rollout_fused");
    return TRXL_OK;
}
