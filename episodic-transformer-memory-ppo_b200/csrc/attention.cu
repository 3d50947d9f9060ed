// Fused episodic-memory window attention (query length 1), forward and backward.
//
// Reference path (transformer.py:31-86,129-137,237-249 + utils.py:52-75 + buffer.py:90):
//   gather whole episodes (mb,M,B,D) -> gather window (mb,L,B,D) -> +PE -> [LayerNorm_kv] -> K = x Wk^T,
//   V = x Wv^T over mb*L rows -> energy = q.K -> masked_fill(-1e20) -> softmax(/sqrt(D)) -> P.V
//
// This kernel never materialises any of those tensors.  Because the query length is always 1 the
// projections fold onto the query side exactly (up to fp32 re-association):
//   energy[h,l] = sum_{d in head h} Q[d] (Wk[d,:] . x_l)          = qk[h,:] . x_l ,  qk[h,:] = Q_h Wk_h
//   out[d]      = sum_l p[h(d),l] (Wv[d,:] . x_l) = Wv[d,:] . ctx[h(d),:],  ctx[h,:] = sum_l p[h,l] x_l
// so the only per-(sample, window-slot) work left is streaming the raw fp32 memory row: H dot
// products + H axpys per row.  qk (N,H,D) comes from two small GEMMs before this kernel and ctx
// (N,H,D) feeds one small GEMM after it.  The pre-LayerNorm case streams the same raw rows and
// normalises on the fly (row mean/rstd from the same pass; gamma/beta are folded into Wk/Wv by the
// caller): energy = rstd*(qkg.x - mu*sum(qkg)) + qkb,  ctx_hat = sum_l p*rstd*(x - mu).
//
// Rows come straight from the episode table (E, slots, B, D) through three index arrays
// (episode id per sample, slot per window position, PE row per window position), so the minibatch
// "memories" and "window" tensors of the reference exist only as addresses.
//
// Masking: a masked slot of a partially-masked sample has softmax weight exactly 0 in the reference
// (exp(-6e18 - max) == 0), so its row is never read.  A fully-masked sample (episode step 0) is the
// exception: the reference yields a uniform distribution over all L slots, and so does this kernel.
//
// Structure (r1 ncu: the single-pass online-softmax version was instruction-issue bound -- ~355
// instructions per row, DRAM 3 %, L2 9 % of peak -- so instructions are traded for L2 reads):
//   pass 1  each warp streams its rows, H dot products per row, one packed multi-value butterfly
//           (the H partial sums share shuffles), energies -> shared memory
//   softmax one warp per head over the L energies in shared memory (each exp evaluated once)
//   pass 2  rows are streamed again (L2 hits) and accumulated with the final weights: no running-max
//           bookkeeping, no per-lane repetition of the softmax scalars
// Algorithmic bytes per (sample, block): 4*L*D (the fp32 window) -> HBM/L2-bandwidth bound.
#include "attention.cuh"

namespace {

constexpr float LN_EPS = 1e-5f;
constexpr unsigned FULL = 0xffffffffu;

template <int NV4, bool FULLD>
__device__ __forceinline__ void load_row(const float* __restrict__ src, const float* __restrict__ pe_row, int D,
                                         int lane, float4 (&x)[NV4]) {
#pragma unroll
    for (int c = 0; c < NV4; ++c) {
        const int col = c * 128 + lane * 4;
        if (FULLD || col < D) {
            x[c] = *reinterpret_cast<const float4*>(src + col);
            if (pe_row) {
                const float4 p = *reinterpret_cast<const float4*>(pe_row + col);
                x[c].x += p.x; x[c].y += p.y; x[c].z += p.z; x[c].w += p.w;
            }
        } else {
            x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Packed warp reduction of NV (power of two <= 8) independent sums.  Each halving stage exchanges
// half of the remaining values, so NV values cost NV-1 + (5 - log2 NV) shuffles instead of 5*NV.
// On return v[0] holds the warp-wide total of value `packed_index<NV>(lane)`; every lane with the same
// index holds the same bits.
template <int NV>
__device__ __forceinline__ void packed_reduce(float (&v)[NV], int lane) {
    int off = 16;
#pragma unroll
    for (int half = NV / 2; half >= 1; half /= 2) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = hi ? v[i + half] : v[i];
            const float send = hi ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(FULL, send, off);
        }
        off >>= 1;
    }
    for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(FULL, v[0], off);
}
template <int NV>
__device__ __forceinline__ int packed_index(int lane) {
    int idx = 0, off = 16;
#pragma unroll
    for (int half = NV / 2; half >= 1; half /= 2) {
        if (lane & off) idx += half;
        off >>= 1;
    }
    return idx;
}
// first lane holding value `idx` after packed_reduce<NV>
template <int NV>
__device__ __forceinline__ int packed_holder(int idx) {
    int lane = 0, off = 16;
#pragma unroll
    for (int half = NV / 2; half >= 1; half /= 2) {
        if (idx & half) lane |= off;
        off >>= 1;
    }
    return lane;
}
__host__ __device__ constexpr int pow2_at_least(int n) { return n <= 1 ? 1 : (n <= 2 ? 2 : (n <= 4 ? 4 : 8)); }

// stage the sample's mask / window / PE index rows in shared memory; returns all_masked
// (row offsets are float offsets inside the sample's episode / inside the PE table: both fit 32 bits)
__device__ __forceinline__ bool stage_meta(const AttnArgs& a, long long row, int* s_win, int* s_pe,
                                           unsigned char* s_vis, int* s_flag) {
    const int tid = threadIdx.x;
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    int any = 0;
    for (int l = tid; l < a.L; l += blockDim.x) {
        const unsigned char m = a.mask ? a.mask[row * a.L + l] : 1;
        s_vis[l] = m;
        any |= m;
        s_win[l] = (int)((a.win_index ? a.win_index[row * a.L + l] : (long long)l) * a.B * a.D);
        s_pe[l] = (int)((a.pe_index ? a.pe_index[row * a.L + l] : 0) * a.D);
    }
    if (any) *s_flag = 1;
    __syncthreads();
    const bool all_masked = (*s_flag == 0);
    if (all_masked) {
        for (int l = tid; l < a.L; l += blockDim.x) s_vis[l] = 1;     // uniform softmax over every slot
        __syncthreads();
    }
    return all_masked;
}

// shared-memory carve shared by forward and backward:
//   win[L] i32x2-padded | pe[L] i32x2-padded | p[HPW][Lp] f32 | acc[HPW][D] f32 | rowstat[2][Lp] f32 | misc[NW*HPW*2 + 4*HPW] f32 | vis[L] u8
__host__ __device__ inline size_t attn_smem_bytes(int L, int D, int HPW, int NW) {
    const int Lp = (L + 3) & ~3;
    return (size_t)L * 16 + (size_t)HPW * Lp * 4 + (size_t)HPW * D * 4 + (size_t)2 * Lp * 4 + (size_t)(NW * HPW * 2 + 4 * HPW) * 4 + L + 16;
}

template <int NV4, int HPW, bool LN, int NW, bool FULLD>
__global__ void __launch_bounds__(NW * 32)
window_attn_fwd_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lp = (a.L + 3) & ~3;
    int* s_win = reinterpret_cast<int*>(smem_raw);
    int* s_pe = s_win + 2 * a.L;
    float* s_p = reinterpret_cast<float*>(s_pe + 2 * a.L);       // energies, then softmax weights [HPW][Lp]
    float* s_ctx = s_p + HPW * Lp;                               // [HPW][D]
    float* s_mu = s_ctx + HPW * a.D;                             // [Lp]   (LN)
    float* s_rstd = s_mu + Lp;                                   // [Lp]   (LN)
    float* s_misc = s_rstd + Lp;                                 // [NW][HPW] c-sums | [HPW] inv_sum
    unsigned char* s_vis = reinterpret_cast<unsigned char*>(s_misc + NW * HPW * 2 + 4 * HPW);
    __shared__ int s_flag;

    const int n = blockIdx.x;
    const int h0 = blockIdx.y * HPW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long row = a.sample_index ? a.sample_index[n] : n;
    const long long ep = a.ep_index ? a.ep_index[row] : row;
    const bool all_masked = stage_meta(a, row, s_win, s_pe, s_vis, &s_flag);
    for (int i = threadIdx.x; i < HPW * a.D; i += blockDim.x) s_ctx[i] = 0.f;

    // folded query vectors for this head group
    float4 qk[HPW][NV4];
    float sg[HPW], qb[HPW];
#pragma unroll
    for (int h = 0; h < HPW; ++h) {
        const bool hv = (h0 + h) < a.H;
        const float* qp = a.qk + ((long long)n * a.H + (hv ? h0 + h : 0)) * a.D;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < NV4; ++c) {
            const int col = c * 128 + lane * 4;
            qk[h][c] = (hv && col < a.D) ? *reinterpret_cast<const float4*>(qp + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += qk[h][c].x + qk[h][c].y + qk[h][c].z + qk[h][c].w;
        }
        sg[h] = 0.f; qb[h] = 0.f;
        if (LN) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
            sg[h] = s;
            qb[h] = (a.qkb && hv) ? a.qkb[(long long)n * a.H + h0 + h] : 0.f;
        }
    }

    const float* tab = a.table + ((ep * a.slots) * a.B + a.blk) * (long long)a.D;     // row l of the window: tab + s_win[l]
        const float invD = 1.f / (float)a.D;
    constexpr int NRED = pow2_at_least(HPW + (LN ? 2 : 0));
    const int my_idx = packed_index<NRED>(lane);
    const bool writer = (lane == packed_holder<NRED>(my_idx)) && (my_idx < HPW);

    // ---------------- pass 1: energies ----------------
    for (int l = w; l < a.L; l += 2 * NW) {
        const int l1 = l + NW;
        const bool v0 = s_vis[l] != 0;
        const bool v1 = (l1 < a.L) && (s_vis[l1] != 0);
        float4 x0[NV4], x1[NV4];
        if (v0) load_row<NV4, FULLD>(tab + s_win[l], a.pe ? a.pe + s_pe[l] : nullptr, a.D, lane, x0);
        if (v1) load_row<NV4, FULLD>(tab + s_win[l1], a.pe ? a.pe + s_pe[l1] : nullptr, a.D, lane, x1);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const bool vis = r ? v1 : v0;
            const int ll = r ? l1 : l;
            if (!vis) continue;                       // warp-uniform
            float4 (&x)[NV4] = r ? x1 : x0;
            float red[NRED];
#pragma unroll
            for (int i = 0; i < NRED; ++i) red[i] = 0.f;
#pragma unroll
            for (int h = 0; h < HPW; ++h) {
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < NV4; ++c) d += dot4(qk[h][c], x[c]);
                red[h] = d;
            }
            if (LN) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int c = 0; c < NV4; ++c) {
                    s1 += x[c].x + x[c].y + x[c].z + x[c].w;
                    s2 += dot4(x[c], x[c]);
                }
                red[HPW] = s1; red[HPW + 1] = s2;
            }
            packed_reduce<NRED>(red, lane);
            float e = red[0];
            if (LN) {
                const float t1 = __shfl_sync(FULL, red[0], packed_holder<NRED>(HPW));
                const float t2 = __shfl_sync(FULL, red[0], packed_holder<NRED>(HPW + 1));
                const float mu = t1 * invD;
                const float rstd = rsqrtf(fmaxf(t2 * invD - mu * mu, 0.f) + LN_EPS);
                if (lane == 0) { s_mu[ll] = mu; s_rstd[ll] = rstd; }
                // sg/qb of this lane's head: select without dynamic register indexing
                float sgh = 0.f, qbh = 0.f;
#pragma unroll
                for (int h = 0; h < HPW; ++h) if (my_idx == h) { sgh = sg[h]; qbh = qb[h]; }
                e = fmaf(rstd, e - mu * sgh, qbh);
            }
            if (writer) s_p[my_idx * Lp + ll] = all_masked ? 0.f : __fdiv_rn(e, a.scale);
        }
    }
    __syncthreads();

    // ---------------- softmax: warp h handles head h (NW >= HPW) ----------------
    if (w < HPW) {
        float* e = s_p + w * Lp;
        float m = -INFINITY;
        for (int l = lane; l < a.L; l += 32) if (s_vis[l]) m = fmaxf(m, e[l]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        float s = 0.f;
        for (int l = lane; l < a.L; l += 32) {
            const float p = s_vis[l] ? __expf(e[l] - m) : 0.f;
            e[l] = p;
            s += p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        const float inv = 1.f / s;
        const bool hv = (h0 + w) < a.H;
        float* pp = a.probs + ((long long)n * a.H + (hv ? h0 + w : 0)) * a.L;
        for (int l = lane; l < a.L; l += 32) {
            const float p = e[l] * inv;
            e[l] = p;
            if (hv) pp[l] = p;
        }
    }
    __syncthreads();

    // ---------------- pass 2: ctx[h,:] = sum_l p[h,l] * x_l (rows come back from L2) ----------------
    float4 acc[HPW][NV4];
    float csum[HPW];
#pragma unroll
    for (int h = 0; h < HPW; ++h) {
        csum[h] = 0.f;
#pragma unroll
        for (int c = 0; c < NV4; ++c) acc[h][c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int l = w; l < a.L; l += 2 * NW) {
        const int l1 = l + NW;
        const bool v0 = s_vis[l] != 0;
        const bool v1 = (l1 < a.L) && (s_vis[l1] != 0);
        float4 x0[NV4], x1[NV4];
        if (v0) load_row<NV4, FULLD>(tab + s_win[l], a.pe ? a.pe + s_pe[l] : nullptr, a.D, lane, x0);
        if (v1) load_row<NV4, FULLD>(tab + s_win[l1], a.pe ? a.pe + s_pe[l1] : nullptr, a.D, lane, x1);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const bool vis = r ? v1 : v0;
            const int ll = r ? l1 : l;
            if (!vis) continue;
            float4 (&x)[NV4] = r ? x1 : x0;
            const float rstd = LN ? s_rstd[ll] : 1.f;
            const float mu = LN ? s_mu[ll] : 0.f;
#pragma unroll
            for (int h = 0; h < HPW; ++h) {
                const float wgt = s_p[h * Lp + ll] * rstd;
                if (LN) csum[h] = fmaf(wgt, mu, csum[h]);
#pragma unroll
                for (int c = 0; c < NV4; ++c) {
                    acc[h][c].x = fmaf(wgt, x[c].x, acc[h][c].x);
                    acc[h][c].y = fmaf(wgt, x[c].y, acc[h][c].y);
                    acc[h][c].z = fmaf(wgt, x[c].z, acc[h][c].z);
                    acc[h][c].w = fmaf(wgt, x[c].w, acc[h][c].w);
                }
            }
        }
    }
    if (LN && lane == 0) {
#pragma unroll
        for (int h = 0; h < HPW; ++h) s_misc[w * HPW + h] = csum[h];
    }
    // deterministic merge: warps add their accumulators in warp order
    for (int ww = 0; ww < NW; ++ww) {
        if (w == ww) {
#pragma unroll
            for (int h = 0; h < HPW; ++h)
#pragma unroll
                for (int c = 0; c < NV4; ++c) {
                    const int col = c * 128 + lane * 4;
                    if (FULLD || col < a.D) {
                        float4* p = reinterpret_cast<float4*>(s_ctx + h * a.D + col);
                        float4 t = *p;
                        t.x += acc[h][c].x; t.y += acc[h][c].y; t.z += acc[h][c].z; t.w += acc[h][c].w;
                        *p = t;
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int h = 0; h < HPW; ++h) {
        if (h0 + h >= a.H) continue;
        float c_all = 0.f;
        if (LN) {
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) c_all += s_misc[ww * HPW + h];
        }
        float* cp = a.ctx + ((long long)n * a.H + h0 + h) * a.D;
        for (int j = threadIdx.x; j < a.D; j += blockDim.x) cp[j] = s_ctx[h * a.D + j] - c_all;
    }
}

// ---------------------------------------------------------------------------------------------
// backward: given d(ctx) (N,H,D) and the saved probs/ctx, produce d(qk) (N,H,D), d(qkb) (N,H) and,
// for a learned positional table, scatter-add d(PE).  The memory rows themselves carry no gradient
// (reference transformer.py:248: memories are detached inputs).  Single pass: the weights are known.
// ---------------------------------------------------------------------------------------------
template <int NV4, int HPW, bool LN, int NW, bool FULLD>
__global__ void __launch_bounds__(NW * 32)
window_attn_bwd_kernel(const AttnArgs a, const AttnBwdArgs g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lp = (a.L + 3) & ~3;
    int* s_win = reinterpret_cast<int*>(smem_raw);
    int* s_pe = s_win + 2 * a.L;
    float* s_p = reinterpret_cast<float*>(s_pe + 2 * a.L);      // probs [HPW][Lp]
    float* s_acc = s_p + HPW * Lp;                              // [HPW][D]
    float* s_unused = s_acc + HPW * a.D;                        // rowstat region (unused here)
    float* s_stat = s_unused + 2 * Lp;                          // [NW][HPW][2]
    unsigned char* s_vis = reinterpret_cast<unsigned char*>(s_stat + NW * HPW * 2 + 4 * HPW);
    __shared__ int s_flag;

    const int n = blockIdx.x;
    const int h0 = blockIdx.y * HPW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long row = a.sample_index ? a.sample_index[n] : n;
    const long long ep = a.ep_index ? a.ep_index[row] : row;
    const bool all_masked = stage_meta(a, row, s_win, s_pe, s_vis, &s_flag);
    const bool need_dx = g.dpe != nullptr;

    for (int i = threadIdx.x; i < HPW * a.L; i += blockDim.x) {
        const int h = i / a.L, l = i % a.L;
        s_p[h * Lp + l] = (h0 + h < a.H) ? a.probs[((long long)n * a.H + h0 + h) * a.L + l] : 0.f;
    }
    for (int i = threadIdx.x; i < HPW * a.D; i += blockDim.x) s_acc[i] = 0.f;

    float4 dc[HPW][NV4], qk[HPW][NV4];
    float dot0[HPW], sdc[HPW];
#pragma unroll
    for (int h = 0; h < HPW; ++h) {
        const bool hv = (h0 + h) < a.H;
        const long long off = ((long long)n * a.H + (hv ? h0 + h : 0)) * a.D;
        float d = 0.f, s = 0.f;
#pragma unroll
        for (int c = 0; c < NV4; ++c) {
            const int col = c * 128 + lane * 4;
            const bool ok = hv && col < a.D;
            dc[h][c] = ok ? *reinterpret_cast<const float4*>(g.dctx + off + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            qk[h][c] = (ok && need_dx) ? *reinterpret_cast<const float4*>(a.qk + off + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                const float4 cx = *reinterpret_cast<const float4*>(a.ctx + off + col);
                d += dot4(dc[h][c], cx);
            }
            s += dc[h][c].x + dc[h][c].y + dc[h][c].z + dc[h][c].w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d += __shfl_xor_sync(FULL, d, o);
            s += __shfl_xor_sync(FULL, s, o);
        }
        dot0[h] = d;
        sdc[h] = LN ? s : 0.f;
    }
    __syncthreads();

    float4 acc[HPW][NV4];
    float accmu[HPW], accb[HPW];
#pragma unroll
    for (int h = 0; h < HPW; ++h) {
        accmu[h] = 0.f; accb[h] = 0.f;
#pragma unroll
        for (int c = 0; c < NV4; ++c) acc[h][c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* tab = a.table + ((ep * a.slots) * a.B + a.blk) * (long long)a.D;     // row l of the window: tab + s_win[l]
        const float invD = 1.f / (float)a.D;
    const float inv_scale = 1.f / a.scale;
    const bool skip_all = all_masked && !need_dx;      // constant energies: no gradient reaches qk
    constexpr int NRED = pow2_at_least(HPW + (LN ? 2 : 0));

    if (!skip_all) {
        for (int l = w; l < a.L; l += 2 * NW) {
            const int l1 = l + NW;
            const bool v0 = s_vis[l] != 0;
            const bool v1 = (l1 < a.L) && (s_vis[l1] != 0);
            float4 x0[NV4], x1[NV4];
            if (v0) load_row<NV4, FULLD>(tab + s_win[l], a.pe ? a.pe + s_pe[l] : nullptr, a.D, lane, x0);
            if (v1) load_row<NV4, FULLD>(tab + s_win[l1], a.pe ? a.pe + s_pe[l1] : nullptr, a.D, lane, x1);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const bool vis = r ? v1 : v0;
                const int ll = r ? l1 : l;
                if (!vis) continue;
                float4 (&x)[NV4] = r ? x1 : x0;
                float red[NRED];
#pragma unroll
                for (int i = 0; i < NRED; ++i) red[i] = 0.f;
#pragma unroll
                for (int h = 0; h < HPW; ++h) {
                    float d = 0.f;
#pragma unroll
                    for (int c = 0; c < NV4; ++c) d += dot4(dc[h][c], x[c]);
                    red[h] = d;
                }
                if (LN) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int c = 0; c < NV4; ++c) {
                        s1 += x[c].x + x[c].y + x[c].z + x[c].w;
                        s2 += dot4(x[c], x[c]);
                    }
                    red[HPW] = s1; red[HPW + 1] = s2;
                }
                packed_reduce<NRED>(red, lane);
                float mu = 0.f, rstd = 1.f;
                if (LN) {
                    const float t1 = __shfl_sync(FULL, red[0], packed_holder<NRED>(HPW));
                    const float t2 = __shfl_sync(FULL, red[0], packed_holder<NRED>(HPW + 1));
                    mu = t1 * invD;
                    rstd = rsqrtf(fmaxf(t2 * invD - mu * mu, 0.f) + LN_EPS);
                }
                float dE[HPW];
#pragma unroll
                for (int h = 0; h < HPW; ++h) {
                    const float dot = __shfl_sync(FULL, red[0], packed_holder<NRED>(h));      // dctx[h] . x_l (raw)
                    const float gdot = LN ? rstd * (dot - mu * sdc[h]) : dot;                 // dctx[h] . x_hat
                    dE[h] = all_masked ? 0.f : s_p[h * Lp + ll] * (gdot - dot0[h]) * inv_scale;
                    const float wgt = dE[h] * rstd;
                    accb[h] += dE[h];
                    if (LN) accmu[h] = fmaf(wgt, mu, accmu[h]);
#pragma unroll
                    for (int c = 0; c < NV4; ++c) {
                        acc[h][c].x = fmaf(wgt, x[c].x, acc[h][c].x); acc[h][c].y = fmaf(wgt, x[c].y, acc[h][c].y);
                        acc[h][c].z = fmaf(wgt, x[c].z, acc[h][c].z); acc[h][c].w = fmaf(wgt, x[c].w, acc[h][c].w);
                    }
                }
                if (need_dx) {
                    // d(x_hat) = sum_h dE[h] qk[h] + p[h,l] dctx[h]; then LayerNorm backward (no affine) if LN
                    float4 dxh[NV4];
                    float t1 = 0.f, t2 = 0.f;
#pragma unroll
                    for (int c = 0; c < NV4; ++c) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int h = 0; h < HPW; ++h) {
                            const float p = s_p[h * Lp + ll];
                            v.x += dE[h] * qk[h][c].x + p * dc[h][c].x; v.y += dE[h] * qk[h][c].y + p * dc[h][c].y;
                            v.z += dE[h] * qk[h][c].z + p * dc[h][c].z; v.w += dE[h] * qk[h][c].w + p * dc[h][c].w;
                        }
                        dxh[c] = v;
                        if (LN) {
                            const int col = c * 128 + lane * 4;
                            if (FULLD || col < a.D) {
                                const float4 xh = make_float4((x[c].x - mu) * rstd, (x[c].y - mu) * rstd,
                                                              (x[c].z - mu) * rstd, (x[c].w - mu) * rstd);
                                t1 += v.x + v.y + v.z + v.w;
                                t2 += dot4(v, xh);
                            }
                        }
                    }
                    if (LN) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            t1 += __shfl_xor_sync(FULL, t1, o);
                            t2 += __shfl_xor_sync(FULL, t2, o);
                        }
                        t1 *= invD; t2 *= invD;
                    }
                    float* dst = g.dpe + s_pe[ll];
#pragma unroll
                    for (int c = 0; c < NV4; ++c) {
                        const int col = c * 128 + lane * 4;
                        if (FULLD || col < a.D) {
                            float4 v = dxh[c];
                            if (LN) {
                                v.x = rstd * (v.x - t1 - (x[c].x - mu) * rstd * t2);
                                v.y = rstd * (v.y - t1 - (x[c].y - mu) * rstd * t2);
                                v.z = rstd * (v.z - t1 - (x[c].z - mu) * rstd * t2);
                                v.w = rstd * (v.w - t1 - (x[c].w - mu) * rstd * t2);
                            }
                            atomicAdd(dst + col, v.x); atomicAdd(dst + col + 1, v.y);
                            atomicAdd(dst + col + 2, v.z); atomicAdd(dst + col + 3, v.w);
                        }
                    }
                }
            }
        }
    }
    // ---- merge warps (deterministic order) ----
    if (lane == 0) {
#pragma unroll
        for (int h = 0; h < HPW; ++h) {
            s_stat[(w * HPW + h) * 2 + 0] = accmu[h];
            s_stat[(w * HPW + h) * 2 + 1] = accb[h];
        }
    }
    for (int ww = 0; ww < NW; ++ww) {
        if (w == ww) {
#pragma unroll
            for (int h = 0; h < HPW; ++h)
#pragma unroll
                for (int c = 0; c < NV4; ++c) {
                    const int col = c * 128 + lane * 4;
                    if (FULLD || col < a.D) {
                        float4* p = reinterpret_cast<float4*>(s_acc + h * a.D + col);
                        float4 t = *p;
                        t.x += acc[h][c].x; t.y += acc[h][c].y; t.z += acc[h][c].z; t.w += acc[h][c].w;
                        *p = t;
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int h = 0; h < HPW; ++h) {
        if (h0 + h >= a.H) continue;
        float mu_sum = 0.f, b_sum = 0.f;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) {
            mu_sum += s_stat[(ww * HPW + h) * 2];
            b_sum += s_stat[(ww * HPW + h) * 2 + 1];
        }
        float* dq = g.dqk + ((long long)n * a.H + h0 + h) * a.D;
        for (int j = threadIdx.x; j < a.D; j += blockDim.x) dq[j] = s_acc[h * a.D + j] - mu_sum;
        if (g.dqkb && threadIdx.x == 0) g.dqkb[(long long)n * a.H + h0 + h] = b_sum;
    }
}

// Warps per sample: 4 when there are enough samples to fill the GPU (training minibatches), 8 for the
// rollout's W-sample forwards, whose latency is the per-warp chain of window rows.
template <int NV4, int HPW, bool LN, int NW>
int launch_fwd_nw(const AttnArgs& a, cudaStream_t st) {
    dim3 grid(a.N, trxl_cdiv(a.H, HPW));
    const size_t smem = attn_smem_bytes(a.L, a.D, HPW, NW);
    const bool full = (a.D == NV4 * 128);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(window_attn_fwd_kernel<NV4, HPW, LN, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(window_attn_fwd_kernel<NV4, HPW, LN, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    trxl_prof_begin(0, a.N, st);
    if (full) window_attn_fwd_kernel<NV4, HPW, LN, NW, true><<<grid, NW * 32, smem, st>>>(a);
    else window_attn_fwd_kernel<NV4, HPW, LN, NW, false><<<grid, NW * 32, smem, st>>>(a);
    trxl_prof_end(0, st);
    TRXL_CHECK_LAUNCH("window_attn_fwd");
    return TRXL_OK;
}
template <int NV4, int HPW>
int launch_fwd(const AttnArgs& a, cudaStream_t st) {
    // few samples (rollout: one per env worker): the latency is each warp's chain of dependent row loads, so
    // spread a sample's window over 16 (or 8) warps; training minibatches fill the GPU with 4-warp CTAs
    const long long ctas = (long long)a.N * trxl_cdiv(a.H, HPW);
    if (NV4 <= 2 && ctas <= 148 && a.L >= 64)
        return a.ln ? launch_fwd_nw<NV4, HPW, true, 16>(a, st) : launch_fwd_nw<NV4, HPW, false, 16>(a, st);
    const bool wide = ctas < 148 * 2 && a.L >= 32;
    if (a.ln) return wide ? launch_fwd_nw<NV4, HPW, true, 8>(a, st) : launch_fwd_nw<NV4, HPW, true, 4>(a, st);
    return wide ? launch_fwd_nw<NV4, HPW, false, 8>(a, st) : launch_fwd_nw<NV4, HPW, false, 4>(a, st);
}

template <int NV4, int HPW, bool LN>
int launch_bwd_ln(const AttnArgs& a, const AttnBwdArgs& g, cudaStream_t st) {
    constexpr int NW = 4;
    dim3 grid(a.N, trxl_cdiv(a.H, HPW));
    const size_t smem = attn_smem_bytes(a.L, a.D, HPW, NW);
    const bool full = (a.D == NV4 * 128);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(window_attn_bwd_kernel<NV4, HPW, LN, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(window_attn_bwd_kernel<NV4, HPW, LN, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    trxl_prof_begin(1, a.N, st);
    if (full) window_attn_bwd_kernel<NV4, HPW, LN, NW, true><<<grid, NW * 32, smem, st>>>(a, g);
    else window_attn_bwd_kernel<NV4, HPW, LN, NW, false><<<grid, NW * 32, smem, st>>>(a, g);
    trxl_prof_end(1, st);
    TRXL_CHECK_LAUNCH("window_attn_bwd");
    return TRXL_OK;
}
template <int NV4, int HPW>
int launch_bwd(const AttnArgs& a, const AttnBwdArgs& g, cudaStream_t st) {
    return a.ln ? launch_bwd_ln<NV4, HPW, true>(a, g, st) : launch_bwd_ln<NV4, HPW, false>(a, g, st);
}

// heads per warp pass: keep the (qk + acc) register arrays <= ~64 float4-lanes
int pick_hpw(int nv4, int H) {
    int hpw = nv4 <= 2 ? 4 : 2;
    if (nv4 > 4) hpw = 1;
    while (hpw > 1 && hpw / 2 >= H) hpw /= 2;
    return hpw;
}

}  // namespace

static int check_args(const AttnArgs& a) {
    TRXL_CHECK_ARG(a.N >= 0 && a.L > 0 && a.H > 0 && a.D > 0, "window_attention: bad dims N=%d L=%d D=%d H=%d", a.N, a.L, a.D, a.H);
    TRXL_CHECK_ARG(a.D % 4 == 0 && a.D <= 1024, "window_attention: embed_dim must be a multiple of 4 and <= 1024 (got %d)", a.D);
    TRXL_CHECK_ARG(a.D % a.H == 0, "window_attention: embed_dim %d not divisible by heads %d", a.D, a.H);
    TRXL_CHECK_ARG(a.L <= 4096, "window_attention: memory_length %d > 4096 unsupported", a.L);
    TRXL_CHECK_ARG(a.table && a.qk && a.probs && a.ctx, "window_attention: null pointer");
    TRXL_CHECK_ARG(((uintptr_t)a.table % 16 == 0) && ((uintptr_t)a.qk % 16 == 0) && ((uintptr_t)a.ctx % 16 == 0) &&
                   (!a.pe || (uintptr_t)a.pe % 16 == 0), "window_attention: pointers must be 16-byte aligned");
    TRXL_CHECK_ARG(!a.pe || a.pe_index, "window_attention: pe table given without pe_index");
    return TRXL_OK;
}

#define DISPATCH(NV4, FN, ...)                                                   \
    do {                                                                        \
        const int hpw = pick_hpw(NV4, a.H);                                     \
        if (hpw == 4) return FN<NV4, 4>(__VA_ARGS__);                           \
        if (hpw == 2) return FN<NV4, 2>(__VA_ARGS__);                           \
        return FN<NV4, 1>(__VA_ARGS__);                                         \
    } while (0)

int trxl_window_attn_fwd(const AttnArgs& a, cudaStream_t st) {
    TRXL_PROPAGATE(check_args(a));
    if (a.N == 0) return TRXL_OK;
    const int nv4 = (a.D + 127) / 128;
    switch (nv4) {
        case 1: DISPATCH(1, launch_fwd, a, st);
        case 2: DISPATCH(2, launch_fwd, a, st);
        case 3: DISPATCH(3, launch_fwd, a, st);
        case 4: DISPATCH(4, launch_fwd, a, st);
        default: return launch_fwd<8, 1>(a, st);
    }
}

int trxl_window_attn_bwd(const AttnArgs& a, const AttnBwdArgs& g, cudaStream_t st) {
    TRXL_PROPAGATE(check_args(a));
    TRXL_CHECK_ARG(g.dctx && g.dqk, "window_attention_bwd: null pointer");
    if (a.N == 0) return TRXL_OK;
    const int nv4 = (a.D + 127) / 128;
    switch (nv4) {
        case 1: DISPATCH(1, launch_bwd, a, g, st);
        case 2: DISPATCH(2, launch_bwd, a, g, st);
        case 3: DISPATCH(3, launch_bwd, a, g, st);
        case 4: DISPATCH(4, launch_bwd, a, g, st);
        default: return launch_bwd<8, 1>(a, g, st);
    }
}
