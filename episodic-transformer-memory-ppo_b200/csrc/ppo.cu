// PPO data-path kernels: GAE, minibatch gathers, rollout bookkeeping, the clipped loss (forward and
// backward in one pass), and the fused global-norm clip + AdamW update over the flat parameter arena.
#include "ppo.cuh"

namespace {

// ------------------------------------------------------------------------------------ GAE
// Reference buffer.py:95-113.  One warp per worker.  The recurrence is evaluated in exactly the
// reference's order with unfused fp32 multiplies/adds (so advantages are bit-identical to the CPU
// path); lanes only share the coalesced loads/stores of 32 timesteps at a time.
__global__ void gae_kernel(const float* __restrict__ rewards, const unsigned char* __restrict__ dones,
                           const float* __restrict__ values, const float* __restrict__ last_value,
                           float* __restrict__ adv, int W, int T, float gamma, float gamma_lambda) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= W) return;
    float lv = last_value[w];
    float la = 0.f;
    const long long base = (long long)w * T;
    for (int end = T; end > 0; end -= 32) {
        const int t = end - 32 + lane;
        float r = 0.f, v = 0.f, m = 0.f;
        if (t >= 0) {
            r = rewards[base + t];
            v = values[base + t];
            m = dones[base + t] ? 0.f : 1.f;
        }
        float mine = 0.f;
        for (int j = 31; j >= 0; --j) {
            if (end - 32 + j < 0) break;                       // warp-uniform
            const float rj = __shfl_sync(0xffffffffu, r, j);
            const float vj = __shfl_sync(0xffffffffu, v, j);
            const float mj = __shfl_sync(0xffffffffu, m, j);
            lv = __fmul_rn(lv, mj);
            la = __fmul_rn(la, mj);
            const float delta = __fsub_rn(__fadd_rn(rj, __fmul_rn(gamma, lv)), vj);
            la = __fadd_rn(delta, __fmul_rn(gamma_lambda, la));
            if (lane == j) mine = la;
            lv = vj;
        }
        if (t >= 0) adv[base + t] = mine;
    }
}

// ------------------------------------------------------------------------------------ gathers / scatters
// dst[n, :] = src[idx[n], :]
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst,
                                   long long row_floats, int vec) {
    const long long n = blockIdx.y;
    const float* s = src + idx[n] * row_floats;
    float* d = dst + n * row_floats;
    if (vec) {
        const long long nv = row_floats >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_floats; i += (long long)gridDim.x * blockDim.x)
            d[i] = s[i];
    }
}
// out[n, l, :] = in[n, idx[n, l], :]       (reference utils.py:52-75 with dim=1)
__global__ void gather_window_kernel(const float* __restrict__ in, const long long* __restrict__ idx, float* __restrict__ out,
                                     int L, long long slots, long long inner) {
    const long long nl = blockIdx.x;
    const long long n = nl / L;
    const float* s = in + (n * slots + idx[nl]) * inner;
    float* d = out + nl * inner;
    for (long long i = threadIdx.x; i < inner; i += blockDim.x) d[i] = s[i];
}
// table[ep[w], step[w], :, :] = new_mem[w, :, :]      (reference trainer.py:174)
__global__ void memory_scatter_kernel(float* __restrict__ table, const long long* __restrict__ ep,
                                      const long long* __restrict__ step, const float* __restrict__ new_mem, long long slots,
                                      long long inner) {
    const long long w = blockIdx.x;
    float* d = table + (ep[w] * slots + step[w]) * inner;
    const float* s = new_mem + w * inner;
    for (long long i = threadIdx.x; i < inner; i += blockDim.x) d[i] = s[i];
}
// The stores that end a rollout step, in one kernel: the memory scatter above; optionally the same rows plus their slot's
// positional row into table_pe (a copy of the table that already carries the positional rows, which the fused rollout forward
// prefetches window rows from); optionally value_dst[w * value_stride] = value[w] (trainer.py:186)
__global__ void rollout_store_kernel(float* __restrict__ table, float* __restrict__ table_pe, const float* __restrict__ pe_table,
                                     const long long* __restrict__ ep, const long long* __restrict__ step,
                                     const float* __restrict__ new_mem, long long slots, int blocks, int D,
                                     const float* __restrict__ value, float* __restrict__ value_dst, long long value_stride) {
    const long long w = blockIdx.x, inner = (long long)blocks * D;
    const long long s_ = step[w], off = (ep[w] * slots + s_) * inner;
    const float* s = new_mem + w * inner;
    const float* pe = table_pe ? pe_table + s_ * D : nullptr;
    for (long long i = threadIdx.x; i < inner; i += blockDim.x) {
        const float v = s[i];
        table[off + i] = v;
        if (table_pe) table_pe[off + i] = v + pe[i % D];
    }
    if (value_dst && threadIdx.x == 0) value_dst[w * value_stride] = value[w];
}
// mask_out[w, :] = mask_table[min(step, L-1), :]; idx_out[w, :] = index_table[step, :]; ep_out[w] = ep[w]
// (reference trainer.py:165-166; tables are the bit-exact integer tables built on the host)
__global__ void rollout_prepare_kernel(const long long* __restrict__ step, const long long* __restrict__ ep,
                                       const unsigned char* __restrict__ mask_table, const long long* __restrict__ index_table,
                                       unsigned char* __restrict__ mask_out, long long mask_stride,
                                       long long* __restrict__ idx_out, long long idx_stride, long long* __restrict__ ep_out,
                                       long long ep_stride, int L) {
    const int w = blockIdx.x;
    const long long s = step[w];
    const long long ms = s < L - 1 ? s : L - 1;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        mask_out[w * mask_stride + l] = mask_table[ms * L + l];
        idx_out[w * idx_stride + l] = index_table[s * L + l];
    }
    if (threadIdx.x == 0 && ep_out) ep_out[w * ep_stride] = ep[w];
}

// Fetch one rollout step's inputs in ONE kernel: observations (n, obs_floats) and the episode cursors (n,) from `*_src`
// -- which may be pinned / cudaHostRegister'ed HOST memory read in place over PCIe (zero-copy; no copy-engine node, no
// engine switch inside the captured step) -- into device staging buffers, and the observations also into their strided rows
// of the rollout buffer (reference trainer.py:163).  Every thread keeps 4 independent 16-byte loads in flight.
__global__ void rollout_fetch_kernel(const float4* __restrict__ obs_src, long long obs_vec4, const long long* __restrict__ step_src,
                                     const long long* __restrict__ ep_src, float4* __restrict__ obs_dev, float4* __restrict__ obs_store,
                                     long long store_stride_vec4, long long* __restrict__ step_dev, long long* __restrict__ ep_dev, int n) {
    const long long total = (long long)n * obs_vec4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        step_dev[i] = step_src[i];
        ep_dev[i] = ep_src[i];
    }
    for (; i < total; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long j = i + k * stride;
            if (j < total) v[k] = obs_src[j];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long j = i + k * stride;
            if (j < total) {
                obs_dev[j] = v[k];
                const long long w = j / obs_vec4;
                obs_store[w * store_stride_vec4 + (j - w * obs_vec4)] = v[k];
            }
        }
    }
}

// scalar variant for observation sizes that are not a multiple of 4 floats (small vector observations)
__global__ void rollout_fetch_scalar_kernel(const float* __restrict__ obs_src, long long obs_floats, const long long* __restrict__ step_src,
                                            const long long* __restrict__ ep_src, float* __restrict__ obs_dev, float* __restrict__ obs_store,
                                            long long store_stride, long long* __restrict__ step_dev, long long* __restrict__ ep_dev, int n) {
    const long long total = (long long)n * obs_floats;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        step_dev[i] = step_src[i];
        ep_dev[i] = ep_src[i];
    }
    for (; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float v = obs_src[i];
        obs_dev[i] = v;
        const long long w = i / obs_floats;
        obs_store[w * store_stride + (i - w * obs_floats)] = v;
    }
}

// ------------------------------------------------------------------------------------ action sampling
// Inverse-CDF sampling from softmax(logits) per branch with caller-provided uniforms (torch RNG).
__device__ __forceinline__ void sample_one(const float* __restrict__ logits, int sumA, const float* __restrict__ u,
                                           const long long* __restrict__ forced, const BranchSpec& bs, long long* __restrict__ act_out,
                                           long long act_stride, float* __restrict__ logp_out, long long logp_stride,
                                           long long* __restrict__ act_compact, int W, int i, long long tag = 0) {
    if (i >= W * bs.n) return;
    const int w = i / bs.n, k = i % bs.n;
    const float* z = logits + (long long)w * sumA + bs.off[k];
    const int A = bs.size[k];
    float m = -INFINITY;
    for (int j = 0; j < A; ++j) m = fmaxf(m, z[j]);
    float s = 0.f;
    for (int j = 0; j < A; ++j) s += expf(z[j] - m);
    const float lse = m + logf(s);
    int a = A - 1;
    if (forced) {
        a = (int)forced[i];                                    // replay a recorded trajectory (parity tests)
    } else {
        const float target = u[i];
        float cdf = 0.f;
        for (int j = 0; j < A; ++j) {
            cdf += expf(z[j] - lse);
            if (target < cdf) { a = j; break; }
        }
    }
    act_out[w * act_stride + k] = a;
    logp_out[w * logp_stride + k] = z[a] - lse;
    if (act_compact) act_compact[i] = tag | (long long)a;
}

__global__ void sample_actions_kernel(const float* __restrict__ logits, int sumA, const float* __restrict__ u,
                                      const long long* __restrict__ forced, BranchSpec bs, long long* __restrict__ act_out, long long act_stride, float* __restrict__ logp_out,
                                      long long logp_stride, long long* __restrict__ act_compact, int W,
                                      long long* __restrict__ done_counter, volatile long long* __restrict__ done_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (done_counter) {
        // single-block launches only (checked by the host wrapper).  Every compact action word carries the launch sequence number
        // in its upper half: word = seq << 32 | action.  A host that polls the words themselves (8-byte stores are single-copy
        // atomic) needs no flag behind a system-wide fence -- two of those cost ~12 us of a 17 us kernel when the words live in
        // pinned host memory (measured).  done_flag, if given, is still published behind fences.
        const long long seq = *done_counter + 1;
        sample_one(logits, sumA, u, forced, bs, act_out, act_stride, logp_out, logp_stride, act_compact, W, i, seq << 32);
        if (done_flag) __threadfence_system();
        __syncthreads();                          // every thread has read the counter
        if (threadIdx.x == 0) {
            *done_counter = seq;
            if (done_flag) {
                __threadfence_system();
                *done_flag = seq;
            }
        }
        return;
    }
    sample_one(logits, sumA, u, forced, bs, act_out, act_stride, logp_out, logp_stride, act_compact, W, i);
}

// ------------------------------------------------------------------------------------ advantage statistics
// out[0] = sum a, out[1] = sum a^2, out[2] = count  (double; all-reducible across ranks)
__global__ void adv_stats_kernel(const float* __restrict__ adv, const long long* __restrict__ sidx, int N, double* __restrict__ out) {
    __shared__ double sm[32];
    double s = 0.0, q = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double a = adv[sidx ? sidx[n] : n];
        s += a;
        q += a * a;
    }
    s = block_sum_d(s, sm);
    q = block_sum_d(q, sm);
    if (threadIdx.x == 0) { out[0] = s; out[1] = q; out[2] = (double)N; }
}

// ------------------------------------------------------------------------------------ PPO loss fwd+bwd
// Reference trainer.py:277-304,315-316.  One thread per sample; gradients w.r.t. logits and value are
// written in the same pass (the loss is a mean, so each sample's gradient is local once the advantage
// statistics are known).  Tie rules of torch.min/torch.max backward (half to each side) are kept.
__global__ void ppo_loss_kernel(PpoLossArgs a) {
    __shared__ float sm[32];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // policy, vf, entropy, kl, clipfrac, (unused)
    if (n < a.N) {
        const long long row = a.sidx ? a.sidx[n] : n;
        const double cnt = a.advstats[2];
        const double mean_d = a.advstats[0] / cnt;
        const double var_d = (a.advstats[1] - cnt * mean_d * mean_d) / (cnt - 1.0);
        const float mean = (float)mean_d;
        const float stdv = (float)sqrt(var_d > 0.0 ? var_d : 0.0);
        const float adv = a.adv[row];
        const float nadv = (adv - mean) / (stdv + 1e-8f);
        const float inv_nk = 1.f / ((float)cnt * (float)a.bs.n);
        const float inv_n = 1.f / (float)cnt;
        float ent_sum = 0.f;
        for (int k = 0; k < a.bs.n; ++k) {
            const float* z = a.logits + (long long)n * a.sumA + a.bs.off[k];
            float* dz = a.dlogits ? a.dlogits + (long long)n * a.sumA + a.bs.off[k] : nullptr;
            const int A = a.bs.size[k];
            float m = -INFINITY;
            for (int j = 0; j < A; ++j) m = fmaxf(m, z[j]);
            float s = 0.f;
            for (int j = 0; j < A; ++j) s += expf(z[j] - m);
            const float lse = m + logf(s);
            float ent = 0.f;
            for (int j = 0; j < A; ++j) {
                const float lp = z[j] - lse;
                ent -= expf(lp) * lp;
            }
            const int act = (int)a.actions[row * a.bs.n + k];
            const float lp_a = z[act] - lse;
            const float log_ratio = lp_a - a.old_logp[row * a.bs.n + k];
            const float ratio = expf(log_ratio);
            const float s1 = ratio * nadv;
            const float rc = fminf(fmaxf(ratio, a.clip_lo), a.clip_hi);
            const float s2 = rc * nadv;
            acc[0] += fminf(s1, s2);
            acc[3] += (ratio - 1.f) - log_ratio;
            acc[4] += (fabsf(ratio - 1.f) > a.clip) ? 1.f : 0.f;
            ent_sum += ent;
            if (dz) {
                const bool inside = (ratio >= a.clip_lo) && (ratio <= a.clip_hi);
                float dmin = 0.f;                               // d min(s1,s2) / d ratio
                if (inside) dmin = nadv;
                else if (s1 < s2) dmin = nadv;
                else if (s1 == s2) dmin = 0.5f * nadv;
                const float g_lp = -inv_nk * dmin * ratio;      // d loss / d log_prob(action)
                const float gb = a.beta * inv_n;
                for (int j = 0; j < A; ++j) {
                    const float lp = z[j] - lse;
                    const float p = expf(lp);
                    dz[j] = g_lp * ((j == act ? 1.f : 0.f) - p) + gb * p * (lp + ent);
                }
            }
        }
        acc[2] = ent_sum;
        const float v = a.value[n];
        const float v_old = a.old_values[row];
        const float ret = v_old + adv;
        const float dv = v - v_old;
        const float cv = v_old + fminf(fmaxf(dv, -a.clip), a.clip);
        const float e1 = (v - ret) * (v - ret), e2 = (cv - ret) * (cv - ret);
        acc[1] = fmaxf(e1, e2);
        if (a.dvalue) {
            const bool inside = (dv >= -a.clip) && (dv <= a.clip);
            float g = 0.f;
            if (inside) g = 2.f * (v - ret);
            else if (e1 > e2) g = 2.f * (v - ret);
            else if (e1 == e2) g = (v - ret);
            a.dvalue[n] = a.vf_coef * inv_n * g;
        }
    }
    for (int i = 0; i < 5; ++i) {
        const float t = block_sum(acc[i], sm);
        if (threadIdx.x == 0) a.partial[(long long)blockIdx.x * 5 + i] = t;
    }
}
__global__ void ppo_loss_final_kernel(const float* __restrict__ partial, int nblocks, const double* __restrict__ advstats,
                                      int nbranch, float beta, float vf_coef, float* __restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < nblocks; ++b)
        for (int i = 0; i < 5; ++i) s[i] += partial[(long long)b * 5 + i];
    const float n = (float)advstats[2];
    const float pl = s[0] / (n * nbranch), vl = s[1] / n, ent = s[2] / n;
    stats[0] = pl;
    stats[1] = vl;
    stats[2] = -(pl - vf_coef * vl + beta * ent);
    stats[3] = ent;
    stats[4] = s[3] / (n * nbranch);
    stats[5] = s[4] / (n * nbranch);
}

// ------------------------------------------------------------------------------------ clip + AdamW
// chunk table (built once on the host from the parameter layout): {start, length, group}
__global__ void sumsq_chunks_kernel(const float* __restrict__ g, const long long* __restrict__ chunks, float* __restrict__ part) {
    __shared__ float sm[32];
    const long long start = chunks[blockIdx.x * 3], len = chunks[blockIdx.x * 3 + 1];
    float s = 0.f;
    for (long long i = threadIdx.x; i < len; i += blockDim.x) {
        const float v = g[start + i];
        s = fmaf(v, v, s);
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
// norms[0..G-1] = per-group L2 norms (unclipped), norms[G] = total norm, norms[G+1] = clip coefficient
// One warp per group (warp G = the total): lanes stride over the chunk partials, then a fixed-order
// butterfly -> deterministic.  Launched with (G + 1) warps.
__global__ void clip_finalize_kernel(const float* __restrict__ part, const long long* __restrict__ chunks, int nchunks, int G,
                                     float max_norm, float* __restrict__ norms) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int gidx = warp; gidx <= G; gidx += nwarps) {
        float s = 0.f;
        for (int c = lane; c < nchunks; c += 32)
            if (gidx == G || chunks[c * 3 + 2] == gidx) s += part[c];
        s = warp_sum(s);
        if (lane == 0) {
            const float nrm = sqrtf(s);
            norms[gidx] = nrm;
            if (gidx == G) {
                const float coef = max_norm / (nrm + 1e-6f);
                norms[G + 1] = coef < 1.f ? coef : 1.f;
            }
        }
    }
}
__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long total, const float* __restrict__ coef_ptr, AdamWArgs h) {
    const float coef = coef_ptr ? *coef_ptr : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float gg = g[i] * coef;
        g[i] = gg;                                             // clipped gradient stays visible, as in torch
        float pp = p[i] * h.decay;
        float mm = m[i];
        mm = mm + h.one_minus_b1 * (gg - mm);
        float vv = v[i] * h.b2 + h.one_minus_b2 * gg * gg;
        const float denom = sqrtf(vv) / h.bc2_sqrt + h.eps;
        pp = pp - h.step_size * (mm / denom);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace

int ppo_gae(cudaStream_t st, const float* rewards, const unsigned char* dones, const float* values, const float* last_value,
            float* adv, int W, int T, double gamma, double lamda) {
    TRXL_CHECK_ARG(W >= 0 && T >= 0, "gae: bad dims");
    if (W == 0 || T == 0) return TRXL_OK;
    gae_kernel<<<trxl_cdiv(W, 4), 128, 0, st>>>(rewards, dones, values, last_value, adv, W, T, (float)gamma, (float)(gamma * lamda));
    TRXL_CHECK_LAUNCH("gae");
    return TRXL_OK;
}

int ppo_gather_rows(cudaStream_t st, const float* src, const long long* idx, float* dst, long long rows, long long row_floats) {
    if (rows == 0 || row_floats == 0) return TRXL_OK;
    const int vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0) && (row_floats % 4 == 0);
    const long long units = vec ? row_floats / 4 : row_floats;
    int bx = trxl_cdiv(units, 256);
    if (bx > 64) bx = 64;
    TRXL_CHECK_ARG(rows <= 65535LL * 1024, "gather_rows: too many rows");
    // grid.y is limited to 65535: loop in slabs
    for (long long r0 = 0; r0 < rows; r0 += 65535) {
        const long long nr = rows - r0 < 65535 ? rows - r0 : 65535;
        gather_rows_kernel<<<dim3(bx, (unsigned)nr), 256, 0, st>>>(src, idx + r0, dst + r0 * row_floats, row_floats, vec);
        TRXL_CHECK_LAUNCH("gather_rows");
    }
    return TRXL_OK;
}

int ppo_gather_window(cudaStream_t st, const float* in, const long long* idx, float* out, long long N, int L, long long slots,
                      long long inner) {
    if (N == 0) return TRXL_OK;
    TRXL_CHECK_ARG(N * L < (1LL << 31), "gather_window: too many rows");
    gather_window_kernel<<<(unsigned)(N * L), 128, 0, st>>>(in, idx, out, L, slots, inner);
    TRXL_CHECK_LAUNCH("gather_window");
    return TRXL_OK;
}

int ppo_memory_scatter(cudaStream_t st, float* table, const long long* ep, const long long* step, const float* new_mem, int W,
                       long long slots, long long inner) {
    if (W == 0) return TRXL_OK;
    memory_scatter_kernel<<<W, 128, 0, st>>>(table, ep, step, new_mem, slots, inner);
    TRXL_CHECK_LAUNCH("memory_scatter");
    return TRXL_OK;
}

int ppo_rollout_store(cudaStream_t st, float* table, float* table_pe, const float* pe_table, const long long* ep, const long long* step,
                      const float* new_mem, int W, long long slots, int blocks, int D, const float* value, float* value_dst,
                      long long value_stride) {
    if (W == 0) return TRXL_OK;
    rollout_store_kernel<<<W, 256, 0, st>>>(table, table_pe, pe_table, ep, step, new_mem, slots, blocks, D, value, value_dst, value_stride);
    TRXL_CHECK_LAUNCH("rollout_store");
    return TRXL_OK;
}

int ppo_rollout_prepare(cudaStream_t st, const long long* step, const long long* ep, const unsigned char* mask_table,
                        const long long* index_table, unsigned char* mask_out, long long mask_stride, long long* idx_out,
                        long long idx_stride, long long* ep_out, long long ep_stride, int W, int L) {
    if (W == 0) return TRXL_OK;
    rollout_prepare_kernel<<<W, 64, 0, st>>>(step, ep, mask_table, index_table, mask_out, mask_stride, idx_out, idx_stride,
                                            ep_out, ep_stride, L);
    TRXL_CHECK_LAUNCH("rollout_prepare");
    return TRXL_OK;
}

int ppo_rollout_fetch(cudaStream_t st, const float* obs_src, long long obs_floats, const long long* step_src, const long long* ep_src,
                      float* obs_dev, float* obs_store, long long store_stride_floats, long long* step_dev, long long* ep_dev, int n) {
    if (n == 0) return TRXL_OK;
    const bool vec = obs_floats % 4 == 0 && store_stride_floats % 4 == 0 && ((uintptr_t)obs_src % 16 == 0) &&
                     ((uintptr_t)obs_dev % 16 == 0) && ((uintptr_t)obs_store % 16 == 0);
    if (!vec) {
        int sblocks = trxl_cdiv((long long)n * obs_floats, 256);
        if (sblocks > 148 * 4) sblocks = 148 * 4;
        if (sblocks < trxl_cdiv(n, 256)) sblocks = trxl_cdiv(n, 256);
        rollout_fetch_scalar_kernel<<<sblocks, 256, 0, st>>>(obs_src, obs_floats, step_src, ep_src, obs_dev, obs_store,
                                                            store_stride_floats, step_dev, ep_dev, n);
        TRXL_CHECK_LAUNCH("rollout_fetch");
        return TRXL_OK;
    }
    const long long vec4 = obs_floats / 4, total = vec4 * n;
    int blocks = trxl_cdiv(total, 256 * 4);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    rollout_fetch_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(obs_src), vec4, step_src, ep_src,
                                                reinterpret_cast<float4*>(obs_dev), reinterpret_cast<float4*>(obs_store),
                                                store_stride_floats / 4, step_dev, ep_dev, n);
    TRXL_CHECK_LAUNCH("rollout_fetch");
    return TRXL_OK;
}

int ppo_sample_actions(cudaStream_t st, const float* logits, int sumA, const float* u, const long long* forced,
                       const BranchSpec& bs, long long* act_out,
                       long long act_stride, float* logp_out, long long logp_stride, long long* act_compact, int W,
                       long long* done_counter, long long* done_flag) {
    if (W == 0) return TRXL_OK;
    TRXL_CHECK_ARG(!done_flag || done_counter, "sample_actions: the completion flag needs a counter");
    TRXL_CHECK_ARG(!done_counter || W * bs.n <= 1024, "sample_actions: sequence-tagged actions need W * branches <= 1024");
    const int threads = done_counter ? ((W * bs.n + 31) / 32 * 32) : 128;
    sample_actions_kernel<<<done_counter ? 1 : trxl_cdiv(W * bs.n, 128), threads, 0, st>>>(logits, sumA, u, forced, bs, act_out, act_stride, logp_out,
                                                                   logp_stride, act_compact, W, done_counter, done_flag);
    TRXL_CHECK_LAUNCH("sample_actions");
    return TRXL_OK;
}

int ppo_adv_stats(cudaStream_t st, const float* adv, const long long* sidx, int N, double* out) {
    adv_stats_kernel<<<1, 1024, 0, st>>>(adv, sidx, N, out);
    TRXL_CHECK_LAUNCH("adv_stats");
    return TRXL_OK;
}

int ppo_loss(cudaStream_t st, PpoLossArgs a, float* stats) {
    TRXL_CHECK_ARG(a.N > 0 && a.bs.n >= 1 && a.bs.n <= TRXL_MAX_BRANCHES, "ppo_loss: bad sizes N=%d branches=%d", a.N, a.bs.n);
    const int nb = trxl_cdiv(a.N, 128);
    ppo_loss_kernel<<<nb, 128, 0, st>>>(a);
    TRXL_CHECK_LAUNCH("ppo_loss");
    ppo_loss_final_kernel<<<1, 32, 0, st>>>(a.partial, nb, a.advstats, a.bs.n, a.beta, a.vf_coef, stats);
    TRXL_CHECK_LAUNCH("ppo_loss_final");
    return TRXL_OK;
}
long long ppo_loss_partial_floats(int N) { return (long long)trxl_cdiv(N, 128) * 5 + 8; }

int ppo_clip_adamw(cudaStream_t st, float* params, float* grads, float* m, float* v, long long total, const long long* chunks,
                   int nchunks, int ngroups, float max_norm, float* partial, float* norms, const AdamWArgs& h) {
    if (nchunks > 0) {
        sumsq_chunks_kernel<<<nchunks, 256, 0, st>>>(grads, chunks, partial);
        TRXL_CHECK_LAUNCH("sumsq_chunks");
        clip_finalize_kernel<<<1, 32 * (ngroups + 1 < 32 ? ngroups + 1 : 32), 0, st>>>(partial, chunks, nchunks, ngroups, max_norm, norms);
        TRXL_CHECK_LAUNCH("clip_finalize");
    }
    int blocks = trxl_cdiv(total, 256 * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    adamw_kernel<<<blocks, 256, 0, st>>>(params, grads, m, v, total, nchunks > 0 ? norms + ngroups + 1 : nullptr, h);
    TRXL_CHECK_LAUNCH("adamw");
    return TRXL_OK;
}
