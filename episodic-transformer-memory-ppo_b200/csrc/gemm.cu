// fp32 SIMT GEMM with strided/batched operands and a fused epilogue.
//
//   C[b][m, n] (+)= epi( sum_k A[b](m, k) * B[b](k, n) )        epi: +bias[n], ReLU, +R[m, n]
//
// A is either "k-contiguous" (a_kc=1: A[m*lda + k], a row-major (M, K) matrix) or "m-contiguous"
// (a_kc=0: A[k*lda + m], i.e. the transpose of a row-major (K, M) matrix).  Same for B with n.
// The three combinations the model needs:
//   forward  y = x W^T      : A = x (kc),   B = W  (kc: B(k, n) = W[n*ldw + k])
//   dgrad    dx = dy W      : A = dy (kc),  B = W  (nc: B(k, n) = W[k*ldw + n])
//   wgrad    dW = dy^T x    : A = dy (mc),  B = x  (nc)
// fp32 accumulate in a fixed k order (deterministic).  The parity contract of this engine is 1e-4
// against the reference's fp32 CPU path, so the linears stay in full fp32 on the CUDA cores here.
#include <stdlib.h>

#include "gemm.cuh"

namespace {

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN, bool AKC, bool BKC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const GemmArgs g) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int b = blockIdx.z / g.ksplit, split = blockIdx.z % g.ksplit;
    const int k_begin = split * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const float* __restrict__ A = g.A + (long long)b * g.sA;
    const float* __restrict__ B = g.B + (long long)b * g.sB;
    float* __restrict__ C = g.C + (long long)b * g.sC;
    const float* __restrict__ bias = g.bias ? g.bias + (long long)b * g.sBias : nullptr;
    const float* __restrict__ R = g.R ? g.R + (long long)b * g.sR : nullptr;

    constexpr int A_V4 = BM * BK / 4, B_V4 = BN * BK / 4;
    constexpr int A_PER = (A_V4 + NT - 1) / NT, B_PER = (B_V4 + NT - 1) / NT;
    float4 ra0[A_PER], rb0[B_PER], ra1[A_PER], rb1[B_PER];      // two register stages: tiles kt+1 and kt+2 in flight

    auto load_a = [&](int k0, float4 (&ra)[A_PER]) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int v = tid + i * NT;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < A_V4) {
                if (AKC) {
                    const int m = m0 + v / (BK / 4), k = k0 + (v % (BK / 4)) * 4;
                    if (m < g.M) {
                        const float* p = A + (long long)m * g.lda + k;
                        if (g.vecA && k + 3 < k_end) val = *reinterpret_cast<const float4*>(p);
                        else {
                            if (k < k_end) val.x = p[0];
                            if (k + 1 < k_end) val.y = p[1];
                            if (k + 2 < k_end) val.z = p[2];
                            if (k + 3 < k_end) val.w = p[3];
                        }
                    }
                } else {
                    const int k = k0 + v / (BM / 4), m = m0 + (v % (BM / 4)) * 4;
                    if (k < k_end) {
                        const float* p = A + (long long)k * g.lda + m;
                        if (g.vecA && m + 3 < g.M) val = *reinterpret_cast<const float4*>(p);
                        else {
                            if (m < g.M) val.x = p[0];
                            if (m + 1 < g.M) val.y = p[1];
                            if (m + 2 < g.M) val.z = p[2];
                            if (m + 3 < g.M) val.w = p[3];
                        }
                    }
                }
            }
            ra[i] = val;
        }
    };
    auto load_b = [&](int k0, float4 (&rb)[B_PER]) {
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int v = tid + i * NT;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < B_V4) {
                if (BKC) {
                    const int n = n0 + v / (BK / 4), k = k0 + (v % (BK / 4)) * 4;
                    if (n < g.N) {
                        const float* p = B + (long long)n * g.ldb + k;
                        if (g.vecB && k + 3 < k_end) val = *reinterpret_cast<const float4*>(p);
                        else {
                            if (k < k_end) val.x = p[0];
                            if (k + 1 < k_end) val.y = p[1];
                            if (k + 2 < k_end) val.z = p[2];
                            if (k + 3 < k_end) val.w = p[3];
                        }
                    }
                } else {
                    const int k = k0 + v / (BN / 4), n = n0 + (v % (BN / 4)) * 4;
                    if (k < k_end) {
                        const float* p = B + (long long)k * g.ldb + n;
                        if (g.vecB && n + 3 < g.N) val = *reinterpret_cast<const float4*>(p);
                        else {
                            if (n < g.N) val.x = p[0];
                            if (n + 1 < g.N) val.y = p[1];
                            if (n + 2 < g.N) val.z = p[2];
                            if (n + 3 < g.N) val.w = p[3];
                        }
                    }
                }
            }
            rb[i] = val;
        }
    };
    auto store_tiles = [&](int buf, const float4 (&ra)[A_PER], const float4 (&rb)[B_PER]) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int v = tid + i * NT;
            if (v < A_V4) {
                if (AKC) {
                    const int m = v / (BK / 4), k = (v % (BK / 4)) * 4;
                    As[buf][k][m] = ra[i].x; As[buf][k + 1][m] = ra[i].y;
                    As[buf][k + 2][m] = ra[i].z; As[buf][k + 3][m] = ra[i].w;
                } else {
                    const int k = v / (BM / 4), m = (v % (BM / 4)) * 4;
                    *reinterpret_cast<float4*>(&As[buf][k][m]) = ra[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int v = tid + i * NT;
            if (v < B_V4) {
                if (BKC) {
                    const int n = v / (BK / 4), k = (v % (BK / 4)) * 4;
                    Bs[buf][k][n] = rb[i].x; Bs[buf][k + 1][n] = rb[i].y;
                    Bs[buf][k + 2][n] = rb[i].z; Bs[buf][k + 3][n] = rb[i].w;
                } else {
                    const int k = v / (BN / 4), n = (v % (BN / 4)) * 4;
                    *reinterpret_cast<float4*>(&Bs[buf][k][n]) = rb[i];
                }
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (k_end - k_begin + BK - 1) / BK;
    auto compute = [&](int buf) {
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], bb[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[buf][k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bb[j] = Bs[buf][k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    };
    // software pipeline: tile kt is computed from shared memory while kt+1 and kt+2 are in flight in registers
    load_a(k_begin, ra0); load_b(k_begin, rb0);
    store_tiles(0, ra0, rb0);
    if (nk > 1) { load_a(k_begin + BK, ra1); load_b(k_begin + BK, rb1); }
    __syncthreads();
    for (int kt = 0; kt < nk; kt += 2) {
        // even step: shared buffer 0, tile kt+1 waits in stage 1, tile kt+2 is issued into stage 0
        if (kt + 2 < nk) { load_a(k_begin + (kt + 2) * BK, ra0); load_b(k_begin + (kt + 2) * BK, rb0); }
        compute(0);
        if (kt + 1 < nk) {
            store_tiles(1, ra1, rb1);
            __syncthreads();
            // odd step: shared buffer 1, tile kt+2 waits in stage 0, tile kt+3 is issued into stage 1
            if (kt + 3 < nk) { load_a(k_begin + (kt + 3) * BK, ra1); load_b(k_begin + (kt + 3) * BK, rb1); }
            compute(1);
            if (kt + 2 < nk) {
                store_tiles(0, ra0, rb0);
                __syncthreads();
            }
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            if (g.ksplit > 1) {      // raw partial; the epilogue runs in splitk_reduce_kernel
                g.ws[((long long)blockIdx.z * g.M + m) * g.N + n] = acc[i][j];
                continue;
            }
            float v = acc[i][j] * g.alpha;
            if (bias) v += bias[n];
            if (g.relu) v = fmaxf(v, 0.f);
            if (R) v += R[(long long)m * g.ldr + n];
            float* cp = C + (long long)m * g.ldc + n;
            if (g.accumulate) v += *cp;
            *cp = v;
        }
    }
}

// sums the split-K partials in split order (deterministic) and applies the epilogue
__global__ void splitk_reduce_kernel(const GemmArgs g) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long mn = (long long)g.M * g.N;
    if (i >= mn * g.batch) return;
    const int b = (int)(i / mn);
    const int m = (int)((i % mn) / g.N), n = (int)(i % g.N);
    float v = 0.f;
    for (int s = 0; s < g.ksplit; ++s) v += g.ws[((long long)(b * g.ksplit + s) * g.M + m) * g.N + n];
    v *= g.alpha;
    if (g.bias) v += g.bias[(long long)b * g.sBias + n];
    if (g.relu) v = fmaxf(v, 0.f);
    if (g.R) v += g.R[(long long)b * g.sR + (long long)m * g.ldr + n];
    float* cp = g.C + (long long)b * g.sC + (long long)m * g.ldc + n;
    if (g.accumulate) v += *cp;
    *cp = v;
}


// ------------------------------------------------------------------------------------------------
// Skinny GEMM for the rollout path: M <= 32 rows (one per env worker).  The tiled kernel above is
// latency-bound there (8 CTAs x 16 dependent k-iterations); here a CTA owns 16 output columns, pulls
// its whole K panel (<= 256 deep per pass) into shared memory with every load in flight at once, and
// then runs a register-blocked (1 row x 4 columns per thread) inner product out of shared memory.
// A must be k-contiguous; B may be k- or n-contiguous.  Split-K and the fused epilogue work as above.
// ------------------------------------------------------------------------------------------------
constexpr int SK_KC = 256, SK_BN = 16, SK_LD = SK_KC + 4;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}

template <bool BKC>
__global__ void __launch_bounds__(128) skinny_gemm_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) float sk_smem[];
    float* As = sk_smem;                    // [32][SK_LD]
    float* Bs = sk_smem + 32 * SK_LD;       // [SK_BN][SK_LD]
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * SK_BN;
    const int b = blockIdx.z / g.ksplit, split = blockIdx.z % g.ksplit;
    const int k_begin = split * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const float* __restrict__ A = g.A + (long long)b * g.sA;
    const float* __restrict__ B = g.B + (long long)b * g.sB;
    const int m = tid & 31, ng = tid >> 5;            // this thread: row m, columns n0 + 4*ng .. +3
    float acc[4] = {0.f, 0.f, 0.f, 0.f};

    for (int k0 = k_begin; k0 < k_end; k0 += SK_KC) {
        const int kc = min(SK_KC, k_end - k0);
        const int kc4 = (kc + 3) & ~3;
        // Every copy of the panel is issued before anything waits (cp.async, 16 B each, zero-filled
        // past the matrix edge): the CTA pays one L2 round trip per K panel, not one per element.
        // ---- A panel: 32 rows x kc, k-contiguous ----
        if (g.vecA && (k0 & 3) == 0) {
#pragma unroll
            for (int i = 0; i < (32 * SK_KC / 4) / 128; ++i) {
                const int v = tid + i * 128;
                const int r = v / (SK_KC / 4), k = (v % (SK_KC / 4)) * 4;
                if (k < kc4) {
                    const int valid = (r < g.M) ? max(0, min(4, kc - k)) : 0;
                    const float* p = valid ? A + (long long)r * g.lda + k0 + k : A;
                    cp_async16(As + r * SK_LD + k, p, valid * 4);
                }
            }
        } else {
            for (int v = tid; v < 32 * kc4; v += 128) {
                const int r = v / kc4, k = v % kc4;
                As[r * SK_LD + k] = (r < g.M && k < kc) ? A[(long long)r * g.lda + k0 + k] : 0.f;
            }
        }
        // ---- B panel: 16 columns x kc ----
        if (BKC) {
            if (g.vecB && (k0 & 3) == 0) {
#pragma unroll
                for (int i = 0; i < (SK_BN * SK_KC / 4) / 128; ++i) {
                    const int v = tid + i * 128;
                    const int c = v / (SK_KC / 4), k = (v % (SK_KC / 4)) * 4;
                    if (k < kc4) {
                        const int valid = (n0 + c < g.N) ? max(0, min(4, kc - k)) : 0;
                        const float* p = valid ? B + (long long)(n0 + c) * g.ldb + k0 + k : B;
                        cp_async16(Bs + c * SK_LD + k, p, valid * 4);
                    }
                }
            } else {
                for (int v = tid; v < SK_BN * kc4; v += 128) {
                    const int c = v / kc4, k = v % kc4;
                    Bs[c * SK_LD + k] = (n0 + c < g.N && k < kc) ? B[(long long)(n0 + c) * g.ldb + k0 + k] : 0.f;
                }
            }
        } else {
            // 16 consecutive n per k row (64 B segments); 4-byte async copies keep every load in flight
#pragma unroll 8
            for (int v = tid; v < SK_BN * kc4; v += 128) {
                const int k = v / SK_BN, c = v % SK_BN;
                const bool ok = (n0 + c < g.N) && (k < kc);
                cp_async4(Bs + c * SK_LD + k, ok ? B + (long long)(k0 + k) * g.ldb + n0 + c : B, ok ? 4 : 0);
            }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float* ar = As + m * SK_LD;
        const float* b0 = Bs + (4 * ng) * SK_LD;
#pragma unroll 4
        for (int k = 0; k < kc4; k += 4) {
            const float4 a = *reinterpret_cast<const float4*>(ar + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = *reinterpret_cast<const float4*>(b0 + j * SK_LD + k);
                acc[j] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[j]))));
            }
        }
        __syncthreads();
    }
    if (m >= g.M) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + 4 * ng + j;
        if (n >= g.N) continue;
        if (g.ksplit > 1) {
            g.ws[((long long)blockIdx.z * g.M + m) * g.N + n] = acc[j];
            continue;
        }
        float v = acc[j] * g.alpha;
        if (g.bias) v += g.bias[(long long)b * g.sBias + n];
        if (g.relu) v = fmaxf(v, 0.f);
        if (g.R) v += g.R[(long long)b * g.sR + (long long)m * g.ldr + n];
        float* cp = g.C + (long long)b * g.sC + (long long)m * g.ldc + n;
        if (g.accumulate) v += *cp;
        *cp = v;
    }
}

int launch_skinny(const GemmArgs& g, cudaStream_t st) {
    static bool attr_set = false;
    constexpr int smem = (32 + SK_BN) * SK_LD * (int)sizeof(float);
    if (!attr_set) {
        cudaFuncSetAttribute(skinny_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(skinny_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    dim3 grid(trxl_cdiv(g.N, SK_BN), 1, g.batch * g.ksplit);
    if (g.b_kc) skinny_gemm_kernel<true><<<grid, 128, smem, st>>>(g);
    else skinny_gemm_kernel<false><<<grid, 128, smem, st>>>(g);
    TRXL_CHECK_LAUNCH("skinny_gemm");
    return TRXL_OK;
}

template <int BM, int BN, int TM, int TN>
int launch_cfg(const GemmArgs& g, cudaStream_t st) {
    dim3 grid(trxl_cdiv(g.N, BN), trxl_cdiv(g.M, BM), g.batch * g.ksplit);
    dim3 block((BM / TM) * (BN / TN));
    if (g.a_kc && g.b_kc) sgemm_kernel<BM, BN, TM, TN, true, true><<<grid, block, 0, st>>>(g);
    else if (g.a_kc && !g.b_kc) sgemm_kernel<BM, BN, TM, TN, true, false><<<grid, block, 0, st>>>(g);
    else if (!g.a_kc && !g.b_kc) sgemm_kernel<BM, BN, TM, TN, false, false><<<grid, block, 0, st>>>(g);
    else sgemm_kernel<BM, BN, TM, TN, false, true><<<grid, block, 0, st>>>(g);
    TRXL_CHECK_LAUNCH("sgemm");
    return TRXL_OK;
}

}  // namespace

// tc_gemm.cu: TMA + tcgen05 3xTF32 path
int trxl_tc_gemm(const GemmArgs& g, int bn, cudaStream_t st);
bool trxl_tc_gemm_eligible(const GemmArgs& g);
int trxl_tc_gemm_tile_n(const GemmArgs& g);

// TRXL_TCGEN05=0 disables the tensor-core GEMM (falls back to the SIMT kernels above, same results to ~1e-6)
static bool tc_enabled() {
    static int state = -1;
    if (state < 0) {
        const char* e = getenv("TRXL_TCGEN05");
        state = (e && e[0] == '0') ? 0 : 1;
    }
    return state == 1;
}

static thread_local float* tl_ws = nullptr;
static thread_local long long tl_ws_floats = 0;
void trxl_gemm_set_workspace(float* ws, long long floats) { tl_ws = ws; tl_ws_floats = floats; }

int trxl_gemm(GemmArgs g, cudaStream_t st) {
    TRXL_CHECK_ARG(g.M >= 0 && g.N >= 0 && g.K >= 0 && g.batch >= 1, "gemm: bad dims M=%d N=%d K=%d batch=%d", g.M, g.N, g.K, g.batch);
    if (g.M == 0 || g.N == 0) return TRXL_OK;
    auto aligned = [](const void* p, long long ld, long long bs) {
        return ((uintptr_t)p % 16 == 0) && (ld % 4 == 0) && (bs % 4 == 0);
    };
    g.vecA = aligned(g.A, g.lda, g.sA);
    g.vecB = aligned(g.B, g.ldb, g.sB);
    if (g.alpha == 0.f) g.alpha = 1.f;
    if (!g.ws && tl_ws) { g.ws = tl_ws; g.ws_floats = tl_ws_floats; }
    // tile choice: small problems get small tiles so more CTAs are in flight
    const long long tiles64 = (long long)trxl_cdiv(g.M, 64) * trxl_cdiv(g.N, 64) * g.batch;
    const bool small = (g.M <= 32 || g.N <= 32 || tiles64 < 96);
    const long long tiles = small ? (long long)trxl_cdiv(g.M, 32) * trxl_cdiv(g.N, 32) * g.batch : tiles64;
    // split-K: shapes with few output tiles and a long reduction (weight gradients: K = samples; the
    // rollout's lin_hidden: M = n_workers, K = 3136) would otherwise occupy a handful of SMs
    g.ksplit = 1;
    g.k_per_split = g.K;
    if (g.ws && g.K >= 512 && tiles * 2 <= 148) {
        int want = (int)((148 * 2 + tiles - 1) / tiles);
        int maxs = g.K / 64;
        int s = want < maxs ? want : maxs;
        if (s > 32) s = 32;
        while (s > 1 && (long long)s * g.batch * g.M * g.N > g.ws_floats) --s;
        if (s > 1) {
            g.k_per_split = ((g.K + s - 1) / s + BK - 1) / BK * BK;
            g.ksplit = (g.K + g.k_per_split - 1) / g.k_per_split;
        }
    }
    int rc;
    if (tc_enabled() && g.M >= 64 && g.N >= 16 && g.K >= 16 && trxl_tc_gemm_eligible(g)) {
        // tensor-core path (tc_gemm.cu): 128 x {32,64,128} tiles, TMA-staged operands, 3xTF32 in TMEM; split K when few tiles
        GemmArgs t = g;
        const int bn = trxl_tc_gemm_tile_n(t);
        const long long tc_tiles = (long long)trxl_cdiv(t.M, 128) * trxl_cdiv(t.N, bn) * t.batch;
        t.ksplit = 1;
        t.k_per_split = t.K;
        if (t.ws && t.K >= 256 && tc_tiles * 2 <= 148) {
            int s = (int)((148 + tc_tiles - 1) / tc_tiles);
            const int maxs = t.K / 64;
            if (s > maxs) s = maxs;
            if (s > 32) s = 32;
            while (s > 1 && (long long)s * t.batch * t.M * t.N > t.ws_floats) --s;
            if (s > 1) {
                t.k_per_split = ((t.K + s - 1) / s + 31) / 32 * 32;
                t.ksplit = (t.K + t.k_per_split - 1) / t.k_per_split;
            }
        }
        // up to 8 splits: the CTAs of a tile form a cluster and reduce through distributed shared memory -- no workspace round
        // trip and no reduce kernel (TRXL_TC_CLUSTER_SPLITK=0: workspace + splitk_reduce_kernel as for the SIMT path)
        static int cluster_off = -1;
        if (cluster_off < 0) { const char* e = getenv("TRXL_TC_CLUSTER_SPLITK"); cluster_off = (e && atoi(e) == 0) ? 1 : 0; }
        if (!cluster_off && t.K >= 256 && tc_tiles * 2 <= 148 && bn <= 128) {
            int s = (int)((148 + tc_tiles - 1) / tc_tiles);
            s = s >= 8 ? 8 : (s >= 4 ? 4 : 2);
            while (s > 1 && t.K / s < 64) s >>= 1;
            if (s > 1) {
                t.k_per_split = ((t.K + s - 1) / s + 31) / 32 * 32;
                if ((t.K + t.k_per_split - 1) / t.k_per_split == s) { t.ksplit = s; t.cluster_reduce = 1; }
            }
        }
        rc = trxl_tc_gemm(t, bn, st);
        if (rc != TRXL_ERR_UNSUPPORTED) {
            if (rc != TRXL_OK || t.ksplit == 1 || t.cluster_reduce) return rc;
            const long long total_t = (long long)t.M * t.N * t.batch;
            splitk_reduce_kernel<<<trxl_cdiv(total_t, 256), 256, 0, st>>>(t);
            TRXL_CHECK_LAUNCH("splitk_reduce");
            return TRXL_OK;
        }
        // a tensor map could not be encoded for these operands: SIMT path below
    }
    if (false) {
    } else if (g.M <= 32 && g.a_kc && g.N >= 8) {
        // rollout-sized batch: one CTA per 16 output columns; split K only when the panel is deep
        g.ksplit = 1;
        g.k_per_split = g.K;
        const long long ctas = (long long)trxl_cdiv(g.N, SK_BN) * g.batch;
        if (g.ws && g.K > 2 * SK_KC && ctas < 148) {
            int s = (int)((148 * 2 + ctas - 1) / ctas);
            const int maxs = trxl_cdiv(g.K, SK_KC);
            if (s > maxs) s = maxs;
            while (s > 1 && (long long)s * g.batch * g.M * g.N > g.ws_floats) --s;
            if (s > 1) {
                g.k_per_split = ((g.K + s - 1) / s + 3) / 4 * 4;
                g.ksplit = (g.K + g.k_per_split - 1) / g.k_per_split;
            }
        }
        rc = launch_skinny(g, st);
    } else if (small) rc = launch_cfg<32, 32, 2, 2>(g, st);
    else rc = launch_cfg<64, 64, 4, 4>(g, st);
    if (rc != TRXL_OK || g.ksplit == 1) return rc;
    const long long total = (long long)g.M * g.N * g.batch;
    splitk_reduce_kernel<<<trxl_cdiv(total, 256), 256, 0, st>>>(g);
    TRXL_CHECK_LAUNCH("splitk_reduce");
    return TRXL_OK;
}
