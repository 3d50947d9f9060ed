// tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy: 3xTF32.
//
//   C[b][m, n] (+)= epi( sum_k A[b](m, k) * B[b](k, n) )     same contract as the SIMT kernel in gemm.cu
//
// The parity contract of this engine is fp32 (1e-4 vs the reference's CPU path); a single TF32 MMA
// (10-bit mantissa) does not hold that through 4 blocks.  Every operand element is therefore split
// into hi = rna_tf32(x) and lo = rna_tf32(x - hi) and three MMAs accumulate hi*hi + lo*hi + hi*lo in
// an fp32 TMEM accumulator (the dropped lo*lo term is ~2^-22 relative).
//
// Structure (one CTA = one 128 x BN output tile, 160 threads):
//   warps 0-3  producers: coalesced fp32 global loads of the A and B tiles for one k-block (32 deep),
//              hi/lo split in registers, st.shared into the canonical UMMA K-major no-swizzle layout
//              (8-row x 16-byte core matrices; any operand orientation is just a different gather),
//              fence.proxy.async, mbarrier arrive.  After the main loop the same warps run the
//              epilogue: tcgen05.ld of their 32 TMEM lanes, bias / ReLU / residual, global stores.
//   warp 4     allocates TMEM; one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) x 4
//              k-steps x 3 split terms per stage and tcgen05.commit's the stage back to the producers.
// Stage ring of mbarriers (full: producers -> MMA, empty: MMA -> producers, tmem_full: MMA -> epilogue).
// Every mbarrier wait is bounded and traps instead of hanging the GPU.
#include "gemm.cuh"
#include "tc_common.cuh"

long long g_trxl_tc_launches = 0;

namespace {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_THREADS = 160;

using namespace tc;      // mbarrier / tcgen05 / descriptor primitives (tc_common.cuh)

// gather a 4-wide k-chunk of row r of a strided operand: kc=1 -> X[r*ld + k], kc=0 -> X[k*ld + r]
__device__ __forceinline__ float4 load_chunk(const float* __restrict__ X, long long ld, int kc, int vec, int r, int R, int k, int Kend) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= R) return v;
    if (kc) {
        const float* p = X + (long long)r * ld + k;
        if (vec && k + 3 < Kend) return *reinterpret_cast<const float4*>(p);
        if (k < Kend) v.x = p[0];
        if (k + 1 < Kend) v.y = p[1];
        if (k + 2 < Kend) v.z = p[2];
        if (k + 3 < Kend) v.w = p[3];
    } else {
        const float* p = X + (long long)k * ld + r;
        if (k < Kend) v.x = p[0];
        if (k + 1 < Kend) v.y = p[ld];
        if (k + 2 < Kend) v.z = p[2 * ld];
        if (k + 3 < Kend) v.w = p[3 * ld];
    }
    return v;
}
__device__ __forceinline__ void split_store(float* hi, float* lo, const float4& v) {
    uint4 h, l;
    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
    l.x = to_tf32(v.x - __uint_as_float(h.x)); l.y = to_tf32(v.y - __uint_as_float(h.y));
    l.z = to_tf32(v.z - __uint_as_float(h.z)); l.w = to_tf32(v.w - __uint_as_float(h.w));
    *reinterpret_cast<uint4*>(hi) = h;
    *reinterpret_cast<uint4*>(lo) = l;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const GemmArgs g) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    constexpr int A_FLOATS = TC_BM * TC_BK, B_FLOATS = BN * TC_BK;
    constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    float* tiles = reinterpret_cast<float*>(tc_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_FLOATS);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    const int b = blockIdx.z / g.ksplit, split = blockIdx.z % g.ksplit;
    const int k_begin = split * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int kblocks = (k_end - k_begin + TC_BK - 1) / TC_BK;
    const float* __restrict__ A = g.A + (long long)b * g.sA;
    const float* __restrict__ B = g.B + (long long)b * g.sB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ---------------- producers ----------------
        // Register-level software pipeline, two k-blocks deep: the global loads of k-block kb+2 are in flight
        // while kb is split/stored and kb+1 waits in registers, so L2 latency is hidden behind the MMAs.
        const int r = threadIdx.x;                               // tile row owned by this thread (A: 0..127, B: 0..BN-1)
        constexpr int NC = TC_BK / 4;                            // 16-byte k-chunks per row per k-block
        const int row_off = (r >> 3) * 32 + (r & 7) * 4;         // core-matrix layout: chunk c at c*(ROWS*4) + (row/8)*32 + (row%8)*4
        float4 ra[2][NC], rb[2][NC];
        auto issue = [&](int kb, float4 (&xa)[NC], float4 (&xb)[NC]) {
            const int k0 = k_begin + kb * TC_BK;
#pragma unroll
            for (int c = 0; c < NC; ++c) xa[c] = load_chunk(A, g.lda, g.a_kc, g.vecA, m0 + r, g.M, k0 + 4 * c, k_end);
            if (r < BN) {
#pragma unroll
                for (int c = 0; c < NC; ++c) xb[c] = load_chunk(B, g.ldb, g.b_kc, g.vecB, n0 + r, g.N, k0 + 4 * c, k_end);
            }
        };
        auto commit = [&](int kb, const float4 (&xa)[NC], const float4 (&xb)[NC]) {
            const int s = kb % STAGES;
            mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
            float* a_hi = tiles + s * STAGE_FLOATS;
            float* a_lo = a_hi + A_FLOATS;
            float* b_hi = a_lo + A_FLOATS;
            float* b_lo = b_hi + B_FLOATS;
#pragma unroll
            for (int c = 0; c < NC; ++c) split_store(a_hi + c * (TC_BM * 4) + row_off, a_lo + c * (TC_BM * 4) + row_off, xa[c]);
            if (r < BN) {
#pragma unroll
                for (int c = 0; c < NC; ++c) split_store(b_hi + c * (BN * 4) + row_off, b_lo + c * (BN * 4) + row_off, xb[c]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        };
        issue(0, ra[0], rb[0]);
        if (kblocks > 1) issue(1, ra[1], rb[1]);
        for (int kb = 0; kb < kblocks; kb += 2) {
            commit(kb, ra[0], rb[0]);
            if (kb + 2 < kblocks) issue(kb + 2, ra[0], rb[0]);
            if (kb + 1 < kblocks) {
                commit(kb + 1, ra[1], rb[1]);
                if (kb + 3 < kblocks) issue(kb + 3, ra[1], rb[1]);
            }
        }
        // ---------------- epilogue: TMEM -> registers -> global ----------------
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = m0 + warp * 32 + lane;
        float* C = g.C + (long long)b * g.sC;
        const float* bias = g.bias ? g.bias + (long long)b * g.sBias : nullptr;
        const float* R = g.R ? g.R + (long long)b * g.sR : nullptr;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < g.M) {
                const int nb0 = n0 + j * 32;
                float* dst = (g.ksplit > 1) ? g.ws + ((long long)blockIdx.z * g.M + m) * g.N : C + (long long)m * g.ldc;
                const bool fast = (nb0 + 32 <= g.N) && ((((uintptr_t)(dst + nb0)) & 15) == 0) &&
                                  (g.ksplit > 1 || (!R && !g.accumulate));
                if (fast) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 o;
                        float* of = reinterpret_cast<float*>(&o);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float x = __uint_as_float(v[i + q]);
                            if (g.ksplit == 1) {
                                x *= g.alpha;
                                if (bias) x += bias[nb0 + i + q];
                                if (g.relu) x = fmaxf(x, 0.f);
                            }
                            of[q] = x;
                        }
                        *reinterpret_cast<float4*>(dst + nb0 + i) = o;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int n = nb0 + i;
                        if (n >= g.N) continue;
                        float x = __uint_as_float(v[i]);
                        if (g.ksplit == 1) {
                            x *= g.alpha;
                            if (bias) x += bias[n];
                            if (g.relu) x = fmaxf(x, 0.f);
                            if (R) x += R[(long long)m * g.ldr + n];
                            if (g.accumulate) x += dst[n];
                        }
                        dst[n] = x;
                    }
                }
            }
        }
    } else {
        // ---------------- MMA issuer (warp 4) ----------------
        // instruction descriptor (cute UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format TF32 (2<<7, 2<<10),
        // K-major A and B (bits 15,16 = 0), N>>3 at bit 17, M>>4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(tiles + s * STAGE_FLOATS);
                const uint32_t a_lo = a_hi + A_FLOATS * 4;
                const uint32_t b_hi = a_lo + A_FLOATS * 4;
                const uint32_t b_lo = b_hi + B_FLOATS * 4;
                constexpr uint32_t A_LBO = TC_BM * 16, B_LBO = BN * 16, SBO = 128;
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {            // one UMMA consumes K = 8 tf32 = 2 core matrices
                    const uint64_t dah = make_desc(a_hi + k * 2 * A_LBO, A_LBO, SBO);
                    const uint64_t dal = make_desc(a_lo + k * 2 * A_LBO, A_LBO, SBO);
                    const uint64_t dbh = make_desc(b_hi + k * 2 * B_LBO, B_LBO, SBO);
                    const uint64_t dbl = make_desc(b_lo + k * 2 * B_LBO, B_LBO, SBO);
                    umma_tf32(tmem_base, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, dal, dbh, idesc, 1u);
                    umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                }
                umma_commit(&empty[s]);                          // frees the stage once these MMAs have read it
                if (kb == kblocks - 1) umma_commit(tmem_full);   // accumulator complete -> epilogue
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

template <int BN, int STAGES>
int launch_tc(const GemmArgs& g, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (2 * TC_BM * TC_BK + 2 * BN * TC_BK) * 4 + (2 * STAGES + 1) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { trxl_set_error("tc_gemm: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
        attr_set = true;
    }
    dim3 grid(trxl_cdiv(g.N, BN), trxl_cdiv(g.M, TC_BM), g.batch * g.ksplit);
    tc_gemm_kernel<BN, STAGES><<<grid, TC_THREADS, smem, st>>>(g);
    ++g_trxl_tc_launches;
    TRXL_CHECK_LAUNCH("tc_gemm");
    return TRXL_OK;
}

}  // namespace

// Returns TRXL_OK and sets *handled = 1 if the tensor-core path took the GEMM (g.ksplit/k_per_split already chosen).
int trxl_tc_gemm(const GemmArgs& g, cudaStream_t st) {
    if (g.N > 64) return launch_tc<128, 3>(g, st);
    return launch_tc<64, 4>(g, st);
}
