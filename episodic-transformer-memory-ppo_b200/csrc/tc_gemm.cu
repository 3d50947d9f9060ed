// tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy (3xTF32), operands staged by TMA.
//
//   C[b][m, n] (+)= epi( sum_k A[b](m, k) * B[b](k, n) )     same contract as the SIMT kernel in gemm.cu
//
// Replaces the reference's nn.Linear forward / autograd backward calls on the training path (transformer.py:50-52,82,158,
// 295-298; model.py:97,104-110).  The parity contract is fp32 (1e-4 vs the reference's CPU path); a single TF32 MMA
// (10-bit mantissa) does not hold that through 4 blocks, so every operand element is split into hi = rna_tf32(x) and
// lo = rna_tf32(x - hi) and three MMAs accumulate hi*hi + lo*hi + hi*lo (the dropped lo*lo term is ~2^-22 relative).  The
// tensor core adds into TMEM with truncation, so three accumulators are kept (hi*hi products alternate between two, the
// ~2^-11 smaller cross terms go to the third) and summed with round-to-nearest in the epilogue -- the scheme tc_conv.cu
// established at fp32-SIMT accuracy.
//
// One CTA = one 128 x BN output tile (one split of K), 320 threads:
//   warp 9     TMA producer: one elected lane issues cp.async.bulk.tensor loads of the raw fp32 A and B tiles of a k-block
//              (32 deep) into a STAGES-deep shared-memory ring; completion is tracked by the stage's `full` mbarrier (tx bytes).
//              Out-of-range rows / columns / k are zero-filled by the TMA unit, so no operand needs padding.
//   warps 0-7  converters: split the landed raw tile IN PLACE into hi and write lo to the twin tile (element-wise, so the
//              shared-memory swizzle is irrelevant to them), fence.proxy.async, arrive on the stage's `ready` mbarrier.
//              After the main loop the same warps run the epilogue: tcgen05.ld of their 32 TMEM lanes, bias / ReLU /
//              residual / accumulate, global stores.
//   warp 8     allocates TMEM; one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8): 4 k-steps x 3 split terms
//              per stage, and tcgen05.commit's the stage back to the producer.
// Operand orientations (all three the model needs, no transposed copies anywhere):
//   K-major  (x (M,K) row-major; W (N,K) row-major)  : TMA box {32 k, rows}, SWIZZLE_128B       -> UMMA K-major SW128
//   MN-major (dy (K,M) / x (K,N) / W (K,N) row-major) : TMA boxes {32 mn, 32 k}, SWIZZLE_128B_ATOM_32B
//                                                       -> UMMA MN-major SWIZZLE_128B_BASE32B (the only MN-major TF32 layout)
// Every mbarrier wait is bounded and traps instead of hanging the GPU.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "gemm.cuh"
#include "tc_common.cuh"

long long g_trxl_tc_launches = 0;

namespace {

using namespace tc;

constexpr int TM_BM = 128;
constexpr int TM_CONV_WARPS = 8, TM_MMA_WARP = 8, TM_TMA_WARP = 9, TM_THREADS = 320;
// k-block depth BK (floats): 32 everywhere except the 256-wide tiles of the episode-grouped attention GEMMs, which use 16 -- at
// BN = 256 a 32-deep stage is 96 KB, i.e. a ring of two, and the kernel ran at 4.8 k clocks per k-block against the 2.6 k of its
// shared-memory traffic (TMA latency + conversion + MMAs of one stage serialised); 16-deep stages make it a ring of four.
__host__ __device__ constexpr int tm_bk(int bn) { return bn == 256 ? 16 : 32; }
// A tile: K-major 128 rows x (4 BK) bytes (SWIZZLE_128B at BK = 32, SWIZZLE_64B at BK = 16); MN-major: 4 groups x (BK k-rows x 128 B)
__host__ __device__ constexpr int tm_a_bytes(int bk) { return TM_BM * bk * 4; }
__host__ __device__ constexpr int tm_stage_bytes(int bn) { return 2 * tm_a_bytes(tm_bk(bn)) + 2 * bn * tm_bk(bn) * 4; }
__host__ __device__ constexpr int tm_stages(int bn) { return bn == 256 ? 4 : (bn == 128 ? 3 : (bn == 64 ? 4 : 5)); }
__host__ __device__ constexpr int tm_tmem_cols(int bn) { return bn == 32 ? 128 : (bn == 64 ? 256 : 512); }
// BN = 256 (the episode-grouped attention GEMMs: one CTA covers all 256 slots / all 256 embedding columns of a tile) leaves
// room for two accumulators only: every hi*hi product goes to the first, the cross terms to the second.  The hi*hi chain
// is then K/8 truncating additions long instead of K/16 (<= 32 at K = 256: ~2e-6 relative), which the attention tolerates.
__host__ __device__ constexpr int tm_accumulators(int bn) { return bn == 256 ? 2 : 3; }

struct TmaGemmParams {
    CUtensorMap ta, tb;          // rank 3: {inner, outer, batch}
    GemmArgs g;
};

// ask the TMA unit to pull a box into L2 (no shared-memory destination, no barrier): issued a few k-blocks ahead of the stage
// ring so that the ring's own loads find their rows in L2 instead of paying the DRAM round trip with only 2-5 stages in flight
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// 32 lanes x 32 consecutive fp32 accumulator columns, no wait (the caller waits once for all three accumulators)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// v = (hi*hi even steps + hi*hi odd steps) + cross terms, for 32 accumulator columns of this warp's 32 lanes
template <int BN>
__device__ __forceinline__ void load_accumulators3(uint32_t taddr, float (&out)[32]) {
    uint32_t v[32], u[32];
    tmem_ld32_nowait(taddr, v);
    tmem_ld32_nowait(taddr + BN, u);
    if (tm_accumulators(BN) == 3) {
        uint32_t x[32];
        tmem_ld32_nowait(taddr + 2 * BN, x);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) out[i] = (__uint_as_float(v[i]) + __uint_as_float(u[i])) + __uint_as_float(x[i]);
    } else {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) out[i] = __uint_as_float(v[i]) + __uint_as_float(u[i]);
    }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TM_THREADS, 1) tma_gemm_kernel(const __grid_constant__ TmaGemmParams p) {
    extern __shared__ __align__(1024) unsigned char tm_smem[];
    constexpr int STAGES = tm_stages(BN);
    constexpr int TM_BK = tm_bk(BN), A_TILE_BYTES = tm_a_bytes(TM_BK), GRP_BYTES = TM_BK * 128;    // GRP: 32 m/n x BK k, MN-major
    constexpr int B_TILE_BYTES = BN * TM_BK * 4;
    constexpr int STAGE_BYTES = tm_stage_bytes(BN);
    unsigned char* tiles = tm_smem + ((1024u - (smem_u32(tm_smem) & 1023u)) & 1023u);       // swizzled layouts need an aligned base
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);              // TMA bytes landed
    uint64_t* ready = full + STAGES;                                                         // hi/lo split done
    uint64_t* empty = ready + STAGES;                                                        // MMAs have consumed the stage
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const GemmArgs& g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int m0 = blockIdx.y * TM_BM, m_end = g.M;
    const int n0 = blockIdx.x * BN;
    const int b = blockIdx.z / g.ksplit, split = blockIdx.z % g.ksplit;
    int b_of_B = b;
    if (g.tiles) {                                    // grouped mode: this CTA's rows and B batch element come from the tile table
        const int4 t = g.tiles[blockIdx.y];
        if (t.y == 0) return;                         // padding entry (tables padded to a fixed length for CUDA-graph replay)
        m0 = t.x; m_end = min(g.M, t.x + t.y); b_of_B = t.z;
    }
    const int k_begin = split * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int kblocks = (k_end - k_begin + TM_BK - 1) / TM_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], TM_CONV_WARPS); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init_fence();
    }
    if (warp == TM_TMA_WARP && lane == 0) { prefetch_tmap(&p.ta); prefetch_tmap(&p.tb); }
    if (warp == TM_MMA_WARP) tmem_alloc(tmem_slot, tm_tmem_cols(BN));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = smem_u32(tiles);

    if (warp == TM_TMA_WARP) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            constexpr int AHEAD = STAGES + 3;             // L2 prefetch distance in k-blocks
            auto prefetch_kb = [&](int kb) {
                const int k0 = k_begin + kb * TM_BK;
                if (A_MN) {
#pragma unroll
                    for (int grp = 0; grp < TM_BM / 32; ++grp) tma_prefetch_3d(&p.ta, m0 + 32 * grp, k0, b);
                } else {
                    tma_prefetch_3d(&p.ta, k0, m0, b);
                }
                if (B_MN) {
#pragma unroll
                    for (int grp = 0; grp < BN / 32; ++grp) tma_prefetch_3d(&p.tb, n0 + 32 * grp, k0, b_of_B);
                } else {
                    tma_prefetch_3d(&p.tb, k0, n0, b_of_B);
                }
            };
            for (int kb = STAGES; kb < min(kblocks, AHEAD); ++kb) prefetch_kb(kb);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                if (kb + AHEAD < kblocks) prefetch_kb(kb + AHEAD);
                mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
                const int k0 = k_begin + kb * TM_BK;
                const uint32_t a_dst = smem_base + s * STAGE_BYTES;
                const uint32_t b_dst = a_dst + 2 * A_TILE_BYTES;
                if ((g.debug & 4) && kb >= STAGES) { mbar_arrive(&full[s]); continue; }      // timing experiment: no TMA traffic
                mbar_arrive_expect_tx(&full[s], A_TILE_BYTES + B_TILE_BYTES);
                if (A_MN) {
#pragma unroll
                    for (int grp = 0; grp < TM_BM / 32; ++grp) tma_load_3d(a_dst + grp * GRP_BYTES, &p.ta, &full[s], m0 + 32 * grp, k0, b);
                } else {
                    tma_load_3d(a_dst, &p.ta, &full[s], k0, m0, b);
                }
                if (B_MN) {
#pragma unroll
                    for (int grp = 0; grp < BN / 32; ++grp) tma_load_3d(b_dst + grp * GRP_BYTES, &p.tb, &full[s], n0 + 32 * grp, k0, b_of_B);
                } else {
                    tma_load_3d(b_dst, &p.tb, &full[s], k0, n0, b_of_B);
                }
            }
        }
    } else if (warp == TM_MMA_WARP) {
        // ---------------- MMA issuer ----------------
        // Descriptors are built once; per stage / k-step only their 14-bit start-address field (16-byte units) advances.
        // K-major SW128: a k-step is 32 bytes further along the 128-byte span; MN-major SW128_BASE32B: 8 k-rows = 1024 bytes.
        constexpr uint32_t idesc = idesc_tf32(TM_BM, BN, A_MN, B_MN);
        constexpr uint32_t A_KSTEP = (A_MN ? 1024 : 32) >> 4, B_KSTEP = (B_MN ? 1024 : 32) >> 4;
        // K-major rows are 4 BK bytes wide: SWIZZLE_128B (8-row groups 1024 B apart) at BK = 32, SWIZZLE_64B (512 B apart) at BK = 16
        constexpr uint32_t KM_LAYOUT = TM_BK == 32 ? LAYOUT_SW128 : LAYOUT_SW64, KM_SBO = TM_BK == 32 ? 1024 : 512;
        const uint64_t da0 = A_MN ? make_desc(smem_base, GRP_BYTES, 512, LAYOUT_SW128_BASE32B) : make_desc(smem_base, 16, KM_SBO, KM_LAYOUT);
        const uint64_t db0 = B_MN ? make_desc(smem_base + 2 * A_TILE_BYTES, GRP_BYTES, 512, LAYOUT_SW128_BASE32B)
                                  : make_desc(smem_base + 2 * A_TILE_BYTES, 16, KM_SBO, KM_LAYOUT);
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&ready[s], (kb / STAGES) & 1);
            fence_after_sync();
            if (lane == 0 && (g.debug & 2)) {                    // timing experiment: no MMAs
                umma_commit(&empty[s]);
                if (kb == kblocks - 1) umma_commit(tmem_full);
            } else if (lane == 0) {
                const uint64_t dah0 = da0 + (uint64_t)((s * STAGE_BYTES) >> 4), dal0 = dah0 + (uint64_t)(A_TILE_BYTES >> 4);
                const uint64_t dbh0 = db0 + (uint64_t)((s * STAGE_BYTES) >> 4), dbl0 = dbh0 + (uint64_t)(B_TILE_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < TM_BK / 8; ++k) {            // one MMA consumes K = 8 tf32
                    const uint64_t dah = dah0 + k * A_KSTEP, dal = dal0 + k * A_KSTEP;
                    const uint64_t dbh = dbh0 + k * B_KSTEP, dbl = dbl0 + k * B_KSTEP;
                    const int step = kb * (TM_BK / 8) + k;
                    if (tm_accumulators(BN) == 3) {
                        umma_tf32(tmem_base + (uint32_t)((step & 1) * BN), dah, dbh, idesc, step >= 2 ? 1u : 0u);
                        umma_tf32(tmem_base + 2 * BN, dal, dbh, idesc, step > 0 ? 1u : 0u);
                        umma_tf32(tmem_base + 2 * BN, dah, dbl, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, dah, dbh, idesc, step > 0 ? 1u : 0u);
                        umma_tf32(tmem_base + BN, dal, dbh, idesc, step > 0 ? 1u : 0u);
                        umma_tf32(tmem_base + BN, dah, dbl, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);                          // frees the stage once these MMAs have read it
                if (kb == kblocks - 1) umma_commit(tmem_full);   // accumulators complete -> epilogue
            }
            __syncwarp();
        }
    } else {
        // ---------------- converters (warps 0-7): raw fp32 tile -> hi (in place) + lo (twin tile) ----------------
        constexpr int A_CHUNKS = A_TILE_BYTES / 16, B_CHUNKS = B_TILE_BYTES / 16;
        constexpr int NT = TM_CONV_WARPS * 32;
        constexpr int PER_THREAD = (A_CHUNKS + B_CHUNKS) / NT;
        static_assert((A_CHUNKS + B_CHUNKS) % NT == 0 && A_CHUNKS % NT == 0, "converter tiling");
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&full[s], (kb / STAGES) & 1);
            unsigned char* a_hi = tiles + s * STAGE_BYTES;
            unsigned char* b_hi = a_hi + 2 * A_TILE_BYTES;
            // hi = x rounded to TF32 (nearest, ties away: add half an ulp of the 10-bit mantissa to the magnitude, clear the low 13
            // bits -- what cvt.rna.tf32.f32 computes for finite values, in 2 integer ops instead of its ~10-instruction expansion);
            // lo = x - hi is exact in fp32 and is stored unrounded: the tensor core only reads its top 19 bits, an error of
            // 2^-11 |lo| <= 2^-22 |x|, the size of the lo*lo term that 3xTF32 drops anyway.
            if (!(g.debug & 1)) {
                float4 v[PER_THREAD];
#pragma unroll
                for (int i = 0; i < PER_THREAD; ++i) {           // all loads of the stage in flight first
                    const int c = threadIdx.x + i * NT;          // A_CHUNKS % NT == 0: chunk i of every thread is on the same side
                    v[i] = *reinterpret_cast<const float4*>((c < A_CHUNKS) ? a_hi + c * 16 : b_hi + (c - A_CHUNKS) * 16);
                }
#pragma unroll
                for (int i = 0; i < PER_THREAD; ++i) {
                    const int c = threadIdx.x + i * NT;
                    unsigned char* hp = (c < A_CHUNKS) ? a_hi + c * 16 : b_hi + (c - A_CHUNKS) * 16;
                    uint4 h;
                    float4 l;
                    h.x = (__float_as_uint(v[i].x) + 0x1000u) & 0xffffe000u; h.y = (__float_as_uint(v[i].y) + 0x1000u) & 0xffffe000u;
                    h.z = (__float_as_uint(v[i].z) + 0x1000u) & 0xffffe000u; h.w = (__float_as_uint(v[i].w) + 0x1000u) & 0xffffe000u;
                    l.x = v[i].x - __uint_as_float(h.x); l.y = v[i].y - __uint_as_float(h.y);
                    l.z = v[i].z - __uint_as_float(h.z); l.w = v[i].w - __uint_as_float(h.w);
                    *reinterpret_cast<uint4*>(hp) = h;
                    *reinterpret_cast<float4*>(hp + ((c < A_CHUNKS) ? A_TILE_BYTES : B_TILE_BYTES)) = l;
                }
            }
            fence_async_proxy();                                 // generic-proxy writes -> visible to the MMAs (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);
        }
        // ---------------- epilogue: TMEM -> registers -> global ----------------
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (rows of the tile) and the 32-column chunks j = w / 4, w / 4 + 2, ...
        mbar_wait(tmem_full, 0);
        fence_after_sync();
        const int quarter = warp & 3;
        const int m = m0 + quarter * 32 + lane;
        // tiles without split-K leave through shared memory as well: the direct epilogue below stores thread-per-row (a warp
        // instruction touches 32 different rows, 16 bytes each) -- ~8 k sector writes per CTA for a 128 x 256 tile; measured on the
        // attention GEMMs: 66 -> 60 us per forward
        const bool coalesced = g.ksplit == 1 && g.ldc % 4 == 0 && g.sC % 4 == 0 && ((((uintptr_t)g.C) & 15) == 0) &&
                               (!g.R || (g.ldr % 4 == 0 && g.sR % 4 == 0 && ((((uintptr_t)g.R) & 15) == 0)));
        if ((BN <= 128 && g.cluster_reduce) || coalesced) {
            // ---- split-K inside a cluster, part 1 / coalesced epilogue: the tile parked in the (now idle) stage ring ----
            constexpr int PSTRIDE = BN + 4;
            static_assert(TM_BM * PSTRIDE * 4 <= STAGES * STAGE_BYTES, "a parked tile must fit the stage ring");
            float* prow = reinterpret_cast<float*>(tiles) + (quarter * 32 + lane) * PSTRIDE;
#pragma unroll 1
            for (int j = warp >> 2; j < BN / 32; j += TM_CONV_WARPS / 4) {
                float v[32];
                load_accumulators3<BN>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 32), v);
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(prow + j * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            if (coalesced) {
                // BN / 4 threads per row: every warp instruction reads / writes contiguous runs of whole rows
                asm volatile("bar.sync 1, %0;" ::"n"(TM_CONV_WARPS * 32) : "memory");       // the eight epilogue warps only
                const float* park = reinterpret_cast<const float*>(tiles);
                constexpr int TPR = BN / 4;                                               // threads per row
                const int c4 = (threadIdx.x & (TPR - 1)) * 4, n = n0 + c4;
                float* __restrict__ C = g.C + (long long)b * g.sC;
                const float* __restrict__ bias = g.bias ? g.bias + (long long)b * g.sBias : nullptr;
                const float* __restrict__ R = g.R ? g.R + (long long)b * g.sR : nullptr;
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (bias) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (n + q < g.N) bv[q] = __ldg(bias + n + q);
                }
                for (int r = threadIdx.x / TPR; r < TM_BM; r += TM_CONV_WARPS * 32 / TPR) {
                    const int mr = m0 + r;
                    if (mr >= m_end) break;
                    if (n >= g.N) continue;
                    const float4 pv = *reinterpret_cast<const float4*>(park + r * PSTRIDE + c4);
                    float x[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        x[q] = fmaf(x[q], g.alpha, bv[q]);
                        if (g.relu) x[q] = fmaxf(x[q], 0.f);
                    }
                    float* dst = C + (long long)mr * g.ldc + n;
                    if (n + 3 < g.N) {
                        if (R) {
                            const float4 rv = *reinterpret_cast<const float4*>(R + (long long)mr * g.ldr + n);
                            x[0] += rv.x; x[1] += rv.y; x[2] += rv.z; x[3] += rv.w;
                        }
                        if (g.accumulate) {
                            const float4 cv = *reinterpret_cast<const float4*>(dst);
                            x[0] += cv.x; x[1] += cv.y; x[2] += cv.z; x[3] += cv.w;
                        }
                        *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (n + q >= g.N) continue;
                            float v = x[q];
                            if (R) v += R[(long long)mr * g.ldr + n + q];
                            if (g.accumulate) v += dst[q];
                            dst[q] = v;
                        }
                    }
                }
            }
        } else {
        float* __restrict__ C = g.C + (long long)b * g.sC;
        const float* __restrict__ bias = g.bias ? g.bias + (long long)b * g.sBias : nullptr;
        const float* __restrict__ R = g.R ? g.R + (long long)b * g.sR : nullptr;
#pragma unroll 1
        for (int j = warp >> 2; j < BN / 32; j += TM_CONV_WARPS / 4) {
            float v[32];
            load_accumulators3<BN>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 32), v);
            const int nb0 = n0 + j * 32;
            if (m >= m_end || nb0 >= g.N) continue;
            float* __restrict__ dst = (g.ksplit > 1) ? g.ws + ((long long)blockIdx.z * g.M + m) * g.N : C + (long long)m * g.ldc;
            const bool full_chunk = nb0 + 32 <= g.N;
            if (g.ksplit == 1) {
                const bool bias_vec = bias && full_chunk && ((((uintptr_t)(bias + nb0)) & 15) == 0);
                if (bias_vec) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + nb0 + i));
                        v[i] = fmaf(v[i], g.alpha, bv.x); v[i + 1] = fmaf(v[i + 1], g.alpha, bv.y);
                        v[i + 2] = fmaf(v[i + 2], g.alpha, bv.z); v[i + 3] = fmaf(v[i + 3], g.alpha, bv.w);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] *= g.alpha;
                        if (bias && nb0 + i < g.N) v[i] += __ldg(bias + nb0 + i);
                    }
                }
                if (g.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
            }
            const bool fast = full_chunk && ((((uintptr_t)(dst + nb0)) & 15) == 0) && (g.ksplit > 1 || (!R && !g.accumulate));
            if (fast) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(dst + nb0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = nb0 + i;
                    if (n >= g.N) continue;
                    float x = v[i];
                    if (g.ksplit == 1) {
                        if (R) x += R[(long long)m * g.ldr + n];
                        if (g.accumulate) x += dst[n];
                    }
                    dst[n] = x;
                }
            }
        }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == TM_MMA_WARP) tmem_dealloc(tmem_base, tm_tmem_cols(BN));
    if constexpr (BN <= 128) if (g.cluster_reduce) {
        // ---- part 2: CTA `split` sums the ksplit partials of rows [split 128 / ksplit, ...) in split order through distributed
        //      shared memory and runs the epilogue on them; thread = (row, 1/8 of the columns): coalesced row stores ----
        cluster_sync_all();
        if (warp < TM_CONV_WARPS) {
            constexpr int PSTRIDE = BN + 4, CPT = BN / 8;
            const int ks = g.ksplit, rows_mine = TM_BM / ks, row_lo = split * rows_mine;
            const int c0 = (threadIdx.x & 7) * CPT;
            const uint32_t park_addr = smem_u32(tiles);
            float* __restrict__ C = g.C + (long long)b * g.sC;
            const float* __restrict__ bias = g.bias ? g.bias + (long long)b * g.sBias : nullptr;
            const float* __restrict__ R = g.R ? g.R + (long long)b * g.sR : nullptr;
            for (int rl = threadIdx.x >> 3; rl < rows_mine; rl += TM_CONV_WARPS * 4) {
                const int r = row_lo + rl, m = m0 + r;
                float x[CPT];
#pragma unroll
                for (int q = 0; q < CPT; ++q) x[q] = 0.f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {             // four peers' loads in flight at a time
                    float4 p4[4][CPT / 4];
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const int src = half * 4 + s4;
#pragma unroll
                        for (int q = 0; q < CPT; q += 4) {
                            p4[s4][q / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (src < ks) p4[s4][q / 4] = ld_peer_f4(park_addr + (uint32_t)((r * PSTRIDE + c0 + q) * 4), (uint32_t)src);
                        }
                    }
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        if (half * 4 + s4 >= ks) break;
#pragma unroll
                        for (int q = 0; q < CPT; q += 4) {
                            x[q] += p4[s4][q / 4].x; x[q + 1] += p4[s4][q / 4].y; x[q + 2] += p4[s4][q / 4].z; x[q + 3] += p4[s4][q / 4].w;
                        }
                    }
                }
                if (m >= m_end) continue;
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    const int n = n0 + c0 + q;
                    if (n >= g.N) continue;
                    float v = x[q] * g.alpha;
                    if (bias) v += __ldg(bias + n);
                    if (g.relu) v = fmaxf(v, 0.f);
                    if (R) v += R[(long long)m * g.ldr + n];
                    if (g.accumulate) v += C[(long long)m * g.ldc + n];
                    C[(long long)m * g.ldc + n] = v;
                }
            }
        }
        cluster_sync_all();                      // nobody leaves while a peer may still read its parked tile
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

struct MapKey {
    const void* base; long long inner, outer, batch, ld, sb; int box_inner, box_outer, mn;
    bool operator==(const MapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && batch == o.batch && ld == o.ld && sb == o.sb &&
               box_inner == o.box_inner && box_outer == o.box_outer && mn == o.mn;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.base);
        for (long long v : {k.inner, k.outer, k.batch, k.ld, k.sb, (long long)k.box_inner, (long long)k.box_outer, (long long)k.mn})
            h = h * 1000003u ^ (size_t)v;
        return h;
    }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mutex;

// rank-3 map over a row-major operand: {inner (contiguous), outer (stride ld floats), batch (stride sb floats)};
// box {box_inner, box_outer, 1}.  mn = 1: MN-major operand (32-wide boxes, SWIZZLE_128B_ATOM_32B), else K-major: SWIZZLE_128B for
// 32-deep k-blocks (128-byte rows), SWIZZLE_64B for 16-deep ones.
int get_map(const float* base, long long inner, long long outer, long long batch, long long ld, long long sb, int box_inner,
            int box_outer, int mn, CUtensorMap* out) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return TRXL_ERR_UNSUPPORTED;
    MapKey key{base, inner, outer, batch, ld, sb, box_inner, box_outer, mn};
    {
        std::lock_guard<std::mutex> lock(g_maps_mutex);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return TRXL_OK; }
    }
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)(batch > 0 ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, batch > 1 ? (cuuint64_t)sb * 4 : (cuuint64_t)ld * 4 * (cuuint64_t)outer};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (box_inner == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return TRXL_ERR_UNSUPPORTED;
    {
        std::lock_guard<std::mutex> lock(g_maps_mutex);
        if (g_maps.size() > 4096) g_maps.clear();          // pointers of freed workspaces would otherwise pile up
        g_maps.emplace(key, m);
    }
    *out = m;
    return TRXL_OK;
}

template <int BN, bool A_MN, bool B_MN>
int launch_tma(const TmaGemmParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)tm_stages(BN) * tm_stage_bytes(BN) + (3 * tm_stages(BN) + 1) * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tma_gemm_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { trxl_set_error("tma_gemm: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
        attr_set = true;
    }
    const GemmArgs& g = p.g;
    dim3 grid(trxl_cdiv(g.N, BN), g.tiles ? g.n_tiles : trxl_cdiv(g.M, TM_BM), g.batch * g.ksplit);
    if (g.cluster_reduce) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(TM_THREADS, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = g.ksplit;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, tma_gemm_kernel<BN, A_MN, B_MN>, p);
        if (le != cudaSuccess) { trxl_set_error("tma_gemm: cluster launch failed: %s", cudaGetErrorString(le)); return TRXL_ERR_CUDA; }
    } else {
        tma_gemm_kernel<BN, A_MN, B_MN><<<grid, TM_THREADS, smem, st>>>(p);
    }
    ++g_trxl_tc_launches;
    TRXL_CHECK_LAUNCH("tma_gemm");
    return TRXL_OK;
}

template <int BN>
int launch_orient(const TmaGemmParams& p, cudaStream_t st) {
    const bool a_mn = !p.g.a_kc, b_mn = !p.g.b_kc;
    if (!a_mn && !b_mn) return launch_tma<BN, false, false>(p, st);
    if (!a_mn && b_mn) return launch_tma<BN, false, true>(p, st);
    if (a_mn && b_mn) return launch_tma<BN, true, true>(p, st);
    return launch_tma<BN, true, false>(p, st);
}

}  // namespace

int trxl_tensor_map(const float* base, long long inner, long long outer, long long batch, long long ld, long long sb, int box_outer,
                    int mn, CUtensorMap* out) {
    return get_map(base, inner, outer, batch, ld, sb, 32, box_outer, mn, out);
}

// Can the TMA path take these operands?  (16-byte aligned bases, leading dimensions / batch strides in whole 16-byte units)
bool trxl_tc_gemm_eligible(const GemmArgs& g) {
    auto ok = [](const void* p, long long ld, long long sb, int batch) {
        return ((uintptr_t)p % 16 == 0) && ld > 0 && (ld % 4 == 0) && (batch <= 1 || (sb > 0 && sb % 4 == 0));
    };
    const int nb = g.tiles ? g.b_batch : g.batch;
    return encode_fn() != nullptr && g.M >= 1 && g.N >= 1 && g.K >= 1 && ok(g.A, g.lda, g.sA, g.batch) && ok(g.B, g.ldb, g.sB, nb);
}

// picks the N tile so that small problems still spread over the SMs; the caller has already chosen g.ksplit / g.k_per_split
int trxl_tc_gemm_tile_n(const GemmArgs& g) {
    static int forced = -1;                      // TRXL_TC_BN=32|64|128 pins the tile (tuning experiments)
    if (forced < 0) { const char* e = getenv("TRXL_TC_BN"); forced = e ? atoi(e) : 0; }
    if (forced == 32 || forced == 64 || forced == 128) return (g.N <= 32) ? 32 : ((g.N <= 64 && forced == 128) ? 64 : forced);
    const long long mt = trxl_cdiv(g.M, TM_BM);
    if (g.N <= 32) return 32;
    const long long t128 = mt * trxl_cdiv(g.N, 128) * g.batch, t64 = mt * trxl_cdiv(g.N, 64) * g.batch;
    // measured on B200 (tools/gemm_bench.py): a tile costs ~5 us of fixed latency + ~0.5 us per k-block whatever its width, so
    // the widest tile that still yields ~100 CTAs wins; below that, more (narrower) CTAs beat wider ones
    if (g.N > 64 && t128 >= 96) return 128;
    if (t64 >= 96 || g.N <= 64) return 64;
    return 32;
}

// Returns TRXL_ERR_UNSUPPORTED (without launching anything) if a tensor map cannot be encoded; the caller then uses the SIMT path.
int trxl_tc_gemm(const GemmArgs& g, int bn, cudaStream_t st) {
    TRXL_CHECK_ARG(!g.cluster_reduce || ((g.ksplit == 2 || g.ksplit == 4 || g.ksplit == 8) && bn <= 128 && !g.tiles),
                   "tc_gemm: cluster split-K takes 2, 4 or 8 splits of a plain GEMM with tiles up to 128 wide");
    TmaGemmParams p;
    p.g = g;
    static int debug = -1;
    if (debug < 0) { const char* e = getenv("TRXL_TC_DEBUG"); debug = e ? atoi(e) : 0; }
    p.g.debug = debug;
    int rc;
    const int bk = tm_bk(bn);
    if (g.a_kc) rc = get_map(g.A, g.K, g.M, g.batch, g.lda, g.sA, bk, TM_BM, 0, &p.ta);   // (M, K) row-major: inner = k
    else rc = get_map(g.A, g.M, g.K, g.batch, g.lda, g.sA, 32, bk, 1, &p.ta);              // (K, M) row-major: inner = m
    if (rc != TRXL_OK) return rc;
    const long long nb = g.tiles ? g.b_batch : g.batch;
    if (g.b_kc) rc = get_map(g.B, g.K, g.N, nb, g.ldb, g.sB, bk, bn, 0, &p.tb);            // (N, K) row-major: inner = k
    else rc = get_map(g.B, g.N, g.K, nb, g.ldb, g.sB, 32, bk, 1, &p.tb);                   // (K, N) row-major: inner = n
    if (rc != TRXL_OK) return rc;
    if (bn == 256) return launch_orient<256>(p, st);
    if (bn == 128) return launch_orient<128>(p, st);
    if (bn == 64) return launch_orient<64>(p, st);
    return launch_orient<32>(p, st);
}
