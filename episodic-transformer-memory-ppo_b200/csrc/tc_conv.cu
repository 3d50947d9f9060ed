// Training path of the CNN encoder (reference model.py:27-38,87-94 and its autograd backward) on the 5th-gen tensor
// cores: every convolution pass is an implicit GEMM issued as tcgen05.mma.kind::tf32 with fp32-grade accuracy (3xTF32).
//
// Why: at the reference's c3 workload the encoder is 58 % of a PPO minibatch step on cuDNN's fp32 SIMT kernels
// (profiles/r1_launches_c3.txt).  The fp32 parity contract (1e-4 against the CPU reference) rules out single-pass TF32,
// so every activation / gradient / weight tensor is kept *pre-split* as hi = tf32(x), lo = tf32(x - hi) (two fp32-sized
// planes written by the producing kernel's epilogue) and each MMA step issues hi*hi + lo*hi + hi*lo into the fp32 TMEM
// accumulator.  Pre-splitting is what makes the producers cheap: operands are already in tensor-core format, so a tile is
// filled with plain 16-byte cp.async copies -- no register staging, no conversion in the main loop.
//
// Kernels (geometry in tc_conv_geom.h; one CTA = 160 threads = 4 producer/epilogue warps + 1 MMA warp):
//   tc_conv_gather_kernel  C[rows, BN] = gather(A)[rows, K] * W[BN, K]^T, both operands K-major in shared memory.
//                          forward convs (bias + ReLU epilogue) and data gradients (ReLU-mask epilogue); the stride-2
//                          data gradient runs as four parity classes.  Out-of-image taps are zero-filled by cp.async.
//   tc_conv_wgrad_kernel   dW[K, BN] = gather(A)[rows, K]^T * dY[rows, BN]: the reduction runs over the GEMM rows, so both
//                          operands are MN-major in shared memory (same 128-byte global runs as the forward gather);
//                          rows are split across CTAs and the partials summed in fixed order (deterministic).
// Pipeline per CTA: STAGES-deep ring; producers cp.async a stage and attach `cp.async.mbarrier.arrive.noinc` to the stage's
// `full` mbarrier (no producer ever waits for its own copies, so all stages stay in flight); the MMA warp waits on `full`,
// runs fence.proxy.async, issues the stage's MMAs and tcgen05.commit's the `empty` mbarrier.
#include "tc_conv.cuh"

#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128, BK = TCG_BK, THREADS = 160;
// STAGES = 2 (<= 48 KB per stage): two CTAs of 4 producer warps per SM for the training-sized grids; STAGES = 4: one CTA per
// SM with three k-blocks in flight and EIGHT producer / epilogue warps for rollout-sized grids (fewer CTAs than SMs, pure
// latency: measured on B200, one producer warp per scheduler needs ~1.2 k clocks to issue a k-block's 24 copies per thread,
// and the four-warp epilogue took 12-16 k clocks of a 36-41 k clock kernel)
// The tensor core adds into the TMEM accumulator with truncation, so the rounding error grows with the length of the
// accumulation chain.  Three accumulators keep it at fp32-SIMT level: hi*hi products alternate between two of them (halving
// the chain that carries the magnitude), the ~2^-11 smaller cross terms go to the third, and the epilogue adds them (RN).
__host__ __device__ constexpr int tmem_cols(int bn) { return bn == 32 ? 128 : (bn == 64 ? 256 : 512); }

struct GatherArgs {
    TcgGather g;
    TcgScatter s;
    const float* a_hi; const float* a_lo;       // gathered image, pre-split
    const long long* sample_index;              // optional: image n of the row grid reads image sample_index[n]
    const float* b_hi; const float* b_lo;       // weights [BN][K], pre-split
    CUtensorMap tb_hi, tb_lo;                   // ... and their tensor maps (box {32 k, BN rows}, SWIZZLE_128B) when use_tma
    int use_tma;
    const float* bias;                          // [BN] or null
    int relu;
    const float* mask;                          // optional image with the output geometry: result kept where mask > 0
    float* out_hi; float* out_lo; float* out_plain;     // any may be null
    float* out_feat; int feat_P;                        // optional: feat[n, c * P + p] = result[(n, p), c] (dense forward rows only)
    int classes;                                        // 1: the N columns are 2x2 parity classes x 32 channels (tcg_plan_dgrad2_merged)
    long long M;
    long long* trace;                                   // debug (TRXL_CONV_TRACE): CTA 0's SM clock at phase boundaries
};

struct WgradArgs {
    TcgGather g;
    const float* a_hi; const float* a_lo;
    const long long* sample_index;
    const float* dy_hi; const float* dy_lo;     // [M][BN], pre-split
    long long M;
    int rows_per_split;                         // multiple of 32; every split is non-empty
    float* partial;                             // [splits][K][BN]
};

__device__ __forceinline__ void stage_taps(const TcgGather& g, int* s_tapoff, int* s_dy, int* s_dx) {
    if (threadIdx.x < g.nkb) {
        const int kb = threadIdx.x;
        s_dy[kb] = g.tap_dy[kb];
        s_dx[kb] = g.tap_dx[kb];
        s_tapoff[kb] = (g.tap_dy[kb] * g.img_w + g.tap_dx[kb]) * g.img_c + g.tap_c0[kb];
    }
}

// v = (hi*hi even steps + hi*hi odd steps) + cross terms, for 32 accumulator columns of this warp's 32 lanes
template <int BN>
__device__ __forceinline__ void load_accumulators(uint32_t taddr, uint32_t (&v)[32]) {
    uint32_t u[32], x[32];
    tmem_ld32(taddr, v);
    tmem_ld32(taddr + BN, u);
    tmem_ld32(taddr + 2 * BN, x);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint((__uint_as_float(v[i]) + __uint_as_float(u[i])) + __uint_as_float(x[i]));
}

// ------------------------------------------------------------------------------------------------------------------
// grid (row tiles, KS); with KS > 1 the launch carries cluster dims (1, KS, 1): CTA (tile, kq) accumulates k-blocks
// [kq nkb / KS, (kq + 1) nkb / KS) of the tile -- rollout-sized problems have fewer row tiles than SMs and are bound by one SM's
// copy-issue rate (measured: ~1.3 k clocks per k-block whatever the number of producer warps), so the k-blocks are spread over
// KS SMs.  Epilogue: every CTA parks its partial tile in its own (now idle) stage memory; after a cluster barrier CTA kq sums
// the KS partials of rows [kq 128 / KS, (kq + 1) 128 / KS) in fixed order through distributed shared memory and finishes
// them: thread = (row, 1/8 of the columns), so a row's 128 / 256 output bytes per plane leave as one coalesced run.
template <int BN, int STAGES, int PW>
__global__ void __launch_bounds__((PW + 1) * 32, STAGES == 2 ? 2 : 1) tc_conv_gather_kernel(const __grid_constant__ GatherArgs a) {
    extern __shared__ __align__(1024) unsigned char tcc_smem[];
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr int PSTRIDE = BN + 4;                          // floats per parked row (the pad keeps float4 accesses conflict-free)
    static_assert(2 * BM * PSTRIDE * 4 <= STAGES * STAGE_BYTES, "a parked tile and the finished copy must fit the stage ring");
    unsigned char* tiles = tcc_smem + ((1024u - (smem_u32(tcc_smem) & 1023u)) & 1023u);      // swizzled layouts need an aligned base
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* park = reinterpret_cast<float*>(tiles);                           // [BM][PSTRIDE] partial tile (the idle stage ring)
    float* fin = park + BM * PSTRIDE;                                        // [rows][PSTRIDE] finished values (out_feat only)
    __shared__ int s_tapoff[TCG_MAX_KB], s_dy[TCG_MAX_KB], s_dx[TCG_MAX_KB];
    __shared__ long long s_rowbase[BM], s_rowdst[BM];       // float offsets of a row's anchor pixel / output pixel (-1: past M)
    __shared__ int s_rowyx[BM];                             // anchor pixel (y << 16 | x)
    __shared__ long long s_rowfeat[BM];                     // out_feat: offset of element (n, c = 0, p) of the row's sample

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * BM;
    const int KS = (int)gridDim.y, kq = (int)blockIdx.y;
    const int kb0 = (kq * a.g.nkb) / KS, kb1 = ((kq + 1) * a.g.nkb) / KS, nkb = kb1 - kb0;
    const bool tracing = a.trace != nullptr && blockIdx.x == 0 && kq == 0;
    if (tracing && threadIdx.x == 0) a.trace[0] = clock64();
    stage_taps(a.g, s_tapoff, s_dy, s_dx);
    if (threadIdx.x < BM) {
        const long long m = m0 + threadIdx.x;
        long long base = -1, dst = -1;
        int ay = 0, ax = 0;
        if (m < a.M) {
            int n, ry, rx;
            tcg_row(a.g, m, n, ry, rx);
            const long long img = a.sample_index ? a.sample_index[n] : n;
            ay = ry * a.g.rstride; ax = rx * a.g.rstride;
            base = ((img * a.g.img_h + ay) * a.g.img_w + ax) * a.g.img_c;
            dst = tcg_dst(a.s, n, ry, rx);
        }
        s_rowbase[threadIdx.x] = base;
        s_rowdst[threadIdx.x] = dst;
        s_rowyx[threadIdx.x] = (ay << 16) | ax;
        if (a.out_feat && dst >= 0) {                   // dense forward rows: dst / BN = n * P + p
            const long long np = dst / BN, fn = np / a.feat_P;
            s_rowfeat[threadIdx.x] = fn * BN * a.feat_P + (np - fn * a.feat_P);
        }
    }
    if (threadIdx.x == 0) {
        // `full`: one cp.async arrival per producer thread, plus (TMA weights) the arrive that announces the tile bytes
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], PW * 32 + (a.use_tma ? 1 : 0)); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init_fence();
        if (a.use_tma) { prefetch_tmap(&a.tb_hi); prefetch_tmap(&a.tb_lo); }
    }
    if (warp == PW) tmem_alloc(tmem_slot, tmem_cols(BN));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = smem_u32(tiles);
    if (tracing && threadIdx.x == 0) a.trace[1] = clock64();

    if (warp < PW) {
        // ---------------- producers ----------------
        // Eight consecutive lanes copy the eight 16-byte pieces of one row's 128-byte run (fully coalesced global reads);
        // thread t owns piece t & 7 of rows (t >> 3) + RSTEP i.  Shared tiles are K-major SWIZZLE_128B: row r at
        // (r / 8) * 1024 + (r % 8) * 128, piece c at position c ^ (r % 8), so the eight lanes of a row hit distinct banks.
        constexpr int RSTEP = PW * 4, RPT = BM / RSTEP;      // rows a pass of all producers covers; rows per thread
        const int piece = threadIdx.x & 7, r0 = threadIdx.x >> 3;
        long long rbase[RPT];
        int ryx[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) { rbase[i] = s_rowbase[r0 + RSTEP * i]; ryx[i] = s_rowyx[r0 + RSTEP * i]; }
        const uint32_t dst0 = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((piece ^ (r0 & 7)) * 16));
        const int K = a.g.nkb * BK;
        for (int it = 0; it < nkb; ++it) {
            const int s = it % STAGES, kb = kb0 + it;
            mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
            const int tdy = s_dy[kb], tdx = s_dx[kb];
            const long long toff = s_tapoff[kb] + piece * 4;
            const uint32_t dst = smem_base + s * STAGE_BYTES + dst0;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int sy = (ryx[i] >> 16) + tdy, sx = (ryx[i] & 0xffff) + tdx;
                const bool ok = rbase[i] >= 0 && sy >= 0 && sy < a.g.img_h && sx >= 0 && sx + a.g.run_px <= a.g.img_w;
                const long long off = ok ? rbase[i] + toff : 0;
                const uint32_t nbytes = ok ? 16u : 0u;
                cp_async16(dst + i * (RSTEP * 128), a.a_hi + off, nbytes);
                cp_async16(dst + A_BYTES + i * (RSTEP * 128), a.a_lo + off, nbytes);
            }
            if (a.use_tma) {
                // the weight tile is a plain 2-D box: one elected thread hands it to the TMA unit (no LSU copy slots -- the
                // gathered rows are bound by the 16-byte copy-issue rate, measured ~13 clocks per warp instruction)
                if (threadIdx.x == 0) {
                    const uint32_t bdst = smem_base + s * STAGE_BYTES + 2 * A_BYTES;
                    mbar_arrive_expect_tx(&full[s], 2 * B_BYTES);
                    tma_load_3d(bdst, &a.tb_hi, &full[s], kb * BK, 0, 0);
                    tma_load_3d(bdst + B_BYTES, &a.tb_lo, &full[s], kb * BK, 0, 0);
                }
            } else {
                const long long boff = (long long)r0 * K + kb * BK + piece * 4;
                const uint32_t dstb = smem_base + s * STAGE_BYTES + 2 * A_BYTES + dst0;
#pragma unroll
                for (int i = 0; i < BN / RSTEP; ++i) {
                    cp_async16(dstb + i * (RSTEP * 128), a.b_hi + boff + (long long)i * RSTEP * K, 16u);
                    cp_async16(dstb + B_BYTES + i * (RSTEP * 128), a.b_lo + boff + (long long)i * RSTEP * K, 16u);
                }
            }
            cp_async_arrive(&full[s]);           // fires when this thread's copies of the stage have landed; nobody waits
            if (tracing && threadIdx.x == 0 && it < 36) a.trace[2 + it] = clock64();
        }
        // ---------------- epilogue, part 1: TMEM -> registers -> this CTA's partial tile, parked in stage 0 ----------------
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (rows of the tile); with eight warps, w / 4 picks the 32-column chunks
        mbar_wait(tmem_full, 0);
        fence_after_sync();
        if (tracing && threadIdx.x == 0) a.trace[80] = clock64();
        const int quarter = warp & 3;
        float* prow = park + (quarter * 32 + lane) * PSTRIDE;
#pragma unroll 1
        for (int j = warp >> 2; j < BN / 32; j += PW / 4) {
            uint32_t v[32];
            load_accumulators<BN>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 32), v);
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(prow + j * 32 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                            __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
    } else {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = idesc_tf32(BM, BN, false, false);
        constexpr uint32_t SBO = 1024;
        for (int it = 0; it < nkb; ++it) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            fence_async_proxy();                     // the landed copies (generic proxy) -> visible to the MMAs this warp issues
            fence_after_sync();
            if (tracing && lane == 0 && it < 36) a.trace[40 + it] = clock64();
            if (lane == 0) {
                const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = a_lo + A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {               // one MMA consumes K = 8 tf32 = 32 bytes of the 128-byte span
                    const uint64_t dah = make_desc(a_hi + k * 32, 16, SBO, LAYOUT_SW128);
                    const uint64_t dal = make_desc(a_lo + k * 32, 16, SBO, LAYOUT_SW128);
                    const uint64_t dbh = make_desc(b_hi + k * 32, 16, SBO, LAYOUT_SW128);
                    const uint64_t dbl = make_desc(b_lo + k * 32, 16, SBO, LAYOUT_SW128);
                    const int step = it * (BK / 8) + k;
                    umma_tf32(tmem_base + (uint32_t)((step & 1) * BN), dah, dbh, idesc, step >= 2 ? 1u : 0u);
                    umma_tf32(tmem_base + 2 * BN, dal, dbh, idesc, step > 0 ? 1u : 0u);
                    umma_tf32(tmem_base + 2 * BN, dah, dbl, idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (it == nkb - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    }
    fence_before_sync();
    __syncthreads();                                 // partial tile parked, TMEM drained
    if (warp == PW) tmem_dealloc(tmem_base, tmem_cols(BN));
    if (KS > 1) cluster_sync_all();                  // every CTA's partial is visible cluster-wide
    if (warp < PW) {
        // ---------------- epilogue, part 2: sum the partials of this CTA's rows, bias / ReLU / mask, pre-split NHWC stores ----------------
        constexpr int NT = PW * 32, CPT = BN / 8;    // a row's columns are shared by 8 threads
        const int rows_mine = BM / KS, row_lo = kq * rows_mine;
        const int g8 = threadIdx.x & 7, c0 = g8 * CPT;
        const bool bias_vec = (((uintptr_t)a.bias) & 15) == 0;
        float bv[CPT];
#pragma unroll
        for (int q = 0; q < CPT; ++q) bv[q] = 0.f;
        if (a.bias) {
            if (bias_vec) {
#pragma unroll
                for (int q = 0; q < CPT; q += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + q));
                    bv[q] = b4.x; bv[q + 1] = b4.y; bv[q + 2] = b4.z; bv[q + 3] = b4.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < CPT; ++q) bv[q] = __ldg(a.bias + c0 + q);
            }
        }
        const uint32_t park_addr = smem_u32(park);
        for (int rl = threadIdx.x >> 3; rl < rows_mine; rl += NT / 8) {
            const int r = row_lo + rl;
            long long dst = s_rowdst[r];
            int oc0 = c0;                                      // first output channel of this thread's columns
            if (a.classes && dst >= 0) {                       // columns = (py, px, channel): the pixel of this thread's class
                const int cls = c0 >> 5, py = cls >> 1, px = cls & 1;
                const int yx = s_rowyx[r];
                if (2 * (yx >> 16) + py >= a.s.out_h || 2 * (yx & 0xffff) + px >= a.s.out_w) dst = -1;
                else dst += ((long long)py * a.s.out_w + px) * a.s.out_c;
                oc0 = c0 & 31;
            }
            // the ReLU-mask values of the row first: their global-memory latency overlaps the partial sums below (issued inside the
            // store loop they serialised into CPT / 4 round trips per row, which a lone CTA per SM cannot hide)
            float4 mk4[CPT / 4];
#pragma unroll
            for (int q = 0; q < CPT; q += 4) {
                mk4[q / 4] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (a.mask && dst >= 0) mk4[q / 4] = __ldg(reinterpret_cast<const float4*>(a.mask + dst + oc0 + q));
            }
            float x[CPT];
#pragma unroll
            for (int q = 0; q < CPT; ++q) x[q] = 0.f;
            // all peers' loads in flight together; summed in CTA order (the result does not depend on which CTA finishes first)
            float4 p4[4][CPT / 4];
#pragma unroll
            for (int src = 0; src < 4; ++src) {
#pragma unroll
                for (int q = 0; q < CPT; q += 4) {
                    const uint32_t off = (uint32_t)((r * PSTRIDE + c0 + q) * 4);
                    p4[src][q / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (KS > 1) { if (src < KS) p4[src][q / 4] = ld_peer_f4(park_addr + off, (uint32_t)src); }
                    else if (src == 0) p4[src][q / 4] = *reinterpret_cast<const float4*>(park + r * PSTRIDE + c0 + q);
                }
            }
#pragma unroll
            for (int src = 0; src < 4; ++src) {
                if (src >= KS) break;
#pragma unroll
                for (int q = 0; q < CPT; q += 4) {
                    x[q] += p4[src][q / 4].x; x[q + 1] += p4[src][q / 4].y; x[q + 2] += p4[src][q / 4].z; x[q + 3] += p4[src][q / 4].w;
                }
            }
            float hi[CPT], lo[CPT];
#pragma unroll
            for (int q = 0; q < CPT; q += 4) {
                const float4 mk = mk4[q / 4];
                const float mkv[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float t = x[q + e] + bv[q + e];
                    if (a.relu) t = fmaxf(t, 0.f);
                    if (!(mkv[e] > 0.f)) t = 0.f;
                    x[q + e] = t;
                    split_tf32(t, hi[q + e], lo[q + e]);
                }
                if (dst >= 0) {
                    if (a.out_hi) *reinterpret_cast<float4*>(a.out_hi + dst + oc0 + q) = make_float4(hi[q], hi[q + 1], hi[q + 2], hi[q + 3]);
                    if (a.out_lo) *reinterpret_cast<float4*>(a.out_lo + dst + oc0 + q) = make_float4(lo[q], lo[q + 1], lo[q + 2], lo[q + 3]);
                    if (a.out_plain) *reinterpret_cast<float4*>(a.out_plain + dst + oc0 + q) = make_float4(x[q], x[q + 1], x[q + 2], x[q + 3]);
                }
                if (a.out_feat) *reinterpret_cast<float4*>(fin + rl * PSTRIDE + c0 + q) = make_float4(x[q], x[q + 1], x[q + 2], x[q + 3]);
            }
        }
        if (a.out_feat) {
            // the reference flattens NCHW (model.py:94): feat[n, c * P + p] = result[(n, p), c]; consecutive lanes take consecutive
            // rows (= consecutive p within a sample), so every store instruction writes runs of the feature row
            asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory");       // the producer / epilogue warps only
            for (int idx = threadIdx.x; idx < rows_mine * BN; idx += NT) {
                const int rl = idx % rows_mine, c = idx / rows_mine;
                if (s_rowdst[row_lo + rl] < 0) continue;
                a.out_feat[s_rowfeat[row_lo + rl] + (long long)c * a.feat_P] = fin[rl * PSTRIDE + c];
            }
        }
    }
    if (tracing && threadIdx.x == 0) a.trace[81] = clock64();
    if (KS > 1) cluster_sync_all();                  // nobody leaves while a peer may still read its parked tile
}

// ------------------------------------------------------------------------------------------------------------------
// grid (K tiles of 128, row splits).  Shared-memory tiles are MN-major in the SWIZZLE_128B_BASE32B layout (the only one
// TF32 has for MN-major): [32-element M/N group][row / 4][row % 4][128 bytes], 32-byte chunks XOR-swizzled with row % 4.
// One 128-byte global run (32 consecutive k of one row, or 32 consecutive channels of one dY row) is one shared row.
template <int BN, int STAGES>
__global__ void __launch_bounds__(THREADS, 2) tc_conv_wgrad_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ __align__(1024) unsigned char tcc_smem[];
    constexpr int A_BYTES = 128 * 32 * 4, B_BYTES = BN * 32 * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    unsigned char* tiles = tcc_smem + ((1024u - (smem_u32(tcc_smem) & 1023u)) & 1023u);      // swizzled layouts need an aligned base
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    __shared__ int s_tapoff[TCG_MAX_KB], s_dy[TCG_MAX_KB], s_dx[TCG_MAX_KB];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x, split = blockIdx.y;
    const long long m_begin = (long long)split * a.rows_per_split;
    const long long m_end = (m_begin + a.rows_per_split < a.M) ? m_begin + a.rows_per_split : a.M;
    const int nst = (int)((m_end - m_begin + 31) / 32);
    const int K = a.g.nkb * BK;
    stage_taps(a.g, s_tapoff, s_dy, s_dx);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init_fence();
    }
    if (warp == 4) tmem_alloc(tmem_slot, tmem_cols(BN));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = smem_u32(tiles);

    if (warp < 4) {
        // ---------------- producers ----------------
        // Eight consecutive lanes copy one 128-byte run (a row's 32 consecutive k, or 32 channels of its dY row); thread t owns
        // piece t & 7 of rows ml0 = t >> 3 and ml0 + 16 of every 32-row stage, for all four runs of the K tile.
        const int piece = threadIdx.x & 7, ml0 = threadIdx.x >> 3;
        int tdy[4], tdx[4], toff[4];
        bool kb_ok[4];
#pragma unroll
        for (int run = 0; run < 4; ++run) {
            const int kb = kt * 4 + run;
            kb_ok[run] = kb < a.g.nkb;
            tdy[run] = kb_ok[run] ? s_dy[kb] : 0; tdx[run] = kb_ok[run] ? s_dx[kb] : 0; toff[run] = kb_ok[run] ? s_tapoff[kb] : 0;
        }
        // row iterators (advance 32 rows per stage without divisions)
        long long rm[2];
        int rn[2], rry[2], rrx[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            rm[h] = m_begin + ml0 + 16 * h;
            tcg_row(a.g, rm[h], rn[h], rry[h], rrx[h]);
        }
        const uint32_t sw = (uint32_t)(ml0 & 3);
        const uint32_t off0 = (uint32_t)((ml0 >> 2) * 512 + (ml0 & 3) * 128 + (((uint32_t)(piece >> 1) ^ sw) * 32) + (piece & 1) * 16);
        for (int it = 0; it < nst; ++it) {
            {
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                const uint32_t dsta = smem_base + s * STAGE_BYTES + off0;
                const uint32_t dstb = dsta + 2 * A_BYTES;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool row_ok = rm[h] < m_end;
                    const long long img = (row_ok && a.sample_index) ? a.sample_index[rn[h]] : rn[h];
                    const int ay = rry[h] * a.g.rstride, ax = rrx[h] * a.g.rstride;
                    const long long base = ((img * a.g.img_h + ay) * a.g.img_w + ax) * a.g.img_c + piece * 4;
#pragma unroll
                    for (int run = 0; run < 4; ++run) {
                        const int sy = ay + tdy[run], sx = ax + tdx[run];
                        const bool ok = row_ok && kb_ok[run] && sy >= 0 && sy < a.g.img_h && sx >= 0 && sx + a.g.run_px <= a.g.img_w;
                        const long long off = ok ? base + toff[run] : 0;
                        const uint32_t nbytes = ok ? 16u : 0u;
                        cp_async16(dsta + h * 2048 + run * 4096, a.a_hi + off, nbytes);
                        cp_async16(dsta + A_BYTES + h * 2048 + run * 4096, a.a_lo + off, nbytes);
                    }
                    const long long doff = row_ok ? rm[h] * BN + piece * 4 : 0;
                    const uint32_t dbytes = row_ok ? 16u : 0u;
#pragma unroll
                    for (int grp = 0; grp < BN / 32; ++grp) {
                        cp_async16(dstb + h * 2048 + grp * 4096, a.dy_hi + doff + grp * 32, dbytes);
                        cp_async16(dstb + B_BYTES + h * 2048 + grp * 4096, a.dy_lo + doff + grp * 32, dbytes);
                    }
                    rm[h] += 32;
                    rrx[h] += 32;
                    while (rrx[h] >= a.g.rw) { rrx[h] -= a.g.rw; ++rry[h]; }
                    while (rry[h] >= a.g.rh) { rry[h] -= a.g.rh; ++rn[h]; }
                }
                cp_async_arrive(&full[s]);
            }
        }
        // ---------------- epilogue: this split's partial dW tile ----------------
        mbar_wait(tmem_full, 0);
        fence_after_sync();
        const int k = kt * 128 + warp * 32 + lane;
        float* dst = a.partial + ((long long)split * K + k) * BN;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t v[32];
            load_accumulators<BN>(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
            if (k >= K) continue;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dst + j * 32 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                           __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
    } else {
        constexpr uint32_t idesc = idesc_tf32(128, BN, true, true);
        constexpr uint32_t LBO = 4096, SBO = 512, STEP = 1024;     // M/N group stride, K atom stride, 8 rows per MMA
        for (int st = 0; st < nst; ++st) {
            const int s = st % STAGES;
            mbar_wait(&full[s], (st / STAGES) & 1);
            fence_async_proxy();
            fence_after_sync();
            if (lane == 0) {
                const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = a_lo + A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int g = 0; g < 4; ++g) {                    // 8 rows of the stage per MMA
                    const uint64_t dah = make_desc(a_hi + g * STEP, LBO, SBO, LAYOUT_SW128_BASE32B);
                    const uint64_t dal = make_desc(a_lo + g * STEP, LBO, SBO, LAYOUT_SW128_BASE32B);
                    const uint64_t dbh = make_desc(b_hi + g * STEP, LBO, SBO, LAYOUT_SW128_BASE32B);
                    const uint64_t dbl = make_desc(b_lo + g * STEP, LBO, SBO, LAYOUT_SW128_BASE32B);
                    const int step = st * 4 + g;
                    umma_tf32(tmem_base + (uint32_t)((step & 1) * BN), dah, dbh, idesc, step >= 2 ? 1u : 0u);
                    umma_tf32(tmem_base + 2 * BN, dal, dbh, idesc, step > 0 ? 1u : 0u);
                    umma_tf32(tmem_base + 2 * BN, dah, dbl, idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (st == nst - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, tmem_cols(BN));
}

// ------------------------------------------------------------------------------------------------------------------
// small elementwise kernels around the GEMMs

// observation rows (NCHW, selected through sample_index) -> NHWC padded to 4 channels, pre-split
__global__ void obs_to_nhwc_split_kernel(const float* __restrict__ obs, const long long* __restrict__ sample_index, long long n,
                                         int C, int H, int W, float* __restrict__ hi, float* __restrict__ lo) {
    const long long total = n * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long long img = i / ((long long)W * H);
        const long long src = sample_index ? sample_index[img] : img;
        float h[4], l[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v = c < C ? obs[((src * C + c) * H + y) * W + x] : 0.f;
            split_tf32(v, h[c], l[c]);
        }
        *reinterpret_cast<float4*>(hi + i * 4) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(lo + i * 4) = make_float4(l[0], l[1], l[2], l[3]);
    }
}
// kind 0: forward weights of `layer`; 1: conv3 data-gradient weights; 2: conv2 data-gradient weights of class (py, px)
__global__ void pack_weights_kernel(const float* __restrict__ w, int kind, int layer, int py, int px, int C, int rows, int K,
                                    float* __restrict__ hi, float* __restrict__ lo) {
    const int total = rows * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int row = i / K, k = i - row * K;
        const int idx = kind == 0 ? tcg_wfwd_index(layer, C, row, k) : (kind == 1 ? tcg_wdgrad3_index(row, k) : tcg_wdgrad2_index(py, px, row, k));
        const float v = idx >= 0 ? w[idx] : 0.f;
        float h, l;
        split_tf32(v, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}
// dy3[(n, p), c] = dfeat[n, c*P + p] where the conv3 output was positive (ReLU backward), pre-split
__global__ void dfeat_to_dy3_kernel(const float* __restrict__ dfeat, const float* __restrict__ y3_hi, long long n, int P,
                                    float* __restrict__ hi, float* __restrict__ lo) {
    const long long total = n * P * 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % 64), p = (int)((i / 64) % P);
        const long long img = i / ((long long)P * 64);
        const float v = y3_hi[i] > 0.f ? dfeat[(img * 64 + c) * P + p] : 0.f;
        float h, l;
        split_tf32(v, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}
// bias gradients: column sums of a pre-split [M][BN] matrix, two deterministic stages
template <int BN>
__global__ void colsum_split_partial_kernel(const float* __restrict__ hi, const float* __restrict__ lo, long long M, int rows_per_block,
                                            float* __restrict__ partial) {
    constexpr int LANES = 256 / BN;
    __shared__ float s[256];
    const int col = threadIdx.x % BN, rl = threadIdx.x / BN;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
    float acc = 0.f;
    for (long long r = r0 + rl; r < r1; r += LANES) acc += hi[r * BN + col] + lo[r * BN + col];
    s[threadIdx.x] = acc;
    __syncthreads();
    if (rl == 0) {
        float t = 0.f;
        for (int q = 0; q < LANES; ++q) t += s[q * BN + col];
        partial[(long long)blockIdx.x * BN + col] = t;
    }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int blocks, int BN, float* __restrict__ out) {
    const int col = threadIdx.x;
    if (col >= BN) return;
    float t = 0.f;
    for (int b = 0; b < blocks; ++b) t += partial[(long long)b * BN + col];
    out[col] = t;
}
// dW (reference layout) = sum over the row splits of the partial [K][BN] tiles.  A block owns 32 consecutive outputs; its eight
// warps sum splits w, w + 8, ... (coalesced rows of the partial tiles) and the eight partial sums are added in warp order
// (deterministic; conv1 has ~800 splits, which a single thread per output used to walk alone)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int K, int BN, int layer, int C, float* __restrict__ dw) {
    __shared__ float s_part[8][32];
    const int total = K * BN, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 32 + lane;
    float t = 0.f;
    if (i < total)
        for (int s = warp; s < splits; s += 8) t += partial[(long long)s * total + i];
    s_part[warp][lane] = t;
    __syncthreads();
    if (warp == 0 && i < total) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_part[w][lane];
        const int k = i / BN, oc = i - k * BN;
        const int idx = tcg_wfwd_index(layer, C, oc, k);
        if (idx >= 0) dw[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
int grid_for(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}
long long align64(long long v) { return (v + 63) / 64 * 64; }

constexpr int COLSUM_BLOCKS = 592;

struct Ws {
    float *x0[2], *y1[2], *y2[2], *y3[2], *d3[2], *d2[2], *d1[2];
    float *wp1[2], *wp2[2], *wp3[2], *wd3[2], *wd2[2];
    float *partial, *colpart;
    long long partial_floats, total;
};
void plan_split(long long M, int ktiles, int* rows_per_split, int* splits) {
    long long target = 296 / ktiles;
    if (target < 1) target = 1;
    long long rps = (M + target - 1) / target;
    rps = (rps + 31) / 32 * 32;
    if (rps < 32) rps = 32;
    if (rps > 1024) rps = 1024;      // bounds the TMEM accumulation chain (the tensor core accumulates with truncation)
    *rows_per_split = (int)rps;
    *splits = (int)((M + rps - 1) / rps);
}
void carve(const TcgEncoder& e, long long n, float* base, Ws& w) {
    long long off = 0;
    auto take = [&](long long floats) { float* p = base ? base + off : nullptr; off += align64(floats); return p; };
    const long long m1 = n * e.h1 * e.w1, m2 = n * e.h2 * e.w2, m3 = n * e.h3 * e.w3;
    for (int i = 0; i < 2; ++i) w.x0[i] = take(n * e.H * e.W * 4);
    for (int i = 0; i < 2; ++i) w.y1[i] = take(m1 * 32);
    for (int i = 0; i < 2; ++i) w.y2[i] = take(m2 * 64);
    for (int i = 0; i < 2; ++i) w.y3[i] = take(m3 * 64);
    for (int i = 0; i < 2; ++i) w.d3[i] = take(m3 * 64);
    for (int i = 0; i < 2; ++i) w.d2[i] = take(m2 * 64);
    for (int i = 0; i < 2; ++i) w.d1[i] = take(m1 * 32);
    for (int i = 0; i < 2; ++i) w.wp1[i] = take(32 * 256);
    for (int i = 0; i < 2; ++i) w.wp2[i] = take(64 * 512);
    for (int i = 0; i < 2; ++i) w.wp3[i] = take(64 * 576);
    for (int i = 0; i < 2; ++i) w.wd3[i] = take(64 * 576);
    for (int i = 0; i < 2; ++i) w.wd2[i] = take(4 * 32 * 256);
    long long pf = 0;
    const long long Ms[3] = {m1, m2, m3};
    const int Ks[3] = {256, 512, 576}, Ns[3] = {32, 64, 64};
    for (int l = 0; l < 3; ++l) {
        int rps, splits;
        plan_split(Ms[l] > 0 ? Ms[l] : 1, (Ks[l] + 127) / 128, &rps, &splits);
        const long long f = (long long)splits * Ks[l] * Ns[l];
        if (f > pf) pf = f;
    }
    w.partial_floats = pf;
    w.partial = take(pf);
    w.colpart = take(COLSUM_BLOCKS * 64);
    w.total = off;
}

template <int BN, int STAGES>
int launch_gather_stages(cudaStream_t st, const GatherArgs& a) {
    constexpr size_t smem = (size_t)STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + (2 * STAGES + 1) * 8 + 16 + 1024;
    constexpr int PW = STAGES == 2 ? 4 : 8;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_gather_kernel<BN, STAGES, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { trxl_set_error("tc_conv: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
        attr_set = true;
    }
    // TRXL_CONV_TRACE present at the first call: later launches of at most one wave whose current value is > 0 print CTA 0's
    // phase clocks (synchronises the stream; debugging only)
    static int trace_armed = -1;
    static long long* trace_dev = nullptr;
    if (trace_armed < 0) trace_armed = getenv("TRXL_CONV_TRACE") ? 1 : 0;
    bool tracing = false;
    if (trace_armed && trxl_cdiv(a.M, BM) <= 148) {
        const char* e = getenv("TRXL_CONV_TRACE");
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        tracing = e && atoi(e) > 0 && cap == cudaStreamCaptureStatusNone;
    }
    GatherArgs b = a;
    {
        static int tma_off = -1;                 // TRXL_CONV_TMA=0: weight tiles through cp.async like the gathered rows
        if (tma_off < 0) { const char* e = getenv("TRXL_CONV_TMA"); tma_off = (e && atoi(e) == 0) ? 1 : 0; }
        const long long K = (long long)a.g.nkb * BK;
        b.use_tma = !tma_off && trxl_tensor_map(a.b_hi, K, BN, 1, K, 0, BN, 0, &b.tb_hi) == TRXL_OK &&
                    trxl_tensor_map(a.b_lo, K, BN, 1, K, 0, BN, 0, &b.tb_lo) == TRXL_OK;
    }
    if (tracing) {
        if (!trace_dev) cudaMalloc(&trace_dev, 96 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, 96 * sizeof(long long), st);
        b.trace = trace_dev;
    }
    // rollout-sized grids: spread a tile's k-blocks over a cluster of KS CTAs so that ~half the SMs or more pull operands
    const int tiles_m = (int)trxl_cdiv(a.M, BM);
    int KS = 1;
    if (STAGES == 4) {
        static int forced = -1;                  // TRXL_CONV_KSPLIT=1|2|4 pins the split (tuning experiments)
        if (forced < 0) { const char* e = getenv("TRXL_CONV_KSPLIT"); forced = e ? atoi(e) : 0; }
        if (forced == 1 || forced == 2 || forced == 4) KS = forced;       // (the kernel sums at most 4 partials)
        else KS = (tiles_m * 4 <= 148 && a.g.nkb >= 8) ? 4 : ((tiles_m * 2 <= 148 && a.g.nkb >= 4) ? 2 : 1);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles_m, KS, 1);
    cfg.blockDim = dim3((PW + 1) * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = KS; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = KS > 1 ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, tc_conv_gather_kernel<BN, STAGES, PW>, b);
    if (le != cudaSuccess) { trxl_set_error("tc_conv_gather: launch failed: %s", cudaGetErrorString(le)); return TRXL_ERR_CUDA; }
    TRXL_CHECK_LAUNCH("tc_conv_gather");
    if (tracing) {
        long long h[96];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[conv-trace] BN=%d stages=%d M=%lld nkb=%d ctas=%d ksplit=%d: setup %lld, epilogue at %lld, end %lld clk\n", BN, STAGES,
                a.M, a.g.nkb, tiles_m, KS, h[1] - h[0], h[80] - h[0], h[81] - h[0]);
        fprintf(stderr, "[conv-trace]   producer issue:");
        for (int i = 0; i < a.g.nkb && i < 36; ++i) fprintf(stderr, " %lld", h[2 + i] - h[0]);
        fprintf(stderr, "\n[conv-trace]   stage landed: ");
        for (int i = 0; i < a.g.nkb && i < 36; ++i) fprintf(stderr, " %lld", h[40 + i] - h[0]);
        fprintf(stderr, "\n");
    }
    return TRXL_OK;
}
template <int BN>
int launch_gather(cudaStream_t st, const GatherArgs& a) {
    if (a.M <= 0) return TRXL_OK;
    if constexpr (BN == 128) return launch_gather_stages<BN, 3>(st, a);      // 64 KB stages: one CTA per SM, ring of three
    else return trxl_cdiv(a.M, BM) <= 148 ? launch_gather_stages<BN, 4>(st, a) : launch_gather_stages<BN, 2>(st, a);
}
template <int BN>
int launch_wgrad(cudaStream_t st, WgradArgs a, int layer, int C, float* dw, const Ws& w) {
    constexpr int STAGES = 2;
    constexpr size_t smem = (size_t)STAGES * (2 * 128 * 32 * 4 + 2 * BN * 32 * 4) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { trxl_set_error("tc_conv: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return TRXL_ERR_CUDA; }
        attr_set = true;
    }
    const int K = a.g.nkb * BK, ktiles = (K + 127) / 128;
    int splits;
    plan_split(a.M, ktiles, &a.rows_per_split, &splits);
    TRXL_CHECK_ARG((long long)splits * K * BN <= w.partial_floats, "tc_conv: wgrad partial buffer too small");
    a.partial = w.partial;
    tc_conv_wgrad_kernel<BN, STAGES><<<dim3(ktiles, splits), THREADS, smem, st>>>(a);
    TRXL_CHECK_LAUNCH("tc_conv_wgrad");
    wgrad_reduce_kernel<<<trxl_cdiv(K * BN, 32), 256, 0, st>>>(w.partial, splits, K, BN, layer, C, dw);
    TRXL_CHECK_LAUNCH("wgrad_reduce");
    return TRXL_OK;
}
template <int BN>
int launch_colsum(cudaStream_t st, const float* hi, const float* lo, long long M, float* colpart, float* out) {
    int blocks = (int)((M + 255) / 256);
    if (blocks > COLSUM_BLOCKS) blocks = COLSUM_BLOCKS;
    if (blocks < 1) blocks = 1;
    const int rpb = (int)((M + blocks - 1) / blocks);
    blocks = (int)((M + rpb - 1) / rpb);
    colsum_split_partial_kernel<BN><<<blocks, 256, 0, st>>>(hi, lo, M, rpb, colpart);
    TRXL_CHECK_LAUNCH("colsum_split_partial");
    colsum_final_kernel<<<1, 64, 0, st>>>(colpart, blocks, BN, out);
    TRXL_CHECK_LAUNCH("colsum_final");
    return TRXL_OK;
}

}  // namespace

int tc_conv_supported(int C, int H, int W) {
    TcgEncoder e;
    return tcg_encoder(C, H, W, e) ? 1 : 0;
}

long long tc_conv_workspace_floats(int N, int C, int H, int W) {
    TcgEncoder e;
    if (!tcg_encoder(C, H, W, e) || N < 0) return -1;
    Ws w;
    carve(e, N, nullptr, w);
    return w.total;
}

int tc_conv_pack_weights(cudaStream_t st, const float* const* p, int N, int C, int H, int W, float* ws) {
    TcgEncoder e;
    TRXL_CHECK_ARG(tcg_encoder(C, H, W, e), "tc_conv: unsupported observation shape (%d, %d, %d)", C, H, W);
    Ws w;
    carve(e, N, ws, w);
    // weights in tensor-core format (forward + data-gradient forms; the backward pass reuses them)
    pack_weights_kernel<<<grid_for(32 * 256), 256, 0, st>>>(p[0], 0, 1, 0, 0, C, 32, 256, w.wp1[0], w.wp1[1]);
    TRXL_CHECK_LAUNCH("pack_weights");
    pack_weights_kernel<<<grid_for(64 * 512), 256, 0, st>>>(p[2], 0, 2, 0, 0, C, 64, 512, w.wp2[0], w.wp2[1]);
    TRXL_CHECK_LAUNCH("pack_weights");
    pack_weights_kernel<<<grid_for(64 * 576), 256, 0, st>>>(p[4], 0, 3, 0, 0, C, 64, 576, w.wp3[0], w.wp3[1]);
    TRXL_CHECK_LAUNCH("pack_weights");
    pack_weights_kernel<<<grid_for(64 * 576), 256, 0, st>>>(p[4], 1, 3, 0, 0, C, 64, 576, w.wd3[0], w.wd3[1]);
    TRXL_CHECK_LAUNCH("pack_weights");
    for (int cls = 0; cls < 4; ++cls) {
        pack_weights_kernel<<<grid_for(32 * 256), 256, 0, st>>>(p[2], 2, 2, cls >> 1, cls & 1, C, 32, 256, w.wd2[0] + cls * 32 * 256,
                                                                w.wd2[1] + cls * 32 * 256);
        TRXL_CHECK_LAUNCH("pack_weights");
    }
    return TRXL_OK;
}

int tc_conv_forward(cudaStream_t st, const float* const* p, const float* obs, const long long* sample_index, int N, int C, int H,
                    int W, float* ws, float* feat, int repack) {
    TcgEncoder e;
    TRXL_CHECK_ARG(tcg_encoder(C, H, W, e), "tc_conv: unsupported observation shape (%d, %d, %d)", C, H, W);
    if (N == 0) return TRXL_OK;
    Ws w;
    carve(e, N, ws, w);
    const long long m1 = (long long)N * e.h1 * e.w1, m2 = (long long)N * e.h2 * e.w2, m3 = (long long)N * e.h3 * e.w3;
    if (repack) TRXL_PROPAGATE(tc_conv_pack_weights(st, p, N, C, H, W, ws));
    obs_to_nhwc_split_kernel<<<grid_for((long long)N * H * W), 256, 0, st>>>(obs, sample_index, N, C, H, W, w.x0[0], w.x0[1]);
    TRXL_CHECK_LAUNCH("obs_to_nhwc_split");

    GatherArgs a{};
    tcg_plan_forward(e, 1, a.g, a.s);
    a.a_hi = w.x0[0]; a.a_lo = w.x0[1]; a.b_hi = w.wp1[0]; a.b_lo = w.wp1[1]; a.bias = p[1]; a.relu = 1;
    a.out_hi = w.y1[0]; a.out_lo = w.y1[1]; a.M = m1;
    TRXL_PROPAGATE(launch_gather<32>(st, a));
    a = GatherArgs{};
    tcg_plan_forward(e, 2, a.g, a.s);
    a.a_hi = w.y1[0]; a.a_lo = w.y1[1]; a.b_hi = w.wp2[0]; a.b_lo = w.wp2[1]; a.bias = p[3]; a.relu = 1;
    a.out_hi = w.y2[0]; a.out_lo = w.y2[1]; a.M = m2;
    TRXL_PROPAGATE(launch_gather<64>(st, a));
    a = GatherArgs{};
    tcg_plan_forward(e, 3, a.g, a.s);
    a.a_hi = w.y2[0]; a.a_lo = w.y2[1]; a.b_hi = w.wp3[0]; a.b_lo = w.wp3[1]; a.bias = p[5]; a.relu = 1;
    a.out_hi = w.y3[0]; a.out_lo = w.y3[1]; a.out_feat = feat; a.feat_P = e.h3 * e.w3; a.M = m3;     // flatten in the epilogue
    TRXL_PROPAGATE(launch_gather<64>(st, a));
    return TRXL_OK;
}

int tc_conv_backward(cudaStream_t st, float* const* g, int N, int C, int H, int W, float* ws, const float* dfeat) {
    TcgEncoder e;
    TRXL_CHECK_ARG(tcg_encoder(C, H, W, e), "tc_conv: unsupported observation shape (%d, %d, %d)", C, H, W);
    if (N == 0) return TRXL_OK;
    Ws w;
    carve(e, N, ws, w);
    const long long m1 = (long long)N * e.h1 * e.w1, m2 = (long long)N * e.h2 * e.w2, m3 = (long long)N * e.h3 * e.w3;
    // ---- layer 3 ----
    dfeat_to_dy3_kernel<<<grid_for(m3 * 64), 256, 0, st>>>(dfeat, w.y3[0], N, e.h3 * e.w3, w.d3[0], w.d3[1]);
    TRXL_CHECK_LAUNCH("dfeat_to_dy3");
    TRXL_PROPAGATE(launch_colsum<64>(st, w.d3[0], w.d3[1], m3, w.colpart, g[5]));
    WgradArgs wa{};
    TcgScatter unused;
    tcg_plan_forward(e, 3, wa.g, unused);
    wa.a_hi = w.y2[0]; wa.a_lo = w.y2[1]; wa.dy_hi = w.d3[0]; wa.dy_lo = w.d3[1]; wa.M = m3;
    TRXL_PROPAGATE(launch_wgrad<64>(st, wa, 3, C, g[4], w));
    GatherArgs a{};
    tcg_plan_dgrad3(e, a.g, a.s);
    a.a_hi = w.d3[0]; a.a_lo = w.d3[1]; a.b_hi = w.wd3[0]; a.b_lo = w.wd3[1]; a.mask = w.y2[0];
    a.out_hi = w.d2[0]; a.out_lo = w.d2[1]; a.M = m2;
    TRXL_PROPAGATE(launch_gather<64>(st, a));
    // ---- layer 2 ----
    TRXL_PROPAGATE(launch_colsum<64>(st, w.d2[0], w.d2[1], m2, w.colpart, g[3]));
    wa = WgradArgs{};
    tcg_plan_forward(e, 2, wa.g, unused);
    wa.a_hi = w.y1[0]; wa.a_lo = w.y1[1]; wa.dy_hi = w.d2[0]; wa.dy_lo = w.d2[1]; wa.M = m2;
    TRXL_PROPAGATE(launch_wgrad<64>(st, wa, 2, C, g[2], w));
    static int per_class = -1;                  // TRXL_CONV_DGRAD2_CLASSES=1: one launch per parity class (A/B timing)
    if (per_class < 0) { const char* ev = getenv("TRXL_CONV_DGRAD2_CLASSES"); per_class = (ev && atoi(ev) > 0) ? 1 : 0; }
    if (!per_class) {
        // the four parity classes gather the same 2x2 neighbourhood: one GEMM, weights stacked along N (4 x 32 columns)
        a = GatherArgs{};
        tcg_plan_dgrad2_merged(e, a.g, a.s);
        a.classes = 1;
        a.a_hi = w.d2[0]; a.a_lo = w.d2[1]; a.b_hi = w.wd2[0]; a.b_lo = w.wd2[1]; a.mask = w.y1[0];
        a.out_hi = w.d1[0]; a.out_lo = w.d1[1]; a.M = (long long)N * a.g.rh * a.g.rw;
        TRXL_PROPAGATE(launch_gather<128>(st, a));
    }
    for (int cls = 0; cls < 4 && per_class; ++cls) {
        a = GatherArgs{};
        tcg_plan_dgrad2(e, cls >> 1, cls & 1, a.g, a.s);
        a.a_hi = w.d2[0]; a.a_lo = w.d2[1]; a.b_hi = w.wd2[0] + cls * 32 * 256; a.b_lo = w.wd2[1] + cls * 32 * 256; a.mask = w.y1[0];
        a.out_hi = w.d1[0]; a.out_lo = w.d1[1]; a.M = (long long)N * a.g.rh * a.g.rw;
        TRXL_PROPAGATE(launch_gather<32>(st, a));
    }
    // ---- layer 1 ----
    TRXL_PROPAGATE(launch_colsum<32>(st, w.d1[0], w.d1[1], m1, w.colpart, g[1]));
    wa = WgradArgs{};
    tcg_plan_forward(e, 1, wa.g, unused);
    wa.a_hi = w.x0[0]; wa.a_lo = w.x0[1]; wa.dy_hi = w.d1[0]; wa.dy_lo = w.d1[1]; wa.M = m1;
    TRXL_PROPAGATE(launch_wgrad<32>(st, wa, 1, C, g[0], w));
    return TRXL_OK;
}
