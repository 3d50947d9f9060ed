// Inference path of the Atari-style CNN encoder (reference model.py:27-38,87-94) as im2col + GEMM.
//
// Used for the rollout forwards (batch = n_workers): cuDNN needs ~150 us for the three tiny
// convolutions at batch 32, most of it launch/latency; here each layer is one gather (im2col) plus one
// of this library's GEMMs with the bias + ReLU fused, and activations stay channels-last between
// layers so every gather reads contiguous channel runs.  Training minibatches keep cuDNN (autograd).
#include "conv.cuh"

#include "gemm.cuh"

constexpr long long CONV_SPLITK_FLOATS = 4LL << 20;      // split-K partials of the conv2/conv3 GEMMs at rollout batch sizes

namespace {

struct WsGuard {
    WsGuard(float* p, long long n) { trxl_gemm_set_workspace(p, n); }
    ~WsGuard() { trxl_gemm_set_workspace(nullptr, 0); }
};

// cols[(n, oy, ox), (c, ky, kx)] = x[n, c, oy*s+ky, ox*s+kx]          (NCHW input: the observation)
// one thread per (row, c, ky): KW contiguous inputs -> KW contiguous outputs, 32-bit index math
__global__ void im2col_nchw_kernel(const float* __restrict__ x, float* __restrict__ cols, int N, int C, int H, int W, int KH,
                                   int KW, int S, int OH, int OW) {
    const int total = N * OH * OW * C * KH;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ky = i % KH, c = (i / KH) % C, m = i / (KH * C);
        const int ox = m % OW, oy = (m / OW) % OH, n = m / (OW * OH);
        const float* src = x + (((long long)n * C + c) * H + oy * S + ky) * W + ox * S;
        float* dst = cols + ((long long)m * C + c) * KH * KW + ky * KW;
        for (int kx = 0; kx < KW; ++kx) dst[kx] = src[kx];
    }
}
// channels-last input y[(n, iy, ix), c] (the previous layer's GEMM output) -> cols[(n, oy, ox), (ky, kx, c)]:
// whole channel runs move as float4 (C % 4 == 0), reads and writes both contiguous
__global__ void im2col_nhwc_kernel(const float* __restrict__ y, float* __restrict__ cols, int N, int C, int H, int W, int KH,
                                   int KW, int S, int OH, int OW) {
    const int C4 = C >> 2;
    const int total = N * OH * OW * KH * KW * C4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c4 = i % C4, kx = (i / C4) % KW, ky = (i / (C4 * KW)) % KH, m = i / (C4 * KW * KH);
        const int ox = m % OW, oy = (m / OW) % OH, n = m / (OW * OH);
        const float4 v = *reinterpret_cast<const float4*>(y + (((long long)n * H + oy * S + ky) * W + ox * S + kx) * C + c4 * 4);
        *reinterpret_cast<float4*>(cols + (((long long)m * KH + ky) * KW + kx) * C + c4 * 4) = v;
    }
}
// w[oc][c][ky][kx] -> wp[oc][ky][kx][c]  (matches the (ky, kx, c) patch order above)
__global__ void permute_weight_kernel(const float* __restrict__ w, float* __restrict__ wp, int OC, int C, int KH, int KW) {
    const int total = OC * C * KH * KW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % C, kx = (i / C) % KW, ky = (i / (C * KW)) % KH, oc = i / (C * KW * KH);
        wp[i] = w[((oc * C + c) * KH + ky) * KW + kx];
    }
}
// feat[n, c*P + p] = y[(n, p), c]  -- the reference flattens NCHW (model.py:94)
__global__ void nhwc_to_flat_nchw_kernel(const float* __restrict__ y, float* __restrict__ feat, int N, int P, int C) {
    const long long total = (long long)N * P * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % P), c = (int)((i / P) % C), n = (int)(i / ((long long)P * C));
        feat[i] = y[((long long)n * P + p) * C + c];
    }
}

int grid_for(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

long long conv_encoder_workspace_floats(int N, int C, int H, int W) {
    const int oh1 = (H - 8) / 4 + 1, ow1 = (W - 8) / 4 + 1, oh2 = (oh1 - 4) / 2 + 1, ow2 = (ow1 - 4) / 2 + 1, oh3 = oh2 - 2, ow3 = ow2 - 2;
    if (oh3 <= 0 || ow3 <= 0) return -1;
    const long long cols1 = (long long)N * oh1 * ow1 * C * 64, y1 = (long long)N * oh1 * ow1 * 32;
    const long long cols2 = (long long)N * oh2 * ow2 * 512, y2 = (long long)N * oh2 * ow2 * 64;
    const long long cols3 = (long long)N * oh3 * ow3 * 576, y3 = (long long)N * oh3 * ow3 * 64;
    long long cols = cols1 > cols2 ? cols1 : cols2;
    if (cols3 > cols) cols = cols3;
    return cols + y1 + y2 + y3 + 64 + CONV_SPLITK_FLOATS + 64 * 512 + 64 * 576;
}

int conv_encoder_forward(cudaStream_t st, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                         const float* b3, const float* obs, int N, int C, int H, int W, float* ws, float* feat) {
    const int oh1 = (H - 8) / 4 + 1, ow1 = (W - 8) / 4 + 1, oh2 = (oh1 - 4) / 2 + 1, ow2 = (ow1 - 4) / 2 + 1, oh3 = oh2 - 2, ow3 = ow2 - 2;
    TRXL_CHECK_ARG(oh3 > 0 && ow3 > 0, "conv_encoder: observation %dx%d too small for the 8/4, 4/2, 3/1 stack", H, W);
    if (N == 0) return TRXL_OK;
    const long long m1 = (long long)N * oh1 * ow1, m2 = (long long)N * oh2 * ow2, m3 = (long long)N * oh3 * ow3;
    const int k1 = C * 64, k2 = 32 * 16, k3 = 64 * 9;
    long long colsz = m1 * k1;
    if (m2 * k2 > colsz) colsz = m2 * k2;
    if (m3 * k3 > colsz) colsz = m3 * k3;
    float* cols = ws;
    float* y1 = cols + (colsz + 3) / 4 * 4;
    float* y2 = y1 + m1 * 32;
    float* y3 = y2 + m2 * 64;
    float* w2p = y3 + m3 * 64;                       // conv2 / conv3 weights in (ky, kx, c) order
    float* w3p = w2p + 64 * k2;
    WsGuard guard(w3p + 64 * k3, CONV_SPLITK_FLOATS);
    permute_weight_kernel<<<grid_for(64 * k2), 256, 0, st>>>(w2, w2p, 64, 32, 4, 4);
    TRXL_CHECK_LAUNCH("permute_weight");
    permute_weight_kernel<<<grid_for(64 * k3), 256, 0, st>>>(w3, w3p, 64, 64, 3, 3);
    TRXL_CHECK_LAUNCH("permute_weight");
    im2col_nchw_kernel<<<grid_for(m1 * C * 8), 256, 0, st>>>(obs, cols, N, C, H, W, 8, 8, 4, oh1, ow1);
    TRXL_CHECK_LAUNCH("im2col_nchw");
    TRXL_PROPAGATE(gemm_nt(st, (int)m1, 32, k1, cols, k1, w1, k1, y1, 32, b1, 1));
    im2col_nhwc_kernel<<<grid_for(m2 * k2 / 4), 256, 0, st>>>(y1, cols, N, 32, oh1, ow1, 4, 4, 2, oh2, ow2);
    TRXL_CHECK_LAUNCH("im2col_nhwc");
    TRXL_PROPAGATE(gemm_nt(st, (int)m2, 64, k2, cols, k2, w2p, k2, y2, 64, b2, 1));
    im2col_nhwc_kernel<<<grid_for(m3 * k3 / 4), 256, 0, st>>>(y2, cols, N, 64, oh2, ow2, 3, 3, 1, oh3, ow3);
    TRXL_CHECK_LAUNCH("im2col_nhwc");
    TRXL_PROPAGATE(gemm_nt(st, (int)m3, 64, k3, cols, k3, w3p, k3, y3, 64, b3, 1));
    nhwc_to_flat_nchw_kernel<<<grid_for(m3 * 64), 256, 0, st>>>(y3, feat, N, oh3 * ow3, 64);
    TRXL_CHECK_LAUNCH("nhwc_to_flat_nchw");
    return TRXL_OK;
}
