#pragma once
#include "common.cuh"

struct GemmArgs {
    int M = 0, N = 0, K = 0;
    const float* A = nullptr; long long lda = 0; int a_kc = 1;
    const float* B = nullptr; long long ldb = 0; int b_kc = 1;
    float* C = nullptr; long long ldc = 0;
    const float* bias = nullptr;          // (N,) added before ReLU
    const float* R = nullptr; long long ldr = 0;   // residual added after ReLU
    int relu = 0;
    int accumulate = 0;                   // C += result
    float alpha = 1.f;
    int batch = 1; long long sA = 0, sB = 0, sC = 0, sBias = 0, sR = 0;
    float* ws = nullptr; long long ws_floats = 0;   // optional split-K workspace
    int vecA = 0, vecB = 0, ksplit = 1, k_per_split = 0;   // filled by trxl_gemm
    int debug = 0;                        // TRXL_TC_DEBUG bitmask (timing experiments only; results are wrong when set)
    int cluster_reduce = 0;               // tensor-core path: the ksplit (2, 4 or 8) CTAs of a tile form a cluster and sum their
                                          // partials through distributed shared memory (no workspace, no reduce kernel)
    // grouped mode (tensor-core path only): output row tile t covers rows [tiles[t].x, +tiles[t].y) of A / C and multiplies
    // with batch element tiles[t].z of B (b_batch of them, stride sB).  grid.y = n_tiles; batch must be 1.
    const int4* tiles = nullptr; int n_tiles = 0; int b_batch = 1;
};

int trxl_gemm(GemmArgs g, cudaStream_t st);
// default split-K workspace for GEMMs issued by this host thread (set around a model forward/backward)
void trxl_gemm_set_workspace(float* ws, long long floats);

// convenience wrappers (row-major x (M,K), W (N,K))
static inline int gemm_nt(cudaStream_t st, int M, int N, int K, const float* x, long long ldx, const float* W, long long ldw,
                          float* y, long long ldy, const float* bias = nullptr, int relu = 0, const float* R = nullptr,
                          long long ldr = 0) {
    GemmArgs g; g.M = M; g.N = N; g.K = K; g.A = x; g.lda = ldx; g.a_kc = 1; g.B = W; g.ldb = ldw; g.b_kc = 1;
    g.C = y; g.ldc = ldy; g.bias = bias; g.relu = relu; g.R = R; g.ldr = ldr;
    return trxl_gemm(g, st);
}
// dx (M,K) = dy (M,N) * W (N,K)
static inline int gemm_nn(cudaStream_t st, int M, int K, int N, const float* dy, long long lddy, const float* W, long long ldw,
                          float* dx, long long lddx, int accumulate = 0) {
    GemmArgs g; g.M = M; g.N = K; g.K = N; g.A = dy; g.lda = lddy; g.a_kc = 1; g.B = W; g.ldb = ldw; g.b_kc = 0;
    g.C = dx; g.ldc = lddx; g.accumulate = accumulate;
    return trxl_gemm(g, st);
}
// dW (N,K) = dy^T (N,M) * x (M,K)
static inline int gemm_tn(cudaStream_t st, int N, int K, int M, const float* dy, long long lddy, const float* x, long long ldx,
                          float* dW, long long lddw, int accumulate = 0, float* ws = nullptr, long long ws_floats = 0) {
    GemmArgs g; g.M = N; g.N = K; g.K = M; g.A = dy; g.lda = lddy; g.a_kc = 0; g.B = x; g.ldb = ldx; g.b_kc = 0;
    g.C = dW; g.ldc = lddw; g.accumulate = accumulate; g.ws = ws; g.ws_floats = ws_floats;
    return trxl_gemm(g, st);
}
