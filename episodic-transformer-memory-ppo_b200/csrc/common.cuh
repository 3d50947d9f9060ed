// Shared helpers for the trxl-ppo sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TRXL_OK 0
#define TRXL_ERR_ARG -1
#define TRXL_ERR_CUDA -2
#define TRXL_ERR_UNSUPPORTED -3

// thread-local last-error string (set by host wrappers, read through trxl_last_error()).
void trxl_set_error(const char* fmt, ...);
// number of kernels this library has launched (bench.py reports it as gpu_launches)
extern long long g_trxl_launches;
// optional per-launch timing of the attention kernels (CUDA events on the launch stream); see api.cu
void trxl_prof_begin(int kind, int n, cudaStream_t st);
void trxl_prof_end(int kind, cudaStream_t st);
void trxl_prof_aux(int kind, long long aux);      // attach a count (e.g. tiles) to the slot opened by trxl_prof_begin

#define TRXL_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            trxl_set_error(__VA_ARGS__);          \
            return TRXL_ERR_ARG;                  \
        }                                         \
    } while (0)

#define TRXL_CHECK_LAUNCH(what)                                                       \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        ++g_trxl_launches;                                                            \
        if (e__ != cudaSuccess) {                                                     \
            trxl_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e__)); \
            return TRXL_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

#define TRXL_PROPAGATE(expr)          \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != TRXL_OK) return rc__; \
    } while (0)

static inline int trxl_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum NV values across the 32 lanes of a warp with a transposing butterfly: every lane ends up
// holding every total, using fewer shuffles than NV independent butterflies when NV > 1.
// (NV independent xor-butterflies cost 5*NV shuffles; for the small NV used here that is fine and
// keeps the result bitwise identical on every lane, which the softmax bookkeeping relies on.)
template <int NV>
__device__ __forceinline__ void warp_sum_multi(float (&v)[NV]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
}

// block-wide sum through shared memory (blockDim.x multiple of 32, <= 1024). Result valid on all threads.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    float t = (lane < nw) ? smem32[lane] : 0.f;
    t = warp_sum(t);
    return t;
}
__device__ __forceinline__ double block_sum_d(double v, double* smem32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    double t = (lane < nw) ? smem32[lane] : 0.0;
    t = warp_sum_d(t);
    return t;
}
#endif
