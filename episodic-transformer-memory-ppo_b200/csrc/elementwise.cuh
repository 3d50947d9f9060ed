#pragma once
#include "common.cuh"

long long ew_scratch_floats(int rows, int cols);

int ew_layernorm_fwd(cudaStream_t st, const float* x, long long ldx, const float* x2, long long ldx2, const float* gamma,
                     const float* beta, float* y, long long ldy, float* sum_out, long long lds, float* mean, float* rstd,
                     int rows, int D);
int ew_layernorm_bwd(cudaStream_t st, const float* dy, long long lddy, const float* x, long long ldx, const float* mean,
                     const float* rstd, const float* gamma, float* dx, long long lddx, int accumulate_dx, float* dgamma,
                     float* dbeta, int accumulate_params, float* scratch, int rows, int D);
int ew_colsum(cudaStream_t st, const float* a, long long lda, float* out, int rows, int cols, float alpha, int accumulate,
              float* scratch);
int ew_add(cudaStream_t st, const float* a, long long lda, const float* b, long long ldb, float* out, long long ldo, int rows,
           int cols);
int ew_relu_bwd(cudaStream_t st, const float* dy, long long lddy, const float* y, long long ldy, float* dx, long long lddx,
                int rows, int cols, int accumulate);
int ew_gate_fwd_a(cudaStream_t st, const float* G1, const float* G2, const float* bg, const float* x, long long ldx, float* r,
                  float* z, float* rx, int N, int D);
int ew_gate_fwd_b(cudaStream_t st, const float* G1, const float* G3, const float* x, long long ldx, const float* z, float* hc,
                  float* out, long long ldo, int N, int D);
int ew_gate_bwd_a(cudaStream_t st, const float* dout, long long lddo, const float* x, long long ldx, const float* z,
                  const float* hc, float* dA1, float* dz, float* dx, long long lddx, int accumulate_dx, int N, int D);
int ew_gate_bwd_b(cudaStream_t st, const float* drx, const float* x, long long ldx, const float* r, const float* z,
                  const float* dz, float* dA1, float* dx, long long lddx, int N, int D);
int ew_scale_cols(cudaStream_t st, const float* W, const float* gamma, float* Wg, int rows, int cols);
int ew_matvec(cudaStream_t st, const float* W, const float* v, float* out, int rows, int cols);
int ew_head_dot(cudaStream_t st, const float* Q, const float* kb, float* qkb, int N, int H, int dh);
int ew_head_dot_bwd(cudaStream_t st, const float* Q, const float* dqkb, const float* kb, float* dQ, float* dkb, int N, int H,
                    int dh, float* scratch);
int ew_unfold(cudaStream_t st, const float* dWg, const float* db, const float* W, const float* gamma, const float* beta,
              float* dW, float* dgamma, float* dbeta, int rows, int cols, int accumulate);
