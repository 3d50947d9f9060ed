// Row-wise / element-wise kernels around the GEMMs: LayerNorm (fwd/bwd), GRU-gate pieces,
// bias-gradient column sums, ReLU backward and the pre-LayerNorm weight folds.
// All of these touch (N, D)-sized activations or (D, D) weights: HBM/latency-bound, one pass each.
#include "elementwise.cuh"

namespace {

constexpr float LN_EPS = 1e-5f;

// ---------------------------------------------------------------- LayerNorm forward (warp per row)
// y = LN(x [+ x2]) * gamma + beta ; optionally stores the pre-norm sum; saves mean / rstd.
__global__ void ln_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ x2, long long ldx2,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                              long long ldy, float* __restrict__ sum_out, long long lds, float* __restrict__ mean,
                              float* __restrict__ rstd, int rows, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * ldx;
    const float* x2r = x2 ? x2 + (long long)row * ldx2 : nullptr;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s += xr[j] + (x2r ? x2r[j] : 0.f);
    const float mu = warp_sum(s) / (float)D;
    float v = 0.f;
    for (int j = lane; j < D; j += 32) {
        const float d = xr[j] + (x2r ? x2r[j] : 0.f) - mu;
        v = fmaf(d, d, v);
    }
    const float rs = rsqrtf(warp_sum(v) / (float)D + LN_EPS);
    if (lane == 0) {
        if (mean) mean[row] = mu;
        if (rstd) rstd[row] = rs;
    }
    for (int j = lane; j < D; j += 32) {
        const float t = xr[j] + (x2r ? x2r[j] : 0.f);
        if (sum_out) sum_out[(long long)row * lds + j] = t;
        y[(long long)row * ldy + j] = (t - mu) * rs * gamma[j] + beta[j];
    }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)) ; optionally accumulated into dx
__global__ void ln_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
                              const float* __restrict__ mean, const float* __restrict__ rstd,
                              const float* __restrict__ gamma, float* __restrict__ dx, long long lddx, int accumulate,
                              int rows, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float mu = mean[row], rs = rstd[row];
    const float* dyr = dy + (long long)row * lddy;
    const float* xr = x + (long long)row * ldx;
    float a = 0.f, b = 0.f;
    for (int j = lane; j < D; j += 32) {
        const float gg = dyr[j] * gamma[j];
        a += gg;
        b = fmaf(gg, (xr[j] - mu) * rs, b);
    }
    a = warp_sum(a) / (float)D;
    b = warp_sum(b) / (float)D;
    for (int j = lane; j < D; j += 32) {
        const float xh = (xr[j] - mu) * rs;
        const float v = rs * (dyr[j] * gamma[j] - a - xh * b);
        float* p = dx + (long long)row * lddx + j;
        *p = accumulate ? *p + v : v;
    }
}

// dgamma[j] = sum_rows dy*xhat ; dbeta[j] = sum_rows dy.  One thread per column, rows in chunks
// across blockIdx.y with a second pass over the chunk partials (deterministic).
__global__ void ln_bwd_params_partial(const float* __restrict__ dy, long long lddy, const float* __restrict__ x,
                                      long long ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                                      float* __restrict__ part, int rows, int D, int rows_per_chunk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float g = 0.f, b = 0.f;
    for (int r = r0; r < r1; ++r) {
        const float d = dy[(long long)r * lddy + j];
        g = fmaf(d, (x[(long long)r * ldx + j] - mean[r]) * rstd[r], g);
        b += d;
    }
    part[((long long)blockIdx.y * 2 + 0) * D + j] = g;
    part[((long long)blockIdx.y * 2 + 1) * D + j] = b;
}
__global__ void ln_bwd_params_final(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int D, int chunks, int accumulate) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    float g = 0.f, b = 0.f;
    for (int c = 0; c < chunks; ++c) {
        g += part[((long long)c * 2 + 0) * D + j];
        b += part[((long long)c * 2 + 1) * D + j];
    }
    dgamma[j] = accumulate ? dgamma[j] + g : g;
    dbeta[j] = accumulate ? dbeta[j] + b : b;
}

// ---------------------------------------------------------------- column sums (bias gradients)
__global__ void colsum_partial(const float* __restrict__ a, long long lda, float* __restrict__ part, int rows, int cols,
                               int rows_per_chunk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += a[(long long)r * lda + j];
    part[(long long)blockIdx.y * cols + j] = s;
}
__global__ void colsum_final(const float* __restrict__ part, float* __restrict__ out, int cols, int chunks, float alpha,
                             int accumulate) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += part[(long long)c * cols + j];
    s *= alpha;
    out[j] = accumulate ? out[j] + s : s;
}

// ---------------------------------------------------------------- simple element-wise
__global__ void add_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b, long long ldb,
                           float* __restrict__ out, long long ldo, int rows, int cols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    out[(long long)r * ldo + c] = a[(long long)r * lda + c] + b[(long long)r * ldb + c];
}
// dx = dy * (y > 0), optionally accumulated
__global__ void relu_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ y, long long ldy,
                                float* __restrict__ dx, long long lddx, int rows, int cols, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = y[(long long)r * ldy + c] > 0.f ? dy[(long long)r * lddy + c] : 0.f;
    float* p = dx + (long long)r * lddx + c;
    *p = accumulate ? *p + v : v;
}

// ---------------------------------------------------------------- GRU gate (reference transformer.py:295-298)
// G1 = y [Wr;Wz;Wg]^T (N,3D), G2 = x [Ur;Uz]^T (N,2D)
__global__ void gate_fwd_a_kernel(const float* __restrict__ G1, const float* __restrict__ G2, const float* __restrict__ bg,
                                  const float* __restrict__ x, long long ldx, float* __restrict__ r, float* __restrict__ z,
                                  float* __restrict__ rx, int N, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * D) return;
    const int n = (int)(i / D), j = (int)(i % D);
    const float ar = G1[(long long)n * 3 * D + j] + G2[(long long)n * 2 * D + j];
    const float az = G1[(long long)n * 3 * D + D + j] + G2[(long long)n * 2 * D + D + j] - bg[j];
    const float rr = 1.f / (1.f + expf(-ar));
    const float zz = 1.f / (1.f + expf(-az));
    r[i] = rr;
    z[i] = zz;
    rx[i] = rr * x[(long long)n * ldx + j];
}
__global__ void gate_fwd_b_kernel(const float* __restrict__ G1, const float* __restrict__ G3, const float* __restrict__ x,
                                  long long ldx, const float* __restrict__ z, float* __restrict__ hc,
                                  float* __restrict__ out, long long ldo, int N, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * D) return;
    const int n = (int)(i / D), j = (int)(i % D);
    const float h = tanhf(G1[(long long)n * 3 * D + 2 * D + j] + G3[i]);
    hc[i] = h;
    const float zz = z[i];
    out[(long long)n * ldo + j] = (1.f - zz) * x[(long long)n * ldx + j] + zz * h;
}
// from d(out): d(a_g) -> dA1[:, 2D:3D], dz (scratch), dx = dout*(1-z)
__global__ void gate_bwd_a_kernel(const float* __restrict__ dout, long long lddo, const float* __restrict__ x, long long ldx,
                                  const float* __restrict__ z, const float* __restrict__ hc, float* __restrict__ dA1,
                                  float* __restrict__ dz, float* __restrict__ dx, long long lddx, int accumulate_dx, int N,
                                  int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * D) return;
    const int n = (int)(i / D), j = (int)(i % D);
    const float d = dout[(long long)n * lddo + j];
    const float h = hc[i], zz = z[i], xv = x[(long long)n * ldx + j];
    dz[i] = d * (h - xv);
    dA1[(long long)n * 3 * D + 2 * D + j] = d * zz * (1.f - h * h);
    float* p = dx + (long long)n * lddx + j;
    const float v = d * (1.f - zz);
    *p = accumulate_dx ? *p + v : v;
}
// from d(rx): da_r -> dA1[:, 0:D], da_z -> dA1[:, D:2D], dx += drx * r
__global__ void gate_bwd_b_kernel(const float* __restrict__ drx, const float* __restrict__ x, long long ldx,
                                  const float* __restrict__ r, const float* __restrict__ z, const float* __restrict__ dz,
                                  float* __restrict__ dA1, float* __restrict__ dx, long long lddx, int N, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * D) return;
    const int n = (int)(i / D), j = (int)(i % D);
    const float rr = r[i], zz = z[i], d = drx[i];
    dA1[(long long)n * 3 * D + j] = d * x[(long long)n * ldx + j] * rr * (1.f - rr);
    dA1[(long long)n * 3 * D + D + j] = dz[i] * zz * (1.f - zz);
    dx[(long long)n * lddx + j] += d * rr;
}

// ---------------------------------------------------------------- pre-LayerNorm folds
// Wg[d, j] = W[d, j] * gamma[j]
__global__ void scale_cols_kernel(const float* __restrict__ W, const float* __restrict__ gamma, float* __restrict__ Wg,
                                  int rows, int cols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    Wg[i] = W[i] * gamma[i % cols];
}
// out[d] = W[d, :] . v   (warp per row)
__global__ void matvec_kernel(const float* __restrict__ W, const float* __restrict__ v, float* __restrict__ out, int rows,
                              int cols) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int j = lane; j < cols; j += 32) s = fmaf(W[(long long)row * cols + j], v[j], s);
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}
// qkb[n, h] = sum_{d in head h} Q[n, d] kb[d]
__global__ void head_dot_kernel(const float* __restrict__ Q, const float* __restrict__ kb, float* __restrict__ qkb, int N,
                                int H, int dh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * H) return;
    const int n = i / H, h = i % H;
    float s = 0.f;
    for (int d = 0; d < dh; ++d) s = fmaf(Q[(long long)n * H * dh + h * dh + d], kb[h * dh + d], s);
    qkb[i] = s;
}
// dQ[n, d] += dqkb[n, h(d)] * kb[d]
__global__ void head_dot_bwd_q_kernel(const float* __restrict__ dqkb, const float* __restrict__ kb, float* __restrict__ dQ,
                                      int N, int H, int dh) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int D = H * dh;
    if (i >= (long long)N * D) return;
    const int n = (int)(i / D), d = (int)(i % D);
    dQ[i] += dqkb[(long long)n * H + d / dh] * kb[d];
}
// dkb[d] = sum_n Q[n, d] dqkb[n, h(d)]   (column reduction, partial over row chunks)
__global__ void head_dot_bwd_kb_partial(const float* __restrict__ Q, const float* __restrict__ dqkb, float* __restrict__ part,
                                        int N, int H, int dh, int rows_per_chunk) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int D = H * dh;
    if (d >= D) return;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(N, r0 + rows_per_chunk);
    float s = 0.f;
    for (int n = r0; n < r1; ++n) s = fmaf(Q[(long long)n * D + d], dqkb[(long long)n * H + d / dh], s);
    part[(long long)blockIdx.y * D + d] = s;
}
// Undo Wg = W*gamma, b = W beta:  dW = dWg*gamma + db (x) beta ;  dgamma[j] (+)= sum_d dWg[d,j] W[d,j] ;
// dbeta[j] (+)= sum_d db[d] W[d,j].   One thread per column j.
__global__ void unfold_kernel(const float* __restrict__ dWg, const float* __restrict__ db, const float* __restrict__ W,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dW,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int cols, int accumulate) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    const float gj = gamma[j], bj = beta[j];
    float sg = 0.f, sb = 0.f;
    for (int d = 0; d < rows; ++d) {
        const float w = W[(long long)d * cols + j], dg = dWg[(long long)d * cols + j], dbv = db[d];
        sg = fmaf(dg, w, sg);
        sb = fmaf(dbv, w, sb);
        dW[(long long)d * cols + j] = fmaf(dg, gj, dbv * bj);
    }
    dgamma[j] = accumulate ? dgamma[j] + sg : sg;
    dbeta[j] = accumulate ? dbeta[j] + sb : sb;
}

// row chunks of the two-pass column reductions: enough CTAs to cover the SMs at training batch sizes
int chunks_for(int rows) { return rows >= 1024 ? 64 : (rows >= 256 ? 16 : (rows >= 32 ? 4 : 1)); }

}  // namespace

int ew_layernorm_fwd(cudaStream_t st, const float* x, long long ldx, const float* x2, long long ldx2, const float* gamma,
                     const float* beta, float* y, long long ldy, float* sum_out, long long lds, float* mean, float* rstd,
                     int rows, int D) {
    if (rows == 0) return TRXL_OK;
    ln_fwd_kernel<<<trxl_cdiv(rows, 4), 128, 0, st>>>(x, ldx, x2, ldx2, gamma, beta, y, ldy, sum_out, lds, mean, rstd, rows, D);
    TRXL_CHECK_LAUNCH("ln_fwd");
    return TRXL_OK;
}

int ew_layernorm_bwd(cudaStream_t st, const float* dy, long long lddy, const float* x, long long ldx, const float* mean,
                     const float* rstd, const float* gamma, float* dx, long long lddx, int accumulate_dx, float* dgamma,
                     float* dbeta, int accumulate_params, float* scratch, int rows, int D) {
    if (rows == 0) return TRXL_OK;
    ln_bwd_kernel<<<trxl_cdiv(rows, 4), 128, 0, st>>>(dy, lddy, x, ldx, mean, rstd, gamma, dx, lddx, accumulate_dx, rows, D);
    TRXL_CHECK_LAUNCH("ln_bwd");
    if (dgamma) {
        const int chunks = chunks_for(rows), rpc = trxl_cdiv(rows, chunks);
        dim3 grid(trxl_cdiv(D, 128), chunks);
        ln_bwd_params_partial<<<grid, 128, 0, st>>>(dy, lddy, x, ldx, mean, rstd, scratch, rows, D, rpc);
        TRXL_CHECK_LAUNCH("ln_bwd_params_partial");
        ln_bwd_params_final<<<trxl_cdiv(D, 128), 128, 0, st>>>(scratch, dgamma, dbeta, D, chunks, accumulate_params);
        TRXL_CHECK_LAUNCH("ln_bwd_params_final");
    }
    return TRXL_OK;
}
long long ew_scratch_floats(int rows, int cols) { return (long long)chunks_for(rows) * 2 * cols + 64; }

int ew_colsum(cudaStream_t st, const float* a, long long lda, float* out, int rows, int cols, float alpha, int accumulate,
              float* scratch) {
    if (cols == 0) return TRXL_OK;
    const int chunks = chunks_for(rows), rpc = trxl_cdiv(rows, chunks);
    dim3 grid(trxl_cdiv(cols, 128), chunks);
    colsum_partial<<<grid, 128, 0, st>>>(a, lda, scratch, rows, cols, rpc);
    TRXL_CHECK_LAUNCH("colsum_partial");
    colsum_final<<<trxl_cdiv(cols, 128), 128, 0, st>>>(scratch, out, cols, chunks, alpha, accumulate);
    TRXL_CHECK_LAUNCH("colsum_final");
    return TRXL_OK;
}

int ew_add(cudaStream_t st, const float* a, long long lda, const float* b, long long ldb, float* out, long long ldo, int rows,
           int cols) {
    const long long n = (long long)rows * cols;
    if (n == 0) return TRXL_OK;
    add_kernel<<<trxl_cdiv(n, 256), 256, 0, st>>>(a, lda, b, ldb, out, ldo, rows, cols);
    TRXL_CHECK_LAUNCH("add");
    return TRXL_OK;
}

int ew_relu_bwd(cudaStream_t st, const float* dy, long long lddy, const float* y, long long ldy, float* dx, long long lddx,
                int rows, int cols, int accumulate) {
    const long long n = (long long)rows * cols;
    if (n == 0) return TRXL_OK;
    relu_bwd_kernel<<<trxl_cdiv(n, 256), 256, 0, st>>>(dy, lddy, y, ldy, dx, lddx, rows, cols, accumulate);
    TRXL_CHECK_LAUNCH("relu_bwd");
    return TRXL_OK;
}

int ew_gate_fwd_a(cudaStream_t st, const float* G1, const float* G2, const float* bg, const float* x, long long ldx, float* r,
                  float* z, float* rx, int N, int D) {
    if (N == 0) return TRXL_OK;
    gate_fwd_a_kernel<<<trxl_cdiv((long long)N * D, 256), 256, 0, st>>>(G1, G2, bg, x, ldx, r, z, rx, N, D);
    TRXL_CHECK_LAUNCH("gate_fwd_a");
    return TRXL_OK;
}
int ew_gate_fwd_b(cudaStream_t st, const float* G1, const float* G3, const float* x, long long ldx, const float* z, float* hc,
                  float* out, long long ldo, int N, int D) {
    if (N == 0) return TRXL_OK;
    gate_fwd_b_kernel<<<trxl_cdiv((long long)N * D, 256), 256, 0, st>>>(G1, G3, x, ldx, z, hc, out, ldo, N, D);
    TRXL_CHECK_LAUNCH("gate_fwd_b");
    return TRXL_OK;
}
int ew_gate_bwd_a(cudaStream_t st, const float* dout, long long lddo, const float* x, long long ldx, const float* z,
                  const float* hc, float* dA1, float* dz, float* dx, long long lddx, int accumulate_dx, int N, int D) {
    if (N == 0) return TRXL_OK;
    gate_bwd_a_kernel<<<trxl_cdiv((long long)N * D, 256), 256, 0, st>>>(dout, lddo, x, ldx, z, hc, dA1, dz, dx, lddx,
                                                                      accumulate_dx, N, D);
    TRXL_CHECK_LAUNCH("gate_bwd_a");
    return TRXL_OK;
}
int ew_gate_bwd_b(cudaStream_t st, const float* drx, const float* x, long long ldx, const float* r, const float* z,
                  const float* dz, float* dA1, float* dx, long long lddx, int N, int D) {
    if (N == 0) return TRXL_OK;
    gate_bwd_b_kernel<<<trxl_cdiv((long long)N * D, 256), 256, 0, st>>>(drx, x, ldx, r, z, dz, dA1, dx, lddx, N, D);
    TRXL_CHECK_LAUNCH("gate_bwd_b");
    return TRXL_OK;
}

int ew_scale_cols(cudaStream_t st, const float* W, const float* gamma, float* Wg, int rows, int cols) {
    scale_cols_kernel<<<trxl_cdiv((long long)rows * cols, 256), 256, 0, st>>>(W, gamma, Wg, rows, cols);
    TRXL_CHECK_LAUNCH("scale_cols");
    return TRXL_OK;
}
int ew_matvec(cudaStream_t st, const float* W, const float* v, float* out, int rows, int cols) {
    matvec_kernel<<<trxl_cdiv(rows, 4), 128, 0, st>>>(W, v, out, rows, cols);
    TRXL_CHECK_LAUNCH("matvec");
    return TRXL_OK;
}
int ew_head_dot(cudaStream_t st, const float* Q, const float* kb, float* qkb, int N, int H, int dh) {
    if (N == 0) return TRXL_OK;
    head_dot_kernel<<<trxl_cdiv((long long)N * H, 128), 128, 0, st>>>(Q, kb, qkb, N, H, dh);
    TRXL_CHECK_LAUNCH("head_dot");
    return TRXL_OK;
}
int ew_head_dot_bwd(cudaStream_t st, const float* Q, const float* dqkb, const float* kb, float* dQ, float* dkb, int N, int H,
                    int dh, float* scratch) {
    if (N == 0) return TRXL_OK;
    const int D = H * dh;
    head_dot_bwd_q_kernel<<<trxl_cdiv((long long)N * D, 256), 256, 0, st>>>(dqkb, kb, dQ, N, H, dh);
    TRXL_CHECK_LAUNCH("head_dot_bwd_q");
    const int chunks = chunks_for(N), rpc = trxl_cdiv(N, chunks);
    dim3 grid(trxl_cdiv(D, 128), chunks);
    head_dot_bwd_kb_partial<<<grid, 128, 0, st>>>(Q, dqkb, scratch, N, H, dh, rpc);
    TRXL_CHECK_LAUNCH("head_dot_bwd_kb_partial");
    colsum_final<<<trxl_cdiv(D, 128), 128, 0, st>>>(scratch, dkb, D, chunks, 1.f, 0);
    TRXL_CHECK_LAUNCH("head_dot_bwd_kb_final");
    return TRXL_OK;
}
int ew_unfold(cudaStream_t st, const float* dWg, const float* db, const float* W, const float* gamma, const float* beta,
              float* dW, float* dgamma, float* dbeta, int rows, int cols, int accumulate) {
    unfold_kernel<<<trxl_cdiv(cols, 64), 64, 0, st>>>(dWg, db, W, gamma, beta, dW, dgamma, dbeta, rows, cols, accumulate);
    TRXL_CHECK_LAUNCH("unfold");
    return TRXL_OK;
}
