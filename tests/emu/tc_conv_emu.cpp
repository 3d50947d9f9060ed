// CPU emulation of the implicit-GEMM CNN encoder (test infrastructure).  It walks exactly the geometry tables of
// csrc/tc_conv_geom.h -- the ones the sm_100a producers use -- with scalar loops, so `-m "not gpu"` tests can pin the
// index math (taps, parity classes, packed-weight permutations, scatter) against torch's convolutions without a GPU.
#include <cstring>
#include <vector>

#include "tc_conv_geom.h"

namespace {

using std::vector;

// out[dst(row) + j] = epi(sum_k A(row, k) * wp[j][k])        (the K-major gather GEMM of tc_conv.cu)
void gather_gemm(const TcgGather& g, const TcgScatter& s, const float* img, const long long* sample_index, long long M,
                 const float* wp, int bn, const float* bias, int relu, const float* mask, float* out) {
    const int K = g.nkb * TCG_BK;
    vector<float> a(K);
    for (long long m = 0; m < M; ++m) {
        int n, ry, rx;
        tcg_row(g, m, n, ry, rx);
        const long long src_img = sample_index ? sample_index[n] : n;
        for (int kb = 0; kb < g.nkb; ++kb) {
            const long long off = tcg_src(g, src_img, ry, rx, kb);
            for (int j = 0; j < TCG_BK; ++j) a[kb * TCG_BK + j] = off < 0 ? 0.f : img[off + j];
        }
        const long long dst = tcg_dst(s, n, ry, rx);
        for (int j = 0; j < bn; ++j) {
            double acc = 0.0;
            for (int k = 0; k < K; ++k) acc += (double)a[k] * wp[(long long)j * K + k];
            float v = (float)acc + (bias ? bias[j] : 0.f);
            if (relu && v < 0.f) v = 0.f;
            if (mask && !(mask[dst + j] > 0.f)) v = 0.f;
            out[dst + j] = v;
        }
    }
}
// dwp[k][j] = sum_m A(m, k) * dy[m][j]                       (the MN-major reduce-over-rows GEMM of tc_conv.cu)
void wgrad(const TcgGather& g, const float* img, const long long* sample_index, long long M, const float* dy, int bn,
           vector<double>& dwp) {
    const int K = g.nkb * TCG_BK;
    dwp.assign((size_t)K * bn, 0.0);
    for (long long m = 0; m < M; ++m) {
        int n, ry, rx;
        tcg_row(g, m, n, ry, rx);
        const long long src_img = sample_index ? sample_index[n] : n;
        for (int kb = 0; kb < g.nkb; ++kb) {
            const long long off = tcg_src(g, src_img, ry, rx, kb);
            if (off < 0) continue;
            for (int j = 0; j < TCG_BK; ++j) {
                const double a = img[off + j];
                for (int o = 0; o < bn; ++o) dwp[(size_t)(kb * TCG_BK + j) * bn + o] += a * dy[m * bn + o];
            }
        }
    }
}
void colsum(const float* dy, long long M, int bn, float* db) {
    for (int o = 0; o < bn; ++o) {
        double s = 0.0;
        for (long long m = 0; m < M; ++m) s += dy[m * bn + o];
        db[o] = (float)s;
    }
}

struct Encoder {
    TcgEncoder e;
    long long n;
    vector<float> x0, y1, y2, y3, wp1, wp2, wp3;
    void forward(const float* obs, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3) {
        x0.assign((size_t)n * e.H * e.W * 4, 0.f);
        for (long long i = 0; i < n; ++i)
            for (int c = 0; c < e.C; ++c)
                for (int y = 0; y < e.H; ++y)
                    for (int x = 0; x < e.W; ++x) x0[((i * e.H + y) * e.W + x) * 4 + c] = obs[((i * e.C + c) * e.H + y) * e.W + x];
        wp1.resize(32 * 256); wp2.resize(64 * 512); wp3.resize(64 * 576);
        for (int oc = 0; oc < 32; ++oc)
            for (int k = 0; k < 256; ++k) { const int i = tcg_wfwd_index(1, e.C, oc, k); wp1[oc * 256 + k] = i < 0 ? 0.f : w1[i]; }
        for (int oc = 0; oc < 64; ++oc)
            for (int k = 0; k < 512; ++k) wp2[oc * 512 + k] = w2[tcg_wfwd_index(2, e.C, oc, k)];
        for (int oc = 0; oc < 64; ++oc)
            for (int k = 0; k < 576; ++k) wp3[oc * 576 + k] = w3[tcg_wfwd_index(3, e.C, oc, k)];
        TcgGather g; TcgScatter s;
        y1.assign((size_t)n * e.h1 * e.w1 * 32, 0.f); y2.assign((size_t)n * e.h2 * e.w2 * 64, 0.f); y3.assign((size_t)n * e.h3 * e.w3 * 64, 0.f);
        tcg_plan_forward(e, 1, g, s); gather_gemm(g, s, x0.data(), nullptr, n * e.h1 * e.w1, wp1.data(), 32, b1, 1, nullptr, y1.data());
        tcg_plan_forward(e, 2, g, s); gather_gemm(g, s, y1.data(), nullptr, n * e.h2 * e.w2, wp2.data(), 64, b2, 1, nullptr, y2.data());
        tcg_plan_forward(e, 3, g, s); gather_gemm(g, s, y2.data(), nullptr, n * e.h3 * e.w3, wp3.data(), 64, b3, 1, nullptr, y3.data());
    }
};

}  // namespace

extern "C" {

int emu_feature_size(int C, int H, int W) {
    TcgEncoder e;
    if (!tcg_encoder(C, H, W, e)) return -1;
    return 64 * e.h3 * e.w3;
}

// feat (n, 64*h3*w3) in the reference's NCHW flatten order
int emu_forward(int C, int H, int W, long long n, const float* obs, const float* w1, const float* b1, const float* w2, const float* b2,
                const float* w3, const float* b3, float* feat) {
    Encoder enc;
    if (!tcg_encoder(C, H, W, enc.e)) return -1;
    enc.n = n;
    enc.forward(obs, w1, b1, w2, b2, w3, b3);
    const int P = enc.e.h3 * enc.e.w3;
    for (long long i = 0; i < n; ++i)
        for (int p = 0; p < P; ++p)
            for (int c = 0; c < 64; ++c) feat[i * 64 * P + c * P + p] = enc.y3[(i * P + p) * 64 + c];
    return 0;
}

int emu_backward(int C, int H, int W, long long n, const float* obs, const float* w1, const float* b1, const float* w2, const float* b2,
                 const float* w3, const float* b3, const float* dfeat, float* dw1, float* db1, float* dw2, float* db2, float* dw3,
                 float* db3) {
    Encoder enc;
    if (!tcg_encoder(C, H, W, enc.e)) return -1;
    enc.n = n;
    enc.forward(obs, w1, b1, w2, b2, w3, b3);
    const TcgEncoder& e = enc.e;
    const int P = e.h3 * e.w3;
    const long long m1 = n * e.h1 * e.w1, m2 = n * e.h2 * e.w2, m3 = n * P;
    vector<float> dy3((size_t)m3 * 64), dy2((size_t)m2 * 64, 0.f), dy1((size_t)m1 * 32, 0.f);
    for (long long i = 0; i < n; ++i)
        for (int p = 0; p < P; ++p)
            for (int c = 0; c < 64; ++c) {
                const long long o = (i * P + p) * 64 + c;
                dy3[o] = enc.y3[o] > 0.f ? dfeat[i * 64 * P + c * P + p] : 0.f;
            }
    TcgGather g; TcgScatter s;
    vector<double> dwp;
    // layer 3
    colsum(dy3.data(), m3, 64, db3);
    tcg_plan_forward(e, 3, g, s);
    wgrad(g, enc.y2.data(), nullptr, m3, dy3.data(), 64, dwp);
    for (int oc = 0; oc < 64; ++oc)
        for (int k = 0; k < 576; ++k) dw3[tcg_wfwd_index(3, e.C, oc, k)] = (float)dwp[(size_t)k * 64 + oc];
    vector<float> wd3(64 * 576);
    for (int c = 0; c < 64; ++c)
        for (int k = 0; k < 576; ++k) wd3[c * 576 + k] = w3[tcg_wdgrad3_index(c, k)];
    tcg_plan_dgrad3(e, g, s);
    gather_gemm(g, s, dy3.data(), nullptr, m2, wd3.data(), 64, nullptr, 0, enc.y2.data(), dy2.data());
    // layer 2
    colsum(dy2.data(), m2, 64, db2);
    tcg_plan_forward(e, 2, g, s);
    wgrad(g, enc.y1.data(), nullptr, m2, dy2.data(), 64, dwp);
    for (int oc = 0; oc < 64; ++oc)
        for (int k = 0; k < 512; ++k) dw2[tcg_wfwd_index(2, e.C, oc, k)] = (float)dwp[(size_t)k * 64 + oc];
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            vector<float> wd2(32 * 256);
            for (int c = 0; c < 32; ++c)
                for (int k = 0; k < 256; ++k) wd2[c * 256 + k] = w2[tcg_wdgrad2_index(py, px, c, k)];
            tcg_plan_dgrad2(e, py, px, g, s);
            gather_gemm(g, s, dy2.data(), nullptr, n * g.rh * g.rw, wd2.data(), 32, nullptr, 0, enc.y1.data(), dy1.data());
        }
    // layer 1
    colsum(dy1.data(), m1, 32, db1);
    tcg_plan_forward(e, 1, g, s);
    wgrad(g, enc.x0.data(), nullptr, m1, dy1.data(), 32, dwp);
    for (int oc = 0; oc < 32; ++oc)
        for (int k = 0; k < 256; ++k) {
            const int i = tcg_wfwd_index(1, e.C, oc, k);
            if (i >= 0) dw1[i] = (float)dwp[(size_t)k * 32 + oc];
        }
    return 0;
}

}  // extern "C"
