// Geometry of the implicit-GEMM formulation of the CNN encoder (reference model.py:27-38,87-94: conv 8/4 -> 4/2 -> 3/1
// with ReLUs) used by the tcgen05 kernels in tc_conv.cu.  Plain C++ without CUDA types: the same functions drive the
// device producers and the CPU emulation in tests/emu that pins the index math against torch's convolutions.
//
// Every convolution pass (forward, data gradient, weight gradient) is a GEMM whose A operand is *gathered* from an NHWC
// image: GEMM row m = (image n, ry, rx) on a per-image row grid, and K is cut into blocks of 32 floats, each of which is
// one contiguous 128-byte run of the image starting at pixel (ry*rstride + tap_dy, rx*rstride + tap_dx), channel tap_c0.
// A run that starts outside the image reads as zeros (only the data-gradient passes have such taps).
//
//   forward  l : image = input activations,   rows = output pixels,                  K = (ky, kx, c)
//   dgrad    3 : image = dL/d(conv3 output),  rows = conv3 input pixels,             K = (ky, kx, oc), taps (-ky, -kx)
//   dgrad    2 : image = dL/d(conv2 output),  rows = conv2 input pixels of one parity class (py, px) (stride 2 makes
//                the set of contributing kernel taps depend on the pixel parity),     K = (ty, tx, oc), taps (-ty, -tx),
//                kernel tap (ky, kx) = (py + 2 ty, px + 2 tx)
//   wgrad    l : same gather as forward l, reduced over the rows instead of over K
#pragma once

#ifdef __CUDACC__
#define TCG_HD __host__ __device__ __forceinline__
#else
#define TCG_HD inline
#endif

constexpr int TCG_BK = 32;          // floats per K block
constexpr int TCG_MAX_KB = 24;

struct TcgGather {
    int img_h, img_w, img_c;        // NHWC source image, img_c floats per pixel
    int rh, rw, rstride;            // row grid per image; anchor pixel of row (ry, rx) = (ry*rstride, rx*rstride)
    int nkb;                        // K = nkb * 32
    int run_px;                     // pixels covered by one 32-float run (>= 1)
    signed char tap_dy[TCG_MAX_KB], tap_dx[TCG_MAX_KB];
    short tap_c0[TCG_MAX_KB];
};

struct TcgScatter {                 // row (n, ry, rx) of the GEMM result -> pixel of the NHWC output image (out_c floats)
    int out_h, out_w, out_c, ostride, oy0, ox0;
};

struct TcgEncoder {
    int C, Cp;                      // observation channels, padded to 4
    int H, W, h1, w1, h2, w2, h3, w3;
};

TCG_HD void tcg_row(const TcgGather& g, long long m, int& n, int& ry, int& rx) {
    const int per = g.rh * g.rw;
    n = (int)(m / per);
    const int r = (int)(m - (long long)n * per);
    ry = r / g.rw;
    rx = r - ry * g.rw;
}
// float offset of the 32-float run of (row, K block) inside the image tensor, or -1 when the run is outside (zeros)
TCG_HD long long tcg_src(const TcgGather& g, long long img, int ry, int rx, int kb) {
    const int sy = ry * g.rstride + g.tap_dy[kb], sx = rx * g.rstride + g.tap_dx[kb];
    if (sy < 0 || sy >= g.img_h || sx < 0 || sx + g.run_px > g.img_w) return -1;
    return ((img * g.img_h + sy) * g.img_w + sx) * g.img_c + g.tap_c0[kb];
}
TCG_HD long long tcg_dst(const TcgScatter& s, long long n, int ry, int rx) {
    return ((n * s.out_h + ry * s.ostride + s.oy0) * s.out_w + rx * s.ostride + s.ox0) * s.out_c;
}

// false when the tensor-core path does not cover the shape (then the caller keeps the cuDNN / SIMT encoders)
inline bool tcg_encoder(int C, int H, int W, TcgEncoder& e) {
    e.C = C; e.Cp = 4; e.H = H; e.W = W;
    if (C < 1 || C > 4 || H < 36 || W < 36) return false;
    e.h1 = (H - 8) / 4 + 1; e.w1 = (W - 8) / 4 + 1;
    e.h2 = (e.h1 - 4) / 2 + 1; e.w2 = (e.w1 - 4) / 2 + 1;
    e.h3 = e.h2 - 2; e.w3 = e.w2 - 2;
    return e.h3 >= 1 && e.w3 >= 1;
}
inline int tcg_out_channels(int layer) { return layer == 1 ? 32 : 64; }

inline void tcg_plan_forward(const TcgEncoder& e, int layer, TcgGather& g, TcgScatter& s) {
    if (layer == 1) {
        g.img_h = e.H; g.img_w = e.W; g.img_c = e.Cp; g.rh = e.h1; g.rw = e.w1; g.rstride = 4; g.nkb = 8; g.run_px = 8;
        for (int ky = 0; ky < 8; ++ky) { g.tap_dy[ky] = (signed char)ky; g.tap_dx[ky] = 0; g.tap_c0[ky] = 0; }     // run = 8 px x 4 ch
        s.out_h = e.h1; s.out_w = e.w1; s.out_c = 32;
    } else if (layer == 2) {
        g.img_h = e.h1; g.img_w = e.w1; g.img_c = 32; g.rh = e.h2; g.rw = e.w2; g.rstride = 2; g.nkb = 16; g.run_px = 1;
        for (int t = 0; t < 16; ++t) { g.tap_dy[t] = (signed char)(t / 4); g.tap_dx[t] = (signed char)(t % 4); g.tap_c0[t] = 0; }
        s.out_h = e.h2; s.out_w = e.w2; s.out_c = 64;
    } else {
        g.img_h = e.h2; g.img_w = e.w2; g.img_c = 64; g.rh = e.h3; g.rw = e.w3; g.rstride = 1; g.nkb = 18; g.run_px = 1;
        for (int t = 0; t < 18; ++t) {
            g.tap_dy[t] = (signed char)((t / 2) / 3); g.tap_dx[t] = (signed char)((t / 2) % 3); g.tap_c0[t] = (short)((t & 1) * 32);
        }
        s.out_h = e.h3; s.out_w = e.w3; s.out_c = 64;
    }
    s.ostride = 1; s.oy0 = 0; s.ox0 = 0;
}
// d(conv3 input): rows = every pixel of the (h2, w2) image, gathered from dL/d(conv3 output) (h3, w3, 64)
inline void tcg_plan_dgrad3(const TcgEncoder& e, TcgGather& g, TcgScatter& s) {
    g.img_h = e.h3; g.img_w = e.w3; g.img_c = 64; g.rh = e.h2; g.rw = e.w2; g.rstride = 1; g.nkb = 18; g.run_px = 1;
    for (int t = 0; t < 18; ++t) {
        g.tap_dy[t] = (signed char)(-((t / 2) / 3)); g.tap_dx[t] = (signed char)(-((t / 2) % 3)); g.tap_c0[t] = (short)((t & 1) * 32);
    }
    s.out_h = e.h2; s.out_w = e.w2; s.out_c = 64; s.ostride = 1; s.oy0 = 0; s.ox0 = 0;
}
// d(conv2 input) for the pixels (2a + py, 2b + px), gathered from dL/d(conv2 output) (h2, w2, 64)
inline void tcg_plan_dgrad2(const TcgEncoder& e, int py, int px, TcgGather& g, TcgScatter& s) {
    g.img_h = e.h2; g.img_w = e.w2; g.img_c = 64; g.rh = (e.h1 - py + 1) / 2; g.rw = (e.w1 - px + 1) / 2; g.rstride = 1;
    g.nkb = 8; g.run_px = 1;
    for (int t = 0; t < 8; ++t) {
        g.tap_dy[t] = (signed char)(-((t / 2) / 2)); g.tap_dx[t] = (signed char)(-((t / 2) % 2)); g.tap_c0[t] = (short)((t & 1) * 32);
    }
    s.out_h = e.h1; s.out_w = e.w1; s.out_c = 32; s.ostride = 2; s.oy0 = py; s.ox0 = px;
}

// All four parity classes at once: the pixels (2a + py, 2b + px), py, px in {0, 1}, gather the SAME 2x2 neighbourhood
// {a - 1, a} x {b - 1, b} of dL/d(conv2 output) -- only the weights differ.  One GEMM with the four classes' weight matrices
// stacked along N (4 x 32 columns): the gathered operand is fetched once instead of four times and the MMAs are 128 wide.
// The scatter's (oy0, ox0) are those of class (0, 0); the epilogue adds the class offset and drops pixels outside the image.
inline void tcg_plan_dgrad2_merged(const TcgEncoder& e, TcgGather& g, TcgScatter& s) {
    tcg_plan_dgrad2(e, 0, 0, g, s);
    g.rh = (e.h1 + 1) / 2; g.rw = (e.w1 + 1) / 2;
}

// ---- packed K-major weight matrices: element (row, k) -> index into the reference (oc, c, ky, kx) tensor, -1 = zero ----
TCG_HD int tcg_wfwd_index(int layer, int C, int oc, int k) {
    if (layer == 1) {
        const int c = k & 3, kx = (k >> 2) & 7, ky = k >> 5;
        return c < C ? ((oc * C + c) * 8 + ky) * 8 + kx : -1;
    }
    if (layer == 2) {
        const int c = k & 31, kx = (k >> 5) & 3, ky = k >> 7;
        return ((oc * 32 + c) * 4 + ky) * 4 + kx;
    }
    const int c = k & 63, t = k >> 6, kx = t % 3, ky = t / 3;
    return ((oc * 64 + c) * 3 + ky) * 3 + kx;
}
// rows = conv3 input channel c, k = (ky, kx, oc)
TCG_HD int tcg_wdgrad3_index(int c, int k) {
    const int oc = k & 63, t = k >> 6, kx = t % 3, ky = t / 3;
    return ((oc * 64 + c) * 3 + ky) * 3 + kx;
}
// rows = conv2 input channel c, k = (ty, tx, oc) of parity class (py, px)
TCG_HD int tcg_wdgrad2_index(int py, int px, int c, int k) {
    const int oc = k & 63, t = k >> 6, tx = t & 1, ty = t >> 1;
    return ((oc * 32 + c) * 4 + (py + 2 * ty)) * 4 + (px + 2 * tx);
}
