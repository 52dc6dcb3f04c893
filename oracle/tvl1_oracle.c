/*
 * tvl1_oracle.c -- CPU restatement of OpenCV's cv::optflow::DualTVL1OpticalFlow (CPU variant).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (tee_optical_flow_b200/) may call, link or
 * load this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / CPU baseline.
 *
 * What it restates
 * ----------------
 * The reference repo (nquach/TEE_optical_flow) holds no TV-L1 arithmetic of its own: it calls
 *     cv2.optflow.createOptFlow_DualTVL1()            optical_flow/calculate_optical_flow.py:577
 *     OF_model.setLambda(config.lambda_value)         optical_flow/calculate_optical_flow.py:578
 *     OF_model.calc(saliency_1, saliency_2, None)     optical_flow/calculate_optical_flow.py:642
 * i.e. the solver lives in the un-vendored third-party dependency `opencv-contrib-python>=4.5.0`
 * (requirements.txt:7), module optflow, file modules/optflow/src/tvl1flow.cpp.  That source is absent
 * from /root/reference and the cv2 build in this image has no `optflow` module, so this file restates
 * the published algorithm of that module (4.x) function by function:
 *     OpticalFlowDual_TVL1::calc, ::procOneScale, centeredGradient, forwardGradient, divergence,
 *     buildFlowMap, calcGradRho, estimateV, estimateU, estimateDualVariables,
 * together with the three imgproc primitives it calls: cv::resize(INTER_LINEAR), cv::remap(INTER_CUBIC,
 * BORDER_CONSTANT 0) and cv::medianBlur(ksize 5, float).
 *
 * Pinning status
 * --------------
 *  - The imgproc primitives ARE pinned: oracle/probe_remap.py and oracle/probe_resize.py show float32
 *    models with exactly the operation order used here to be BIT-EXACT against the genuine
 *    cv2.remap / cv2.resize (IPP off) / cv2.medianBlur of the OpenCV 4.13 in this image, and
 *    tests/test_oracle.py re-checks the compiled functions against committed cv2 outputs
 *    (tests/golden/primitives_cv2.npz).
 *  - The solver control flow and the pointwise formulae are restated from the published source and are
 *    cross-checked against an independent numpy + real-cv2 composition (oracle/tvl1_cv2ref.py), but there
 *    is no golden vector from a real cv2.optflow: "PARITY UNPINNED" for the solver as a whole.
 *    tools/dump_golden.py lets anybody with opencv-contrib produce such vectors.
 *
 * All arithmetic is IEEE float32 in the operation order of the C++ source; build with
 * -ffp-contract=off (no FMA contraction) -- see oracle/Makefile.
 */
#include <float.h>
#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TF_EXPORT __attribute__((visibility("default")))

#ifdef _OPENMP
#include <omp.h>
#endif
/* number of OpenMP threads the oracle uses (launchers such as torchrun export OMP_NUM_THREADS=1); returns the
 * value in effect */
TF_EXPORT int oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

typedef struct {
    double tau;            /* 0.25 */
    double lambda;         /* 0.15 */
    double theta;          /* 0.3  */
    double epsilon;        /* 0.01 */
    double scale_step;     /* 0.8  */
    int nscales;           /* 5 */
    int warps;             /* 5 */
    int inner_iterations;  /* 30 */
    int outer_iterations;  /* 10 */
    int median_filtering;  /* 5 (<=1: off; 3 or 5) */
    int err_mode;          /* 0: serial float32 accumulation in raster order (OpenCV estimateU);
                              1: float64 accumulation of the same float32 terms (what the GPU engine does) */
} tvl1_oracle_params;

/* ------------------------------------------------------------------------------------------------
 * cv::resize(src, dst, INTER_LINEAR) for CV_32FC1  (imgproc/resize.cpp: resizeGeneric_, HResizeLinear,
 * VResizeLinear).  scale_x/scale_y are the source-per-destination scales as hal::resize computes them:
 * 1./inv_scale.  Horizontal coefficients are clamped (fx=0 at the clamp), vertical ones are NOT (row
 * indices are clipped instead) -- verified bit-exact against cv2.resize, oracle/probe_resize.py.
 * ------------------------------------------------------------------------------------------------ */
TF_EXPORT void oracle_resize_linear(const float* src, int sH, int sW, float* dst, int dH, int dW,
                                    double scale_x, double scale_y) {
    int* xofs = (int*)malloc(sizeof(int) * (size_t)dW);
    float* alpha = (float*)malloc(sizeof(float) * 2 * (size_t)dW);
    for (int dx = 0; dx < dW; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= (float)sx;
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= sW - 1) { fx = 0.f; sx = sW - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = 1.f - fx;
        alpha[2 * dx + 1] = fx;
    }
#pragma omp parallel for schedule(static)
    for (int dy = 0; dy < dH; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= (float)sy;
        const float b0 = 1.f - fy, b1 = fy;
        int y0 = sy < 0 ? 0 : (sy > sH - 1 ? sH - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sH - 1 ? sH - 1 : sy + 1);
        const float* S0 = src + (size_t)y0 * sW;
        const float* S1 = src + (size_t)y1 * sW;
        float* D = dst + (size_t)dy * dW;
        for (int dx = 0; dx < dW; ++dx) {
            const int sx = xofs[dx];
            const int sx1 = sx + 1 < sW ? sx + 1 : sW - 1; /* weight is 0 there */
            const float a0 = alpha[2 * dx], a1 = alpha[2 * dx + 1];
            const float r0 = S0[sx] * a0 + S0[sx1] * a1;
            const float r1 = S1[sx] * a0 + S1[sx1] * a1;
            D[dx] = r0 * b0 + r1 * b1;
        }
    }
    free(xofs);
    free(alpha);
}

/* dsize = Size(saturate_cast<int>(cols*f), saturate_cast<int>(rows*f)) -- cvRound, half to even */
TF_EXPORT void oracle_scaled_size(int sH, int sW, double f, int* dH, int* dW) {
    *dW = (int)lrint(sW * f);
    *dH = (int)lrint(sH * f);
}

/* ------------------------------------------------------------------------------------------------
 * cv::remap(src, dst, mapx, mapy, INTER_CUBIC, BORDER_CONSTANT, 0) for CV_32FC1
 * (imgproc/imgwarp.cpp: remap -> convert float maps to fixed point with INTER_BITS=5, remapBicubic).
 * Verified bit-exact against cv2.remap, oracle/probe_remap.py.
 * ------------------------------------------------------------------------------------------------ */
static float g_cubic[32][4];
static int g_cubic_ready = 0;

static void cubic_tab_init(void) {
    if (g_cubic_ready) return;
    const float A = -0.75f;
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; ++i) {
        float x = i * scale;
        float* c = g_cubic[i];
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
    }
    g_cubic_ready = 1;
}

TF_EXPORT void oracle_cubic_table(float* out128) {
    cubic_tab_init();
    memcpy(out128, g_cubic, sizeof(g_cubic));
}

static inline int cv_round_f(float v) {
    /* cvRound: round half to even; out-of-range / NaN -> INT_MIN like cvtss2si */
    if (!(v > -2147483648.f && v < 2147483648.f)) return INT32_MIN;
    return (int)lrintf(v);
}

static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

static inline float remap_cubic_px(const float* src, int H, int W, float mx, float my) {
    const int ix = cv_round_f(mx * 32.f), iy = cv_round_f(my * 32.f);
    const int sx = sat_short(ix >> 5) - 1, sy = sat_short(iy >> 5) - 1;
    const float* wx = g_cubic[ix & 31];
    const float* wy = g_cubic[iy & 31];
    if ((unsigned)sx < (unsigned)(W - 3 > 0 ? W - 3 : 0) && (unsigned)sy < (unsigned)(H - 3 > 0 ? H - 3 : 0)) {
        const float* S = src + (size_t)sy * W + sx;
        float sum = S[0] * (wy[0] * wx[0]) + S[1] * (wy[0] * wx[1]) + S[2] * (wy[0] * wx[2]) + S[3] * (wy[0] * wx[3]);
        S += W;
        sum += S[0] * (wy[1] * wx[0]) + S[1] * (wy[1] * wx[1]) + S[2] * (wy[1] * wx[2]) + S[3] * (wy[1] * wx[3]);
        S += W;
        sum += S[0] * (wy[2] * wx[0]) + S[1] * (wy[2] * wx[1]) + S[2] * (wy[2] * wx[2]) + S[3] * (wy[2] * wx[3]);
        S += W;
        sum += S[0] * (wy[3] * wx[0]) + S[1] * (wy[3] * wx[1]) + S[2] * (wy[3] * wx[2]) + S[3] * (wy[3] * wx[3]);
        return sum;
    }
    if (sx + 3 < 0 || sx >= W || sy + 3 < 0 || sy >= H) return 0.f;
    float sum = 0.f; /* cval * ONE */
    for (int i = 0; i < 4; ++i) {
        const int yi = sy + i;
        if (yi < 0 || yi >= H) continue;
        const float* S = src + (size_t)yi * W;
        for (int j = 0; j < 4; ++j) {
            const int xj = sx + j;
            if (xj >= 0 && xj < W) sum += S[xj] * (wy[i] * wx[j]); /* (S - cval) * w, cval = 0 */
        }
    }
    return sum;
}

TF_EXPORT void oracle_remap_cubic(const float* src, int H, int W, const float* mapx, const float* mapy,
                                  float* dst, int dH, int dW) {
    cubic_tab_init();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dH; ++y)
        for (int x = 0; x < dW; ++x)
            dst[(size_t)y * dW + x] = remap_cubic_px(src, H, W, mapx[(size_t)y * dW + x], mapy[(size_t)y * dW + x]);
}

/* ------------------------------------------------------------------------------------------------
 * cv::medianBlur(src, dst, ksize) for CV_32FC1, ksize 3 or 5: exact median, BORDER_REPLICATE.
 * ------------------------------------------------------------------------------------------------ */
#define CSWAP(a, b) do { const float lo_ = v[a] < v[b] ? v[a] : v[b]; const float hi_ = v[a] < v[b] ? v[b] : v[a]; v[a] = lo_; v[b] = hi_; } while (0)

/* selection network for the median of 25: 99 compare-exchanges (exhaustively verified on all 2^25 0/1 inputs by
 * tests/test_oracle.py::test_median_networks_zero_one) */
#define MEDIAN25_NETWORK \
    CSWAP(0, 1); CSWAP(3, 4); CSWAP(2, 4); CSWAP(2, 3); CSWAP(6, 7); CSWAP(5, 7); CSWAP(5, 6); CSWAP(9, 10); \
    CSWAP(8, 10); CSWAP(8, 9); CSWAP(12, 13); CSWAP(11, 13); CSWAP(11, 12); CSWAP(15, 16); CSWAP(14, 16); \
    CSWAP(14, 15); CSWAP(18, 19); CSWAP(17, 19); CSWAP(17, 18); CSWAP(21, 22); CSWAP(20, 22); CSWAP(20, 21); \
    CSWAP(23, 24); CSWAP(2, 5); CSWAP(3, 6); CSWAP(0, 6); CSWAP(0, 3); CSWAP(4, 7); CSWAP(1, 7); CSWAP(1, 4); \
    CSWAP(11, 14); CSWAP(8, 14); CSWAP(8, 11); CSWAP(12, 15); CSWAP(9, 15); CSWAP(9, 12); CSWAP(13, 16); \
    CSWAP(10, 16); CSWAP(10, 13); CSWAP(20, 23); CSWAP(17, 23); CSWAP(17, 20); CSWAP(21, 24); CSWAP(18, 24); \
    CSWAP(18, 21); CSWAP(19, 22); CSWAP(8, 17); CSWAP(9, 18); CSWAP(0, 18); CSWAP(0, 9); CSWAP(10, 19); \
    CSWAP(1, 19); CSWAP(1, 10); CSWAP(11, 20); CSWAP(2, 20); CSWAP(2, 11); CSWAP(12, 21); CSWAP(3, 21); \
    CSWAP(3, 12); CSWAP(13, 22); CSWAP(4, 22); CSWAP(4, 13); CSWAP(14, 23); CSWAP(5, 23); CSWAP(5, 14); \
    CSWAP(15, 24); CSWAP(6, 24); CSWAP(6, 15); CSWAP(7, 16); CSWAP(7, 19); CSWAP(13, 21); CSWAP(15, 23); \
    CSWAP(7, 13); CSWAP(7, 15); CSWAP(1, 9); CSWAP(3, 11); CSWAP(5, 17); CSWAP(11, 17); CSWAP(9, 17); \
    CSWAP(4, 10); CSWAP(6, 12); CSWAP(7, 14); CSWAP(4, 6); CSWAP(4, 7); CSWAP(12, 14); CSWAP(10, 14); \
    CSWAP(6, 7); CSWAP(10, 12); CSWAP(6, 10); CSWAP(6, 17); CSWAP(12, 17); CSWAP(7, 17); CSWAP(7, 10); \
    CSWAP(12, 18); CSWAP(7, 12); CSWAP(10, 18); CSWAP(12, 20); CSWAP(10, 20); CSWAP(10, 12);

#define MEDIAN9_NETWORK \
    CSWAP(1, 2); CSWAP(4, 5); CSWAP(7, 8); CSWAP(0, 1); CSWAP(3, 4); CSWAP(6, 7); CSWAP(1, 2); CSWAP(4, 5); \
    CSWAP(7, 8); CSWAP(0, 3); CSWAP(5, 8); CSWAP(4, 7); CSWAP(3, 6); CSWAP(1, 4); CSWAP(2, 5); CSWAP(4, 7); \
    CSWAP(4, 2); CSWAP(6, 4); CSWAP(4, 2);

/* exported for the exhaustive zero-one test of the networks */
TF_EXPORT float oracle_median25(const float* in) { float v[25]; memcpy(v, in, sizeof(v)); MEDIAN25_NETWORK return v[12]; }
TF_EXPORT float oracle_median9(const float* in) { float v[9]; memcpy(v, in, sizeof(v)); MEDIAN9_NETWORK return v[4]; }
/* number of 0/1 input vectors (out of 2^n) on which the network output differs from the true median */
TF_EXPORT long oracle_median_network_failures(int n) {
    long bad = 0;
    const int half = n / 2;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long m = 0; m < (1L << n); ++m) {
        float v[25];
        int ones = 0;
        for (int k = 0; k < n; ++k) { v[k] = (float)((m >> k) & 1); ones += (int)((m >> k) & 1); }
        float got;
        if (n == 25) { MEDIAN25_NETWORK got = v[12]; } else { MEDIAN9_NETWORK got = v[4]; }
        const float want = ones > half ? 1.f : 0.f;
        bad += got != want;
    }
    return bad;
}

TF_EXPORT int oracle_median_blur(const float* src, float* dst, int H, int W, int ksize) {
    if (ksize != 3 && ksize != 5) return -1;
    const int r = ksize / 2;
    const int PW = W + 2 * r;
    /* BORDER_REPLICATE padded copy, so that the hot loop is branch-free and vectorisable */
    float* pad = (float*)malloc(sizeof(float) * (size_t)(H + 2 * r) * PW);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H + 2 * r; ++y) {
        int yy = y - r; yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);
        float* P = pad + (size_t)y * PW;
        const float* S = src + (size_t)yy * W;
        for (int x = 0; x < r; ++x) P[x] = S[0];
        memcpy(P + r, S, sizeof(float) * (size_t)W);
        for (int x = 0; x < r; ++x) P[r + W + x] = S[W - 1];
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        float* D = dst + (size_t)y * W;
        if (ksize == 5) {
            const float *r0 = pad + (size_t)y * PW, *r1 = r0 + PW, *r2 = r1 + PW, *r3 = r2 + PW, *r4 = r3 + PW;
#pragma omp simd
            for (int x = 0; x < W; ++x) {
                float v[25];
                for (int j = 0; j < 5; ++j) {
                    v[j] = r0[x + j]; v[5 + j] = r1[x + j]; v[10 + j] = r2[x + j]; v[15 + j] = r3[x + j];
                    v[20 + j] = r4[x + j];
                }
                MEDIAN25_NETWORK
                D[x] = v[12];
            }
        } else {
            const float *r0 = pad + (size_t)y * PW, *r1 = r0 + PW, *r2 = r1 + PW;
#pragma omp simd
            for (int x = 0; x < W; ++x) {
                float v[9];
                for (int j = 0; j < 3; ++j) { v[j] = r0[x + j]; v[3 + j] = r1[x + j]; v[6 + j] = r2[x + j]; }
                MEDIAN9_NETWORK
                D[x] = v[4];
            }
        }
    }
    free(pad);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * tvl1flow.cpp pointwise steps
 * ------------------------------------------------------------------------------------------------ */

/* centeredGradient: 0.5*(next-prev); one-sided at the borders WITH the 0.5 factor kept */
TF_EXPORT void oracle_centered_gradient(const float* src, int H, int W, float* dx, float* dy) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
        for (int x = 0; x < W; ++x) {
            const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
            dx[(size_t)y * W + x] = 0.5f * (src[(size_t)y * W + xp] - src[(size_t)y * W + xm]);
            dy[(size_t)y * W + x] = 0.5f * (src[(size_t)yp * W + x] - src[(size_t)ym * W + x]);
        }
    }
}

/* forwardGradient: dx = 0 in the last column, dy = 0 in the last row */
static void forward_gradient(const float* u, int H, int W, float* ux, float* uy) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            ux[i] = x < W - 1 ? u[i + 1] - u[i] : 0.f;
            uy[i] = y < H - 1 ? u[i + W] - u[i] : 0.f;
        }
}

/* divergence: backward differences, out-of-range terms dropped on the first row / column only */
static void divergence(const float* v1, const float* v2, int H, int W, float* div) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            float d;
            if (y > 0 && x > 0) {
                const float v1x = v1[i] - v1[i - 1];
                const float v2y = v2[i] - v2[i - W];
                d = v1x + v2y;
            } else if (y == 0 && x > 0) {
                d = v1[i] - v1[i - 1] + v2[i];
            } else if (x == 0 && y > 0) {
                d = v1[i] + v2[i] - v2[i - W];
            } else {
                d = v1[i] + v2[i];
            }
            div[i] = d;
        }
}

/* buildFlowMap + 3x remap + calcGradRho */
TF_EXPORT void oracle_warp_step(const float* I0, const float* I1, const float* I1x, const float* I1y,
                                const float* u1, const float* u2, int H, int W, float* I1wx, float* I1wy,
                                float* grad, float* rho_c) {
    cubic_tab_init();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const float mx = (float)x + u1[i], my = (float)y + u2[i];
            const float w = remap_cubic_px(I1, H, W, mx, my);
            const float wx = remap_cubic_px(I1x, H, W, mx, my);
            const float wy = remap_cubic_px(I1y, H, W, mx, my);
            const float Ix2 = wx * wx, Iy2 = wy * wy;
            I1wx[i] = wx;
            I1wy[i] = wy;
            grad[i] = Ix2 + Iy2;
            rho_c[i] = (w - wx * u1[i] - wy * u2[i] - I0[i]);
        }
}

static void estimate_v(const float* I1wx, const float* I1wy, const float* u1, const float* u2, const float* grad,
                       const float* rho_c, float* v1, float* v2, float l_t, int H, int W) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const float rho = rho_c[i] + (I1wx[i] * u1[i] + I1wy[i] * u2[i]);
            float d1 = 0.f, d2 = 0.f;
            if (rho < -l_t * grad[i]) {
                d1 = l_t * I1wx[i];
                d2 = l_t * I1wy[i];
            } else if (rho > l_t * grad[i]) {
                d1 = -l_t * I1wx[i];
                d2 = -l_t * I1wy[i];
            } else if (grad[i] > FLT_EPSILON) {
                const float fi = -rho / grad[i];
                d1 = fi * I1wx[i];
                d2 = fi * I1wy[i];
            }
            v1[i] = u1[i] + d1;
            v2[i] = u2[i] + d2;
        }
}

/* estimateU; returns the error as the float OpenCV would compare (mode 0) or the float64 sum (mode 1) */
static double estimate_u(const float* v1, const float* v2, const float* div_p1, const float* div_p2, float* u1,
                         float* u2, float theta, int H, int W, int err_mode, float* term_buf) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const float u1k = u1[i], u2k = u2[i];
            u1[i] = v1[i] + theta * div_p1[i];
            u2[i] = v2[i] + theta * div_p2[i];
            term_buf[i] = (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
        }
    const size_t n = (size_t)H * W;
    if (err_mode == 0) {
        float e = 0.f;
        for (size_t i = 0; i < n; ++i) e += term_buf[i];
        return (double)e;
    }
    double e = 0.0;
    for (size_t i = 0; i < n; ++i) e += (double)term_buf[i];
    return e;
}

static void estimate_dual(const float* u1x, const float* u1y, const float* u2x, const float* u2y, float* p11,
                          float* p12, float* p21, float* p22, float taut, int H, int W) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            /* static_cast<float>(hypot(a, b)); glibc hypotf == (float)sqrt((double)a*a + (double)b*b) */
            const float g1 = (float)sqrt((double)u1x[i] * (double)u1x[i] + (double)u1y[i] * (double)u1y[i]);
            const float g2 = (float)sqrt((double)u2x[i] * (double)u2x[i] + (double)u2y[i] * (double)u2y[i]);
            const float ng1 = 1.0f + taut * g1;
            const float ng2 = 1.0f + taut * g2;
            p11[i] = (p11[i] + taut * u1x[i]) / ng1;
            p12[i] = (p12[i] + taut * u1y[i]) / ng1;
            p21[i] = (p21[i] + taut * u2x[i]) / ng2;
            p22[i] = (p22[i] + taut * u2y[i]) / ng2;
        }
}

typedef struct {
    float *I1x, *I1y, *I1wx, *I1wy, *grad, *rho_c, *v1, *v2, *p11, *p12, *p21, *p22, *div_p1, *div_p2;
    float *u1x, *u1y, *u2x, *u2y, *term;
} scratch_t;

/* one inner iteration on caller-owned planes (exported so that the GPU kernel can be unit-tested) */
TF_EXPORT double oracle_inner_iteration(const float* I1wx, const float* I1wy, const float* grad, const float* rho_c,
                                        float* u1, float* u2, float* p11, float* p12, float* p21, float* p22, int H,
                                        int W, float l_t, float theta, float taut, int err_mode) {
    const size_t n = (size_t)H * W;
    float* buf = (float*)malloc(sizeof(float) * n * 9);
    float *v1 = buf, *v2 = buf + n, *d1 = buf + 2 * n, *d2 = buf + 3 * n, *u1x = buf + 4 * n, *u1y = buf + 5 * n,
          *u2x = buf + 6 * n, *u2y = buf + 7 * n, *term = buf + 8 * n;
    estimate_v(I1wx, I1wy, u1, u2, grad, rho_c, v1, v2, l_t, H, W);
    divergence(p11, p12, H, W, d1);
    divergence(p21, p22, H, W, d2);
    const double err = estimate_u(v1, v2, d1, d2, u1, u2, theta, H, W, err_mode, term);
    forward_gradient(u1, H, W, u1x, u1y);
    forward_gradient(u2, H, W, u2x, u2y);
    estimate_dual(u1x, u1y, u2x, u2y, p11, p12, p21, p22, taut, H, W);
    free(buf);
    return err;
}

/* procOneScale; counters[0..2] += inner iterations, median passes, warps executed */
static int g_trace = 0;
TF_EXPORT void oracle_set_trace(int on) { g_trace = on; }

static void proc_one_scale(const tvl1_oracle_params* P, const float* I0, const float* I1, float* u1, float* u2,
                           int H, int W, scratch_t* S, int* counters) {
    const size_t n = (size_t)H * W;
    const float scaledEpsilon = (float)(P->epsilon * P->epsilon * (double)(H * W));
    oracle_centered_gradient(I1, H, W, S->I1x, S->I1y);
    memset(S->p11, 0, n * sizeof(float));
    memset(S->p12, 0, n * sizeof(float));
    memset(S->p21, 0, n * sizeof(float));
    memset(S->p22, 0, n * sizeof(float));
    const float l_t = (float)(P->lambda * P->theta);
    const float taut = (float)(P->tau / P->theta);
    const float theta = (float)P->theta;
    for (int warpings = 0; warpings < P->warps; ++warpings) {
        oracle_warp_step(I0, I1, S->I1x, S->I1y, u1, u2, H, W, S->I1wx, S->I1wy, S->grad, S->rho_c);
        counters[2]++;
        float error = FLT_MAX;
        for (int n_outer = 0; error > scaledEpsilon && n_outer < P->outer_iterations; ++n_outer) {
            if (P->median_filtering > 1) {
                oracle_median_blur(u1, u1, H, W, P->median_filtering);
                oracle_median_blur(u2, u2, H, W, P->median_filtering);
                counters[1]++;
            }
            for (int n_inner = 0; error > scaledEpsilon && n_inner < P->inner_iterations; ++n_inner) {
                estimate_v(S->I1wx, S->I1wy, u1, u2, S->grad, S->rho_c, S->v1, S->v2, l_t, H, W);
                divergence(S->p11, S->p12, H, W, S->div_p1);
                divergence(S->p21, S->p22, H, W, S->div_p2);
                error = (float)estimate_u(S->v1, S->v2, S->div_p1, S->div_p2, u1, u2, theta, H, W, P->err_mode,
                                          S->term);
                forward_gradient(u1, H, W, S->u1x, S->u1y);
                forward_gradient(u2, H, W, S->u2x, S->u2y);
                estimate_dual(S->u1x, S->u1y, S->u2x, S->u2y, S->p11, S->p12, S->p21, S->p22, taut, H, W);
                counters[0]++;
                /* diagnostics for the engine's speculation policy (tools/spec_policy.py): error / threshold */
                if (g_trace) fprintf(stderr, "TRACE %dx%d warp %d outer %d inner %d err_over_eps %.6g\n", H, W, warpings,
                                     n_outer, n_inner, (double)error / (double)scaledEpsilon);
            }
        }
    }
}

/*
 * OpticalFlowDual_TVL1::calc.  I0/I1: H x W, uint8 (is_f32 = 0, multiplier 1) or float32 (is_f32 = 1,
 * multiplier 255).  flow: H x W x 2 float32 (channel 0 = x displacement, 1 = y).  counters: nscales x 3 int32
 * (inner iterations, median passes, warps) per level, level 0 = finest; may be NULL.  Returns the number of
 * levels actually used (the pyramid stops when a side drops below 16 px), or a negative error code.
 */
TF_EXPORT int tvl1_oracle_calc(const tvl1_oracle_params* P, const void* I0v, const void* I1v, int is_f32, int H,
                               int W, float* flow, int* counters) {
    if (!P || !I0v || !I1v || !flow || H <= 0 || W <= 0) return -1;
    if (P->nscales <= 0 || P->nscales > 32) return -2;
    if (P->median_filtering > 1 && P->median_filtering != 3 && P->median_filtering != 5) return -3;
    int nscales = P->nscales;
    float* I0s[32];
    float* I1s[32];
    float* u1s[32];
    float* u2s[32];
    int Hs[32], Ws[32];
    const size_t n0 = (size_t)H * W;
    Hs[0] = H; Ws[0] = W;
    I0s[0] = (float*)malloc(n0 * sizeof(float));
    I1s[0] = (float*)malloc(n0 * sizeof(float));
    if (is_f32) {
        const float* a = (const float*)I0v; const float* b = (const float*)I1v;
        for (size_t i = 0; i < n0; ++i) { I0s[0][i] = a[i] * 255.0f; I1s[0][i] = b[i] * 255.0f; }
    } else {
        const uint8_t* a = (const uint8_t*)I0v; const uint8_t* b = (const uint8_t*)I1v;
        for (size_t i = 0; i < n0; ++i) { I0s[0][i] = (float)a[i]; I1s[0][i] = (float)b[i]; }
    }
    u1s[0] = (float*)malloc(n0 * sizeof(float));
    u2s[0] = (float*)malloc(n0 * sizeof(float));
    int allocated = 1;
    for (int s = 1; s < nscales; ++s) {
        oracle_scaled_size(Hs[s - 1], Ws[s - 1], P->scale_step, &Hs[s], &Ws[s]);
        if (Hs[s] <= 0 || Ws[s] <= 0) { nscales = s; break; }
        const size_t ns = (size_t)Hs[s] * Ws[s];
        I0s[s] = (float*)malloc(ns * sizeof(float));
        I1s[s] = (float*)malloc(ns * sizeof(float));
        u1s[s] = (float*)malloc(ns * sizeof(float));
        u2s[s] = (float*)malloc(ns * sizeof(float));
        allocated = s + 1;
        const double sc = 1.0 / P->scale_step;
        oracle_resize_linear(I0s[s - 1], Hs[s - 1], Ws[s - 1], I0s[s], Hs[s], Ws[s], sc, sc);
        oracle_resize_linear(I1s[s - 1], Hs[s - 1], Ws[s - 1], I1s[s], Hs[s], Ws[s], sc, sc);
        if (Ws[s] < 16 || Hs[s] < 16) { nscales = s; break; }
    }
    memset(u1s[nscales - 1], 0, (size_t)Hs[nscales - 1] * Ws[nscales - 1] * sizeof(float));
    memset(u2s[nscales - 1], 0, (size_t)Hs[nscales - 1] * Ws[nscales - 1] * sizeof(float));

    scratch_t S;
    float** planes = (float**)&S;
    const int nplanes = (int)(sizeof(scratch_t) / sizeof(float*));
    for (int k = 0; k < nplanes; ++k) planes[k] = (float*)malloc(n0 * sizeof(float));
    if (counters) memset(counters, 0, sizeof(int) * 3 * (size_t)P->nscales);

    int dummy[3];
    for (int s = nscales - 1; s >= 0; --s) {
        proc_one_scale(P, I0s[s], I1s[s], u1s[s], u2s[s], Hs[s], Ws[s], &S, counters ? counters + 3 * s : dummy);
        if (s == 0) break;
        /* resize(u(s), u(s-1), size(s-1), 0, 0, INTER_LINEAR): inv_scale = dsize/ssize, scale = 1./inv_scale */
        const double sx = 1.0 / ((double)Ws[s - 1] / (double)Ws[s]);
        const double sy = 1.0 / ((double)Hs[s - 1] / (double)Hs[s]);
        oracle_resize_linear(u1s[s], Hs[s], Ws[s], u1s[s - 1], Hs[s - 1], Ws[s - 1], sx, sy);
        oracle_resize_linear(u2s[s], Hs[s], Ws[s], u2s[s - 1], Hs[s - 1], Ws[s - 1], sx, sy);
        /* multiply(u, Scalar::all(1 / scaleStep), u) -- scalar converted to float for CV_32F */
        const float mul = (float)(1.0 / P->scale_step);
        const size_t nf = (size_t)Hs[s - 1] * Ws[s - 1];
        for (size_t i = 0; i < nf; ++i) { u1s[s - 1][i] = u1s[s - 1][i] * mul; u2s[s - 1][i] = u2s[s - 1][i] * mul; }
    }
    for (size_t i = 0; i < n0; ++i) { flow[2 * i] = u1s[0][i]; flow[2 * i + 1] = u2s[0][i]; }

    for (int k = 0; k < nplanes; ++k) free(planes[k]);
    for (int s = 0; s < allocated; ++s) { free(I0s[s]); free(I1s[s]); free(u1s[s]); free(u2s[s]); }
    return nscales;
}

TF_EXPORT void tvl1_oracle_default_params(tvl1_oracle_params* P) {
    P->tau = 0.25; P->lambda = 0.15; P->theta = 0.3; P->epsilon = 0.01; P->scale_step = 0.8;
    P->nscales = 5; P->warps = 5; P->inner_iterations = 30; P->outer_iterations = 10;
    P->median_filtering = 5; P->err_mode = 0;
}
