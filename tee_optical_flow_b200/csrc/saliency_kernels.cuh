// saliency_kernels.cuh -- StaticSaliencyFineGrained on the GPU: the reference's input stage when no_saliency=False
// (optical_flow/calculate_optical_flow.py:560, :586: saliency_obj.computeSaliency(frame) per frame, the float32 map
// goes into the TV-L1 solver as the image).  Restates opencv_contrib modules/saliency/src/staticSaliencyFineGrained.cpp
// (PARITY UNPINNED: that module is in neither this image's cv2 nor /root/reference; oracle/saliency_ref.py is the
// same restatement on genuine cv2 primitives).  Integer / byte work is bit-exact against that oracle:
//   gray   = cvtColor(BGR2GRAY) of the RGB frame (the reference's quirk): (3735 c0 + 19235 c1 + 9798 c2 + 2^14) >> 15
//   blur   = 2 x GaussianBlur(5x5, sigma 0): [1 4 6 4 1]/16 per axis, REFLECT_101, (sum + 128) >> 8
//   integ  = integral(gray, CV_32F): row prefixes are exact integers; the column accumulation rounds serially in y
//   scales = six neighbourhoods: mean of the surround from the integral image, on = gray - mean, off = mean - gray
//            (uchar) truncated, summed over the scales; then two max-normalisations with (uchar) truncation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace teeflow {

__global__ void sal_gray_kernel(const uint8_t* __restrict__ rgb, int n_px, uint8_t* __restrict__ gray) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
        const int c0 = rgb[3 * (size_t)i], c1 = rgb[3 * (size_t)i + 1], c2 = rgb[3 * (size_t)i + 2];
        gray[i] = (uint8_t)((c0 * 3735 + c1 * 19235 + c2 * 9798 + (1 << 14)) >> 15);
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {   // BORDER_REFLECT_101, |overshoot| <= 2 < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__global__ void sal_blur5_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n_frames, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (x >= W || y >= H || f >= n_frames) return;
    const uint8_t* S = src + (size_t)f * H * W;
    const int w[5] = {1, 4, 6, 4, 1};
    int acc = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint8_t* row = S + (size_t)reflect101(y + i - 2, H) * W;
        int r = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j) r += w[j] * row[reflect101(x + j - 2, W)];
        acc += w[i] * r;
    }
    dst[(size_t)f * H * W + (size_t)y * W + x] = (uint8_t)((acc + 128) >> 8);
}

// row prefixes (exact: <= 255 * W < 2^24), one warp per image row
__global__ void sal_rowprefix_kernel(const uint8_t* __restrict__ gray, float* __restrict__ prefix, int n_rows, int W) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const uint8_t* g = gray + (size_t)row * W;
    float* p = prefix + (size_t)row * W;
    int carry = 0;
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        int v = x < W ? g[x] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (x < W) p[x] = (float)(carry + v);
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
}

// integ[y+1][x+1] = integ[y][x+1] + prefix[y][x] in float32, serial in y (the only place the integral image rounds)
__global__ void sal_integral_kernel(const float* __restrict__ prefix, float* __restrict__ integ, int n_frames, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (x > W || f >= n_frames) return;
    float* I = integ + (size_t)f * (H + 1) * (W + 1);
    I[x] = 0.f;
    if (x == 0) { for (int y = 1; y <= H; ++y) I[(size_t)y * (W + 1)] = 0.f; return; }
    const float* P = prefix + (size_t)f * H * W + (x - 1);
    float s = 0.f;
    for (int y = 0; y < H; ++y) {
        s = s + P[(size_t)y * W];
        I[(size_t)(y + 1) * (W + 1) + x] = s;
    }
}

struct SalMax { int sum_on, sum_off, on8, off8; };

__device__ __forceinline__ int clampi_s(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// six centre-surround scales per pixel, on / off responses summed over the scales (getIntensityScaled + mixScales' sums)
__global__ void sal_scales_kernel(const uint8_t* __restrict__ gray, const float* __restrict__ integ, int n_frames, int H,
                                  int W, uint16_t* __restrict__ sum_on, uint16_t* __restrict__ sum_off, SalMax* __restrict__ mx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    int son = 0, soff = 0;
    if (x < W && y < H && f < n_frames) {
        const float* I = integ + (size_t)f * (H + 1) * (W + 1);
        const int gi = gray[(size_t)f * H * W + (size_t)y * W + x];
        const float g = (float)gi;
        const int nb[6] = {12, 24, 48, 28, 56, 112};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int n = nb[k];
            const int p1x = clampi_s(x - n + 1, 0, W), p1y = clampi_s(y - n + 1, 0, H);
            const int p2x = clampi_s(x + n + 1, 0, W), p2y = clampi_s(y + n + 1, 0, H);
            float v = __fadd_rn(I[(size_t)p2y * (W + 1) + p2x], I[(size_t)p1y * (W + 1) + p1x]);
            v = __fsub_rn(v, I[(size_t)p2y * (W + 1) + p1x]);
            v = __fsub_rn(v, I[(size_t)p1y * (W + 1) + p2x]);
            v = __fdiv_rn(__fsub_rn(v, g), (float)((p2x - p1x) * (p2y - p1y) - 1));
            const float mon = __fsub_rn(g, v), moff = __fsub_rn(v, g);
            if (mon > 0.f) son += (int)mon;        // (uchar) truncation, <= 255
            if (moff > 0.f) soff += (int)moff;
        }
        sum_on[(size_t)f * H * W + (size_t)y * W + x] = (uint16_t)son;
        sum_off[(size_t)f * H * W + (size_t)y * W + x] = (uint16_t)soff;
    }
    // per-frame maxima of the sums
    for (int o = 16; o > 0; o >>= 1) {
        son = max(son, __shfl_xor_sync(0xffffffffu, son, o));
        soff = max(soff, __shfl_xor_sync(0xffffffffu, soff, o));
    }
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && f < n_frames) {
        if (son > 0) atomicMax(&mx[f].sum_on, son);
        if (soff > 0) atomicMax(&mx[f].sum_off, soff);
    }
}

__device__ __forceinline__ int norm255(int num, int den) {   // (uchar)(255. * (float)(num / (float)den)); den == 0 -> 0
    if (den == 0) return 0;
    const float q = __fdiv_rn((float)num, (float)den);
    return (int)(255.0 * (double)q) & 255;
}

__global__ void sal_mix_kernel(const uint16_t* __restrict__ sum_on, const uint16_t* __restrict__ sum_off, int n_frames,
                               int npx, uint8_t* __restrict__ on8, uint8_t* __restrict__ off8, SalMax* __restrict__ mx) {
    const int f = blockIdx.y;
    if (f >= n_frames) return;
    const int m_on = mx[f].sum_on, m_off = mx[f].sum_off;
    int a_max = 0, b_max = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        const int a = norm255(sum_on[(size_t)f * npx + i], m_on), b = norm255(sum_off[(size_t)f * npx + i], m_off);
        on8[(size_t)f * npx + i] = (uint8_t)a; off8[(size_t)f * npx + i] = (uint8_t)b;
        a_max = max(a_max, a); b_max = max(b_max, b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a_max = max(a_max, __shfl_xor_sync(0xffffffffu, a_max, o));
        b_max = max(b_max, __shfl_xor_sync(0xffffffffu, b_max, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (a_max > 0) atomicMax(&mx[f].on8, a_max);
        if (b_max > 0) atomicMax(&mx[f].off8, b_max);
    }
}

// mixOnOff + convertTo(CV_32F, 1/255.f)
__global__ void sal_final_kernel(const uint8_t* __restrict__ on8, const uint8_t* __restrict__ off8, int n_frames, int npx,
                                 const SalMax* __restrict__ mx, float* __restrict__ out, uint8_t* __restrict__ out_u8) {
    const int f = blockIdx.y;
    if (f >= n_frames) return;
    const int mv = max(mx[f].on8, mx[f].off8);
    const double den = (double)(float)mv;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        int v = 0;
        if (mv > 0) {
            const float s = (float)((int)on8[(size_t)f * npx + i] + (int)off8[(size_t)f * npx + i]);
            v = (int)(255.0 * (double)s / den) & 255;
        }
        if (out) out[(size_t)f * npx + i] = __fmul_rn((float)v, 1.0f / 255.0f);
        if (out_u8) out_u8[(size_t)f * npx + i] = (uint8_t)v;
    }
}

}  // namespace teeflow
